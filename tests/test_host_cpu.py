"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol of
include/almpc.h, the ctypes mirror of almpc_desc has the C layout, the MPC module keeps the
reference's state semantics, and the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import types

import pytest
import torch

from deq_mpc_corl_b200 import _lib
from deq_mpc_corl_b200.al_mpc import MPC
from deq_mpc_corl_b200.al_utils import QuadCost
from deq_mpc_corl_b200.dynamics import make as make_dyn, resolve
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "almpc.h")


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(almpc_[a-z0-9_]+)\s*\(", src))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.almpc_version()


def test_desc_layout_matches_header(tmp_path):
    csrc = tmp_path / "sz.c"
    csrc.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "almpc.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu",'
                    "sizeof(almpc_desc), offsetof(almpc_desc, rho_mult), offsetof(almpc_desc, dyn_params),"
                    "offsetof(almpc_desc, u_lower), offsetof(almpc_desc, flags), offsetof(almpc_desc, select),"
                    "offsetof(almpc_desc, obstacles), offsetof(almpc_desc, profile_ms),"
                    "offsetof(almpc_desc, stop_reduce_user));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(csrc), "-o", str(exe)])
    size, o_mult, o_par, o_lo, o_flags, o_sel, o_obs, o_prof, o_user = map(int, subprocess.check_output([str(exe)]).split())
    assert size == ctypes.sizeof(_lib.Desc)
    assert o_sel == _lib.Desc.select.offset and o_obs == _lib.Desc.obstacles.offset
    assert o_prof == _lib.Desc.profile_ms.offset and o_user == _lib.Desc.stop_reduce_user.offset
    assert o_mult == _lib.Desc.rho_mult.offset and o_par == _lib.Desc.dyn_params.offset
    assert o_lo == _lib.Desc.u_lower.offset and o_flags == _lib.Desc.flags.offset


def test_model_dims_and_sizes():
    lib = _lib.load()
    for name, spec in MODEL_SPECS.items():
        dyn = make_dyn(name)
        nx, nu = ctypes.c_int32(), ctypes.c_int32()
        assert lib.almpc_model_dims(dyn.dyn_id, ctypes.byref(nx), ctypes.byref(nu)) == 0
        assert (nx.value, nu.value) == (spec["nx"], spec["nu"]) == (dyn.nx, dyn.nu)
        d = _lib.Desc()
        d.B, d.T, d.dyn_id = 10, 5, dyn.dyn_id
        n = nx.value + nu.value
        assert lib.almpc_factor_elems(ctypes.byref(d)) == 5 * n * n + 4 * nx.value * n
        assert lib.almpc_workspace_bytes(ctypes.byref(d)) > 0
    assert lib.almpc_model_dims(99, None, None) != 0
    assert b"dyn_id" in lib.almpc_last_error()


def test_bad_descriptor_is_rejected_without_touching_the_gpu():
    lib = _lib.load()
    d = _lib.Desc()
    d.B, d.T, d.dyn_id, d.al_iter, d.max_newton = 4, 1, 0, 2, 4  # T too small
    rc = lib.almpc_forward_f64(ctypes.byref(d), *([None] * 15))
    assert rc < 0 and b"horizon" in lib.almpc_last_error()
    d.T, d.al_iter = 5, 0
    assert lib.almpc_forward_f64(ctypes.byref(d), *([None] * 15)) < 0
    d.al_iter = 2
    assert lib.almpc_forward_f64(ctypes.byref(d), *([None] * 15)) == -10  # null buffers
    d.B = 0
    assert lib.almpc_forward_f64(ctypes.byref(d), *([None] * 15)) == 0  # empty batch is a no-op


def test_quadrotor_param_block_matches_oracle_constants():
    from oracle.dynamics_oracle import FlyingCartpole, RexQuadrotor

    o = RexQuadrotor()
    p = make_dyn("rex_quadrotor").param_block()
    assert p[0] == o.dt and p[1] == o.act_scale and p[3] == o.kf and p[5] == o.bf and p[6] == o.km
    assert p[8:11] == (o.mass * o.g).double().tolist() and p[11:14] == o.Bf.double().tolist()
    assert p[14:23] == o.J.double().flatten().tolist() and p[23:32] == o.Jinv.double().flatten().tolist()
    arm = (o.motor_dist * o.ss.double())[:, :2].flatten().tolist()
    assert p[32:40] == arm
    of = FlyingCartpole()
    pf = make_dyn("flying_cartpole").param_block()
    assert pf[2] == of.u_hover_offset and pf[7] == 1.0 / of.mass and pf[41] == 2.0
    assert abs(MODEL_SPECS["flying_cartpole"]["u_upper"][0] - of.u_upper[0]) == 0.0


def test_mpc_state_semantics_and_cpu_refusal():
    inp = make_inputs("cartpole1l", 5, 6, 1)
    ctrl = MPC(4, 1, 5, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=6, exit_unconverged=False, eps=1e-2,
               backprop=False, verbose=0, solver_type="dense", dtype=torch.float64, state_estimator=False)
    assert ctrl.neq == 20 and ctrl.nineq == 10 and ctrl.al_iter == 2 and ctrl.rho_max == 1e8
    ctrl.reinitialize(inp["x_ref"], None)
    assert ctrl.x_init is None and ctrl.u_init is None and not ctrl.warm_starting
    assert ctrl.rho_prev.shape == (6, 1) and ctrl.rho_prev.dtype == torch.float32 and (ctrl.rho_prev == 1).all()
    assert ctrl.lamda_prev.shape == (6, 30) and (ctrl.lamda_prev == 0).all()
    ctrl.rho_prev = ctrl.rho_prev * 1e6
    ctrl.lamda_prev = ctrl.lamda_prev + 1
    ctrl.warm_start_initialize(inp["x_ref"], inp["u_ref"], types.SimpleNamespace(rho_init_max=1e4))
    assert ctrl.warm_starting and (ctrl.lamda_prev == 0).all() and (ctrl.rho_prev == 1e4).all()
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
    with pytest.raises(_lib.AlmpcError, match="CUDA only"):
        ctrl(inp["x0"], QuadCost(Qd, q, None), make_dyn("cartpole1l"), None)
    with pytest.raises(_lib.AlmpcError):
        MPC(4, 1, 5, u_lower=inp["u_lower"], u_upper=inp["u_upper"], add_goal_constraint=True)
    # state-estimator variant (AL_mpc.py:186-199): the T-1 dynamics blocks only, no bound rows
    se = MPC(4, 1, 5, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=6, state_estimator=True)
    assert se.neq == 16 and se.nineq == 0
    se.reinitialize(inp["x_ref"], None)
    assert se.lamda_prev.shape == (6, 16)


def test_get_cost_matches_reference_formula():
    from deq_mpc_corl_b200.al_utils import compute_cost, compute_cost_gradient

    g = torch.Generator().manual_seed(9)
    B, T, n = 4, 5, 3
    xu, Q, q, f = (torch.randn(B, T, n, generator=g), torch.rand(B, T, n, generator=g), torch.randn(B, T, n, generator=g),
                   torch.randn(B, T, generator=g))
    ref = torch.stack([sum(0.5 * xu[b, t] @ torch.diag(Q[b, t]) @ xu[b, t] + q[b, t] @ xu[b, t] + f[b, t] for t in range(T))
                       for b in range(B)])
    assert torch.allclose(compute_cost(xu, Q, q, f), ref, atol=1e-5)
    assert torch.allclose(compute_cost_gradient(xu, Q, q), Q * xu + q)
    ctrl = MPC(2, 1, T, u_lower=torch.tensor([-1.0]), u_upper=torch.tensor([1.0]), n_batch=B)
    ctrl.x_init, ctrl.u_init = xu[..., :2], xu[..., 2:]
    got = ctrl.get_cost(QuadCost(torch.diag_embed(Q), q, f))
    assert torch.allclose(got, ref, atol=1e-5)


def test_rollout_with_lindx_matches_reference_formula():
    from deq_mpc_corl_b200.al_utils import LinDx

    g = torch.Generator().manual_seed(4)
    B, T, nx, nu = 3, 4, 2, 1
    F, f = torch.randn(B, T - 1, nx, nx + nu, generator=g), torch.randn(B, T - 1, nx, generator=g)
    x0, u = torch.randn(B, nx, generator=g), torch.randn(B, T, nu, generator=g)
    ctrl = MPC(nx, nu, T, u_lower=torch.tensor([-1.0]), u_upper=torch.tensor([1.0]), n_batch=B)
    xs = ctrl.rollout(x0, u, LinDx(F, f))
    assert xs.shape == (B, T, nx) and torch.equal(xs[:, 0], x0)
    for t in range(T - 1):  # AL_mpc.py:527-529: new_x = bmv(F[:, t], cat[x_t, u_t]) + f[:, t]
        ref = torch.einsum("bij,bj->bi", F[:, t], torch.cat([xs[:, t], u[:, t]], -1)) + f[:, t]
        assert torch.allclose(xs[:, t + 1], ref, atol=1e-6)


def test_resolve_recognises_reference_style_objects():
    class RexQuadrotor_dynamics:  # attribute-compatible stand-in for the reference class
        m, motor_dist, kf, bf, km, dt, act_scale = 2.0, 0.28, 0.0244101, -30.48576, 0.00029958, 0.05, 100.0
        J = torch.tensor([[0.01566089, 0.00000318037, 0.0], [0.00000318037, 0.01562078, 0.0], [0.0, 0.0, 0.02226868]]).unsqueeze(0)
        g = torch.tensor([[0.0, 0.0, -9.81]])

    native = resolve(RexQuadrotor_dynamics())
    ours = make_dyn("rex_quadrotor")
    assert native.dyn_id == ours.dyn_id
    assert max(abs(a - b) for a, b in zip(native.param_block(), ours.param_block())) < 1e-9
    with pytest.raises(_lib.AlmpcError, match="no native CUDA functor"):
        resolve(lambda x, u: x)


def test_problem_generator_is_deterministic_and_shaped():
    a, b = make_inputs("rex_quadrotor", 5, 7, 3), make_inputs("rex_quadrotor", 5, 7, 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["x0"].shape == (7, 12) and a["x_ref"].shape == (7, 5, 12) and a["u_ref"].shape == (7, 5, 4)
    assert a["x0"].dtype == torch.float32 and torch.equal(a["x_ref"][:, 0], a["x0"])
    Qd, q, f = tracking_cost(a["Qd"], a["x_ref"], a["u_ref"])
    assert Qd.shape == (7, 5, 16) and q.dtype == torch.float64 and f.shape == (7, 5)


def test_newton_al_context_from_reference_style_closures():
    """NewtonAL learns the dynamics functor and the control bounds from the closure of the reference's own
    lambdas (AL_mpc.py:300-303 close over `self` and `dx_`) or from an explicit almpc_context."""
    from deq_mpc_corl_b200.al_utils import NewtonContext, _context_from

    dyn = make_dyn("cartpole1l")

    class FakeMPC:  # what `self` looks like inside MPC.al_solve
        n_ctrl = 1
        u_lower = torch.tensor([-100.0])
        u_upper = torch.tensor([100.0])

    def al_solve_like(self, dx_):
        x0 = rho = None
        merit_fn = lambda xi, Qi, qi, yi, x0i=x0, rhoi=rho, grad=False: (self, dx_)  # noqa: E731
        dyn_fn = lambda xi: (self, dx_, x0)  # noqa: E731
        return merit_fn, dyn_fn

    merit_fn, dyn_fn = al_solve_like(FakeMPC(), dyn)
    ctx = _context_from(merit_fn, dyn_fn, None)
    assert resolve(ctx.dynamics).dyn_id == dyn.dyn_id
    assert float(torch.as_tensor(ctx.u_upper)[0]) == 100.0 and float(torch.as_tensor(ctx.u_lower)[0]) == -100.0

    def explicit(*a):
        return None

    explicit.almpc_context = NewtonContext(dyn, [-1.0], [1.0])
    assert _context_from(explicit).u_upper == [1.0]
    with pytest.raises(_lib.AlmpcError, match="cannot tell"):
        _context_from(lambda x: x)


def test_tracking_mpc_refuses_non_al_paths():
    from deq_mpc_corl_b200.tracking import Tracking_MPC

    env = types.SimpleNamespace(nx=2, nu=1, nq=1, dt=0.05, dynamics=make_dyn("pendulum"),
                                action_space=types.SimpleNamespace(low=[-2.0], high=[2.0]))
    base = dict(T=5, bsz=4, Q=torch.ones(2), R=torch.ones(1) * 1e-3, dtype="double", device="cpu")
    with pytest.raises(_lib.AlmpcError):
        Tracking_MPC(types.SimpleNamespace(solver_type="ip", env="pendulum", **base), env)
    # "obstacles" in args.env selects Obstacle_MPC (policies.py:1180-1198); the state-estimator variant too
    env_obs = types.SimpleNamespace(obstacle_positions=torch.zeros(40, 3), obstacle_radius=0.5, **vars(env))
    to = Tracking_MPC(types.SimpleNamespace(solver_type="al", env="FlyingCartpole_obstacles", **base), env_obs)
    assert type(to.ctrl).__name__ == "Obstacle_MPC" and to.ctrl.nineq == 5 * (2 + 4)
    se = Tracking_MPC(types.SimpleNamespace(solver_type="al", env="pendulum", **base), env, state_estimator=True)
    assert se.ctrl.state_estimator and se.ctrl.neq == 8 and se.ctrl.nineq == 0
    tm = Tracking_MPC(types.SimpleNamespace(solver_type="al", env="pendulum", **base), env)
    # fp32 R rounded like the reference does (policies.py:1148: cat of fp32 tensors, then .to(dtype))
    assert tm.Q.tolist() == [1.0, 1.0, float(torch.tensor(1e-3))] and tm.ctrl.neq == 10 and tm.ctrl.nineq == 10
    with pytest.raises(_lib.AlmpcError, match="CUDA only"):
        tm(torch.zeros(4, 2), None, torch.zeros(4, 5, 2), torch.zeros(4, 5, 1))
