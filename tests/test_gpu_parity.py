"""GPU parity tests (run on the B200 box): the CUDA path, reached through the product API / C ABI,
against (1) golden outputs of the unmodified reference, (2) the CPU oracle on seeded inputs, and
(3) size-independent properties at BASELINE.json's full batch size.

Tolerances (north_star): 1e-6 relative in fp64, 1e-4 in fp32 *where the problem is well enough
conditioned for two correct implementations to agree that closely*; cond(H) ~ rho/R, so the fp64
budget is widened with the largest penalty used in the call exactly as for the oracle-vs-reference
check (tests/test_oracle_golden.py::tolerances), and fp32 is only asserted in the rho <= 1e2 regime
(SURVEY.md A.5)."""
import os
import types

import pytest
import torch

from golden_util import Golden, list_golden, list_tracking_golden, max_rel_err_rows, rel_err

pytestmark = pytest.mark.gpu

cuda = torch.cuda.is_available()
if cuda:
    import deq_mpc_corl_b200 as pkg
    from deq_mpc_corl_b200 import _lib
    from deq_mpc_corl_b200.al_mpc import MPC
    from deq_mpc_corl_b200.al_utils import QuadCost
    from deq_mpc_corl_b200.dynamics import make as make_dyn
    from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost

DEV = "cuda:0"


def tolerances(rho_out_max, dtype, c=None):
    rho_used = rho_out_max / 10.0
    if dtype != torch.float64:
        # north_star: 1e-4 in fp32.  Asserted as such for the solution in the regime where fp32 can deliver it
        # (rho <= 1e2, i.e. a cold al_iter = 2 call: SURVEY A.5 — from rho ~ 1e4 on H is no longer numerically SPD in
        # fp32 and the reference's own behaviour is backend-specific).  Multipliers and gradients amplify the fp32
        # rounding of the factorisation by cond(H): the fp32 goldens therefore carry the SAME scenario run by the
        # reference in fp64 (keys ref64_*, oracle/gen_golden.py), and the budget is max(1e-4, 2 x the distance
        # between the reference's own fp32 and fp64 results) — measured 2.8e-4 (lam) and 6.5e-4 (q_grad) on
        # quad_T5_f32_cold, where two fp32 runs of the unmodified reference differ by 2.5e-4 in q_grad between
        # themselves.  Measured errors per golden: profiles/r02_parity_errors.md.
        tol = dict(xu=max(1e-4, 1e-6 * rho_used), lam=max(1e-4, 1e-5 * rho_used), grad=max(1e-4, 1e-5 * rho_used))
        if c is not None and "ref64_xu64" in c:
            floor = dict(xu=rel_err(c["xu64"], c["ref64_xu64"]), lam=rel_err(c["lam_out"], c["ref64_lam_out"]),
                         grad=max(rel_err(c["q_grad"], c["ref64_q_grad"]), rel_err(c["Q_grad"], c["ref64_Q_grad"])))
            tol = {k: max(tol[k], 2 * floor[k]) if k != "xu" else tol[k] for k in tol}
            # x alone (fp32 output; the iterate's norm is dominated by the controls): same rule
            tol["x"] = max(tol["xu"], 2 * rel_err(c["x"], c["ref64_x"]))
            tol["reference_fp32_vs_fp64"] = floor
        return tol
    # x, u: the 1e-6 relative bar of the task, flat (measured worst over all goldens: 9e-9).  lam = lam + rho c
    # amplifies an iterate difference by rho, and cond(H) ~ rho / R limits the agreement of any two correct
    # factorisations, so the multiplier and gradient budgets grow with the largest penalty used in the call.  The
    # constants are ~5x the worst measured ratio error / rho over all goldens (profiles/r02_parity_errors.md:
    # lam 1.2e-11 rho on flying_obs c2, grad 2.4e-13 rho on quad_T5_f64_lastqp); round 1 allowed 1e-10 / 1e-11.
    return dict(xu=1e-6, lam=max(1e-6, 5e-11 * rho_used), grad=max(1e-6, 2e-12 * rho_used))


def log_parity(**rec):
    """Measured errors go to $ALMPC_PARITY_LOG (json lines) so that the widening of a tolerance is evidence, not
    allowance: tools/parity_table.py turns the log into profiles/r02_parity_errors.md."""
    path = os.environ.get("ALMPC_PARITY_LOG")
    if path:
        import json

        with open(path, "a") as f:
            f.write(json.dumps(rec) + "\n")


# --------------------------------------------------------------------------------------------------
# dynamics functors vs the oracle models
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["pendulum", "cartpole1l", "cartpole2l", "rex_quadrotor", "flying_cartpole"])
def test_dynamics_matches_oracle(model):
    from oracle.problem_gen import make_oracle_model

    om = make_oracle_model(model)
    dyn = make_dyn(model, dt=om.dt)
    g = torch.Generator().manual_seed(5)
    R = 257
    x = (torch.rand(R, om.nx, generator=g, dtype=torch.float64) * 2 - 1) * 0.8
    lo, hi = torch.tensor(om.u_lower), torch.tensor(om.u_upper)
    u = lo + (hi - lo) * torch.rand(R, om.nu, generator=g, dtype=torch.float64)
    xn_ref, (jx_ref, ju_ref) = om.step_jac(x, u)
    xn, (jx, ju) = dyn.jacobian(x.to(DEV), u.to(DEV))
    xn2 = dyn(x.to(DEV), u.to(DEV))
    assert torch.equal(xn, xn2)
    assert rel_err(xn.cpu(), xn_ref) < 1e-13
    assert rel_err(jx.cpu(), jx_ref) < 1e-12
    assert rel_err(ju.cpu(), ju_ref) < 1e-12
    # fp32 functor
    xn32 = dyn(x.float().to(DEV), u.float().to(DEV))
    assert rel_err(xn32.cpu(), xn_ref) < 5e-6


def test_cartpole2l_known_answer_traj():
    """my_envs/traj.npz (the only on-disk fixture of the reference): X[k+1] = f(X[k], U[k]) at dt=0.03
    to 1e-6.  A copy of the arrays lives in tests/golden/cartpole2l_traj.npz."""
    import numpy as np
    import os

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cartpole2l_traj.npz"))
    X, U = torch.from_numpy(z["X_np"]), torch.from_numpy(z["U_np"])
    dyn = make_dyn("cartpole2l", dt=0.03)
    xn = dyn(X[:-1].contiguous().to(DEV), U.contiguous().to(DEV)).cpu()
    assert (xn - X[1:]).abs().max().item() < 2e-6


# --------------------------------------------------------------------------------------------------
# full solves vs golden outputs of the unmodified reference
# --------------------------------------------------------------------------------------------------
def make_ctrl(gold):
    spec = MODEL_SPECS[gold.model]
    dyn = make_dyn(gold.model, dt=spec["dt"])
    if "obstacle_positions" in gold.z.files:  # Obstacle_MPC scenario (qpth/AL_mpc_custom.py)
        from deq_mpc_corl_b200.al_mpc import Obstacle_MPC

        env = types.SimpleNamespace(obstacle_positions=gold.t("obstacle_positions"),
                                    obstacle_radius=float(gold.z["obstacle_radius"]))
        ctrl = Obstacle_MPC(spec["nx"], spec["nu"], gold.T, u_lower=gold.t("u_lower"), u_upper=gold.t("u_upper"),
                            n_batch=gold.B, dtype=gold.dtype, env=env)
    else:
        ctrl = MPC(spec["nx"], spec["nu"], gold.T, u_lower=gold.t("u_lower"), u_upper=gold.t("u_upper"),
                   n_batch=gold.B, dtype=gold.dtype, state_estimator=gold.state_estimator)
    return ctrl, dyn, spec


@pytest.mark.parametrize("name", list_golden() if cuda else [])
def test_solve_matches_reference_golden(name):
    gold = Golden(name)
    ctrl, dyn, spec = make_ctrl(gold)
    nx = spec["nx"]
    x0 = gold.t("x0").to(DEV)
    Qd1 = gold.t("Qd").to(gold.dtype)
    Qfull0 = torch.diag(Qd1).repeat(gold.B, gold.T, 1, 1).to(DEV)
    for ci in range(gold.n_calls):
        c = gold.call(ci)
        if ci == 0:
            ctrl.reinitialize(c["x_ref"].to(DEV))
        if gold.stream_after == ci:
            # Tracking_MPC.warm_start_initialize (policies.py:1305-1310): last stage of the warm start <- reference
            xi, ui = ctrl.x_init, ctrl.u_init
            ui[:, -1:] = c["u_ref"][:, -1:].to(DEV)
            xi[:, -1:] = c["x_ref"][:, -1:].to(DEV)
            ctrl.warm_start_initialize(xi, ui, types.SimpleNamespace(rho_init_max=gold.rho_init_max))
        if "obs_pos" in c:  # the nearest-obstacle selection must match the reference's
            assert torch.equal(ctrl.obstacles[0].cpu(), c["obs_pos"]), f"{name} call {ci}: obstacle selection"
        # call-level parity: start from the reference's own state
        assert bool(c["warm"]) == ctrl.warm_starting
        ctrl.x_init = c["xu_init"][..., :nx].to(DEV)
        ctrl.u_init = c["xu_init"][..., nx:].to(DEV)
        ctrl.lamda_prev = c["lam_in"].to(DEV)
        ctrl.rho_prev = c["rho_in"].to(DEV)
        ctrl.al_iter = int(c["al_iter"])
        Qleaf = Qfull0.clone().requires_grad_(True)
        q = c["q"].to(DEV).requires_grad_(True)
        x, u, status = ctrl(x0, QuadCost(Qleaf, q, None), dyn, dyn)
        Qg, qg = torch.autograd.grad((x, u), (Qleaf, q), (c["gx"].to(DEV), c["gu"].to(DEV)))
        Qg = Qg.diagonal(dim1=-2, dim2=-1)
        tol = tolerances(float(c["rho_out"].double().max()), gold.dtype, c)
        assert ctrl.newton_steps() == c["nsteps"].tolist(), f"{name} call {ci}: Newton-step counts"
        assert bool(status) == bool(c["status"])
        assert (ctrl.last_status.cpu() == 3).all()
        assert torch.equal(ctrl.rho_prev.cpu().double(), c["rho_out"].double())
        assert ctrl.rho_prev.dtype == c["rho_out"].dtype
        e_xu = rel_err(ctrl.last_xu.cpu(), c["xu64"])
        e_lam = rel_err(ctrl.lamda_prev.cpu(), c["lam_out"])
        e_q = rel_err(qg.cpu(), c["q_grad"])
        e_Q = rel_err(Qg.cpu(), c["Q_grad"])
        print(f"{name} c{ci}: xu {e_xu:.1e} lam {e_lam:.1e} q_grad {e_q:.1e} Q_grad {e_Q:.1e} (tol {tol})")
        log_parity(test="solve_vs_reference_golden", golden=name, call=ci, rho_used=float(c["rho_out"].double().max()) / 10,
                   xu=e_xu, lam=e_lam, q_grad=e_q, Q_grad=e_Q, tol=tol, newton_steps=ctrl.newton_steps())
        assert e_xu < tol["xu"], f"{name} call {ci}: xu {e_xu}"
        assert e_lam < tol["lam"], f"{name} call {ci}: lam {e_lam}"
        assert e_q < tol["grad"] and e_Q < tol["grad"], f"{name} call {ci}: grads {e_q} {e_Q}"
        assert rel_err(x.detach().cpu(), c["x"]) < max(tol.get("x", tol["xu"]), 1e-6)
        assert rel_err(u.detach().cpu(), c["u"]) < max(tol["xu"] * 10, 1e-6)
        assert x.dtype == torch.float32 and u.dtype == torch.float32
        assert rel_err(ctrl.dyn_res_prev.cpu(), c["dyn_res"]) < max(tol["lam"] * 100, 1e-5)


@pytest.mark.parametrize("name", ["quad_T5_f64_cold", "pendulum_T5_f64_chain", "flying_T10_f64_cold", "cartpole1l_T5_f64_stream",
                                  "quad_se_T5_f64_chain", "cartpole1l_se_T9_f64", "quad_T5_f32_cold"] if cuda else [])
def test_first_newton_step_trace(name):
    """Step-level parity: gradient, Newton direction, the 20 merits, argmin and accept of the first
    Newton step against the reference's own intermediate values.  fp64: 1e-12 / 1e-9; fp32 (rho = 1, the regime of
    SURVEY A.5): the north_star bar of 1e-4, argmin and accept exact."""
    from deq_mpc_corl_b200.solver import SolveConfig, raw_forward

    gold = Golden(name)
    c = gold.call(0)
    spec = MODEL_SPECS[gold.model]
    dyn = make_dyn(gold.model, dt=spec["dt"])
    cfg = SolveConfig(dyn=dyn, T=gold.T, u_lower=gold.t("u_lower").tolist(), u_upper=gold.t("u_upper").tolist(),
                      al_iter=1, max_newton=1, debug=True, skip_final_dual=True, state_estimator=gold.state_estimator)
    dt = gold.dtype
    B, T, n = gold.B, gold.T, spec["nx"] + spec["nu"]
    xu = c["xu_init"].to(dt).to(DEV).contiguous()
    Qd = gold.t("Qd").to(dt).expand(B, T, n).contiguous().to(DEV)
    out = raw_forward(cfg, gold.t("x0").to(dt).to(DEV), Qd, c["q"].to(dt).to(DEV).contiguous(), xu,
                      c["lam_in"].to(dt).to(DEV).contiguous(), c["rho_in"].to(dt).reshape(-1).to(DEV).contiguous())
    dbg = out["debug"].cpu()
    N = T * n
    f64 = dt == torch.float64
    e = dict(grad=rel_err(dbg[:, :N], c["step_grad"][0]),
             dir=max_rel_err_rows(dbg[:, N:2 * N], c["step_dir"][0].reshape(B, -1)),
             merits=rel_err(dbg[:, 2 * N:2 * N + 20].T, c["step_merits"][0]),
             merit_in=rel_err(dbg[:, 2 * N + 22], c["step_merit_in"][0]))
    print(f"{name} first Newton step: " + " ".join(f"{k} {v:.1e}" for k, v in e.items()))
    log_parity(test="first_newton_step", golden=name, **e)
    # fp32: the north_star bar is 1e-4; measured 5e-8 / 6e-7 / 9e-8 (profiles/r02_parity_errors.md), asserted at 1e-5
    assert e["grad"] < (1e-12 if f64 else 1e-5)
    assert e["dir"] < (1e-9 if f64 else 1e-5)
    assert e["merits"] < (1e-12 if f64 else 1e-5)
    idx, ref_idx = dbg[:, 2 * N + 20].long(), c["step_idx"][0]
    if f64:
        assert torch.equal(idx, ref_idx)
    else:  # fp32: identical except at ties of the reference's own merits (two fp32 runs of the reference differ there too)
        gm = c["step_merits"][0].double()
        b = torch.arange(B)
        gap = (gm[idx, b] - gm[ref_idx, b]).abs() / gm[ref_idx, b].abs().clamp_min(1e-300)
        assert (gap[idx != ref_idx] < 1e-6).all()
    assert torch.equal(dbg[:, 2 * N + 21].bool(), c["step_accept"][0])
    assert e["merit_in"] < (1e-12 if f64 else 1e-5)


# --------------------------------------------------------------------------------------------------
# CUDA vs the CPU oracle on seeded synthetic batches
# --------------------------------------------------------------------------------------------------
def solve_both(model, T, B, seed, dtype, al_iter=2, state_estimator=False):
    from oracle import almpc_oracle as orc
    from oracle.problem_gen import make_oracle_model

    inp = make_inputs(model, T, B, seed)
    spec = MODEL_SPECS[model]
    if state_estimator:  # the controls are data: drawn inside the bounds (policies.py:378, u_gt_est)
        g1 = torch.Generator().manual_seed(seed + 13)
        lo, hi = torch.as_tensor(inp["u_lower"]), torch.as_tensor(inp["u_upper"])
        inp["u_ref"] = (lo + (hi - lo) * torch.rand(B, T, spec["nu"], generator=g1)).to(inp["u_ref"].dtype)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], dtype)
    # oracle
    om = make_oracle_model(model)
    ompc = orc.OracleMPC(om, T, inp["u_lower"], inp["u_upper"], B, dtype=dtype, al_iter=al_iter,
                         state_estimator=state_estimator)
    ompc.reinitialize(inp["x_ref"])
    ompc.x_init, ompc.u_init = inp["x_ref"].clone(), inp["u_ref"].clone()
    tr = orc.SolveTrace()
    xo, uo, _, xuo = ompc.solve(inp["x0"], Qd, q, trace=tr)
    g = torch.Generator().manual_seed(seed + 1)
    gx, gu = torch.randn(xo.shape, generator=g), torch.randn(uo.shape, generator=g)
    Qg_o, qg_o = ompc.backward(gx, gu)
    # CUDA
    dyn = make_dyn(model, dt=spec["dt"])
    ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, dtype=dtype,
               state_estimator=state_estimator)
    ctrl.reinitialize(inp["x_ref"].to(DEV))
    ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
    ctrl.al_iter = al_iter
    Qdl = Qd.to(DEV).requires_grad_(True)
    ql = q.to(DEV).requires_grad_(True)
    x, u, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qdl, ql, None), dyn, dyn)
    Qg, qg = torch.autograd.grad((x, u), (Qdl, ql), (gx.to(DEV), gu.to(DEV)))
    return dict(o=dict(xu=xuo, lam=ompc.lamda_prev, Qg=Qg_o, qg=qg_o, steps=tr.newton_steps),
                c=dict(xu=ctrl.last_xu.cpu(), lam=ctrl.lamda_prev.cpu(), Qg=Qg.cpu(), qg=qg.cpu(), steps=ctrl.newton_steps()))


@pytest.mark.parametrize("model,T,B", [("rex_quadrotor", 5, 192), ("pendulum", 5, 200), ("cartpole1l", 5, 256),
                                       ("cartpole1l", 10, 64), ("pendulum", 20, 33), ("flying_cartpole", 10, 48),
                                       ("rex_quadrotor", 10, 32), ("cartpole2l", 5, 40),
                                       # every warp-synchronous line-search instantiation ((T-1) | 32) and ragged batches
                                       ("pendulum", 2, 37), ("cartpole1l", 3, 45), ("cartpole2l", 9, 21),
                                       ("pendulum", 17, 19), ("cartpole1l", 33, 5), ("rex_quadrotor", 9, 13),
                                       ("flying_cartpole", 5, 31), ("rex_quadrotor", 3, 7), ("pendulum", 5, 1)])
def test_cuda_matches_oracle_f64(model, T, B):
    r = solve_both(model, T, B, 4242 + T, torch.float64)
    assert r["c"]["steps"] == r["o"]["steps"]
    assert rel_err(r["c"]["xu"], r["o"]["xu"]) < 1e-6
    assert max_rel_err_rows(r["c"]["xu"], r["o"]["xu"]) < 1e-5
    assert rel_err(r["c"]["lam"], r["o"]["lam"]) < 1e-6
    assert rel_err(r["c"]["qg"], r["o"]["qg"]) < 1e-6
    assert rel_err(r["c"]["Qg"], r["o"]["Qg"]) < 1e-6


@pytest.mark.parametrize("model,T,B", [("rex_quadrotor", 5, 96), ("pendulum", 5, 70), ("cartpole1l", 9, 33),
                                       ("cartpole2l", 3, 17), ("flying_cartpole", 10, 24), ("rex_quadrotor", 7, 9)])
def test_state_estimator_matches_oracle_f64(model, T, B):
    """MPC(state_estimator=True) (qpth/al_utils_se.py): dynamics rows only, controls are data."""
    r = solve_both(model, T, B, 515 + T, torch.float64, state_estimator=True)
    assert r["c"]["steps"] == r["o"]["steps"]
    assert rel_err(r["c"]["xu"], r["o"]["xu"]) < 1e-6
    assert max_rel_err_rows(r["c"]["xu"], r["o"]["xu"]) < 1e-5
    assert rel_err(r["c"]["lam"], r["o"]["lam"]) < 1e-6
    assert rel_err(r["c"]["qg"], r["o"]["qg"]) < 1e-6
    assert rel_err(r["c"]["Qg"], r["o"]["Qg"]) < 1e-6
    nx = MODEL_SPECS[model]["nx"]
    # the controls never move
    assert torch.equal(r["c"]["xu"][..., nx:], r["o"]["xu"][..., nx:])


# --------------------------------------------------------------------------------------------------
# generic-callable dynamics mode (arbitrary Python dx / dx_jac)
# --------------------------------------------------------------------------------------------------
def run_mpc(dx, dx_jac, nx, nu, T, inp, Qd, q, gx, gu, dtype, al_iter=2, state_estimator=False):
    B = inp["x0"].shape[0]
    ctrl = MPC(nx, nu, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, dtype=dtype,
               state_estimator=state_estimator)
    ctrl.reinitialize(inp["x_ref"].to(DEV))
    ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
    ctrl.al_iter = al_iter
    Qdl = Qd.to(DEV).requires_grad_(True)
    ql = q.to(DEV).requires_grad_(True)
    x, u, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qdl, ql, None), dx, dx_jac)
    Qg, qg = torch.autograd.grad((x, u), (Qdl, ql), (gx.to(DEV), gu.to(DEV)))
    assert x.dtype == torch.float32 and u.dtype == torch.float32
    return dict(xu=ctrl.last_xu.cpu(), lam=ctrl.lamda_prev.cpu(), rho=ctrl.rho_prev.cpu(), Qg=Qg.cpu(), qg=qg.cpu(),
                steps=ctrl.newton_steps(), status=ctrl.last_status.cpu(), dyn_res=ctrl.dyn_res_prev.cpu())


@pytest.mark.parametrize("model,T,B,se", [("rex_quadrotor", 5, 64, False), ("cartpole1l", 7, 33, False),
                                          ("pendulum", 5, 50, False), ("flying_cartpole", 5, 17, False),
                                          ("cartpole2l", 4, 9, False), ("cartpole1l", 5, 21, True)])
def test_generic_callable_mode_matches_fused_path(model, T, B, se):
    """The same dynamics passed (a) as a native functor -> fused kernels and (b) as opaque Python callables ->
    generic mode (torch host loop + almpc_blocktri) must agree: identical Newton-step counts, iterates to 1e-7."""
    spec = MODEL_SPECS[model]
    inp = make_inputs(model, T, B, 9090 + T)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], torch.float64)
    g = torch.Generator().manual_seed(3)
    gx, gu = torch.randn(B, T, spec["nx"], generator=g), torch.randn(B, T, spec["nu"], generator=g)
    native = make_dyn(model, dt=spec["dt"])
    calls = {"dx": 0, "jac": 0}

    def dx(x, u):
        calls["dx"] += 1
        return native(x, u)

    def dx_jac(x, u):
        calls["jac"] += 1
        return native.jacobian(x, u)

    fused = run_mpc(native, native, spec["nx"], spec["nu"], T, inp, Qd, q, gx, gu, torch.float64, state_estimator=se)
    gen = run_mpc(dx, dx_jac, spec["nx"], spec["nu"], T, inp, Qd, q, gx, gu, torch.float64, state_estimator=se)
    assert calls["jac"] == sum(gen["steps"]) and calls["dx"] > 0
    assert gen["steps"] == fused["steps"]
    assert (gen["status"] == 3).all()
    assert torch.equal(gen["rho"], fused["rho"])
    # two different summation orders of the same algorithm: agreement is limited by cond(H) ~ rho / R
    # (cartpole: R = 1e-10), not by either implementation; measured 1e-13 (quadrotor) ... 2e-9 (cartpole T=7)
    assert rel_err(gen["xu"], fused["xu"]) < 1e-7
    assert rel_err(gen["lam"], fused["lam"]) < 1e-6
    assert rel_err(gen["qg"], fused["qg"]) < 1e-6 and rel_err(gen["Qg"], fused["Qg"]) < 1e-6
    assert rel_err(gen["dyn_res"], fused["dyn_res"]) < 1e-6


class _Unicycle:
    """A model the library has no functor for (nx=3, nu=2): Euler-discretised unicycle with drag."""
    nx, nu, dt = 3, 2, 0.1

    def step(self, x, u):
        px, py, th = x[:, 0], x[:, 1], x[:, 2]
        v, om = u[:, 0], u[:, 1]
        return torch.stack([px + self.dt * v * torch.cos(th), py + self.dt * v * torch.sin(th),
                            th + self.dt * (om - 0.1 * torch.sin(th) * v * v)], dim=1)

    def step_jac(self, x, u):
        from torch.func import jacrev, vmap

        f = lambda xi, ui: self.step(xi[None], ui[None])[0]
        Jx = vmap(jacrev(f, 0))(x, u)
        Ju = vmap(jacrev(f, 1))(x, u)
        return self.step(x, u), (Jx, Ju)


@pytest.mark.parametrize("T,B,dtype", [(6, 40, torch.float64), (9, 25, torch.float64), (6, 40, torch.float32)])
def test_generic_callable_mode_user_model_matches_oracle(T, B, dtype):
    """A pure-torch user model with no compiled functor: generic mode vs the CPU oracle on the same callables."""
    from oracle import almpc_oracle as orc

    m = _Unicycle()
    g = torch.Generator().manual_seed(77 + T)
    x0 = torch.randn(B, 3, generator=g, dtype=torch.float64)
    x_ref = x0[:, None, :] * torch.linspace(1.0, 0.5, T, dtype=torch.float64)[None, :, None] + 0.05 * torch.randn(B, T, 3, generator=g, dtype=torch.float64)
    x_ref[:, 0] = x0
    u_ref = torch.zeros(B, T, 2, dtype=torch.float64)
    x0, x_ref, u_ref = x0.to(dtype), x_ref.to(dtype), u_ref.to(dtype)
    inp = dict(x0=x0, x_ref=x_ref, u_ref=u_ref, u_lower=torch.tensor([-1.0, -2.0]), u_upper=torch.tensor([1.0, 2.0]))
    Qd, q, _ = tracking_cost(torch.tensor([10.0, 10.0, 1.0, 1e-2, 1e-2]), x_ref, u_ref, dtype)
    q = q.to(dtype)
    gx, gu = torch.randn(B, T, 3, generator=g), torch.randn(B, T, 2, generator=g)
    ompc = orc.OracleMPC(m, T, inp["u_lower"], inp["u_upper"], B, dtype=dtype, al_iter=2)
    ompc.reinitialize(x_ref)
    ompc.x_init, ompc.u_init = x_ref.clone(), u_ref.clone()
    tr = orc.SolveTrace()
    _, _, _, xuo = ompc.solve(x0, Qd, q, trace=tr)
    Qg_o, qg_o = ompc.backward(gx, gu)
    gen = run_mpc(m.step, m.step_jac, 3, 2, T, inp, Qd, q, gx, gu, dtype)
    if dtype == torch.float64:
        assert gen["steps"] == tr.newton_steps
        assert rel_err(gen["xu"], xuo) < 1e-6 and max_rel_err_rows(gen["xu"], xuo) < 1e-5
        assert rel_err(gen["lam"], ompc.lamda_prev) < 1e-6
        assert rel_err(gen["qg"], qg_o) < 1e-6 and rel_err(gen["Qg"], Qg_o) < 1e-6
    else:
        assert abs(sum(gen["steps"]) - sum(tr.newton_steps)) <= 1
        assert rel_err(gen["xu"], xuo) < 3e-2


@pytest.mark.parametrize("model,T,B", [("rex_quadrotor", 5, 128), ("cartpole1l", 5, 128)])
def test_cuda_matches_oracle_f32(model, T, B):
    """fp32 solver vs the fp32 oracle on 128 seeded problems.  Both sit on the fp32 noise floor of the algorithm:
    the batch-global stop statistic can land on either side of its 1e-3 threshold (one Newton step more or less),
    and a problem whose 20 merits tie in fp32 takes a different step.  So the step counts may differ by one and the
    bulk of the problems must agree to the fp32 bar, with a small budget of tie-broken outliers."""
    r = solve_both(model, T, B, 777, torch.float32)
    assert abs(sum(r["c"]["steps"]) - sum(r["o"]["steps"])) <= 1
    a, b = r["c"]["xu"].double().reshape(B, -1), r["o"]["xu"].double().reshape(B, -1)
    rows = ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-300))
    med, frac_ok, tot = rows.median().item(), (rows < 1e-3).float().mean().item(), rel_err(r["c"]["xu"], r["o"]["xu"])
    print(f"fp32 {model}: steps {r['c']['steps']} vs {r['o']['steps']}, xu median row err {med:.1e}, rows < 1e-3: {frac_ok:.2f}, total {tot:.1e}")
    log_parity(test="cuda_vs_oracle_f32", model=model, median_row=med, frac_rows_below_1e_3=frac_ok, total=tot,
               steps=r["c"]["steps"], oracle_steps=r["o"]["steps"])
    assert tot < 3e-2
    if r["c"]["steps"] == r["o"]["steps"]:  # measured: quadrotor 7e-6 / 1.00, cartpole (R = 1e-10) 6e-5 / 0.88
        assert med < 1e-4 and frac_ok > 0.8


# --------------------------------------------------------------------------------------------------
# edge cases + full-size properties
# --------------------------------------------------------------------------------------------------
def test_empty_and_ragged_batches():
    spec = MODEL_SPECS["rex_quadrotor"]
    dyn = make_dyn("rex_quadrotor")
    for B in (0, 1, 3, 5):  # not a multiple of the 4 problems per CTA
        inp = make_inputs("rex_quadrotor", 5, max(B, 1), 11)
        inp = {k: (v[:B] if v.dim() > 1 else v) for k, v in inp.items()}
        Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
        ctrl = MPC(12, 4, 5, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
        ctrl.reinitialize(inp["x_ref"].to(DEV))
        ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
        x, u, st = ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), q.to(DEV), None), dyn, dyn)
        assert x.shape == (B, 5, 12) and u.shape == (B, 5, 4) and st is False
        assert torch.isfinite(x).all() and torch.isfinite(u).all()


def test_horizon_limits():
    """Longest supported horizon (T = 64, check_desc) against the oracle, and the loud failure one past it."""
    r = solve_both("pendulum", 64, 3, 321, torch.float64)
    assert r["c"]["steps"] == r["o"]["steps"]
    assert rel_err(r["c"]["xu"], r["o"]["xu"]) < 1e-6 and rel_err(r["c"]["qg"], r["o"]["qg"]) < 1e-6
    r = solve_both("cartpole1l", 2, 9, 322, torch.float64)  # shortest horizon: one dynamics block
    assert r["c"]["steps"] == r["o"]["steps"] and rel_err(r["c"]["xu"], r["o"]["xu"]) < 1e-6
    inp = make_inputs("pendulum", 65, 2, 5)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
    ctrl = MPC(2, 1, 65, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=2)
    ctrl.reinitialize(inp["x_ref"].to(DEV))
    ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
    with pytest.raises(_lib.AlmpcError, match="horizon"):
        ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), q.to(DEV), None), make_dyn("pendulum"), None)


def test_unregistered_dynamics_and_cpu_tensors_fail_loudly():
    inp = make_inputs("pendulum", 5, 4, 3)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
    ctrl = MPC(2, 1, 5, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=4)
    ctrl.reinitialize(inp["x_ref"])
    ctrl.x_init, ctrl.u_init = inp["x_ref"], inp["u_ref"]
    with pytest.raises(_lib.AlmpcError):
        ctrl(inp["x0"], QuadCost(Qd, q, None), make_dyn("pendulum"), None)  # CPU tensors
    with pytest.raises(_lib.AlmpcError):
        ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), q.to(DEV), None), lambda x, u: x, None)


def test_full_size_properties_quadrotor_64k():
    """BASELINE.json headline size (quadrotor T=5, B=65536, fp64): properties that do not need the
    oracle — determinism, replicated problems give bit-identical rows, KKT residual of the adjoint
    solve, linearity of the backward — plus an oracle check on a sample of rows with the Newton-step
    counts forced to the ones the GPU batch realised (the stop rule is batch-global)."""
    model, T, B = "rex_quadrotor", 5, 65536
    spec = MODEL_SPECS[model]
    base = make_inputs(model, T, 4096, 99)
    rep = B // 4096
    inp = {k: (v.repeat(rep, *([1] * (v.dim() - 1))) if v.dim() > 1 else v) for k, v in base.items()}
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
    dyn = make_dyn(model)

    def run():
        ctrl = MPC(12, 4, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
        ctrl.reinitialize(inp["x_ref"].to(DEV))
        ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
        ql = q.to(DEV).requires_grad_(True)
        x, u, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), ql, None), dyn, dyn)
        return ctrl, x, u, ql

    ctrl, x, u, ql = run()
    ctrl2, x2, u2, _ = run()
    assert torch.equal(x, x2) and torch.equal(u, u2) and torch.equal(ctrl.lamda_prev, ctrl2.lamda_prev)  # deterministic
    xu = ctrl.last_xu.view(rep, 4096, T, 16)
    assert (xu == xu[0:1]).all()  # replicas bit-identical
    assert (ctrl.last_status == 3).all()
    steps = ctrl.newton_steps()
    # backward linearity
    g = torch.Generator().manual_seed(1)
    v1 = torch.randn(B, T, 12, generator=g).to(DEV), torch.randn(B, T, 4, generator=g).to(DEV)
    v2 = torch.randn(B, T, 12, generator=g).to(DEV), torch.randn(B, T, 4, generator=g).to(DEV)
    g1, = torch.autograd.grad((x, u), ql, v1, retain_graph=True)
    g2, = torch.autograd.grad((x, u), ql, v2, retain_graph=True)
    g12, = torch.autograd.grad((x, u), ql, (2 * v1[0] - 3 * v2[0], 2 * v1[1] - 3 * v2[1]), retain_graph=True)
    assert rel_err(g12, 2 * g1 - 3 * g2) < 1e-6  # cond(H) ~ rho/R = 1e9 at the end of a cold 2-iteration solve
    # oracle on a sample of rows, forcing the realised step counts
    from oracle import almpc_oracle as orc
    from oracle.problem_gen import make_oracle_model

    rows = torch.arange(0, 4096, 37)[:96]
    om = make_oracle_model(model)
    ompc = orc.OracleMPC(om, T, inp["u_lower"], inp["u_upper"], len(rows), al_iter=2)
    ompc.reinitialize(inp["x_ref"][rows])
    ompc.x_init, ompc.u_init = inp["x_ref"][rows].clone(), inp["u_ref"][rows].clone()
    ompc.force_steps = list(steps)
    _, _, _, xuo = ompc.solve(inp["x0"][rows], Qd[rows], q[rows])
    assert rel_err(ctrl.last_xu[rows.to(DEV)].cpu(), xuo) < 1e-6
    Qg_o, qg_o = ompc.backward(v1[0][rows.to(DEV)].cpu(), v1[1][rows.to(DEV)].cpu())
    assert rel_err(g1[rows.to(DEV)].cpu(), qg_o) < 1e-6


def test_newton_al_function_signature():
    """al_utils.NewtonAL.apply with the reference's 15-argument order (al_utils.py:451-468)."""
    from deq_mpc_corl_b200.al_utils import NewtonAL, NewtonContext
    from oracle import almpc_oracle as orc
    from oracle.problem_gen import make_oracle_model

    model, T, B = "cartpole1l", 5, 64
    inp = make_inputs(model, T, B, 21)
    Qd, q, f = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
    dyn = make_dyn(model)
    xu0 = torch.cat([inp["x_ref"], inp["u_ref"]], -1).double()
    lam = torch.zeros(B, T * 4 + 2 * T, dtype=torch.float64)
    rho = torch.ones(B, 1, dtype=torch.float64)

    def meritfn(*a, **k):
        raise AssertionError("callables are not evaluated by the fused path")

    meritfn.almpc_context = NewtonContext(dyn, inp["u_lower"], inp["u_upper"])
    Ql, ql = Qd.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)
    x_est, status = NewtonAL.apply(meritfn, None, None, None, xu0.to(DEV), inp["x0"].double().to(DEV), lam.to(DEV),
                                   rho.to(DEV), Ql, ql, f.to(DEV), 1e-3, 1e-6, True, 0)
    gQ, gq = torch.autograd.grad(x_est, (Ql, ql), torch.ones_like(x_est))
    pb = orc.Problem(make_oracle_model(model), T, inp["u_lower"], inp["u_upper"])
    xo, st, saved = orc.newton_al(pb, xu0.clone(), inp["x0"].double(), lam, rho, Qd, q)
    Qg_o, qg_o = orc.newton_backward(saved, xo, torch.ones_like(xo))
    assert status.all() and rel_err(x_est.detach().cpu(), xo) < 1e-6
    assert rel_err(gq.cpu(), qg_o) < 1e-6 and rel_err(gQ.cpu(), Qg_o) < 1e-6


def test_tracking_mpc_fused_cost_and_chain_rule():
    """Tracking_MPC (policies.py:1117-1315) with the fused cost build: same solve and same gradients w.r.t.
    x_ref, u_ref and q_scaling as the reference's torch-op route (diag-embedded Q, p = -(Q xu_ref))."""
    from deq_mpc_corl_b200.tracking import Tracking_MPC

    model, T, B = "rex_quadrotor", 5, 96
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    inp = make_inputs(model, T, B, 77)
    dyn = make_dyn(model)
    args = types.SimpleNamespace(T=T, bsz=B, Q=torch.tensor(spec["Q"]), R=torch.tensor(spec["R"]), dtype="double",
                                 device=DEV, solver_type="al", env="rexquadrotor", rho_init_max=1e4)
    env = types.SimpleNamespace(nx=nx, nu=nu, nq=7, dt=spec["dt"], dynamics=dyn, dynamics_derivatives=dyn,
                                action_space=types.SimpleNamespace(low=inp["u_lower"].numpy(), high=inp["u_upper"].numpy()))
    g = torch.Generator().manual_seed(5)
    gx, gu = torch.randn(B, T, nx, generator=g).to(DEV), torch.randn(B, T, nu, generator=g).to(DEV)
    s0 = (0.3 * torch.rand(B, T, generator=g)).to(DEV)
    x0 = inp["x0"].to(DEV)

    def leaves():
        return (inp["x_ref"].to(DEV).requires_grad_(True), inp["u_ref"].to(DEV).requires_grad_(True),
                s0.clone().requires_grad_(True))

    # fused route
    tm = Tracking_MPC(args, env)
    xr, ur, sc = leaves()
    tm.reinitialize(xr, None)
    x, u, st = tm(x0, None, xr, ur, q_scaling=sc, al_iters=2)
    gr = torch.autograd.grad((x, u), (xr, ur, sc), (gx, gu))
    # reference-shaped route through the same solver: diag-embedded Q and torch ops (policies.py:1245-1297)
    ctrl = MPC(nx, nu, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
    xr2, ur2, sc2 = leaves()
    ctrl.reinitialize(xr2, None)
    ctrl.x_init, ctrl.u_init = xr2.detach().clone(), ur2.detach().clone()
    Qfull = torch.diag(torch.tensor(spec["Q"] + spec["R"], dtype=torch.float64)).repeat(B, T, 1, 1).to(DEV)
    qs = sc2 + torch.ones_like(sc2)
    Q = Qfull * qs[:, :, None, None]
    xu_ref = torch.cat([xr2, ur2], -1)
    p = -(Q * xu_ref.unsqueeze(-2)).sum(-1)
    x2, u2, _ = ctrl(x0, QuadCost(Q, p, None), dyn, dyn)
    gr2 = torch.autograd.grad((x2, u2), (xr2, ur2, sc2), (gx, gu))
    assert st is False and torch.equal(x, x2) and torch.equal(u, u2)
    for a, b, nm in zip(gr, gr2, ("x_ref", "u_ref", "q_scaling")):
        assert a.dtype == b.dtype and rel_err(a, b) < 1e-5, (nm, rel_err(a, b))
    # streaming glue keeps working on top (warm start: lam <- 0, rho <- min(rho, rho_init_max))
    tm.warm_start_initialize(xr.detach(), ur.detach())
    x3, u3, st3 = tm(x0, None, xr.detach(), ur.detach(), al_iters=2)
    assert tm.ctrl.warm_starting and isinstance(st3, bool) and torch.isfinite(x3).all()


@pytest.mark.parametrize("name", list_tracking_golden() if cuda else [])
def test_tracking_mpc_matches_reference_golden(name):
    """`tracking.Tracking_MPC` against the reference's OWN `Tracking_MPC.forward` (deqmpc/policies.py:1236-1283,
    golden written by oracle/gen_golden.py::run_tracking_scenario from the unmodified reference): a chain of cold
    calls, then `warm_start_initialize` and streaming calls up to the rho_max exit; outputs, carried duals /
    penalties and the gradients w.r.t. the fp32 network outputs x_ref, u_ref, q_scaling.  The solver state (lam, rho,
    warm start) is carried by OUR module across the whole chain, never re-synchronised with the golden."""
    from deq_mpc_corl_b200.tracking import Tracking_MPC

    gold = Golden(name)
    spec = MODEL_SPECS[gold.model]
    nx, nu, T, B = spec["nx"], spec["nu"], gold.T, gold.B
    dyn = make_dyn(gold.model, dt=spec["dt"])
    args = types.SimpleNamespace(T=T, bsz=B, Q=gold.t("Q"), R=gold.t("R"), dtype="double" if gold.dtype == torch.float64 else "float",
                                 device=DEV, solver_type="al", env=gold.model, rho_init_max=gold.rho_init_max)
    env = types.SimpleNamespace(nx=nx, nu=nu, nq=None, dt=spec["dt"], dynamics=dyn, dynamics_derivatives=dyn,
                                action_space=types.SimpleNamespace(low=gold.t("u_lower").numpy(), high=gold.t("u_upper").numpy()))
    tm = Tracking_MPC(args, env)
    x0 = gold.t("x0").to(DEV)
    tm.reinitialize(gold.call(0)["x_ref"].to(DEV), None)
    for ci in range(gold.n_calls):
        c = gold.call(ci)
        if gold.stream_after == ci:
            tm.warm_start_initialize(c["x_ref"].to(DEV), c["u_ref"].to(DEV))
        xr = c["x_ref"].to(DEV).requires_grad_(True)
        ur = c["u_ref"].to(DEV).requires_grad_(True)
        qs = c["q_scaling"].to(DEV).requires_grad_(True) if "q_scaling" in c else None
        x, u, status = tm(x0, None, xr, ur, q_scaling=qs, al_iters=int(c["al_iters"]))
        ins = (xr, ur) + ((qs,) if qs is not None else ())
        grads = torch.autograd.grad((x, u), ins, (c["gx"].to(DEV), c["gu"].to(DEV)))
        tol = tolerances(float(c["rho_out"].double().max()), gold.dtype)
        assert bool(status) == bool(c["status"]), f"{name} call {ci}: status"
        assert torch.equal(tm.ctrl.rho_prev.cpu().double(), c["rho_out"].double())
        e = dict(x=rel_err(x.detach().cpu(), c["x"]), u=rel_err(u.detach().cpu(), c["u"]),
                 lam=rel_err(tm.ctrl.lamda_prev.cpu(), c["lam_out"]),
                 g_x_ref=rel_err(grads[0].cpu(), c["g_x_ref"]), g_u_ref=rel_err(grads[1].cpu(), c["g_u_ref"]))
        if qs is not None:
            e["g_q_scaling"] = rel_err(grads[2].cpu(), c["g_q_scaling"])
        print(f"{name} c{ci}: " + " ".join(f"{k} {v:.1e}" for k, v in e.items()) + f" (tol {tol})")
        log_parity(test="tracking_vs_reference_golden", golden=name, call=ci, err=e)
        assert x.dtype == c["x"].dtype and grads[0].dtype == c["g_x_ref"].dtype
        # x, u are fp32 casts of the fp64 iterate: 1e-6 is below one fp32 ulp of the larger entries -> 2e-7 floor of the cast
        assert e["x"] < max(tol["xu"], 1e-6) and e["u"] < max(tol["xu"] * 10, 1e-6), (ci, e)
        assert e["lam"] < tol["lam"], (ci, e)
        assert all(e[k] < max(tol["grad"], 2e-6) for k in e if k.startswith("g_")), (ci, e)


def test_tracking_mpc_obstacle_and_state_estimator_routes():
    """Tracking_MPC picks Obstacle_MPC for `"obstacles" in args.env` (policies.py:1180-1198) and forwards
    state_estimator=True (policies.py:321); both must equal the corresponding MPC called directly."""
    from deq_mpc_corl_b200.al_mpc import Obstacle_MPC
    from deq_mpc_corl_b200.tracking import Tracking_MPC

    model, T, B = "flying_cartpole", 5, 24
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    inp = make_inputs(model, T, B, 91)
    dyn = make_dyn(model)
    g = torch.Generator().manual_seed(8)
    obstacles = 2.0 * torch.randn(40, 3, generator=g, dtype=torch.float64)
    x_ref = inp["x_ref"].clone()
    x_ref[:, :, :3] = obstacles[torch.randint(0, 40, (B,), generator=g)][:, None] + 0.3 * torch.randn(B, T, 3, generator=g, dtype=torch.float64)
    x0 = x_ref[:, 0].clone()
    mk_args = lambda env_name: types.SimpleNamespace(T=T, bsz=B, Q=torch.tensor(spec["Q"]), R=torch.tensor(spec["R"]),
                                                     dtype="double", device=DEV, solver_type="al", env=env_name,
                                                     rho_init_max=1e4)
    env = types.SimpleNamespace(nx=nx, nu=nu, nq=8, dt=spec["dt"], dynamics=dyn, dynamics_derivatives=dyn,
                                obstacle_positions=obstacles, obstacle_radius=0.5,
                                action_space=types.SimpleNamespace(low=inp["u_lower"].numpy(), high=inp["u_upper"].numpy()))
    Qd, q, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"], torch.float64)
    # ---- obstacles
    tm = Tracking_MPC(mk_args("FlyingCartpole_obstacles-v0"), env)
    assert isinstance(tm.ctrl, Obstacle_MPC)
    tm.reinitialize(x_ref.to(DEV), None)
    x, u, _ = tm(x0.to(DEV), None, x_ref.to(DEV), inp["u_ref"].to(DEV), al_iters=2)
    ctrl = Obstacle_MPC(nx, nu, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, env=env)
    ctrl.reinitialize(x_ref.to(DEV), None)
    ctrl.x_init, ctrl.u_init = x_ref.to(DEV), inp["u_ref"].to(DEV)
    x2, u2, _ = ctrl(x0.to(DEV), QuadCost(Qd.to(DEV), q.to(DEV), None), dyn, dyn)
    assert torch.equal(x, x2) and torch.equal(u, u2)
    assert (tm.ctrl.lamda_prev[:, T * nx + T * 2 * nu:] > 0).any(), "no obstacle row became active: weak test"
    # ---- state estimator
    se = Tracking_MPC(mk_args("FlyingCartpole-v0"), env, state_estimator=True)
    assert se.ctrl.state_estimator and se.ctrl.neq == (T - 1) * nx and se.ctrl.nineq == 0
    u_data = (0.1 * torch.rand(B, T, nu, generator=g, dtype=torch.float64)).to(DEV)
    se.reinitialize(x_ref.to(DEV), None)
    xs, us, _ = se(x0.to(DEV), None, x_ref.to(DEV), u_data, al_iters=2)
    assert torch.equal(us, u_data.float()) and torch.isfinite(xs).all()
    assert se.ctrl.lamda_prev.shape == (B, (T - 1) * nx)


@pytest.mark.parametrize("model,T,B,se", [("rex_quadrotor", 5, 131, False), ("pendulum", 5, 64, False),
                                          ("cartpole1l", 9, 33, False), ("cartpole2l", 3, 17, True),
                                          ("rex_quadrotor", 2, 1, False), ("pendulum", 33, 6, False)])
def test_fused_newton_tail_equals_three_kernel_path(model, T, B, se, monkeypatch):
    """desc.select = ALMPC_SELECT_FUSE (k_newton_pair: factor + line search + steering in one kernel, north_star (3))
    against the default three-kernel Newton step over two chained calls (the second runs the fused dual update):
    same Newton-step counts, bit-identical iterates, multipliers and gradients."""
    spec = MODEL_SPECS[model]
    inp = make_inputs(model, T, B, 606 + T)
    dyn = make_dyn(model, dt=spec["dt"])

    def chain(fuse):
        # the fused kernel continues the half-warp-per-problem factor phase: compare with THAT three-kernel path
        monkeypatch.setenv("ALMPC_FAC_PAIR", "1")
        if fuse:
            monkeypatch.setenv("ALMPC_FUSE", "1")
        else:
            monkeypatch.delenv("ALMPC_FUSE", raising=False)
        ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, state_estimator=se)
        ctrl.reinitialize(inp["x_ref"].to(DEV))
        ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
        out, x_ref = [], inp["x_ref"]
        for _ in range(2):
            Qd, q, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"])
            ql = q.to(DEV).requires_grad_(True)
            x, u, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), ql, None), dyn, dyn)
            g, = torch.autograd.grad((x, u), ql, (torch.ones_like(x), torch.ones_like(u)))
            out.append((ctrl.last_xu.clone(), ctrl.lamda_prev.clone(), g.clone(), ctrl.dyn_res_prev.clone(),
                        ctrl.last_status.clone(), ctrl.newton_steps()))
            x_ref = 0.7 * x_ref + 0.3 * x.detach().cpu()
            x_ref[:, 0] = inp["x0"]
        return out

    for a, b in zip(chain(False), chain(True)):
        assert a[5] == b[5]
        for ta, tb in zip(a[:5], b[:5]):
            assert torch.equal(ta, tb)


@pytest.mark.parametrize("model,T,B,se,dtype", [
    ("rex_quadrotor", 5, 131, False, torch.float64), ("rex_quadrotor", 2, 1, False, torch.float64),
    ("rex_quadrotor", 9, 67, True, torch.float64), ("rex_quadrotor", 17, 40, False, torch.float64),
    ("rex_quadrotor", 5, 257, False, torch.float32), ("rex_quadrotor", 4, 33, True, torch.float32),
    ("flying_cartpole", 10, 65, False, torch.float64), ("flying_cartpole", 3, 7, True, torch.float64),
    ("flying_cartpole", 6, 33, False, torch.float32), ("pendulum", 5, 77, False, torch.float64),
    ("cartpole1l", 12, 31, False, torch.float64), ("cartpole2l", 7, 19, True, torch.float64),
    ("cartpole1l", 5, 40, False, torch.float32)])
def test_fragment_layout_factor_kernel_matches_half_warp_kernel(model, T, B, se, dtype, monkeypatch):
    """k_factor_solve_mma (blocked Cholesky in FP64-MMA fragment layout: the default for fp64 blocks wider than 8 —
    quadrotor, flying cartpole — and selectable everywhere, ALMPC_SELECT_FAC_MMA: 1, 2 and 3 column tiles, padded
    blocks) against the right-looking row-per-lane kernels (ALMPC_SELECT_FAC_PAIR, which falls back to k_factor_solve
    where the half-warp mapping does not apply) over two chained calls (the second runs the fused dual update): same
    Newton-step counts and status, iterates / multipliers / gradients equal to rounding.  The kernels order their
    floating-point operations differently, so bit-identity is not expected."""
    spec = MODEL_SPECS[model]
    inp = make_inputs(model, T, B, 909 + T)
    dyn = make_dyn(model, dt=spec["dt"])

    def chain(pair):
        monkeypatch.delenv("ALMPC_FAC_PAIR", raising=False)
        monkeypatch.delenv("ALMPC_FAC_MMA", raising=False)
        monkeypatch.setenv("ALMPC_FAC_PAIR" if pair else "ALMPC_FAC_MMA", "1")  # (fp32 defaults to the half-warp kernel)
        ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, state_estimator=se)
        ctrl.reinitialize(inp["x_ref"].to(DEV).to(dtype))
        ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
        out, x_ref = [], inp["x_ref"]
        for _ in range(2):
            Qd, q, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"], dtype)
            ql = q.to(dtype).to(DEV).requires_grad_(True)
            x, u, _ = ctrl(inp["x0"].to(DEV).to(dtype), QuadCost(Qd.to(dtype).to(DEV), ql, None), dyn, dyn)
            g, = torch.autograd.grad((x, u), ql, (torch.ones_like(x), torch.ones_like(u)))
            out.append((ctrl.last_xu.clone(), ctrl.lamda_prev.clone(), g.clone(), ctrl.last_status.clone(), ctrl.newton_steps()))
            x_ref = 0.7 * x_ref + 0.3 * x.detach().cpu()
            x_ref[:, 0] = inp["x0"]
        return out

    # (rounding differences grow with the conditioning of H: 2e-8 in the multipliers of the pendulum's second call)
    tol = 1e-6 if dtype == torch.float64 else 2e-3
    for ci, (a, b) in enumerate(zip(chain(False), chain(True))):
        assert a[4] == b[4], (ci, a[4], b[4])
        assert torch.equal(a[3], b[3])
        for k, (ta, tb) in enumerate(zip(a[:3], b[:3])):
            if tb.numel():
                assert rel_err(ta.cpu().double(), tb.cpu().double()) < tol, (ci, k)


def test_c_abi_used_directly_like_integration_md():
    """The ctypes stub of INTEGRATION.md section 3, written out independently of deq_mpc_corl_b200/_lib.py: plain
    pointers, sizes and a stream are all the library needs; the result equals the Python drop-in's."""
    import ctypes

    from deq_mpc_corl_b200 import _lib as product_lib

    class Desc(ctypes.Structure):  # struct almpc_desc (include/almpc.h)
        _fields_ = [("B", ctypes.c_int32), ("T", ctypes.c_int32), ("dyn_id", ctypes.c_int32),
                    ("al_iter", ctypes.c_int32), ("max_newton", ctypes.c_int32), ("stream_mode", ctypes.c_int32),
                    ("rho_f32", ctypes.c_int32), ("flags", ctypes.c_int32),
                    ("rho_mult", ctypes.c_double), ("rho_max", ctypes.c_double),
                    ("dyn_params", ctypes.c_double * 48), ("u_lower", ctypes.c_double * 8), ("u_upper", ctypes.c_double * 8),
                    ("n_obs", ctypes.c_int32), ("select", ctypes.c_int32), ("obs_radius", ctypes.c_double),
                    ("obstacles", ctypes.c_void_p), ("profile_ms", ctypes.c_void_p), ("stop_stats", ctypes.c_void_p),
                    ("stop_reduce", ctypes.c_void_p), ("stop_reduce_user", ctypes.c_void_p)]

    lib = ctypes.CDLL(product_lib.LIB_PATH)
    vp = ctypes.c_void_p
    for fn, nargs in (("almpc_forward_f64", 15), ("almpc_backward_f64", 7)):
        getattr(lib, fn).argtypes = [ctypes.POINTER(Desc)] + [vp] * nargs
        getattr(lib, fn).restype = ctypes.c_int
    for fn in ("almpc_factor_elems", "almpc_workspace_bytes"):
        getattr(lib, fn).argtypes = [ctypes.POINTER(Desc)]
        getattr(lib, fn).restype = ctypes.c_size_t
    lib.almpc_last_error.restype = ctypes.c_char_p

    model, T, B = "rex_quadrotor", 5, 100
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    n, m = nx + nu, T * nx + 2 * T * nu
    inp = make_inputs(model, T, B, 202)
    dyn = make_dyn(model)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
    d = Desc(B=B, T=T, dyn_id=dyn.dyn_id, al_iter=2, max_newton=4, stream_mode=0, rho_f32=1, flags=0, rho_mult=10.0,
             rho_max=1e8)
    for i, v in enumerate(dyn.param_block()):
        d.dyn_params[i] = v
    for i in range(nu):
        d.u_lower[i], d.u_upper[i] = float(inp["u_lower"][i]), float(inp["u_upper"][i])
    f64 = dict(dtype=torch.float64, device=DEV)
    x0 = inp["x0"].double().to(DEV).contiguous()
    Qd_d, q_d = Qd.to(DEV).contiguous(), q.to(DEV).contiguous()
    xu = torch.cat([inp["x_ref"], inp["u_ref"]], -1).double().to(DEV).contiguous()
    lam, rho = torch.zeros(B, m, **f64), torch.ones(B, **f64)
    factor = torch.empty(B, lib.almpc_factor_elems(ctypes.byref(d)), **f64)
    ws = torch.empty(lib.almpc_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8, device=DEV)
    xf, uf = torch.empty(B, T, nx, device=DEV), torch.empty(B, T, nu, device=DEV)
    dyn_res, status = torch.empty(B, **f64), torch.empty(B, dtype=torch.int32, device=DEV)
    info = torch.empty(36, dtype=torch.int32, device=DEV)
    p = lambda t: vp(0 if t is None else t.data_ptr())
    stream = vp(torch.cuda.current_stream().cuda_stream)
    rc = lib.almpc_forward_f64(ctypes.byref(d), p(x0), p(Qd_d), p(q_d), p(xu), p(lam), p(rho), p(factor), p(xf), p(uf),
                               p(dyn_res), p(status), p(info), None, p(ws), stream)
    assert rc == 0, lib.almpc_last_error()
    g = torch.Generator().manual_seed(5)
    grad = torch.randn(B, T, n, generator=g, dtype=torch.float64).to(DEV)
    gQ, gq = torch.empty_like(xu), torch.empty_like(xu)
    rc = lib.almpc_backward_f64(ctypes.byref(d), p(factor), p(xu), p(grad), p(status), p(gQ), p(gq), stream)
    assert rc == 0, lib.almpc_last_error()
    torch.cuda.synchronize()
    # the Python drop-in on the same inputs
    ctrl = MPC(nx, nu, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
    ctrl.reinitialize(inp["x_ref"].to(DEV))
    ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
    ql = q_d.clone().requires_grad_(True)
    x, u, _ = ctrl(x0, QuadCost(Qd_d, ql, None), dyn, dyn)
    qg, = torch.autograd.grad((x, u), ql, (grad[..., :nx].float(), grad[..., nx:].float()))
    assert torch.equal(xu, ctrl.last_xu) and torch.equal(xf, x) and torch.equal(uf, u)
    assert torch.equal(lam, ctrl.lamda_prev) and (status == 3).all()
    assert info[0].item() == 2 and info[4:6].tolist() == ctrl.newton_steps()
    # the drop-in feeds fp32 output gradients (AL_mpc.py:337-338 casts x, u to float)
    assert rel_err(gq.cpu(), qg.cpu()) < 1e-6


def test_cpp_host_example_runs():
    """examples/c_host: a C++ / CUDA-runtime host (no Python, no torch) on the same C ABI; built by build()."""
    import re
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "c_host")
    if not os.path.exists(exe):
        pytest.skip("examples/c_host not built (python -m deq_mpc_corl_b200.build)")
    out = subprocess.run([exe, "777"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"AL iterations (\d+), newton steps \[(\d+), (\d+)\].*status ok (\d+)/(\d+)", out.stdout)
    assert m, out.stdout
    assert int(m.group(1)) == 2 and 1 <= int(m.group(2)) <= 4 and 1 <= int(m.group(3)) <= 4
    assert m.group(4) == m.group(5) == "777"


def test_forward_backward_are_cuda_graph_capturable():
    """include/almpc.h promises no host synchronisation and no allocation inside the library: capture a whole
    forward + backward in a CUDA graph, replay it on fresh inputs, compare with the eager result."""
    from deq_mpc_corl_b200.solver import SolveConfig, raw_backward, raw_forward

    model, T, B = "rex_quadrotor", 5, 512
    spec = MODEL_SPECS[model]
    n, m = spec["nx"] + spec["nu"], T * spec["nx"] + 2 * T * spec["nu"]
    dyn = make_dyn(model)
    inpA, inpB = make_inputs(model, T, B, 31), make_inputs(model, T, B, 32)
    cfg = SolveConfig(dyn=dyn, T=T, u_lower=inpA["u_lower"].tolist(), u_upper=inpA["u_upper"].tolist(), al_iter=2, rho_f32=True)

    def tensors(inp):
        Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
        return (inp["x0"].double().to(DEV), Qd.to(DEV), q.to(DEV),
                torch.cat([inp["x_ref"], inp["u_ref"]], -1).double().to(DEV).contiguous())

    def eager(inp):
        x0, Qd, q, xu = tensors(inp)
        lam, rho = torch.zeros(B, m, dtype=torch.float64, device=DEV), torch.ones(B, dtype=torch.float64, device=DEV)
        out = raw_forward(cfg, x0, Qd, q, xu, lam, rho)
        _, gq = raw_backward(cfg, out["factor"], xu, torch.ones_like(xu), need_Q=False)
        return xu, lam, gq, out["info"]

    # static buffers of the graph
    x0, Qd, q, xu0 = tensors(inpA)
    xu, lam, rho = xu0.clone(), torch.zeros(B, m, dtype=torch.float64, device=DEV), torch.ones(B, dtype=torch.float64, device=DEV)
    gsrc = torch.ones_like(xu)
    eager(inpA)  # warm-up outside capture (lazy module loading, attribute calls)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = raw_forward(cfg, x0, Qd, q, xu, lam, rho)
        _, gq = raw_backward(cfg, out["factor"], xu, gsrc, need_Q=False)
    for inp in (inpA, inpB):
        a, b, c, d = tensors(inp)
        x0.copy_(a); Qd.copy_(b); q.copy_(c); xu.copy_(d); lam.zero_(); rho.fill_(1.0)
        graph.replay()
        torch.cuda.synchronize()
        xu_e, lam_e, gq_e, info_e = eager(inp)
        assert torch.equal(xu, xu_e) and torch.equal(lam, lam_e) and torch.equal(gq, gq_e)
        assert torch.equal(out["info"], info_e)


def test_mpc_cuda_graph_mode_equals_eager():
    """MPC.cuda_graph = True replays the forward launch sequence as one CUDA graph: bit-identical to the eager path
    over a chain of calls (duals and penalties carried, fresh inputs each call), gradients included."""
    model, T, B = "cartpole1l", 5, 200
    spec = MODEL_SPECS[model]
    dyn = make_dyn(model, dt=spec["dt"])
    inp = make_inputs(model, T, B, 77)

    def chain(graph):
        ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
        ctrl.cuda_graph = graph
        ctrl.reinitialize(inp["x_ref"].to(DEV))
        ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
        outs = []
        x_ref = inp["x_ref"].clone()
        for it in range(3):
            Qd, q, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"])
            ql = q.to(DEV).requires_grad_(True)
            x, u, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), ql, None), dyn, dyn)
            qg, = torch.autograd.grad((x, u), ql, (torch.ones_like(x), torch.ones_like(u)))
            outs.append((x.clone(), u.clone(), qg.clone(), ctrl.lamda_prev.clone(), ctrl.rho_prev.clone(), ctrl.newton_steps()))
            x_ref = 0.7 * x_ref + 0.3 * x.detach().cpu()
            x_ref[:, 0] = inp["x0"]
        return outs, ctrl

    eager, _ = chain(False)
    graphed, ctrl = chain(True)
    assert len(ctrl._graph_cache.entries) == 1
    for a, b in zip(eager, graphed):
        for ta, tb in zip(a[:5], b[:5]):
            assert torch.equal(ta, tb)
        assert a[5] == b[5]


@pytest.mark.parametrize("model,T,B", [("pendulum", 5, 200), ("rex_quadrotor", 5, 48)])
def test_deq_chain_without_resync_matches_oracle(model, T, B):
    """BASELINE config 1 shape (pendulum, T=5, B=200, fp64, 6 DEQ iterations of al_iter=2 with lam / rho and the
    warm start carried across calls, rho 1 -> 1e12) and the headline model (quadrotor, through the fragment-layout factor
    kernel): the CUDA module and the oracle run the whole chain from the same inputs with NO re-synchronisation; the
    synthetic "network" pulls x_ref towards each side's own solution."""
    from oracle import almpc_oracle as orc
    from oracle.problem_gen import make_oracle_model

    spec = MODEL_SPECS[model]
    inp = make_inputs(model, T, B, 1000)
    dyn = make_dyn(model)
    ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
    ompc = orc.OracleMPC(make_oracle_model(model), T, inp["u_lower"], inp["u_upper"], B)
    xr_c, xr_o = inp["x_ref"].to(DEV), inp["x_ref"].clone()
    ctrl.reinitialize(xr_c, None)
    ompc.reinitialize(xr_o)
    ctrl.x_init, ctrl.u_init = xr_c.clone(), inp["u_ref"].to(DEV)
    ompc.x_init, ompc.u_init = xr_o.clone(), inp["u_ref"].clone()
    for it in range(6):
        Qd, q_c, _ = tracking_cost(inp["Qd"].to(DEV), xr_c, inp["u_ref"].to(DEV))
        _, q_o, _ = tracking_cost(inp["Qd"], xr_o, inp["u_ref"])
        x_c, u_c, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qd, q_c, None), dyn, dyn)
        tr = orc.SolveTrace()
        x_o, u_o, _, xu_o = ompc.solve(inp["x0"], Qd.cpu(), q_o, trace=tr)
        assert ctrl.newton_steps() == tr.newton_steps, f"DEQ iteration {it}"
        assert torch.equal(ctrl.rho_prev.cpu(), ompc.rho_prev) and ctrl.rho_prev.dtype == torch.float32
        err = rel_err(ctrl.last_xu.cpu(), xu_o)
        # the chain amplifies differences with rho: tight early, documented budget later
        assert err < (1e-8 if it < 3 else 1e-5), (it, err)
        xr_c = 0.7 * xr_c + 0.3 * x_c.detach()
        xr_c[:, 0] = inp["x0"].to(DEV)
        xr_o = 0.7 * xr_o + 0.3 * x_o
        xr_o[:, 0] = inp["x0"]
    assert float(ctrl.rho_prev.max()) == float(torch.tensor(1e12, dtype=torch.float32))


@pytest.mark.parametrize("pair", [True, False])
def test_fp32_high_penalty_flags_instead_of_failing(pair, monkeypatch):
    """fp32 loses positive definiteness from rho ~ 1e4 on.  The solver must keep running, keep the
    iterate finite and report the broken pivots per problem (ALMPC_STATUS_CHOL_OK cleared) — never hang or NaN.
    pair=True: the half-warp-per-problem kernel (the fp32 default), which factors in fp32 and does lose pivots;
    pair=False: the fragment-layout kernel (ALMPC_SELECT_FAC_MMA) keeps the blocks in FP64 accumulators (fp32 only at the interface), so its pivots survive."""
    monkeypatch.setenv("ALMPC_FAC_PAIR" if pair else "ALMPC_FAC_MMA", "1")
    model, T, B = "rex_quadrotor", 5, 256
    inp = make_inputs(model, T, B, 5)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], torch.float32)
    dyn = make_dyn(model)
    ctrl = MPC(12, 4, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, dtype=torch.float32)
    ctrl.reinitialize(inp["x_ref"].to(DEV), None)
    ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
    flagged = 0
    for it in range(4):  # rho: 1 -> 1e8
        x, u, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), q.to(DEV), None), dyn, dyn)
        st = ctrl.last_status.cpu()
        assert torch.isfinite(x).all() and torch.isfinite(u).all()
        assert ((st & _lib.STATUS_FINITE) == _lib.STATUS_FINITE).all()
        flagged = int(((st & _lib.STATUS_CHOL_OK) == 0).sum())
    if pair:
        assert flagged > 0  # at rho = 1e6..1e7 fp32 pivots do go non-positive for some problems


def test_nan_propagation_and_dual_shape_check():
    """Round-1 advisor items.  (1) A NaN that no dynamics row sees (last-stage control) must reach ||c+|| like
    torch.clamp(min=0) propagates it (AL_mpc.py:317, al_utils.py:419), not vanish in max(x, 0).  (2) A NaN in the
    UPSTREAM gradient propagates to that problem's q_grad (the reference's cholesky_solve would) instead of turning
    into an exact zero; the other problems are untouched.  (3) `lamda_prev` of the wrong shape raises instead of
    silently restarting the duals from zero."""
    model, T, B = "rex_quadrotor", 5, 12
    spec = MODEL_SPECS[model]
    inp = make_inputs(model, T, B, 31)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
    dyn = make_dyn(model)

    def solve(u_ref, gx_mod=None):
        ctrl = MPC(12, 4, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
        ctrl.reinitialize(inp["x_ref"].to(DEV))
        ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), u_ref.to(DEV)
        ql = q.to(DEV).requires_grad_(True)
        x, u, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), ql, None), dyn, dyn)
        gx, gu = torch.ones_like(x), torch.ones_like(u)
        if gx_mod is not None:
            gx_mod(gx)
        qg, = torch.autograd.grad((x, u), ql, (gx, gu))
        return ctrl, x, u, qg

    # (1)
    u_nan = inp["u_ref"].clone()
    u_nan[0, -1, 2] = float("nan")
    ctrl, x, u, _ = solve(u_nan)
    assert torch.isnan(ctrl.dyn_res_prev[0]) and torch.isfinite(ctrl.dyn_res_prev[1:]).all()
    assert torch.isfinite(x[1:]).all() and torch.isfinite(u[1:]).all()
    # (2)
    def poison(gx):
        gx[3, 2, 5] = float("nan")
    ctrl, x, u, qg = solve(inp["u_ref"], poison)
    assert (ctrl.last_status == 3).all()
    assert torch.isnan(qg[3]).any() and torch.isfinite(qg[:3]).all() and torch.isfinite(qg[4:]).all()
    # (3)
    ctrl.lamda_prev = torch.zeros(B + 1, ctrl.neq + ctrl.nineq, dtype=torch.float64, device=DEV)
    with pytest.raises(_lib.AlmpcError, match="lamda_prev"):
        ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), q.to(DEV), None), dyn, dyn)


@pytest.mark.parametrize("model,T,B", [("rex_quadrotor", 5, 8192 + 5), ("flying_cartpole", 5, 8192 + 3), ("pendulum", 9, 8192)])
def test_scratch_ring_is_transparent(model, T, B, monkeypatch):
    """Batches >= 8192 write the block factors of non-final AL iterations into the L2-resident scratch ring instead of
    `factor` (both factor kernels: half-warp-per-problem and row-per-lane).  Results, incl. the gradients that come
    from the factor of the LAST step, must be bit-identical with the ring switched off."""
    spec = MODEL_SPECS[model]
    inp = make_inputs(model, T, B, 808)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
    dyn = make_dyn(model, dt=spec["dt"])

    def run(no_ring):
        if no_ring:
            monkeypatch.setenv("ALMPC_NO_RING", "1")
        else:
            monkeypatch.delenv("ALMPC_NO_RING", raising=False)
        ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
        ctrl.reinitialize(inp["x_ref"].to(DEV))
        ctrl.x_init, ctrl.u_init = inp["x_ref"].to(DEV), inp["u_ref"].to(DEV)
        ctrl.al_iter = 3
        ql = q.to(DEV).requires_grad_(True)
        x, u, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qd.to(DEV), ql, None), dyn, dyn)
        g, = torch.autograd.grad((x, u), ql, (torch.ones_like(x), torch.ones_like(u)))
        return ctrl.last_xu.clone(), ctrl.lamda_prev.clone(), g.clone(), ctrl.newton_steps(), ctrl.last_status.clone()

    a, b = run(False), run(True)
    assert a[3] == b[3] and (a[4] == 3).all()
    for ta, tb in zip(a[:3], b[:3]):
        assert torch.equal(ta, tb)


@pytest.mark.parametrize("name", list_tracking_golden() if cuda else [])
def test_reference_policies_module_runs_on_our_mpc(name, monkeypatch):
    """The one-import swap of INTEGRATION.md §1, exercised for real: the UNMODIFIED `deqmpc/policies.py` of the
    reference (imported from its staged copy, oracle/ref_import.py) with `policies.al_mpc` / `policies.al_utils`
    pointing at this package.  The reference's own `Tracking_MPC` code then constructs OUR `MPC` with its own
    keyword list (policies.py:1200-1216), hands it the reference's TorchScript dynamics modules
    (`dynamics.resolve` maps them to the CUDA functor), its diag-embedded [B,T,n,n] cost and its call sequence
    (reinitialize / forward / warm_start_initialize) — and must reproduce the golden that the same class produced
    on top of the reference's solver."""
    from oracle import ref_import

    if not ref_import.available():
        pytest.skip("reference copy not staged (make -C oracle pyref)")
    ref = ref_import.import_reference(with_policies=True)
    from deq_mpc_corl_b200 import al_mpc as our_al_mpc, al_utils as our_al_utils

    monkeypatch.setattr(ref.policies, "al_mpc", our_al_mpc)
    monkeypatch.setattr(ref.policies, "al_utils", our_al_utils)
    gold = Golden(name)
    spec = MODEL_SPECS[gold.model]
    T, B = gold.T, gold.B
    env = ref.rex_quadrotor.RexQuadrotor(bsz=B)
    args = types.SimpleNamespace(T=T, device=DEV, qp_iter=2, eps=1e-2, warm_start=True, bsz=B, Q=gold.t("Q"), R=gold.t("R"),
                                 dtype="double", solver_type="al", env=gold.model, rho_init_max=gold.rho_init_max)
    tm = ref.policies.Tracking_MPC(args, env)
    assert type(tm.ctrl) is our_al_mpc.MPC
    x0 = gold.t("x0").to(DEV)
    tm.reinitialize(gold.call(0)["x_ref"].to(DEV), None)
    for ci in range(gold.n_calls):
        c = gold.call(ci)
        if gold.stream_after == ci:
            tm.warm_start_initialize(c["x_ref"].to(DEV), c["u_ref"].to(DEV))
        xr = c["x_ref"].to(DEV).requires_grad_(True)
        ur = c["u_ref"].to(DEV).requires_grad_(True)
        qs = c["q_scaling"].to(DEV).requires_grad_(True) if "q_scaling" in c else None
        x, u, status = tm(x0, None, xr, ur, q_scaling=qs, al_iters=int(c["al_iters"]))
        ins = (xr, ur) + ((qs,) if qs is not None else ())
        grads = torch.autograd.grad((x, u), ins, (c["gx"].to(DEV), c["gu"].to(DEV)))
        tol = tolerances(float(c["rho_out"].double().max()), gold.dtype)
        assert bool(status) == bool(c["status"]), f"call {ci}"
        assert torch.equal(tm.ctrl.rho_prev.cpu().double(), c["rho_out"].double())
        assert rel_err(x.detach().cpu(), c["x"]) < 1e-6 and rel_err(u.detach().cpu(), c["u"]) < 1e-5, f"call {ci}"
        assert rel_err(tm.ctrl.lamda_prev.cpu(), c["lam_out"]) < tol["lam"], f"call {ci}"
        assert rel_err(grads[0].cpu(), c["g_x_ref"]) < max(tol["grad"], 2e-6), f"call {ci}"
        assert rel_err(grads[1].cpu(), c["g_u_ref"]) < max(tol["grad"], 2e-6), f"call {ci}"
        if qs is not None:
            assert rel_err(grads[2].cpu(), c["g_q_scaling"]) < max(tol["grad"], 2e-6), f"call {ci}"
