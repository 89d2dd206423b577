"""Runs the UNMODIFIED reference (`qpth.AL_mpc.MPC.forward` + autograd backward) on the host CPU.

TEST INFRASTRUCTURE ONLY: the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) and tests.  The
reference is imported through oracle/ref_import.py (from /root/reference in the build container, from the
staged copy oracle/_ref/py on the GPU box).  Nothing of libalmpc is touched here.

BASELINE.md §3 recipe: same synthetic inputs as the GPU arm (deq_mpc_corl_b200.problems), batch chunked to
<= 4096 problems (dense H, J and the 20x line-search temporaries), one warm-up call (TorchScript profiling
executor), timed repeats, `torch.set_num_threads(os.cpu_count())`.
"""
import os
import time
import types

import torch

from oracle import ref_import
from oracle.problem_gen import MODEL_SPECS, make_inputs


def _make_env(ref, model, B):
    """-> (dx, dx_jac) callables of the reference for `model` (TorchScript modules / compiled CasADi C)."""
    if model == "rex_quadrotor":
        env = ref.rex_quadrotor.RexQuadrotor(bsz=B)
        return env.dynamics, env.dynamics_derivatives
    if model == "flying_cartpole":
        env = ref.flying_cartpole2d.FlyingCartpole(bsz=B)
        return env.dynamics, env.dynamics_derivatives
    from oracle.gen_golden import CasadiDx

    pkg, nq = {"pendulum": ("pendulum1l", 1), "cartpole1l": ("cartpole1l", 2), "cartpole2l": ("cartpole2l", 3)}[model]
    cd = CasadiDx(pkg, nq, MODEL_SPECS[model]["dt"])
    return cd.dx, cd.dx_jac


def solve_chunk(ref, model, T, inp, lo, hi, dtype, al_iter=2, grads=None, stream_calls=0, rho_init_max=1e4):
    """One cold-start `MPC.forward` (+ `stream_calls` warm-started streaming calls, policies.py:1305-1310) and
    the implicit backward for problems [lo, hi).  Returns (x, u, q_grad, newton-step counts per AL iteration)."""
    spec = MODEL_SPECS[model]
    nx, nu = spec["nx"], spec["nu"]
    B = hi - lo
    AL_mpc, al_utils = ref.AL_mpc, ref.al_utils
    dx, dx_jac = _make_env(ref, model, B)
    x0 = inp["x0"][lo:hi]
    x_ref, u_ref = inp["x_ref"][lo:hi], inp["u_ref"][lo:hi]
    xu_ref = torch.cat([x_ref, u_ref], -1)
    Qfull = torch.diag(inp["Qd"].to(dtype)).repeat(B, T, 1, 1)
    p = (-(Qfull * xu_ref.unsqueeze(-2)).sum(-1)).detach().requires_grad_(True)  # policies.py:1295
    f = 0.5 * (xu_ref * (Qfull * xu_ref.unsqueeze(-2)).sum(-1)).sum(-1)
    ctrl = AL_mpc.MPC(nx, nu, T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, dtype=dtype)
    ctrl.reinitialize(x_ref, None)
    ctrl.x_init, ctrl.u_init = x_ref.detach().clone(), u_ref.detach().clone()
    ctrl.al_iter = al_iter
    # realised Newton-step counts: count line-search calls between NewtonAL.apply calls (no reference edits)
    counts, cur = [], [0]
    orig_ls, orig_newton = al_utils.line_search_newton, al_utils.NewtonAL

    def ls(*a, **k):
        cur[0] += 1
        return orig_ls(*a, **k)

    class NewtonProxy:
        @staticmethod
        def apply(*a):
            cur[0] = 0
            out = orig_newton.apply(*a)
            counts.append(cur[0])
            return out

    al_utils.line_search_newton, al_utils.NewtonAL = ls, NewtonProxy
    try:
        x, u, _ = ctrl(x0, al_utils.QuadCost(Qfull, p, f), dx, dx_jac)
        for _ in range(stream_calls):
            xi, ui = ctrl.x_init, ctrl.u_init
            ui[:, -1:] = u_ref[:, -1:]
            xi[:, -1:] = x_ref[:, -1:]
            ctrl.warm_start_initialize(xi, ui, types.SimpleNamespace(rho_init_max=rho_init_max))
            x, u, status = ctrl(x0, al_utils.QuadCost(Qfull, p, f), dx, dx_jac)
            if status:
                break
    finally:
        al_utils.line_search_newton, al_utils.NewtonAL = orig_ls, orig_newton
    gx = torch.ones_like(x) if grads is None else grads[0][lo:hi]
    gu = torch.ones_like(u) if grads is None else grads[1][lo:hi]
    qg, = torch.autograd.grad((x, u), p, (gx, gu))
    return x.detach(), u.detach(), qg, counts


def run(model, T, B, dtype, seed, chunk=4096, al_iter=2, threads=None, stream_calls=0):
    """-> dict(seconds, solves_per_s, newton_steps (first chunk), cores)."""
    ref = ref_import.import_reference()
    threads = threads or os.cpu_count()
    if torch.get_num_threads() != threads:
        torch.set_num_threads(threads)
    inp = make_inputs(model, T, B, seed)
    t0 = time.perf_counter()
    steps = None
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        _, _, _, counts = solve_chunk(ref, model, T, inp, lo, hi, dtype, al_iter=al_iter, stream_calls=stream_calls)
        steps = steps or counts
    dt = time.perf_counter() - t0
    return dict(seconds=dt, solves_per_s=B / dt, newton_steps=steps, cores=threads, root=ref.root)
