#!/usr/bin/env python
"""BASELINE.json `configs` as a table: device-timed (CUDA events) through the product API (`al_mpc.MPC`), with
the realised Newton-step counts.  Config 2 reproduces the streaming schedule of SURVEY.md §8d: a 6-call cold chain
(rho 1 -> 1e12), then 2 streaming steps of warm-started calls that end on the rho_max exit (2+2+1 AL iterations)."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from deq_mpc_corl_b200.al_mpc import MPC
from deq_mpc_corl_b200.al_utils import QuadCost
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost

dev = "cuda:0"


def setup(model, T, B, dtype, seed=1003):
    spec = MODEL_SPECS[model]
    inp = make_inputs(model, T, B, seed)
    dyn = make_dyn(model, dt=spec["dt"])
    ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, dtype=dtype)
    return spec, {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}, dyn, ctrl


def timed(fn, reps=3):
    best = 1e9
    out = None
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        if r > 0: best = min(best, e0.elapsed_time(e1))
    return best, out


def cold_fwd_bwd(model, T, B, dtype):
    spec, inp, dyn, ctrl = setup(model, T, B, dtype)
    Qd, q0, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], dtype)
    g = torch.Generator().manual_seed(1)
    gx = torch.randn(B, T, spec["nx"], generator=g).to(dev); gu = torch.randn(B, T, spec["nu"], generator=g).to(dev)

    def step():
        ctrl.reinitialize(inp["x_ref"], None)
        ctrl.x_init, ctrl.u_init = inp["x_ref"], inp["u_ref"]
        q = q0.clone().requires_grad_(True)
        x, u, _ = ctrl(inp["x0"], QuadCost(Qd, q, None), dyn, dyn)
        torch.autograd.grad((x, u), q, (gx, gu))
        return ctrl
    ms, c = timed(step)
    return ms, c.newton_steps(), 1


def streaming_chain(model, T, B, dtype, deq_iter=6, streaming_steps=2):
    spec, inp, dyn, ctrl = setup(model, T, B, dtype)
    args = types.SimpleNamespace(rho_init_max=1e4)
    log = []

    def chain():
        log.clear()
        x_ref = inp["x_ref"].clone()
        Qd, _, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"], dtype)
        ctrl.reinitialize(x_ref, None)
        ctrl.x_init, ctrl.u_init = x_ref.clone(), inp["u_ref"].clone()
        calls = 0
        for i in range(deq_iter):
            _, q, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"], dtype)
            x, u, st = ctrl(inp["x0"], QuadCost(Qd, q, None), dyn, dyn)
            x_ref = 0.7 * x_ref + 0.3 * x.detach(); x_ref[:, 0] = inp["x0"]
            calls += 1
        for s in range(streaming_steps):
            ctrl.warm_start_initialize(ctrl.x_init, ctrl.u_init, args)
            n_al = []
            for i in range(deq_iter):
                _, q, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"], dtype)
                x, u, st = ctrl(inp["x0"], QuadCost(Qd, q, None), dyn, dyn)
                n_al.append(len(ctrl.newton_steps()))
                x_ref = 0.7 * x_ref + 0.3 * x.detach(); x_ref[:, 0] = inp["x0"]
                calls += 1
                if st:
                    break
            log.append(n_al)
        return calls
    ms, calls = timed(chain, reps=2)
    return ms, log, calls


if __name__ == "__main__":
    rows = []
    rows.append(("pendulum deq-mpc-nn T=5 B=200 fp64 (cold fwd+bwd)",) + cold_fwd_bwd("pendulum", 5, 200, torch.float64))
    rows.append(("cartpole-1l T=5 B=4096 fp64: 6-call cold chain + 2 streaming steps",) + streaming_chain("cartpole1l", 5, 4096, torch.float64))
    rows.append(("flying_cartpole T=10 B=16384 fp64 (cold fwd+bwd)",) + cold_fwd_bwd("flying_cartpole", 10, 16384, torch.float64))
    rows.append(("rex_quadrotor T=5 B=65536 fp64 (cold fwd+bwd)",) + cold_fwd_bwd("rex_quadrotor", 5, 65536, torch.float64))
    rows.append(("rex_quadrotor T=5 B=65536 fp32 (cold fwd+bwd)",) + cold_fwd_bwd("rex_quadrotor", 5, 65536, torch.float32))
    print("| config | ms (device, through al_mpc.MPC) | MPC calls | solves/s (calls x B / time) | Newton steps / AL iterations |")
    print("|---|---|---|---|---|")
    for name, ms, steps, calls in rows:
        B = int(name.split("B=")[1].split()[0])
        print(f"| {name} | {ms:.3f} | {calls} | {calls * B / ms * 1e3:.3e} | {steps} |")
