#!/usr/bin/env python
"""Synthetic AL-MPC batch sweep (BASELINE.json config 5): B in 2^10..2^20, nx in {2,4,12}, T in {5,10,20},
fp64 and fp32, forward(al_iter=2)+backward, device-resident inputs, CUDA events.  Prints a markdown table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from deq_mpc_corl_b200 import _lib
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost
from deq_mpc_corl_b200.solver import SolveConfig, raw_backward, raw_forward

dev = "cuda:0"

def run(model, T, B, dtype, reps=3):
    spec = MODEL_SPECS[model]
    base = min(B, 8192)
    inp = make_inputs(model, T, base, 1003)
    rep = B // base
    inp = {k: (v.repeat(rep, *([1] * (v.dim() - 1))) if v.dim() > 1 else v) for k, v in inp.items()}
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], dtype)
    dyn = make_dyn(model, dt=spec["dt"]); n = spec["nx"] + spec["nu"]; m = T * spec["nx"] + 2 * T * spec["nu"]
    cfg = SolveConfig(dyn=dyn, T=T, u_lower=inp["u_lower"].tolist(), u_upper=inp["u_upper"].tolist(), al_iter=2, rho_f32=(dtype == torch.float64))
    x0 = inp["x0"].to(dtype).to(dev); Qd = Qd.to(dev); q = q.to(dev)
    xu0 = torch.cat([inp["x_ref"], inp["u_ref"]], -1).to(dtype).to(dev).contiguous()
    g = torch.randn(B, T, n, dtype=dtype, device=dev)
    best = 1e9
    for r in range(reps + 1):
        xu = xu0.clone(); lam = torch.zeros(B, m, dtype=dtype, device=dev); rho = torch.ones(B, dtype=dtype, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = raw_forward(cfg, x0, Qd, q, xu, lam, rho); raw_backward(cfg, out["factor"], xu, g); e1.record()
        torch.cuda.synchronize()
        if r > 0: best = min(best, e0.elapsed_time(e1))
    info = out["info"].cpu().tolist()
    ok = int((out["status"] == 3).sum().item())
    return best, info[_lib.INFO_NEWTON0:_lib.INFO_NEWTON0 + info[0]], ok

if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    print("| model | nx | T | dtype | B | ms (fwd+bwd) | Newton steps | solves/s | status ok |")
    print("|---|---|---|---|---|---|---|---|---|")
    models = [("pendulum", 2), ("cartpole1l", 4), ("rex_quadrotor", 12)]
    for model, nx in models:
        for T in (5, 10, 20):
            for dtype in (torch.float64, torch.float32):
                for lb in ((10, 14, 16) if quick else (10, 12, 14, 16, 18, 20)):
                    B = 1 << lb
                    if model == "rex_quadrotor" and (B * T > (1 << 20) * 5):   # keep the factor buffer < ~20 GB
                        continue
                    try:
                        ms, steps, ok = run(model, T, B, dtype)
                        print(f"| {model} | {nx} | {T} | {'f64' if dtype == torch.float64 else 'f32'} | {B} | {ms:.3f} | {steps} | {B / ms * 1e3:.3e} | {ok}/{B} |", flush=True)
                    except Exception as ex:
                        print(f"| {model} | {nx} | {T} | {dtype} | {B} | ERROR {str(ex)[-150:]} | | | |", flush=True)
                    torch.cuda.empty_cache()
