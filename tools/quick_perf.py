#!/usr/bin/env python
"""Fast GPU iteration loop: a parity spot-check against two reference goldens + forward/backward timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from deq_mpc_corl_b200 import _lib
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost
from deq_mpc_corl_b200.solver import SolveConfig, raw_backward, raw_forward
from golden_util import Golden, rel_err

dev = "cuda:0"

def parity(name):
    g = Golden(name); c = g.call(0); spec = MODEL_SPECS[g.model]
    dyn = make_dyn(g.model, dt=spec["dt"]); n = spec["nx"] + spec["nu"]
    cfg = SolveConfig(dyn=dyn, T=g.T, u_lower=g.t("u_lower").tolist(), u_upper=g.t("u_upper").tolist(), al_iter=int(c["al_iter"]), rho_f32=True)
    xu = c["xu_init"].to(g.dtype).to(dev).contiguous()
    lam = c["lam_in"].to(g.dtype).to(dev).contiguous(); rho = c["rho_in"].to(g.dtype).reshape(-1).to(dev).contiguous()
    Qd = g.t("Qd").to(g.dtype).expand(g.B, g.T, n).contiguous().to(dev)
    out = raw_forward(cfg, g.t("x0").to(g.dtype).to(dev), Qd, c["q"].to(g.dtype).to(dev).contiguous(), xu, lam, rho)
    gxu = torch.cat([c["gx"], c["gu"]], -1).to(g.dtype).to(dev)
    gQ, gq = raw_backward(cfg, out["factor"], xu, gxu)
    info = out["info"].cpu().tolist()
    steps = info[_lib.INFO_NEWTON0:_lib.INFO_NEWTON0 + info[0]]
    print(f"parity {name}: steps {steps} vs {c['nsteps'].tolist()} xu {rel_err(xu.cpu(), c['xu64']):.1e} lam {rel_err(lam.cpu(), c['lam_out']):.1e} q_grad {rel_err(gq.cpu(), c['q_grad']):.1e}")

def timing(model, T, B, dtype=torch.float64, reps=5):
    spec = MODEL_SPECS[model]; inp = make_inputs(model, T, B, 1003)
    Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], dtype)
    dyn = make_dyn(model, dt=spec["dt"]); n = spec["nx"] + spec["nu"]; m = T * spec["nx"] + 2 * T * spec["nu"]
    cfg = SolveConfig(dyn=dyn, T=T, u_lower=inp["u_lower"].tolist(), u_upper=inp["u_upper"].tolist(), al_iter=2, rho_f32=True)
    x0 = inp["x0"].to(dtype).to(dev); Qd = Qd.to(dev); q = q.to(dev)
    xu0 = torch.cat([inp["x_ref"], inp["u_ref"]], -1).to(dtype).to(dev).contiguous()
    g = torch.randn(B, T, n, dtype=dtype, device=dev)
    best_f = best_b = 1e9
    for r in range(reps):
        xu = xu0.clone(); lam = torch.zeros(B, m, dtype=dtype, device=dev); rho = torch.ones(B, dtype=dtype, device=dev)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); out = raw_forward(cfg, x0, Qd, q, xu, lam, rho); e[1].record()
        raw_backward(cfg, out["factor"], xu, g); e[2].record(); torch.cuda.synchronize()
        if r > 0:
            best_f = min(best_f, e[0].elapsed_time(e[1])); best_b = min(best_b, e[1].elapsed_time(e[2]))
    info = out["info"].cpu().tolist(); ns = info[_lib.INFO_TOTAL_NEWTON]
    print(f"timing {model} T={T} B={B} {str(dtype)[-7:]}: fwd {best_f:.3f} ms ({ns} newton steps -> {best_f/ns:.3f} ms/step) bwd {best_b:.3f} ms  => {B/(best_f+best_b)*1e3/1e6:.3f} M solves/s")

if __name__ == "__main__":
    for nm in ["quad_T5_f64_cold", "flying_T10_f64_cold", "pendulum_T5_f64_chain", "quad_T5_f32_cold"]:
        parity(nm)
    timing("rex_quadrotor", 5, 16384)
    timing("rex_quadrotor", 5, 65536)
    if len(sys.argv) > 1:
        timing("rex_quadrotor", 5, 65536, torch.float32)
        timing("flying_cartpole", 10, 16384)
        timing("cartpole1l", 5, 65536)
        timing("pendulum", 5, 65536)
