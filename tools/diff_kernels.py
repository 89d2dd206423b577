"""Differential stress test: the default factor kernel (fragment layout, k_factor_solve_mma, where it applies) against
the row-per-lane kernel (ALMPC_FAC_SINGLE) — or, with a second argument `pair`, against the half-warp-per-problem
kernel (ALMPC_FAC_PAIR) — on random (model, T, B, variant) draws, two chained calls each (the second runs the fused
dual update): same Newton-step counts and status, iterates / multipliers / gradients equal to rounding.
python tools/diff_kernels.py [n_draws [pair|single] [mma]]   (mma: force the fragment-layout kernel everywhere)"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from deq_mpc_corl_b200.al_mpc import MPC
from deq_mpc_corl_b200.al_utils import QuadCost
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost

dev = "cuda:0"


def run(model, T, B, seed, se, single):
    for k in ("ALMPC_FAC_SINGLE", "ALMPC_FAC_PAIR", "ALMPC_FAC_MMA"):
        os.environ.pop(k, None)
    if single:
        os.environ[OTHER] = "1"
    elif FORCE_MMA:  # the fragment-layout kernel also where it is not the default (one-tile models)
        os.environ["ALMPC_FAC_MMA"] = "1"
    spec = MODEL_SPECS[model]
    inp = make_inputs(model, T, B, seed)
    dyn = make_dyn(model, dt=spec["dt"])
    ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B, state_estimator=se)
    ctrl.reinitialize(inp["x_ref"].to(dev))
    ctrl.x_init, ctrl.u_init = inp["x_ref"].to(dev), inp["u_ref"].to(dev)
    out = []
    x_ref = inp["x_ref"]
    for call in range(2):
        Qd, q, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"])
        ql = q.to(dev).requires_grad_(True)
        x, u, _ = ctrl(inp["x0"].to(dev), QuadCost(Qd.to(dev), ql, None), dyn, dyn)
        g, = torch.autograd.grad((x, u), ql, (torch.ones_like(x), torch.ones_like(u)))
        out.append((ctrl.last_xu.clone(), ctrl.lamda_prev.clone(), g.clone(), ctrl.newton_steps(), ctrl.last_status.clone()))
        x_ref = 0.7 * x_ref + 0.3 * x.detach().cpu()
        x_ref[:, 0] = inp["x0"]
    return out


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
OTHER = "ALMPC_FAC_PAIR" if len(sys.argv) > 2 and sys.argv[2] == "pair" else "ALMPC_FAC_SINGLE"
FORCE_MMA = "mma" in sys.argv[2:]
rng = random.Random(2024)
worst = 0.0
for i in range(n):
    model = rng.choice(["pendulum", "cartpole1l", "cartpole2l", "rex_quadrotor", "flying_cartpole"])
    T = rng.choice([2, 3, 4, 5, 6, 7, 9, 12, 17, 24])
    B = rng.choice([1, 2, 3, 5, 31, 64, 127, 500, 1025])
    se = rng.random() < 0.25
    seed = rng.randrange(10 ** 6)
    a, b = run(model, T, B, seed, se, False), run(model, T, B, seed, se, True)
    for ci, (pa, pb) in enumerate(zip(a, b)):
        assert pa[3] == pb[3], (model, T, B, se, ci, pa[3], pb[3])
        assert torch.equal(pa[4], pb[4])
        e = max(rel(pa[0], pb[0]), rel(pa[1], pb[1]) if pb[1].numel() else 0.0, rel(pa[2], pb[2]))
        worst = max(worst, e)
        assert e < 1e-6, (model, T, B, se, ci, e)
    print(f"{i:3d} {model:14s} T={T:2d} B={B:5d} se={int(se)} steps {a[0][3]} {a[1][3]} ok")
print(f"all {n} draws agree; worst relative difference {worst:.2e}")
