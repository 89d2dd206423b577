"""Accuracy of the factor kernels along a 6-iteration DEQ chain (rho 1 -> 1e12, no re-synchronisation) against the
CPU oracle: relative error of the iterate per iteration, `=`/`!` = Newton-step counts equal / different.
python tools/chain_accuracy.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from deq_mpc_corl_b200.al_mpc import MPC
from deq_mpc_corl_b200.al_utils import QuadCost
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost
from golden_util import rel_err
from oracle import almpc_oracle as orc
from oracle.problem_gen import make_oracle_model
DEV = "cuda:0"
def run(model, T, B, env):
    for k in ("ALMPC_FAC_PAIR", "ALMPC_FAC_MMA", "ALMPC_FAC_SINGLE"): os.environ.pop(k, None)
    if env: os.environ[env] = "1"
    spec = MODEL_SPECS[model]; inp = make_inputs(model, T, B, 1000); dyn = make_dyn(model)
    ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
    ompc = orc.OracleMPC(make_oracle_model(model), T, inp["u_lower"], inp["u_upper"], B)
    xr_c, xr_o = inp["x_ref"].to(DEV), inp["x_ref"].clone()
    ctrl.reinitialize(xr_c, None); ompc.reinitialize(xr_o)
    ctrl.x_init, ctrl.u_init = xr_c.clone(), inp["u_ref"].to(DEV)
    ompc.x_init, ompc.u_init = xr_o.clone(), inp["u_ref"].clone()
    out = []
    for it in range(6):
        Qd, q_c, _ = tracking_cost(inp["Qd"].to(DEV), xr_c, inp["u_ref"].to(DEV))
        _, q_o, _ = tracking_cost(inp["Qd"], xr_o, inp["u_ref"])
        x_c, u_c, _ = ctrl(inp["x0"].to(DEV), QuadCost(Qd, q_c, None), dyn, dyn)
        tr = orc.SolveTrace()
        x_o, u_o, _, xu_o = ompc.solve(inp["x0"], Qd.cpu(), q_o, trace=tr)
        out.append((ctrl.newton_steps() == tr.newton_steps, rel_err(ctrl.last_xu.cpu(), xu_o)))
        xr_c = 0.7 * xr_c + 0.3 * x_c.detach(); xr_c[:, 0] = inp["x0"].to(DEV)
        xr_o = 0.7 * xr_o + 0.3 * x_o; xr_o[:, 0] = inp["x0"]
    return out
for model, T, B in (("pendulum", 5, 200), ("rex_quadrotor", 5, 64), ("cartpole1l", 5, 100)):
    for env in ("ALMPC_FAC_MMA", "ALMPC_FAC_PAIR"):
        r = run(model, T, B, env)
        print(model, env, " ".join(f"{'=' if a else '!'}{e:.1e}" for a, e in r), flush=True)
