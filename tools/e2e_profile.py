"""GPU-side op list of one end-to-end step through the module API (MPC.reinitialize + forward + autograd backward) at
the headline size: what the drop-in adds around the library's kernels.  python tools/e2e_profile.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from deq_mpc_corl_b200.al_mpc import MPC
from deq_mpc_corl_b200.al_utils import QuadCost
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost

model, T, B, dev = "rex_quadrotor", 5, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, "cuda:0"
spec = MODEL_SPECS[model]
inp = make_inputs(model, T, B, 1003)
Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
dyn = make_dyn(model, dt=spec["dt"])
d = {k: v.to(dev) for k, v in dict(x0=inp["x0"].double(), Qd=Qd, q=q, x_ref=inp["x_ref"], u_ref=inp["u_ref"]).items()}
d["gx"] = torch.randn(B, T, spec["nx"], device=dev); d["gu"] = torch.randn(B, T, spec["nu"], device=dev)
ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)

def step():
    ctrl.reinitialize(d["x_ref"], None)
    ctrl.x_init, ctrl.u_init = d["x_ref"], d["u_ref"]
    ql = d["q"].detach().requires_grad_(True)
    x, u, _ = ctrl(d["x0"], QuadCost(d["Qd"], ql, None), dyn, dyn)
    qg, = torch.autograd.grad((x, u), ql, (d["gx"], d["gu"]))
    return qg

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
