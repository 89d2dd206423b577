"""Host-side cost of one small MPC.forward + backward (pendulum, B=200): wall time per call and a cProfile listing."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from deq_mpc_corl_b200.al_mpc import MPC
from deq_mpc_corl_b200.al_utils import QuadCost
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost

model, T, B, dev = "pendulum", 5, 200, "cuda:0"
spec = MODEL_SPECS[model]
inp = make_inputs(model, T, B, 3)
inp = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() > 1 else v) for k, v in inp.items()}
dyn = make_dyn(model, dt=spec["dt"])
ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=inp["u_lower"], u_upper=inp["u_upper"], n_batch=B)
ctrl.cuda_graph = "--graph" in sys.argv
Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
Qd, q = Qd.to(dev), q.to(dev)
gx, gu = torch.randn(B, T, spec["nx"], device=dev), torch.randn(B, T, spec["nu"], device=dev)

def call():
    ctrl.reinitialize(inp["x_ref"], None)
    ctrl.x_init, ctrl.u_init = inp["x_ref"], inp["u_ref"]
    ql = q.detach().requires_grad_(True)
    x, u, _ = ctrl(inp["x0"], QuadCost(Qd, ql, None), dyn, dyn)
    torch.autograd.grad((x, u), ql, (gx, gu))

for _ in range(20): call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): call()
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / 200 * 1e3:.3f} ms per forward+backward call (wall, B={B})")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): call()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
