"""Forward+backward of the 64K quadrotor batch: eager launches vs one CUDA-graph replay (device ms)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost
from deq_mpc_corl_b200.solver import SolveConfig, raw_backward, raw_forward

model, T, B, DEV = "rex_quadrotor", 5, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, "cuda:0"
spec = MODEL_SPECS[model]
m = T * spec["nx"] + 2 * T * spec["nu"]
inp = make_inputs(model, T, B, 31)
cfg = SolveConfig(dyn=make_dyn(model), T=T, u_lower=inp["u_lower"].tolist(), u_upper=inp["u_upper"].tolist(), al_iter=2, rho_f32=True)
Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"])
x0, Qd, q = inp["x0"].double().to(DEV), Qd.to(DEV), q.to(DEV)
xu0 = torch.cat([inp["x_ref"], inp["u_ref"]], -1).double().to(DEV).contiguous()
xu, lam, rho = xu0.clone(), torch.zeros(B, m, dtype=torch.float64, device=DEV), torch.ones(B, dtype=torch.float64, device=DEV)
g = torch.ones_like(xu)

def step():
    out = raw_forward(cfg, x0, Qd, q, xu, lam, rho)
    raw_backward(cfg, out["factor"], xu, g, need_Q=False)

def reset():
    xu.copy_(xu0); lam.zero_(); rho.fill_(1.0)

def timeit(fn, reps=10):
    best = 1e9
    for _ in range(reps):
        reset(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

reset(); step(); torch.cuda.synchronize()
e = timeit(step)
reset()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
gr = timeit(graph.replay)
print(f"B={B}: eager {e:.3f} ms, graph replay {gr:.3f} ms")
