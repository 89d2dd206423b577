"""Per-kernel CUDA-event times of one 64K quadrotor forward (desc.profile_ms): tools/kernel_times.py [model T B [f32]]."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from deq_mpc_corl_b200 import _lib
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost
from deq_mpc_corl_b200.dynamics import make as make_dyn
from deq_mpc_corl_b200.solver import SolveConfig, raw_forward

model = sys.argv[1] if len(sys.argv) > 1 else "rex_quadrotor"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
DT = torch.float32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else torch.float64
dev = "cuda:0"
spec = MODEL_SPECS[model]
inp = make_inputs(model, T, B, 1003)
Qd, q, _ = tracking_cost(inp["Qd"], inp["x_ref"], inp["u_ref"], DT)
dyn = make_dyn(model, dt=spec["dt"])
cfg = SolveConfig(dyn=dyn, T=T, u_lower=list(map(float, inp["u_lower"])), u_upper=list(map(float, inp["u_upper"])))
x0 = inp["x0"].to(DT).to(dev); Qd = Qd.to(dev).contiguous(); q = q.to(dev).contiguous()
xu0 = torch.cat([inp["x_ref"], inp["u_ref"]], -1).to(DT).to(dev).contiguous()
m = T * spec["nx"] + T * 2 * spec["nu"]
lib = _lib.load()
res = []
for it in range(4):
    xu = xu0.clone(); lam = torch.zeros(B, m, dtype=DT, device=dev); rho = torch.ones(B, dtype=DT, device=dev)
    cfg.profile = bool(it)
    out = raw_forward(cfg, x0, Qd, q, xu, lam, rho)
    torch.cuda.synchronize()
    if it:
        res.append(out["profile_ms"])
res.sort(key=lambda r: r[1])
r = res[len(res) // 2]
print(f"{model} T={T} B={B}: newton steps {out['info'][_lib.INFO_TOTAL_NEWTON].item()}  k_jacobian {r[0]/max(r[4],1):.3f} ms  "
      f"factor (or fused factor+line search) {r[1]/max(r[5],1):.3f} ms  k_line_search {r[2]/max(r[6],1):.3f} ms  (launch counts {r[4]:.0f}/{r[5]:.0f}/{r[6]:.0f})")
