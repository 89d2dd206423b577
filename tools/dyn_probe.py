import sys; sys.path.insert(0,'/root/repo')
import torch
from deq_mpc_corl_b200.dynamics import make
dev='cuda:0'
dyn = make('rex_quadrotor')
for R,jac in [(65536*4, True), (65536*80, False)]:
    x = (torch.rand(R,12,dtype=torch.float64,device=dev)-0.5); u = 11.5+6*torch.rand(R,4,dtype=torch.float64,device=dev)
    for it in range(3):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        if jac: dyn.jacobian(x,u)
        else: dyn(x,u)
        e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)
    flops = R*(16*3*980 if jac else 980)
    print(f"R={R} jac={jac}: {ms:.3f} ms  -> {flops/ms/1e9:.2f} TFLOP/s (dual counted as 3x primal)")
