import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import quick_perf, torch
quick_perf.parity(sys.argv[1] if len(sys.argv) > 1 else "quad_T5_f64_cold")
torch.cuda.synchronize()
