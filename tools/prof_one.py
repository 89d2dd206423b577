"""One 64K quadrotor solve for ncu captures: `ncu -k regex:k_factor_solve -s 3 -c 1 ... python tools/prof_one.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import quick_perf, torch
quick_perf.timing("rex_quadrotor", 5, 65536, reps=1)
torch.cuda.synchronize()
