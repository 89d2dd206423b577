#!/usr/bin/env python
"""One launcher for profiler runs (replaces prof_one.py / prof_one2.py / one_solve.py):

    python tools/prof_solve.py [--reps R] [--model M --T T --batch B --dtype f64|f32]   timed forward+backward solves
    python tools/prof_solve.py --golden quad_T5_f64_cold                                one golden scenario (parity print)

Typical captures (one GPU, after the command has exited 0 without ncu):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python tools/prof_solve.py --reps 1
    ncu --set full --clock-control none --import-source on -k regex:k_factor_solve -s 3 -c 1 -o gpurun_out/fac \\
        python tools/prof_solve.py --reps 1
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import quick_perf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--golden", default=None)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--model", default="rex_quadrotor")
ap.add_argument("--T", type=int, default=5)
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
if a.golden:
    quick_perf.parity(a.golden)
else:
    quick_perf.timing(a.model, a.T, a.batch, torch.float64 if a.dtype == "f64" else torch.float32, reps=a.reps)
torch.cuda.synchronize()
