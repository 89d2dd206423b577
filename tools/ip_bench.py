#!/usr/bin/env python
"""Interior-point path timing (bench.py::ip_entry) for a few batch sizes, with a watchdog traceback.
    python tools/ip_bench.py [B ...]"""
import faulthandler
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
faulthandler.dump_traceback_later(90, repeat=True, file=sys.stderr)
import torch  # noqa: E402

import bench  # noqa: E402

for B in [int(a) for a in sys.argv[1:]] or [256, 1024, 4096]:
    t0 = time.perf_counter()
    e = bench.ip_entry(B, torch.device("cuda", 0), 64, "--no-cpu" in sys.argv)
    print(f"B={B}: {e['value']:.0f} QP solves/s ({e['ms']:.1f} ms fwd+bwd), cpu {e.get('cpu_baseline', {}).get('value')}, "
          f"wall {time.perf_counter() - t0:.1f} s", flush=True)
