import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import quick_perf
quick_perf.timing("rex_quadrotor", 5, 65536, reps=5)
