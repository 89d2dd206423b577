#!/usr/bin/env python
"""Turns an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
ONE forward+backward solve (tools/prof_solve.py --reps 1) into the per-kernel DRAM traffic table bench.py reads
(profiles/r02_ncu_traffic.json, keyed by workload):

    python tools/ncu_traffic.py gpurun_out/launches.csv rex_quadrotor/T5/B65536/f64 [--merge profiles/r02_ncu_traffic.json]

Per kernel: median DRAM bytes per executed launch (no-op launches after the solve finished are dropped) and the
launch count; `solve_dram_bytes_per_problem` = all launches of the solve summed / problems.
"""
import csv
import json
import statistics
import sys
from collections import OrderedDict

path, key = sys.argv[1], sys.argv[2]
merge = sys.argv[sys.argv.index("--merge") + 1] if "--merge" in sys.argv else None
B = int(key.split("/")[2][1:])
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
ui = hdr.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
launches = OrderedDict()
for r in rows[rows.index(hdr) + 1:]:
    try:
        val = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    except ValueError:
        continue
    name = r[ki].split("(")[0].split("<")[0].split("::")[-1].strip()
    launches.setdefault((r[ii], name), {})[r[mi]] = val
per = OrderedDict()
total = 0.0
for (_, name), m in launches.items():
    if not name.startswith("k_"):
        continue
    byts = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    if m.get("gpu__time_duration.sum", 0.0) < 2e4:  # < 20 us at 64K problems: a launch that found MODE_DONE
        continue
    per.setdefault(name, []).append((byts, m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0), m.get("gpu__time_duration.sum", 0.0)))
    total += byts
alias = {"k_factor_solve_pair": "k_factor_solve", "k_factor_solve_mma": "k_factor_solve", "k_line_search_warp": "k_line_search"}
out = {"source": path, "solve_dram_bytes_per_problem": total / B, "solve_dram_bytes": total}
for name, v in per.items():
    out[alias.get(name, name)] = statistics.median(x[0] for x in v)
    out[alias.get(name, name) + "_detail"] = {"kernel": name, "launches": len(v), "read_median": statistics.median(x[1] for x in v),
                                             "write_median": statistics.median(x[2] for x in v),
                                             "ncu_duration_us_median": statistics.median(x[3] for x in v) / 1e3,
                                             "per_launch_bytes": [x[0] for x in v]}
print(json.dumps({key: out}, indent=1))
if merge:
    try:
        tab = json.load(open(merge))
    except Exception:
        tab = {}
    tab[key] = out
    json.dump(tab, open(merge, "w"), indent=1)
