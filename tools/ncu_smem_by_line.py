#!/usr/bin/env python
"""Per-source-line shared-memory wavefronts / stall samples / instructions of one kernel, per problem:
ncu source page (ncu -i X.ncu-rep --page source --csv) joined with `nvdisasm -g` line info of the same cubin.
    tools/ncu_smem_by_line.py src.csv all.sass <mangled-kernel-substring> <problems> [top_n]"""
import csv
import sys
from collections import defaultdict

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from ncu_lines import sass_lines  # noqa: E402

src, sass, ksub, nprob = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 30
rows = list(csv.reader(open(src)))
hdr = next(r for r in rows if "Address" in r and "# Samples" in r)
i0 = rows.index(hdr)
names = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"]
col = {n: hdr.index(n) for n in names}
inst = [r for r in rows[i0 + 1:] if len(r) > max(col.values()) and r[0]]
lines = sass_lines(sass, ksub)
f = lambda x: int(float(x)) if x not in ("", "n/a") else 0
agg = defaultdict(lambda: [0, 0, 0, 0])
for k in range(min(len(inst), len(lines))):
    key = lines[k][1][:2] if lines[k][1] else ("?", 0)
    for q, n in enumerate(names):
        agg[key][q] += f(inst[k][col[n]])
tot = [sum(v[q] for v in agg.values()) for q in range(4)]
print(f"# sass rows: ncu {len(inst)} nvdisasm {len(lines)}; per problem: {tot[1] / nprob:.0f} warp instructions, "
      f"{tot[2] / nprob:.0f} shared wavefronts (ideal {tot[3] / nprob:.0f})")
srcs = {}


def text(key):
    fn, ln = key
    if fn not in srcs:
        try:
            srcs[fn] = open("deq_mpc_corl_b200/csrc/" + fn).read().splitlines()
        except OSError:
            srcs[fn] = []
    return srcs[fn][ln - 1].strip()[:100] if 0 < ln <= len(srcs[fn]) else ""


print(" smem wf   ideal  stall%   instr  file:line  source   (per problem)")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:topn]:
    print(f"{v[2] / nprob:8.1f} {v[3] / nprob:7.1f} {100.0 * v[0] / max(tot[0], 1):7.1f} {v[1] / nprob:7.1f}  {key[0]}:{key[1]}  {text(key)}")
print("# top stall lines")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn // 2]:
    print(f"{v[2] / nprob:8.1f} {v[3] / nprob:7.1f} {100.0 * v[0] / max(tot[0], 1):7.1f} {v[1] / nprob:7.1f}  {key[0]}:{key[1]}  {text(key)}")
