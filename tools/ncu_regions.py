#!/usr/bin/env python
"""Per-region table (stall samples, instructions, smem wavefronts, global sectors) of one kernel from an ncu
source page joined with nvdisasm -g line info.  usage: ncu_regions.py src.csv all.sass kernel_sub nproblems name:lo-hi ..."""
import csv, sys
from collections import defaultdict
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from ncu_lines import sass_lines
src, sass, ksub, nprob = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
regions = [(a.split(":")[0], *map(int, a.split(":")[1].split("-"))) for a in sys.argv[5:]]
rows = list(csv.reader(open(src)))
hdr = next(r for r in rows if 'Address' in r and '# Samples' in r)
i0 = rows.index(hdr)
col = {n: hdr.index(n) for n in ['# Samples', 'Instructions Executed', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal', 'L2 Theoretical Sectors Global']}
inst = [r for r in rows[i0 + 1:] if len(r) > max(col.values()) and r[0]]
lines = sass_lines(sass, ksub)
f = lambda x: int(float(x)) if x not in ('', 'n/a') else 0
agg = defaultdict(lambda: [0] * 5)
for k in range(min(len(inst), len(lines))):
    key = lines[k][1]
    name = 'other'
    if key and key[0].startswith('almpc_'):
        for nm, lo, hi in regions:
            if lo <= key[1] <= hi:
                name = nm
                break
    elif key:
        name = key[0]
    for q, n in enumerate(col):
        agg[name][q] += f(inst[k][col[n]])
tot = [sum(v[q] for v in agg.values()) for q in range(5)]
print(f"# rows ncu {len(inst)} sass {len(lines)}; per problem: instr {tot[1]/nprob:.0f} smem wavefronts {tot[2]/nprob:.0f} (ideal {tot[3]/nprob:.0f}) global sectors {tot[4]/nprob:.0f}")
print(f"{'region':24s} {'stall%':>7s} {'instr':>7s} {'smem wf':>8s} {'ideal':>7s} {'gl.sect':>8s}   (per problem)")
for nm, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{nm:24s} {100*v[0]/tot[0]:7.1f} {v[1]/nprob:7.0f} {v[2]/nprob:8.0f} {v[3]/nprob:7.0f} {v[4]/nprob:8.0f}")
