#!/usr/bin/env python
"""Join an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line
info of the same cubin and aggregate warp-stall samples / executed instructions per source line and
per kernel phase.  Usage: tools/ncu_lines.py src.csv all.sass <mangled-kernel-substring> [top_n]"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, kernel_sub):
    out, cur, active = [], None, False
    for ln in open(path, errors="ignore"):
        if ln.startswith(".text."):
            active = kernel_sub in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    src_csv, sass, ksub = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = next(r for r in rows if "Address" in r and "# Samples" in r)
    i0 = rows.index(hdr)
    ai, si, ii, so = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    inst = [r for r in rows[i0 + 1:] if len(r) > si and r[ai]]
    lines = sass_lines(sass, ksub)
    n = min(len(inst), len(lines))
    print(f"# sass rows: ncu {len(inst)} nvdisasm {len(lines)}")
    per_line = defaultdict(lambda: [0, 0])
    tot = 0
    for k in range(n):
        s = int(inst[k][si] or 0)
        e = int(inst[k][ii] or 0)
        key = lines[k][1][:2] if lines[k][1] else ("?", 0)
        per_line[key][0] += s
        per_line[key][1] += e
        tot += s
    print(f"# total samples {tot}")
    for key, (s, e) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:topn]:
        print(f"{100.0 * s / tot:6.2f}%  samples {s:8d}  inst {e:10d}  {key[0]}:{key[1]}")
    # phases by line ranges of almpc_kernels.cu (argument 5.. as name:lo-hi)
    if len(sys.argv) > 5:
        ph = [(a.split(":")[0], *map(int, a.split(":")[1].split("-"))) for a in sys.argv[5:]]
        agg = defaultdict(lambda: [0, 0])
        for key, (s, e) in per_line.items():
            name = "other/" + key[0]
            if key[0].startswith("almpc_"):
                for nm, lo, hi in ph:
                    if lo <= key[1] <= hi:
                        name = nm
                        break
            agg[name][0] += s
            agg[name][1] += e
        print("# phases")
        for nm, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"{100.0 * s / tot:6.2f}%  inst {e:10d}  {nm}")


if __name__ == "__main__":
    main()
