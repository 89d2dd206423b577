#!/usr/bin/env python
"""Summarise an `ncu --metrics ... --csv` launch list: median per kernel of each metric."""
import csv, sys, statistics
from collections import OrderedDict, defaultdict
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=next(r for r in rows if 'Kernel Name' in r)
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d=OrderedDict()
for r in rows[rows.index(hdr)+1:]:
    
    try: val=float(r[vi].replace(',',''))
    except ValueError: val=float('nan')
    d.setdefault((r[ii], r[ki].split('(')[0][-60:]), {})[r[mi]]=val
agg=defaultdict(lambda: defaultdict(list))
for (i,k),m in d.items():
    if m.get('gpu__time_duration.sum',0) < 100000 and 'k_' in k: continue   # skip no-op launches
    for kk,v in m.items(): agg[k][kk].append(v)
for k,m in agg.items():
    print(f"{k:55s} n={len(m['gpu__time_duration.sum']):3d} " + ' '.join(f"{kk.split('__')[-1].replace('.sum','').replace('.avg','')[:24]}={statistics.median(v):.4g}" for kk,v in m.items()))
