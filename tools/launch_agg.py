"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_agg.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        try:
            agg[r[ki].split('(')[0][:40]].append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:42s} n={len(v):5d} total={sum(v)/1e6:9.3f} ms mean={sum(v)/len(v)/1e3:9.2f} us min={min(v)/1e3:.2f} max={max(v)/1e3:.2f}")
