"""Summarise an ncu launch list (gpu__time_duration.sum per launch): per-kernel shares and the conv launches in order."""
import csv, sys
from collections import defaultdict
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = [(x['Kernel Name'], x['Grid Size'], float(x['Metric Value'].replace(',', ''))) for x in csv.DictReader(lines)]
tot = sum(t for _, _, t in rows)
print(f"{len(rows)} launches, {tot/1e6:.3f} ms total (cold-cache, serialised)")
agg = defaultdict(lambda: [0, 0])
for n, g, t in rows:
    k = n.split('(')[0][:48]; agg[k][0] += t; agg[k][1] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  {k:48s} {v[1]:4d} launches {v[0]/1e6:9.3f} ms {100*v[0]/tot:5.1f}%")
if len(sys.argv) > 2:
    i = 0
    for n, g, t in rows:
        if 'conv_' in n and 'post' not in n:
            print(i, n.split("(")[0][-14:], g, f"{t/1e3:9.1f} us"); i += 1
