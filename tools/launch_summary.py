"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / mean / min / max (us)."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) == len(hdr):
        d[r[kn][:70]].append(float(r[mv].replace(',', '')) / 1e3)
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print(f"{k:70s} n={len(v):4d} mean={sum(v)/len(v):9.2f} us min={min(v):9.2f} max={max(v):9.2f} share={sum(v)/tot:6.1%}")
