"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r'\(.*', '', r[ik]).replace('void <unnamed>::', '').replace('void rb::gemm::', '')
    v = float(r[iv].replace(',', ''))
    u = r[iu]
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e3:.1f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:9.2f} ms {100 * t / tot:5.1f}%  n={n:5d} avg={t / n:8.1f} us  {k}")
