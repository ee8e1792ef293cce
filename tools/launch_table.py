"""Per-kernel table (avg us, DRAM read/write MB) from an ncu csv with gpu__time_duration + dram bytes metrics."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ik, im, iv, iu, iid = (hdr.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
per = {}
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    k = (int(r[iid]), re.sub(r'\(.*', '', r[ik]).replace('void <unnamed>::', '').replace('void rb::', ''))
    v = float(r[iv].replace(',', '')); u = r[iu]
    if r[im].startswith('gpu__time'):
        per.setdefault(k, {})['us'] = v / 1e3 if u in ('ns', 'nsecond') else v * 1e3 if u in ('ms', 'msecond') else v
    else:
        per.setdefault(k, {})[r[im][:16]] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (i, name), d in sorted(per.items()):
    if i < skip:
        continue
    a = agg[name]; a[0] += 1; a[1] += d.get('us', 0); a[2] += d.get('dram__bytes_read', 0); a[3] += d.get('dram__bytes_writ', 0)
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e3:.2f} ms")
for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1e3:8.3f} ms {100*t/tot:5.1f}%  n={n:3d} avg {t/n:8.1f} us  rd {rd/n/1e6:8.1f} MB wr {wr/n/1e6:8.1f} MB  {(rd+wr)/max(t,1e-9)/1e3:6.2f} TB/s  {k}")
