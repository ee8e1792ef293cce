"""Print key raw metrics + hottest SASS lines of the first kernel in an .ncu-rep."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:3 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]:
    for k in ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
              'launch__block_size', 'launch__registers_per_thread', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
              'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
              'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
              'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']:
        if k in hdr:
            print(f"{k:72s} {r[hdr.index(k)][:60]} {units[hdr.index(k)]}")
    print()
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
seg = rows[starts[0] + 1: starts[1] if len(starts) > 1 else None]
h = seg[0]
ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
sc = [(i, x) for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
data = [r for r in seg[1:] if len(r) > isamp and r[isamp].isdigit()]
print('total samples', sum(int(r[isamp]) for r in data))
for r in sorted(data, key=lambda r: -int(r[isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    st = sorted([(int(r[i] or 0), x) for i, x in sc], reverse=True)[:2]
    print(f"{int(r[isamp]):6d} ex={r[iex]:>8s} {r[ia][:66]:66s} {st}")
