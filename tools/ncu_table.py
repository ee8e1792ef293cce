"""One line per profiled launch of an .ncu-rep (ncu --set full): duration, DRAM bytes, achieved DRAM GB/s, pipe / memory
throughput percentages.  usage: python tools/ncu_table.py report.ncu-rep"""
import csv
import io
import re
import subprocess
import sys

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def col(r, k, default=''):
    return r[hdr.index(k)] if k in hdr else default


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return float('nan')


def to_bytes(v, u):
    m = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    return num(v) * m.get(u, 1)


def to_us(v, u):
    m = {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6}
    return num(v) * m.get(u, 1)


print(f"{'kernel':44s} {'grid':>6s} {'regs':>4s} {'us':>9s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'dram%':>6s} {'lts%':>6s} "
      f"{'l1%':>6s} {'tensor%':>7s} {'warps%':>6s}")
for r in rows[2:]:
    name = re.sub(r'\(.*', '', col(r, 'Kernel Name')).replace('void ', '').replace('<unnamed>::', '')[:44]
    t = to_us(col(r, 'gpu__time_duration.sum'), units[hdr.index('gpu__time_duration.sum')])
    rd = to_bytes(col(r, 'dram__bytes_read.sum'), units[hdr.index('dram__bytes_read.sum')])
    wr = to_bytes(col(r, 'dram__bytes_write.sum'), units[hdr.index('dram__bytes_write.sum')])
    print(f"{name:44s} {col(r, 'launch__grid_size'):>6s} {col(r, 'launch__registers_per_thread'):>4s} {t:9.1f} {rd / 1e6:8.1f} "
          f"{wr / 1e6:8.1f} {(rd + wr) / t / 1e3:7.0f} {num(col(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
          f"{num(col(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
          f"{num(col(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
          f"{num(col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):7.1f} "
          f"{num(col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')):6.1f}")
