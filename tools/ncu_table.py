"""Per-launch table from an `ncu --page raw --csv` dump: duration, DRAM bytes / throughput, tensor pipe %, issue %, occupancy.
    ncu -i X.ncu-rep --page raw --csv > x.csv ; python tools/ncu_table.py x.csv [HBM_GBs]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hbm = float(sys.argv[2]) if len(sys.argv) > 2 else 6538.0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units = rows[hi], rows[hi + 1]
def col(r, k):
    return r[hdr.index(k)] if k in hdr else ""
def f(x):
    try: return float(x.replace(",", ""))
    except Exception: return float("nan")
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}
print(f"kernel | grid x block | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s (% of {hbm:.0f}) | tensor pipe % | issue active % | warps active % | regs")
for r in rows[hi + 2:]:
    if len(r) < len(hdr): continue
    name = re.sub(r"\(.*", "", col(r, "Kernel Name")).replace("<unnamed>::", "")
    t = f(col(r, "gpu__time_duration.sum")) * SC.get(units[hdr.index("gpu__time_duration.sum")], 1)
    rd = f(col(r, "dram__bytes_read.sum")) * SC.get(units[hdr.index("dram__bytes_read.sum")], 1) / 1e6
    wr = f(col(r, "dram__bytes_write.sum")) * SC.get(units[hdr.index("dram__bytes_write.sum")], 1) / 1e6
    gbs = (rd + wr) * 1e6 / (t * 1e-6) / 1e9 if t > 0 else 0
    print(f"{name} | {col(r,'Grid Size')} x {col(r,'Block Size')} | {t:.1f} | {rd:.2f} | {wr:.2f} | {gbs:.0f} ({100*gbs/hbm:.1f}%) | "
          f"{f(col(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):.1f} | {f(col(r,'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
          f"{f(col(r,'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | {col(r,'launch__registers_per_thread')}")
