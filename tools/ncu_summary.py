"""Summarise an `ncu --csv --page raw` capture (tools/prof_kernels.py under `ncu --set full`) into the markdown table committed
under profiles/ and the per-kernel DRAM traffic JSON that bench.py reports as `roofline.traffic`.
    python tools/ncu_summary.py gpurun_out/x/ncu_full.csv profiles/r1_ncu_full_cmdm_kernels_final.md profiles/r1_ncu_traffic.json"""
import csv, json, re, sys
src, md, js = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units = rows[hi], rows[hi + 1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "sm__cycles_elapsed.max"]
cols = [(w, hdr.index(w)) for w in want if w in hdr]
ni = hdr.index("Kernel Name")
def short(n):
    m = re.search(r"([A-Za-z_0-9]+_kernel(?:<[^(]*>)?)", n)
    return m.group(1) if m else n[:60]
def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None
out = ["# ncu --set full --clock-control none, tools/prof_kernels.py (one CMDM layer at B=32, S=326, d=512): per launch, cold cache, serialised",
       "kernel | " + " | ".join(f"{w} [{units[i]}]" for w, i in cols)]
traffic = {}
for r in rows[hi + 2:]:
    if len(r) <= ni:
        continue
    name = short(r[ni])
    out.append(name + " | " + " | ".join(r[i] for _, i in cols))
    rd, wr = num(r[hdr.index("dram__bytes_read.sum")]), num(r[hdr.index("dram__bytes_write.sum")])
    if rd is not None and wr is not None:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index("dram__bytes_read.sum")], 1)
        key = "linear_tc" if "gemm_tc" in name else ("mha_tc_fwd" if "mha_tc" in name else ("layernorm" if "layernorm" in name else name))
        traffic.setdefault(key, []).append((rd + wr) * scale)
open(md, "w").write("\n".join(out) + "\n")
json.dump({k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_sampled": len(v)} for k, v in traffic.items()}, open(js, "w"), indent=1)
print("\n".join(out))
