"""Per-role clock64 timeline of CTA (0,0) of the tcgen05 GEMM (debug hook am_tc_set_debug_)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200"))
import torch
from amb200 import ops, lib
dev = "cuda:0"
M = 32 * 326
N, K = int(os.environ.get("TC_N", "1536")), int(os.environ.get("TC_K", "512"))
g = torch.Generator(device=dev).manual_seed(0)
x, w = torch.randn(M, K, device=dev, generator=g), torch.randn(N, K, device=dev, generator=g) / 22
xs, wsp = ops.split_bf16(x, M, K), ops.split_bf16(w, N, K)
y = torch.empty(M, N, device=dev)
L = lib.load()
L.am_tc_set_debug_.argtypes = [ctypes.c_void_p]
case = os.environ.get("TC_CASE", "plain")
bias = torch.randn(N, device=dev, generator=g)
res = torch.randn(M, N, device=dev, generator=g)
y2 = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=dev)
def run():
    if case == "qkv":      # bias -> bf16 split only (fast mode 1)
        ops.linear_tc(xs, wsp, M, N, K, y2=y2, bias=bias, Np2=N)
    elif case == "out":    # bias + residual -> fp32 (fast mode 2)
        ops.linear_tc(xs, wsp, M, N, K, y=y, bias=bias, residual=res)
    else:
        ops.linear_tc(xs, wsp, M, N, K, y=y)
for _ in range(3):
    run()
dbg = torch.zeros(256, dtype=torch.int64, device=dev)
dbg[255] = int(os.environ.get("TC_DFLAGS", "0"))
L.am_tc_set_debug_(dbg.data_ptr())
run()
torch.cuda.synchronize()
L.am_tc_set_debug_(None)
d = dbg.cpu().tolist()
t0 = d[0]
if os.environ.get("AMB200_TC_VARIANT", "persistent") == "persistent":
    print("tile  mma_committed  epi_start  epi_end   (cycles since CTA 0 start; the CTA-pair kernel records commit and epi_end only)")
    last = 0
    for it in range(6):
        print(it, *[(d[o + it] - t0 if d[o + it] else None) for o in (8, 40, 72)])
        last = d[72 + it] - t0 if d[72 + it] else last
    if d[1] and d[2]:
        print(f"CTA 0: {last} cycles in {d[2] - d[1]} ns -> SM clock {1e3 * last / (d[2] - d[1]):.0f} MHz during this launch")
    reps = 50
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.am_tc_set_debug_(None)
    for _ in range(5):
        run()
    s.record()
    for _ in range(reps):
        run()
    e.record(); torch.cuda.synchronize()
    print(f"back-to-back launch period: {1e3 * s.elapsed_time(e) / reps:.1f} us")
    sys.exit(0)
nkb = K // (64 if os.environ.get("AMB200_TC_VARIANT") == "64x3" else 32)
print("prologue sync done +", d[1] - t0, " tmem_full seen +", d[2] - t0, " epilogue done +", d[3] - t0, " teardown +", d[4] - t0)
print("kb  tma_issue  full_seen  mma_issued(commit)")
for kb in range(nkb):
    print(kb, d[8 + kb] - t0, d[72 + kb] - t0, d[136 + kb] - t0)
