"""Time (CUDA events) + check the tcgen05 GEMM at the CMDM layer shapes for the variant in AMB200_TC_VARIANT."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200"))
import torch
from amb200 import ops
dev = "cuda:0"
M = 32 * 326
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for (N, K) in ((1536, 512), (512, 512), (1024, 512), (512, 1024)):
    x, w, b = r(M, K), r(N, K) / K ** 0.5, r(N)
    xs, wsp = ops.split_bf16(x, M, K), ops.split_bf16(w, N, K)
    y = torch.empty(M, N, device=dev)
    ops.linear_tc(xs, wsp, M, N, K, y=y, bias=b)
    ref = x.double() @ w.double().T + b.double()
    err = (y.double() - ref).abs().max().item()
    ts = []
    for _ in range(5):
        flush.fill_(0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.linear_tc(xs, wsp, M, N, K, y=y, bias=b); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts) // 2]
    print(f"variant={os.environ.get('AMB200_TC_VARIANT','32x3')} M={M} N={N} K={K}: {t*1e3:.1f} us  {2*M*N*K/t/1e9:.1f} TFLOP/s fp32-equivalent ({6*M*N*K/t/1e9:.0f} bf16-issued)  max_err={err:.2e}")
