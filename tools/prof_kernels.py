"""Micro driver for ncu: runs the CMDM hot kernels at BASELINE config-2 shapes (B=32, S=326, d=512) a few times.
Usage (under gpurun):  ncu --set full --clock-control none --import-source on -k regex:<pat> -s <skip> -c <n> -o out python tools/prof_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200"))
import torch
from amb200 import ops

dev = "cuda:0"
B, S, D, FF, H = 32, 326, 512, 1024, 8
M = B * S
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
x, win, wout, w1, w2 = r(M, D), r(3 * D, D) / 22, r(D, D) / 22, r(FF, D) / 22, r(D, FF) / 32
xs, wins, wouts, w1s, w2s = (ops.split_bf16(t, t.shape[0], t.shape[1]) for t in (x, win, wout, w1, w2))
qkvs = torch.zeros(M, 6 * D, dtype=torch.bfloat16, device=dev)
tmp, y1 = torch.empty(M, D, device=dev), torch.empty(M, D, device=dev)
ffs = torch.zeros(M, 2 * FF, dtype=torch.bfloat16, device=dev)
atts, y1s = torch.zeros(M, 2 * D, dtype=torch.bfloat16, device=dev), torch.zeros(M, 2 * D, dtype=torch.bfloat16, device=dev)
bias3, bias1, biasf = r(3 * D), r(D), r(FF)
gam, bet = r(D), r(D)
pad = torch.zeros(B, S, dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):   # exactly the per-layer sequence of amb200.cmdm_engine._forward_tc
    ops.linear_tc(xs, wins, M, 3 * D, D, y2=qkvs, bias=bias3, Np2=3 * D)                 # in_proj -> bf16 (hi|lo) QKV
    ops.mha_tc_fwd(qkvs, None, atts, pad, B, S, H, 64, 0.125)                           # tcgen05 attention
    ops.linear_tc(atts, wouts, M, D, D, y=tmp, bias=bias1, residual_split=xs)           # out_proj + residual (bf16 hi|lo stream)
    ops.layernorm(tmp, gam, bet, None, M, D, y2=y1s)                                    # LN1 -> bf16 (hi|lo) only
    ops.linear_tc(y1s, w1s, M, FF, D, y2=ffs, bias=biasf, act="gelu", Np2=FF)           # FFN1 + GELU (split out only)
    ops.linear_tc(ffs, w2s, M, D, FF, y=tmp, bias=bias1, residual_split=y1s)            # FFN2 + residual
    ops.layernorm(tmp, gam, bet, None, M, D, y2=xs)                                     # LN2
torch.cuda.synchronize()
print("done")
