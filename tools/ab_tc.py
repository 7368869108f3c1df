"""A/B timing of the tcgen05 kernels of one CMDM layer (B=32, S=326, d=512) with a given library build:
    AMB200_LIB=/path/to/libamb200.so python tools/ab_tc.py [reps]
CUDA events per kernel, L2-warm back-to-back launches after warm-up; prints mean us per launch.  Used to separate code changes from
box-to-box differences (power cap) when a bench number moves between rounds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200"))
import torch
from amb200 import lib, ops
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
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
ks = {
    "in_proj  [10432x1536x512]": lambda: ops.linear_tc(xs, wins, M, 3 * D, D, y2=qkvs, bias=bias3, Np2=3 * D),
    "mha_tc   [32x8, S=326]   ": lambda: ops.mha_tc_fwd(qkvs, None, atts, pad, B, S, H, 64, 0.125),
    "out_proj [10432x512x512] ": lambda: ops.linear_tc(atts, wouts, M, D, D, y=tmp, bias=bias1, residual_split=xs),
    "layernorm                ": lambda: ops.layernorm(tmp, gam, bet, None, M, D, y2=y1s),
    "ffn1     [10432x1024x512]": lambda: ops.linear_tc(y1s, w1s, M, FF, D, y2=ffs, bias=biasf, act="gelu", Np2=FF),
    "ffn2     [10432x512x1024]": lambda: ops.linear_tc(ffs, w2s, M, D, FF, y=tmp, bias=bias1, residual_split=y1s),
}
if hasattr(lib.load(), "am_linear_ln_tc"):
    ks["out_proj+LN fused        "] = lambda: ops.linear_ln_tc(atts, wouts, M, D, D, bias1, xs, gam, bet, 1e-5, y1s)
    ks["ffn2+LN fused            "] = lambda: ops.linear_ln_tc(ffs, w2s, M, D, FF, bias1, y1s, gam, bet, 1e-5, atts)
print("library:", lib.LIB_PATH)
for _ in range(3):
    for f in ks.values():
        f()
tot = 0.0
for name, f in ks.items():
    for _ in range(5):
        f()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        f()
    e.record()
    torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / reps
    tot += 0.0 if "fused" in name else us * (2 if "layernorm" in name else 1)
    print(f"  {name} {us:8.2f} us")
print(f"  layer total (2 LN) {tot:8.2f} us -> x5 layers {5 * tot / 1e3:.3f} ms")
