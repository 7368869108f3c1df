"""Tile-shape / cluster A/B of the tcgen05 GEMM on the four CMDM trunk shapes (B=32, S=326 -> M=10432), each with the epilogue
the engine uses.  CUDA-event time per launch, 20 back-to-back launches after warm-up (operands L2-resident, as in the step graph)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200"))
import torch
from amb200 import lib, ops
L = lib.load()
L.am_tc_set_tile_.argtypes = [ctypes.c_int, ctypes.c_int]
L.am_tc_set_epi_.argtypes = [ctypes.c_int]
L.am_tc_set_2sm_.argtypes = [ctypes.c_int]
L.am_tc_set_bk_.argtypes = [ctypes.c_int]
L.am_tc_set_mixed_.argtypes = [ctypes.c_int]
dev = "cuda:0"
M, D, FF = 32 * 326, 512, 1024
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
x, xf = r(M, D), r(M, FF)
W = {"qkv": r(3 * D, D) / 22, "out": r(D, D) / 22, "ffn1": r(FF, D) / 22, "ffn2": r(D, FF) / 32}
xs, xfs = ops.split_bf16(x, M, D), ops.split_bf16(xf, M, FF)
Ws = {k: ops.split_bf16(v, v.shape[0], v.shape[1]) for k, v in W.items()}
bias = {k: r(v.shape[0]) for k, v in W.items()}
y32 = torch.empty(M, D, device=dev)
y2q = torch.zeros(M, 6 * D, dtype=torch.bfloat16, device=dev)
y2f = torch.zeros(M, 2 * FF, dtype=torch.bfloat16, device=dev)
cases = {
    "qkv  N=1536 K=512  -> bf16 split": lambda: ops.linear_tc(xs, Ws["qkv"], M, 3 * D, D, y2=y2q, bias=bias["qkv"], Np2=3 * D),
    "out  N=512  K=512  -> fp32 + res": lambda: ops.linear_tc(xs, Ws["out"], M, D, D, y=y32, bias=bias["out"], residual=x),
    "ffn1 N=1024 K=512  -> gelu split": lambda: ops.linear_tc(xs, Ws["ffn1"], M, FF, D, y2=y2f, bias=bias["ffn1"], act="gelu", Np2=FF),
    "ffn2 N=512  K=1024 -> fp32 + res": lambda: ops.linear_tc(xfs, Ws["ffn2"], M, D, FF, y=y32, bias=bias["ffn2"], residual=x),
}
flops = {"qkv": 2 * M * 3 * D * D, "out": 2 * M * D * D, "ffn1": 2 * M * FF * D, "ffn2": 2 * M * D * FF}
ref_out = {}
for name, fn in cases.items():
    row = []
    for bn, cl, lsu, sm2, bk, mixed in ((0, 0, 0, 0, 32, 0), (0, 0, 0, 1, 32, 0), (0, 0, 0, 1, 64, 0), (128, 0, 0, 1, 64, 0), (256, 0, 0, 1, 64, 0), (256, 0, 0, 1, 64, 1), (0, 0, 0, 1, 64, 1)):
        L.am_tc_set_tile_(bn, cl)
        L.am_tc_set_mixed_(mixed)
        L.am_tc_set_bk_(bk)
        L.am_tc_set_epi_(lsu)
        L.am_tc_set_2sm_(sm2)
        for _ in range(5):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            fn()
        e.record(); torch.cuda.synchronize()
        us = 1e3 * s.elapsed_time(e) / 20
        out = (y32 if "fp32" in name else (y2q if "qkv" in name else y2f)).float().clone()
        key = name.split()[0]
        if key not in ref_out:
            ref_out[key] = out
        same = bool(torch.equal(out, ref_out[key]))  # every variant must give bit-identical results (same accumulation order)
        row.append(f"{'2SM' if sm2 else '1SM'} bk={bk} bn={bn or 'auto'} {'mixed' if mixed else 'whole'} {'lsu' if lsu else 'tma'}: {us:6.1f} us ({flops[key] / us / 1e6:4.0f} TF/s){'' if same else ' MISMATCH'}")
    print(name + "\n   " + "\n   ".join(row))
L.am_tc_set_tile_(0, -1)
L.am_tc_set_epi_(-1)
L.am_tc_set_2sm_(-1)
L.am_tc_set_bk_(0)
L.am_tc_set_mixed_(1)
