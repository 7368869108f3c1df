"""clock64 timeline of CTA 0 of the tcgen05 attention kernel (debug hook am_att_set_debug_)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200"))
import torch
from amb200 import ops, lib
dev = "cuda:0"
B, S, H = 32, 326, 8
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B * S, 3 * H * 64, device=dev, generator=g)
qkv2 = ops.split_bf16(qkv, B * S, 3 * H * 64)
out2 = torch.zeros(B * S, 2 * H * 64, dtype=torch.bfloat16, device=dev)
pad = torch.zeros(B, S, dtype=torch.uint8, device=dev)
L = lib.load(); L.am_att_set_debug_.argtypes = [ctypes.c_void_p]
for _ in range(3):
    ops.mha_tc_fwd(qkv2, None, out2, pad, B, S, H, 64, 0.125)
dbg = torch.zeros(256, dtype=torch.int64, device=dev)
L.am_att_set_debug_(dbg.data_ptr())
ops.mha_tc_fwd(qkv2, None, out2, pad, B, S, H, 64, 0.125)
torch.cuda.synchronize(); L.am_att_set_debug_(None)
d = dbg.cpu().tolist(); t0 = d[0]
print("kv_full seen by MMA +", d[1] - t0)
print("tile  q_full  qk_issued  s_full_seen  pass1_done  pass2_done  pv_issued  o_full_seen")
for t in range(3):
    print(t, d[8 + t] - t0, d[16 + t] - t0, d[32 + t] - t0, d[40 + t] - t0, d[48 + t] - t0, d[24 + t] - t0, d[56 + t] - t0)

if os.environ.get("AMB200_ATTN_PIPE", "1") != "0":
    print("tile 1, MMA warp: p_ready[chunk] seen:", [d[64 + c] - t0 for c in range(12) if d[64 + c]])
    print("tile 1, QK(t+1, kt) issued:", [d[88 + k] - t0 if d[88 + k] else None for k in range(3)])
    print("tile 1, block j finished by its softmax group:", [d[96 + c] - t0 for c in range(6) if d[96 + c]])
    print("epilogue (O read, staged) group 0 / group 1 per tile:", [((d[104 + 4 * t] - t0, d[106 + 4 * t] - t0), (d[105 + 4 * t] - t0, d[107 + 4 * t] - t0)) for t in range(3)])
