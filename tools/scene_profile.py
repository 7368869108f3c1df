"""Frozen PointTransformerSeg scene model at HUMANISE shapes: B per GPU, N=8192 -> [B,N,32].
Per-kernel CUDA-event breakdown of one forward (it runs once per batch, hoisted out of the denoise loop)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import torch
from amb200 import ops, synth
from models.scene_models.pointtransformer import pointtransformer_seg_repro
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 8192
seg = pointtransformer_seg_repro(c=3, num_points=N)
seg.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in seg.state_dict().items()}, seed=0), strict=False)
seg.to(dev).eval()
xyz = synth.scene_points(B, N, seed=3).to(dev)
for _ in range(2):
    out = seg((xyz, None))
ops.PROFILER = ops.KernelProfiler()
out = seg((xyz, None))
agg = ops.PROFILER.summary(); ops.PROFILER = None
tot = sum(a["ms"] for a in agg.values())
torch.cuda.synchronize(); t0 = time.perf_counter()
out = seg((xyz, None))
torch.cuda.synchronize(); wall = 1e3 * (time.perf_counter() - t0)
print(f"PointTransformerSeg B={B} N={N}: {tot:.2f} ms sum of kernels, {wall:.2f} ms wall, finite={bool(torch.isfinite(out).all())}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:24s} {a['ms']:8.3f} ms  {a['launches']:4d} launches  {100*a['ms']/tot:5.1f}%")
