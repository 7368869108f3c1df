"""CDM (BASELINE config 3 shapes: B per GPU, N=8192, 100-step DDIM of a 500-step process): per-kernel CUDA-event
breakdown of one denoise step + wall time of a full sampling job."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import torch
from amb200 import ops, synth
from amb200.config import cdm_model_cfg, full_cfg
from models.base import create_model_and_diffusion
from models.functions import set_text_feature_provider
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
BASE = int(sys.argv[2]) if len(sys.argv) > 2 else 500   # base diffusion steps: 500 (reference CDM scripts) or 1000 (stride 10)
N = 8192
model, diff = create_model_and_diffusion(full_cfg(cdm_model_cfg(N), steps=BASE, timestep_respacing="ddim100"), device=dev)
model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
model.to(dev).eval()
txt = synth.text_features(B, seed=3).to(dev)
set_text_feature_provider(lambda raw: txt)
xyz = synth.scene_points(B, N, seed=3).to(dev)
kw = dict(c_text=[f"p{i}" for i in range(B)], c_pc_xyz=xyz, c_pc_feat=None)
cond = model.encode_condition(**kw)
x = torch.randn(B, N, 6, device=dev)
t_dev = torch.full((1,), 50, device=dev, dtype=torch.int32)
out = torch.empty(B, N, 6, device=dev)
for _ in range(3):
    model.engine.forward(x, t_dev, 0, cond, out=out)
ops.PROFILER = ops.KernelProfiler()
for _ in range(3):
    model.engine.forward(x, t_dev, 0, cond, out=out)
agg = ops.PROFILER.summary(); ops.PROFILER = None
tot = sum(a["ms"] for a in agg.values())
print(f"CDM B={B} N={N}: one network evaluation = {tot/3:.3f} ms (sum of kernels)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:24s} {a['ms']/3:8.3f} ms/step  {a['launches']/3:5.1f} launches  {100*a['ms']/tot:5.1f}%")
SHORT = len(sys.argv) > 3 and sys.argv[3] == "short"   # ncu runs: skip the sampling jobs
for rep in range(0 if SHORT else 2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = diff.ddim_sample_loop(model, (B, N, 6), clip_denoised=False, model_kwargs=kw, eta=0.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"ddim100-of-{BASE} job (B={B}): {1e3*(t1-t0):.1f} ms -> {100/(t1-t0):.1f} denoise-steps/s, {B/(t1-t0):.1f} affordance maps/s, finite={bool(torch.isfinite(s).all())}")
