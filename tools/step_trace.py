"""Kernel-level trace of the CMDM sampling loop as it really runs (CUDA-graph replays): per kernel type the in-graph duration
and the idle gap that precedes it (CUPTI activity records via torch.profiler).  Answers: where do the 1.15 ms of a step go?"""
import os, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench as Bn
from amb200 import synth
from amb200.config import cmdm_model_cfg, full_cfg
from models.base import create_model_and_diffusion
from models.functions import set_text_feature_provider
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "cmdm"
nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
if which == "cmdm":   # python tools/step_trace.py [cmdm] [denoise_steps]
    nd = nums[0] if nums else 200
    model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(Bn.NPTS), steps=nd), device=dev)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    model.to(dev).eval()
    host = Bn.synth_host_inputs(0)
    txt = host["text"].to(dev)
    set_text_feature_provider(lambda raw: txt)
    kw = dict(c_text=host["texts"], c_pc_xyz=host["xyz"].to(dev), c_pc_contact=host["contact"].to(dev), x_mask=host["x_mask"].to(dev))
    job = lambda: diff.p_sample_loop(model, (Bn.B, Bn.T, Bn.DM), clip_denoised=False, model_kwargs=kw)
else:                 # python tools/step_trace.py cdm [batch]   (BASELINE config 3: 100-step DDIM of a 500-step process, N=8192)
    from amb200.config import cdm_model_cfg
    Bc = nums[0] if nums else 8
    model, diff = create_model_and_diffusion(full_cfg(cdm_model_cfg(Bn.NPTS), steps=500, timestep_respacing="ddim100"), device=dev)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    model.to(dev).eval()
    txt = synth.text_features(Bc, seed=3).to(dev)
    set_text_feature_provider(lambda raw: txt)
    kw = dict(c_text=[f"p{i}" for i in range(Bc)], c_pc_xyz=synth.scene_points(Bc, Bn.NPTS, seed=3).to(dev), c_pc_feat=None)
    job = lambda: diff.ddim_sample_loop(model, (Bc, Bn.NPTS, 6), clip_denoised=False, model_kwargs=kw, eta=0.0)
for _ in range(2):
    job()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    job()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
evs.sort(key=lambda e: e.time_range.start)
# keep the steady-state middle of the job (graph replays)
n = len(evs)
evs = evs[n // 4: 3 * n // 4]
agg = collections.OrderedDict()
prev_end = None
for e in evs:
    m = re.search(r"([A-Za-z_0-9]+_kernel(?:<[^(]*>)?)", e.name)
    name = (m.group(1) if m else e.name)[:60]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
    if prev_end is not None:
        a[2] += max(0.0, e.time_range.start - prev_end)
    prev_end = max(prev_end or 0, e.time_range.end)
span = evs[-1].time_range.end - evs[0].time_range.start
busy = sum(a[1] for a in agg.values())
gaps = sum(a[2] for a in agg.values())
steps = sum(a[0] for k, a in agg.items() if "sampler_update" in k)
print(f"{len(evs)} kernels over {span/1e3:.2f} ms = {steps} denoise steps -> {span/max(steps,1):.1f} us/step; busy {100*busy/span:.1f}%, gaps {100*gaps/span:.1f}%")
print(f"{'kernel':60s} {'n/step':>7s} {'avg us':>8s} {'us/step':>8s} {'gap before (avg us)':>20s} {'gap us/step':>11s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {a[0]/max(steps,1):7.1f} {a[1]/a[0]:8.2f} {a[1]/max(steps,1):8.1f} {a[2]/a[0]:20.2f} {a[2]/max(steps,1):11.1f}")
