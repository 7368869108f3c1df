"""Probe: does running the CMDM batch as TWO half-batches on two CUDA streams (two engines, same weights) hide the per-launch
tails / wave quantisation of the 42 kernels of a denoise step?  Compares one B=32 job with two concurrent B=16 jobs.
    python tools/dual_stream_probe.py [steps]"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import torch
from amb200 import synth
from amb200.config import cmdm_model_cfg, full_cfg
from models.base import create_model_and_diffusion
from models.functions import set_text_feature_provider
dev = torch.device("cuda:0")
N, T, Dm = 8192, 196, 263
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
txt = synth.text_features(32, seed=3).to(dev)
set_text_feature_provider(lambda raw: txt[: len(raw)])
def mk():
    m, d = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N), steps=steps), device=dev)
    m.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=0), strict=False)
    return m.to(dev).eval(), d
def kw(B, seed):
    return dict(c_text=["p"] * B, c_pc_xyz=synth.scene_points(B, N, seed=seed).to(dev), c_pc_contact=synth.contact_map(B, N, seed=seed).to(dev),
                x_mask=synth.motion_mask(B, T, seed=seed, all_valid=True).to(dev))
m32, d32 = mk()
k32 = kw(32, 1)
def job32():
    return d32.p_sample_loop(m32, (32, T, Dm), clip_denoised=False, model_kwargs=k32)
for _ in range(2):
    job32()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    job32()
torch.cuda.synchronize(); t32 = (time.perf_counter() - t0) / 3
print(f"one stream, B=32: {1e3 * t32:.1f} ms/job -> {steps / t32:.1f} denoise-steps/s")
halves = [mk() + (kw(16, 2 + i), torch.cuda.Stream()) for i in range(2)]
def run_half(i, reps):
    m, d, k, st = halves[i]
    with torch.cuda.stream(st):
        for _ in range(reps):
            d.p_sample_loop(m, (16, T, Dm), clip_denoised=False, model_kwargs=k)
for i in range(2):  # warm-up / graph capture, one at a time
    run_half(i, 2)
torch.cuda.synchronize()
def both(reps):
    th = [threading.Thread(target=run_half, args=(i, reps)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
both(1)
t0 = time.perf_counter(); both(3); t16 = (time.perf_counter() - t0) / 3
print(f"two streams, 2 x B=16: {1e3 * t16:.1f} ms per pair of jobs -> {steps / t16:.1f} denoise-steps/s at batch 32 ({100 * (t32 / t16 - 1):+.1f} %)")
run_half(0, 1); torch.cuda.synchronize(); t0 = time.perf_counter(); run_half(0, 3); torch.cuda.synchronize()
print(f"one stream, B=16 alone: {1e3 * (time.perf_counter() - t0) / 3:.1f} ms/job")
