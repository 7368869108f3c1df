"""Kernel-time table of one CMDM training step (torch.profiler CUDA activities)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from amb200 import synth
from amb200.config import cmdm_model_cfg, full_cfg
from models.base import create_model_and_diffusion
from models.functions import set_text_feature_provider
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N, T, Dm = 8192, 196, 263
dev = torch.device("cuda:0")
model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N)), device=dev)
model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
model.to(dev).train()
txt = synth.text_features(B, seed=0).to(dev)
set_text_feature_provider(lambda raw: txt)
kw = dict(c_text=["p"] * B, c_pc_xyz=synth.scene_points(B, N, seed=0).to(dev), c_pc_contact=synth.contact_map(B, N, seed=0).to(dev),
          x_mask=synth.motion_mask(B, T, seed=0).to(dev))
x0 = synth.motion_noise(B, T, Dm, seed=0).to(dev)
def step():
    model.zero_grad()
    t = torch.randint(0, 1000, (B,), device=dev)
    diff.training_losses(model, x0, t, model_kwargs=kw)["loss"].mean().backward()
step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
