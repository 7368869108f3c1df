"""One short CMDM sampling job (B=32, N=8192, 16 steps) for ncu captures of the once-per-job conditioning kernels (fps_kernel,
knn_kernel, pt_layer_kernel, transition_down_kernel) and the per-step elementwise kernels (sampler_update_kernel, layernorm):
    ncu --set full --clock-control none -k regex:'fps_kernel|knn_kernel|sampler_update|pt_layer|transition_down' -c 24 -o out python tools/prof_misc.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import torch
from amb200 import synth
from amb200.config import cmdm_model_cfg, full_cfg
from models.base import create_model_and_diffusion
from models.functions import set_text_feature_provider
dev = torch.device("cuda:0")
B, N, T, Dm = 32, 8192, 196, 263
model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N), steps=int(sys.argv[1]) if len(sys.argv) > 1 else 16), device=dev)
model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
model.to(dev).eval()
txt = synth.text_features(B, seed=3).to(dev)
set_text_feature_provider(lambda raw: txt)
kw = dict(c_text=["p"] * B, c_pc_xyz=synth.scene_points(B, N, seed=3).to(dev), c_pc_contact=synth.contact_map(B, N, seed=3).to(dev),
          x_mask=synth.motion_mask(B, T, seed=3, all_valid=True).to(dev))
out = diff.p_sample_loop(model, (B, T, Dm), clip_denoised=False, model_kwargs=kw)
torch.cuda.synchronize()
print("finite", bool(torch.isfinite(out).all()))
