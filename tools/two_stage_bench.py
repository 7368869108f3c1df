"""BASELINE config 5: two-stage CDM -> CMDM generation, batch 16 (2 per GPU at 8 GPUs), 100 DDIM steps (of a 500-step process)
+ 1000 DDPM steps, N=8192, with the on-device contact hand-off (amb200.pipeline).
    python tools/two_stage_bench.py [B]      |  torchrun --nproc-per-node N tools/two_stage_bench.py [B_per_gpu]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import torch
from amb200 import dist as amdist, synth
from amb200.config import cdm_model_cfg, cmdm_model_cfg, full_cfg
from amb200.pipeline import two_stage_generate
from models.base import create_model_and_diffusion
from models.functions import set_text_feature_provider

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N, T, Dm = 8192, 196, 263
rank, world, local = amdist.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
amdist.init("nccl", dev)
def mk(cfg, steps, resp=""):
    m, d = create_model_and_diffusion(full_cfg(cfg, steps=steps, timestep_respacing=resp), device=dev)
    m.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=0), strict=False)
    return m.to(dev).eval(), d
cdm, cdiff = mk(cdm_model_cfg(N), 500, "ddim100")
cmdm, mdiff = mk(cmdm_model_cfg(N), 1000)
cdiff.sample_offset = mdiff.sample_offset = rank * B
txt = synth.text_features(B, seed=rank).to(dev)
set_text_feature_provider(lambda raw: txt[: len(raw)])
xyz = synth.scene_points(B, N, seed=rank).to(dev)
x_mask = synth.motion_mask(B, T, seed=rank, all_valid=True).to(dev)
times = []
for it in range(4):
    torch.cuda.synchronize(); amdist.barrier(); t0 = time.perf_counter()
    motion, contact = two_stage_generate(cdm, cdiff, cmdm, mdiff, [f"p{i}" for i in range(B)], xyz, x_mask, (T, Dm), contact_mean=0.2,
                                         contact_std=0.3, ddim=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if it >= 1:
        times.append(dt)
ms = amdist.max_over_ranks(1e3 * sum(times) / len(times), device=dev)
if rank == 0:
    print(f"two-stage CDM(100 DDIM) -> CMDM(1000 DDPM): {B}/GPU x {world} GPU(s): {ms:.1f} ms/job -> {B * world / (ms / 1e3):.1f} motions/s, "
          f"{1100 * world / (ms / 1e3):.0f} denoise-steps/s; finite={bool(torch.isfinite(motion).all())}")
if world > 1:
    torch.distributed.destroy_process_group()
