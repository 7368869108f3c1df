"""CMDM training step (BASELINE config 4 shapes: 32 samples per GPU, T=196, N=8192): fwd + bwd + AdamW, optionally under
torchrun with SyncBatchNorm + DistributedDataParallel exactly as train_ddp.py:63-65 wraps the model.
    python tools/train_bench.py [B] [steps] [ddp|native]   |  python -m torch.distributed.run --nproc-per-node 2 ... tools/train_bench.py
mode "ddp" (default)   : torch.optim.AdamW + SyncBatchNorm + DistributedDataParallel exactly as train_ddp.py:63-65 wraps the model
mode "native"          : amb200.optim.FusedAdamW (one update launch) + SyncBatchNorm + ONE flat-buffer gradient all-reduce, no DDP"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from amb200 import dist as amdist, synth
from amb200.config import cmdm_model_cfg, full_cfg
from diffusion.resample import uniform_sampling
from models.base import create_model_and_diffusion
from models.functions import set_text_feature_provider

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "ddp"
N, T, Dm = 8192, 196, 263
rank, world, local = amdist.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
amdist.init("nccl", dev)
model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N)), device=dev)
model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
model.to(dev)
net = model
if world > 1:  # train_ddp.py:63-65
    net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    if mode == "ddp":
        net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True, broadcast_buffers=False)
net.train()
if mode == "native":
    from amb200.optim import FusedAdamW
    opt = FusedAdamW([p for p in net.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.0)
    overlap = world > 1 and os.environ.get("AMB200_GRAD_OVERLAP", "1") != "0"
    if overlap:
        opt.enable_overlap(3)
else:
    opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.0)
txt = synth.text_features(B, seed=rank).to(dev)
set_text_feature_provider(lambda raw: txt)
xyz, contact = synth.scene_points(B, N, seed=rank).to(dev), synth.contact_map(B, N, seed=rank).to(dev)
x0, x_mask = synth.motion_noise(B, T, Dm, seed=rank).to(dev), synth.motion_mask(B, T, seed=rank).to(dev)
kw = dict(c_text=["p"] * B, c_pc_xyz=xyz, c_pc_contact=contact, x_mask=x_mask)
np.random.seed(2023 + rank)
times, losses = [], []
for it in range(steps + 2):
    torch.cuda.synchronize(); amdist.barrier(); t0 = time.perf_counter()
    opt.zero_grad()
    t = uniform_sampling(B, dev, diff.num_timesteps)
    terms = diff.training_losses(net, x0, t, model_kwargs=kw)
    loss = terms["loss"].mean()
    if mode == "native" and overlap:
        opt.begin_overlap()
    loss.backward()
    if mode == "native":
        opt.all_reduce_grads()
    opt.step()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if it >= 2:
        times.append(dt)
    losses.append(float(loss.detach()))
ms = amdist.max_over_ranks(1e3 * sum(times) / len(times), device=dev)
if rank == 0:
    print(f"CMDM training step [{mode}]: {B}/GPU x {world} GPU(s): {ms:.1f} ms/step -> {B * world / (ms / 1e3):.1f} samples/s; "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB; losses {['%.4f' % l for l in losses]}")
if world > 1:
    dist.destroy_process_group()
