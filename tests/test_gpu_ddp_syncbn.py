"""the reference's data-parallel wrapping (train_ddp.py:63-65: SyncBatchNorm.convert_sync_batchnorm +
DistributedDataParallel(find_unused_parameters=True, broadcast_buffers=False)) around the drop-in CMDM on 2 GPUs over NCCL, and the
native exchange (SyncBatchNorm + ONE flat gradient all-reduce, amb200.optim.FusedAdamW) — both against the single-process
gradients of the same global batch.  With SyncBatchNorm the batch statistics are global, the per-sample masked-MSE losses are
normalised per sample, and DDP averages gradients over ranks, so the 2-rank gradients must equal the 1-rank ones.
Needs >= 2 GPUs: skipped on a single-GPU box (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp_syncbn.py -m gpu`)."""
import os
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
N, T, DM, BG = 1024, 196, 263, 4   # global batch 4 = 2 ranks x 2
WATCH = ("motion_layer.weight", "self_attn_layer.layers.0.self_attn.in_proj_weight", "contact_adapter.weight",
         "contact_encoder.enc1.0.linear.weight", "contact_encoder.enc4.1.transformer2.linear_w.2.weight", "contact_encoder.enc2.0.bn.weight")


def _build(dev):
    from amb200 import synth
    from amb200.config import cmdm_model_cfg, full_cfg
    from models.base import create_model_and_diffusion
    model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N)), device=dev)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    model.to(dev).train()
    for mod in model.modules():  # deterministic comparison: every dropout probability 0
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    return model, diff


def _batch(lo, hi, dev):
    from amb200 import synth
    sl = slice(lo, hi)
    kw = dict(c_text=["a"] * (hi - lo), c_pc_xyz=synth.scene_points(BG, N, seed=71)[sl].contiguous().to(dev),
              c_pc_contact=synth.contact_map(BG, N, seed=71)[sl].contiguous().to(dev), x_mask=synth.motion_mask(BG, T, seed=71)[sl].contiguous().to(dev))
    x0 = synth.motion_noise(BG, T, DM, seed=71)[sl].contiguous().to(dev)
    noise = synth.motion_noise(BG, T, DM, seed=72)[sl].contiguous().to(dev)
    t = torch.tensor([900, 400, 50, 3])[sl].to(dev)
    return x0, t, noise, kw


def _worker(rank, world, port, mode, out_path):
    import torch.distributed as dist
    from amb200 import synth
    from models.functions import set_text_feature_provider
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model, diff = _build(dev)
        txt = synth.text_features(BG, seed=71)
        per = BG // world
        set_text_feature_provider(lambda raw: txt[rank * per: rank * per + len(raw)])
        net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)             # train_ddp.py:63
        if mode == "ddp":
            net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[rank], find_unused_parameters=True,
                                                            broadcast_buffers=False)  # train_ddp.py:64-65
            opt = None
        else:
            from amb200.optim import FusedAdamW
            opt = FusedAdamW([p for p in net.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.0)
            opt.zero_grad()
            opt.enable_overlap(3)   # pieces of the flat buffer are all-reduced while backward is still running
            opt.begin_overlap()
        x0, t, noise, kw = _batch(rank * per, (rank + 1) * per, dev)
        loss = diff.training_losses(net, x0, t, model_kwargs=kw, noise=noise)["loss"].mean()
        loss.backward()
        scale = 1.0
        if mode == "native":
            scale = 1.0 / opt.all_reduce_grads()   # ONE flat all-reduce(sum); the mean is folded into the next step()
        params = dict((net.module if mode == "ddp" else net).named_parameters())
        if rank == 0:
            torch.save({"loss": float(loss), "grads": {k: (params[k].grad * scale).cpu() for k in WATCH},
                        "bn_mean": dict(model.named_buffers())["contact_encoder.enc1.0.bn.running_mean"].cpu()}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ddp", "native"])
def test_two_rank_gradients_equal_single_rank(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from amb200 import synth
    from models.functions import set_text_feature_provider
    # ---- single process, whole global batch (BatchNorm over all 4 samples = what SyncBatchNorm computes over 2 x 2)
    dev = torch.device("cuda", 0)
    model, diff = _build(dev)
    txt = synth.text_features(BG, seed=71)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        x0, t, noise, kw = _batch(0, BG, dev)
        terms = diff.training_losses(model, x0, t, model_kwargs=kw, noise=noise)
        terms["loss"].mean().backward()
    finally:
        set_text_feature_provider(None)
    ref = {k: dict(model.named_parameters())[k].grad.cpu() for k in WATCH}
    ref_bn = dict(model.named_buffers())["contact_encoder.enc1.0.bn.running_mean"].cpu()
    loss_rank0_ref = float(terms["loss"][:2].mean())
    # ---- two ranks
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "r0.pt")
        mp.spawn(_worker, args=(2, 29650 + (os.getpid() % 200), mode, out), nprocs=2, join=True)
        got = torch.load(out)
    assert abs(got["loss"] - loss_rank0_ref) < 2e-5 * max(1.0, abs(loss_rank0_ref))  # rank 0's samples, global BN statistics
    for k in WATCH:
        a, b = got["grads"][k].double(), ref[k].double()
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        assert rel < 2e-3, (k, rel)
    assert (got["bn_mean"] - ref_bn).abs().max().item() < 1e-5  # running statistics follow the GLOBAL batch on every rank
