"""GPU parity of the fused flat-buffer AdamW (utils/training.py:48-50,139,154) against the oracle restatement
(oracle/optim_ref.py, itself pinned to torch.optim.AdamW on CPU) and against torch.optim.AdamW on the real CMDM model."""
import copy

import numpy as np
import pytest
import torch

from amb200 import synth
from amb200.config import cmdm_model_cfg, full_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_fused_adamw_vs_oracle_ragged_sizes():
    from amb200.optim import FusedAdamW
    from oracle.optim_ref import adamw_step
    g = torch.Generator().manual_seed(0)
    shapes = [(33, 7), (5,), (1,), (128, 64), (3, 3, 3)]  # sizes not multiples of 4: every view still starts 16-byte aligned
    p0 = [torch.randn(s, generator=g) for s in shapes]
    params = [torch.nn.Parameter(t.clone().to(DEV)) for t in p0]
    opt = FusedAdamW(params, lr=3e-3, weight_decay=0.02)
    ref = [t.clone() for t in p0]
    m = [torch.zeros_like(t) for t in p0]
    v = [torch.zeros_like(t) for t in p0]
    for step in range(1, 7):
        opt.zero_grad()
        assert all(float(q.grad.abs().max()) == 0.0 for q in params)
        grads = [torch.randn(s, generator=g) for s in shapes]
        loss = sum((q * gr.to(DEV)).sum() for q, gr in zip(params, grads))  # d loss / d q = gr, accumulated in place into the flat buffer
        loss.backward()
        if step == 4:
            for group in opt.param_groups:  # utils/training.py:84-90 anneals the lr through param_groups
                group["lr"] = 1e-3
        opt.step()
        for i in range(len(shapes)):
            adamw_step(ref[i], grads[i], m[i], v[i], step, lr=3e-3 if step < 4 else 1e-3, weight_decay=0.02)
            assert (params[i].detach().cpu() - ref[i]).abs().max().item() < 2e-7, (step, i)
            assert (opt.state[params[i]]["exp_avg_sq"].cpu() - v[i]).abs().max().item() < 1e-7
    # torch-format state_dict round trip into a torch.optim.AdamW and back
    sd = opt.state_dict()
    t_opt = torch.optim.AdamW([torch.nn.Parameter(q.detach().clone()) for q in params], lr=1e-3, weight_decay=0.02)
    t_opt.load_state_dict(copy.deepcopy(sd))
    opt2 = FusedAdamW([torch.nn.Parameter(q.detach().clone()) for q in params], lr=1e-3, weight_decay=0.02)
    opt2.load_state_dict(t_opt.state_dict())
    assert opt2._flat[0]["step"] == 6
    for a, b in zip(opt.param_groups[0]["params"], opt2.param_groups[0]["params"]):
        assert torch.equal(opt.state[a]["exp_avg"], opt2.state[b]["exp_avg"])


def test_fused_adamw_trains_cmdm_like_torch_adamw():
    """Three CMDM training steps (B=2, N=1024) with FusedAdamW vs torch.optim.AdamW from identical weights and RNG state:
    losses and every parameter agree; the sampling engine sees the updated weights (version counters bumped)."""
    from amb200.optim import FusedAdamW
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    B, N, T, Dm = 2, 1024, 196, 263
    txt = synth.text_features(B, seed=31)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        runs = []
        for fused in (False, True):
            model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N)), device=DEV)
            model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
            model.to(DEV).train()
            for mod in model.modules():  # deterministic comparison: no dropout
                if isinstance(mod, torch.nn.Dropout):
                    mod.p = 0.0
                if isinstance(mod, torch.nn.MultiheadAttention):
                    mod.dropout = 0.0
            params = [p for p in model.parameters() if p.requires_grad]
            opt = (FusedAdamW if fused else torch.optim.AdamW)(params, lr=1e-4, weight_decay=0.0)
            kw = dict(c_text=["a"] * B, c_pc_xyz=synth.scene_points(B, N, seed=31).to(DEV), c_pc_contact=synth.contact_map(B, N, seed=31).to(DEV),
                      x_mask=synth.motion_mask(B, T, seed=31).to(DEV))
            x0 = synth.motion_noise(B, T, Dm, seed=31).to(DEV)
            losses = []
            torch.manual_seed(5)
            for it in range(3):
                opt.zero_grad()
                t = torch.tensor([700 - it, 23 + it], device=DEV)
                noise = synth.step_noise((B, T, Dm), 70 + it).to(DEV)
                loss = diff.training_losses(model, x0, t, model_kwargs=kw, noise=noise)["loss"].mean()
                loss.backward()
                opt.step()
                losses.append(float(loss))
            model.eval()
            with torch.no_grad():
                out = model(x0, torch.tensor([10, 10], device=DEV), **kw)  # sampling engine must pick up the updated weights
            runs.append((losses, {n: p.detach().clone() for n, p in model.named_parameters()}, out))
        (l0, p0, o0), (l1, p1, o1) = runs
        assert np.allclose(l0, l1, rtol=1e-5, atol=1e-6), (l0, l1)
        assert l0[0] != l0[2]
        for n in p0:
            assert (p0[n] - p1[n]).abs().max().item() < 2e-6, n
        assert (o0 - o1).abs().max().item() < 1e-4
    finally:
        set_text_feature_provider(None)
