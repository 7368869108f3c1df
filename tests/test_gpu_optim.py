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
    """Three CMDM training steps (B=2, N=1024): model A is driven by torch.optim.AdamW, model B by FusedAdamW.
    (1) B's own backward accumulates straight into the flat gradient buffer and reproduces A's gradients;
    (2) fed the SAME gradients, both optimisers move every parameter identically (Adam's m/sqrt(v) amplifies run-to-run
        rounding noise of tiny gradients, so the trajectories are compared on identical gradients);
    (3) the sampling engine sees the updated weights (version counters bumped by the fused step)."""
    from amb200.optim import FusedAdamW
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    B, N, T, Dm = 2, 1024, 196, 263
    txt = synth.text_features(B, seed=31)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        models, opts = [], []
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
            opts.append((FusedAdamW if fused else torch.optim.AdamW)(params, lr=1e-4, weight_decay=0.0))
            models.append(model)
        (ma, mb), (oa, ob) = models, opts
        kw = dict(c_text=["a"] * B, c_pc_xyz=synth.scene_points(B, N, seed=31).to(DEV), c_pc_contact=synth.contact_map(B, N, seed=31).to(DEV),
                  x_mask=synth.motion_mask(B, T, seed=31).to(DEV))
        x0 = synth.motion_noise(B, T, Dm, seed=31).to(DEV)
        flat_g = ob.flat_grads()[0]
        mb.eval()
        with torch.no_grad():  # builds B's sampling engine (packed / bf16-split weight copies) from the INITIAL weights
            o_init = mb(x0, torch.tensor([10, 10], device=DEV), **kw).clone()
        mb.train()
        losses = []
        for it in range(3):
            t = torch.tensor([700 - it, 23 + it], device=DEV)
            noise = synth.step_noise((B, T, Dm), 70 + it).to(DEV)
            oa.zero_grad()
            ob.zero_grad()
            la = diff.training_losses(ma, x0, t, model_kwargs=kw, noise=noise)["loss"].mean()
            la.backward()
            lb = diff.training_losses(mb, x0, t, model_kwargs=kw, noise=noise)["loss"].mean()
            lb.backward()
            losses.append((float(la.detach()), float(lb.detach())))
            assert abs(float(la.detach()) - float(lb.detach())) < 1e-5 * max(1.0, abs(float(la.detach())))
            gnorm = max(float(pa.grad.norm()) for pa in ma.parameters() if pa.grad is not None)
            for (n, pa), pb in zip(ma.named_parameters(), mb.parameters()):
                if pa.grad is None:
                    continue
                lo, hi = flat_g.data_ptr(), flat_g.data_ptr() + 4 * flat_g.numel()
                assert lo <= pb.grad.data_ptr() < hi, n                      # (1) still a view of the flat buffer
                # same criterion as tests/test_gpu_training.py: relative L2 with a floor (encoder gradients are sums of cancelling
                # terms whose fp32 summation order differs run to run; zero-gradient parameters carry pure rounding noise)
                err = float((pa.grad - pb.grad).double().norm() / (pa.grad.double().norm() + 1e-6 * gnorm))
                assert err < 2e-2, (n, err)
                pb.grad.copy_(pa.grad)                                       # (2) identical gradients from here on
            oa.step()
            ob.step()
            for (n, pa), pb in zip(ma.named_parameters(), mb.parameters()):
                assert float((pa - pb).abs().max()) < 5e-7, (it, n)  # <= 4 ulp at |p| ~ 1 (an lr=1e-4 step is 200x larger)
        assert losses[0][0] != losses[2][0]
        ma.eval(); mb.eval()
        with torch.no_grad():  # (3) sampling engine must pick up the updated weights
            oa_ = ma(x0, torch.tensor([10, 10], device=DEV), **kw)
            ob_ = mb(x0, torch.tensor([10, 10], device=DEV), **kw)
        assert float((oa_ - ob_).abs().max()) < 5e-5  # parameters agree to 5e-7; the trunk carries activations as bf16 pairs (~1e-5)
        assert float((o_init - ob_).abs().max()) > 1e-4  # ... the engine re-packed: the output moved away from the initial weights'
    finally:
        set_text_feature_provider(None)
