"""GPU tests of the sampling pipelines at BASELINE shapes: CDM 100-step DDIM (config 3 shapes, per-GPU shard),
device-resident loop vs step-by-step reference semantics, and the two-stage CDM -> CMDM hand-off (config 5)."""
import numpy as np
import pytest
import torch

from amb200 import ops, synth
from amb200.config import cdm_model_cfg, cmdm_model_cfg, full_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mk(cfg, steps, respacing=""):
    from models.base import create_model_and_diffusion
    model, diff = create_model_and_diffusion(full_cfg(cfg, steps=steps, timestep_respacing=respacing), device=DEV)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    return model.to(DEV).eval(), diff


def test_cdm_ddim_loop_matches_stepwise_oracle():
    """B=2, N=1024, ddim100 of 500: the graph-captured device loop equals the oracle DDIM recursion (eta=0 is
    deterministic given x_T), within the 1e-3 budget after 100 chained network evaluations."""
    from models.functions import set_text_feature_provider
    from oracle import cdm_ref, diffusion_ref as D
    B, N = 2, 1024
    model, diff = _mk(cdm_model_cfg(N), 500, "ddim20")
    xyz = synth.scene_points(B, N, seed=11)
    txt = synth.text_features(B, seed=11)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        xT = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(5))
        kw = dict(c_text=["a"] * B, c_pc_xyz=xyz.to(DEV), c_pc_feat=None)
        out = diff.ddim_sample_loop(model, (B, N, 6), noise=xT.to(DEV), clip_denoised=False, model_kwargs=kw, eta=0.0)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        nb, tmap = D.respaced(D.cosine_betas(500), D.space_timesteps(500, "ddim20"))
        assert diff.timestep_map == tmap
        tab = D.make_tables(nb)
        img = xT.clone()
        for i in range(len(tmap) - 1, -1, -1):
            t = torch.full((B,), i, dtype=torch.long)
            x0 = cdm_ref.cdm_forward(sd, img, torch.tensor([tmap[i]] * B), txt, xyz)
            img = D.ddim_step(tab, x0, img, t, torch.zeros_like(img), eta=0.0)
        err = (out.cpu() - img).abs().max().item()
        assert err < 1e-3, err
    finally:
        set_text_feature_provider(None)


def test_cdm_full_size_ddim_properties():
    """Config-3 per-GPU shard (B=8, N=8192, 100 DDIM steps of 500): finite, deterministic, shard-invariant."""
    from models.functions import set_text_feature_provider
    B, N = 8, 8192
    model, diff = _mk(cdm_model_cfg(N), 500, "ddim100")
    assert diff.num_timesteps == 100 and diff.timestep_map[:3] == [0, 5, 10]
    xyz = synth.scene_points(B, N, seed=12).to(DEV)
    txt = synth.text_features(B, seed=12)
    texts = [f"t{i}" for i in range(B)]
    set_text_feature_provider(lambda raw: torch.stack([txt[int(s[1:])] for s in raw]))
    try:
        kw = dict(c_text=texts, c_pc_xyz=xyz, c_pc_feat=None)
        torch.manual_seed(3)
        a = diff.ddim_sample_loop(model, (B, N, 6), clip_denoised=False, model_kwargs=kw, eta=0.0)
        torch.manual_seed(3)
        b = diff.ddim_sample_loop(model, (B, N, 6), clip_denoised=False, model_kwargs=kw, eta=0.0)
        assert torch.isfinite(a).all() and torch.equal(a, b)
        kw2 = dict(c_text=texts[:2], c_pc_xyz=xyz[:2].contiguous(), c_pc_feat=None)
        torch.manual_seed(3)
        c = diff.ddim_sample_loop(model, (2, N, 6), clip_denoised=False, model_kwargs=kw2, eta=0.0)
        assert (c - a[:2]).abs().max() < 1e-4
    finally:
        set_text_feature_provider(None)


def test_two_stage_cdm_to_cmdm_handoff():
    """Config 5 (reduced steps): CDM sample -> contact map on device (no .npy round trip) -> CMDM sample."""
    from amb200.pipeline import contact_from_cdm_sample, two_stage_generate
    from models.functions import set_text_feature_provider
    B, N, T, Dm = 2, 8192, 196, 263
    cdm, cdiff = _mk(cdm_model_cfg(N), 500, "ddim10")
    cmdm, mdiff = _mk(cmdm_model_cfg(N), 12)
    xyz = synth.scene_points(B, N, seed=13).to(DEV)
    txt = synth.text_features(B, seed=13)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        # the fused hand-off equals the reference's denormalise -> clip -> distance -> re-exponentiate round trip
        s = torch.randn(B, N, 6, device=DEV)
        mean, std, sigma = 0.2, 0.3, 0.8
        c = contact_from_cdm_sample(s, mean, std)
        ref = (s * std + mean).clamp(1e-20, 1.0)
        dist = torch.sqrt(-2 * torch.log(ref) * sigma ** 2)           # utils/evaluate.py:55-66
        back = torch.exp(-0.5 * dist ** 2 / sigma ** 2)               # datasets/humanml3d.py:773-774
        assert (c - back).abs().max() < 1e-5
        torch.manual_seed(1)
        x_mask = synth.motion_mask(B, T, seed=13).to(DEV)
        motion, contact = two_stage_generate(cdm, cdiff, cmdm, mdiff, ["a"] * B, xyz, x_mask, (T, Dm), contact_mean=mean, contact_std=std,
                                             ddim=True)
        assert motion.shape == (B, T, Dm) and contact.shape == (B, N, 6)
        assert torch.isfinite(motion).all() and (contact > 0).all() and (contact <= 1).all()
    finally:
        set_text_feature_provider(None)
