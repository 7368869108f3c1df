"""GPU parity of the rank-collapsed CDM Perceiver point path (csrc/perceiver_tc.cu: SIMT encoder statistic + tcgen05 decoder tile
kernel) against the unfolded CPU oracle (oracle/cdm_ref.py, pinned by tests/golden/cdm_b2_n1024.npz) at BASELINE shapes, against
the general-cin kernels of csrc/perceiver.cu, and the conditioning-cache regression (pointer-keyed cache)."""
import pytest
import torch

from amb200 import ops, synth
from amb200.config import cdm_model_cfg, full_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NET_TOL = 1e-3  # BASELINE.json north_star: within 1e-3 max-abs on fp32


def _mk(N, steps=500, respacing=""):
    from models.base import create_model_and_diffusion
    model, diff = create_model_and_diffusion(full_cfg(cdm_model_cfg(N), steps=steps, timestep_respacing=respacing), device=DEV)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    return model.to(DEV).eval(), diff


@pytest.mark.parametrize("B,N", [(2, 1024), (3, 1000), (1, 77), (8, 8192), (5, 512)])
def test_collapsed_forward_vs_oracle_and_general_path(B, N):
    """Config-1 shape, ragged tile counts (N not a multiple of 128 / smaller than one tile) and the config-3 per-GPU shard
    (B=8, N=8192): collapsed kernels vs the straight oracle; and vs the general (unfolded-per-point) CUDA kernels."""
    from models.functions import set_text_feature_provider
    from oracle import cdm_ref
    model, _ = _mk(N)
    assert model.engine.point_path == "collapsed"
    xyz = synth.scene_points(B, N, seed=21, dup_frac=0.05)
    x = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(21))
    t = torch.tensor([(37 * i + 499) % 500 for i in range(B)])
    txt = synth.text_features(B, seed=21)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        with torch.no_grad():
            kw = dict(c_text=["a"] * B, c_pc_xyz=xyz.to(DEV), c_pc_feat=None)
            out = model(x.to(DEV), t.to(DEV), **kw)
            assert model.engine.K is not None and ("c", B, N, DEV) in model.engine._ws  # the collapsed kernels ran
            ref = cdm_ref.cdm_forward({k: v.detach().cpu() for k, v in model.state_dict().items()}, x, t, txt, xyz)
            err = (out.cpu() - ref).abs().max().item()
            assert err < 1e-4, err  # budget NET_TOL; the collapsed path's measured noise floor is ~3e-6
            assert model.engine.latw is not None and model.engine.latent_path == "fused"  # the cluster kernels ran
            model.engine.latent_path = "layers"   # same collapsed point kernels, latent side as one launch per layer
            out_l = model(x.to(DEV), t.to(DEV), **kw)
            assert (out_l - out).abs().max().item() < 2e-5
            model.engine.point_path = "general"
            model._cond_cache = None
            out_g = model(x.to(DEV), t.to(DEV), **kw)
            assert (out_g - out).abs().max().item() < 1e-4
    finally:
        model.engine.point_path = "collapsed"
        model.engine.latent_path = "fused"
        set_text_feature_provider(None)


def test_collapsed_large_inputs():
    """|u| up to ~8 (late DDIM states, scaled scenes): the LayerNorm quadratic forms stay within budget."""
    from models.functions import set_text_feature_provider
    from oracle import cdm_ref
    B, N = 2, 2048
    model, _ = _mk(N)
    xyz = synth.scene_points(B, N, seed=22) * 2.5
    x = 3.0 * torch.randn(B, N, 6, generator=torch.Generator().manual_seed(22))
    t = torch.tensor([250, 0])
    txt = synth.text_features(B, seed=22)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        with torch.no_grad():
            out = model(x.to(DEV), t.to(DEV), c_text=["a"] * B, c_pc_xyz=xyz.to(DEV), c_pc_feat=None)
            ref = cdm_ref.cdm_forward({k: v.detach().cpu() for k, v in model.state_dict().items()}, x, t, txt, xyz)
        assert (out.cpu() - ref).abs().max().item() < NET_TOL / 2
    finally:
        set_text_feature_provider(None)


def test_cdm_ddim100_chain_n8192_vs_oracle():
    """BASELINE config 3 recursion at its own point count: 100 DDIM steps (ddim100 of 500, eta = 0) at N = 8192, B = 2, graph-
    captured device loop vs the oracle recursion evaluated step by step on the CPU."""
    from models.functions import set_text_feature_provider
    from oracle import cdm_ref, diffusion_ref as D
    B, N = 2, 8192
    model, diff = _mk(N, 500, "ddim100")
    xyz = synth.scene_points(B, N, seed=23)
    txt = synth.text_features(B, seed=23)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        xT = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(23))
        kw = dict(c_text=["a"] * B, c_pc_xyz=xyz.to(DEV), c_pc_feat=None)
        out = diff.ddim_sample_loop(model, (B, N, 6), noise=xT.to(DEV), clip_denoised=False, model_kwargs=kw, eta=0.0)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        nb, tmap = D.respaced(D.cosine_betas(500), D.space_timesteps(500, "ddim100"))
        tab = D.make_tables(nb)
        img = xT.clone()
        with torch.no_grad():
            for i in range(len(tmap) - 1, -1, -1):
                x0 = cdm_ref.cdm_forward(sd, img, torch.tensor([tmap[i]] * B), txt, xyz)
                img = D.ddim_step(tab, x0, img, torch.full((B,), i, dtype=torch.long), torch.zeros_like(img), eta=0.0)
        err = (out.cpu() - img).abs().max().item()
        assert err < NET_TOL, err
    finally:
        set_text_feature_provider(None)


def test_condition_cache_is_not_keyed_on_addresses():
    """a new tensor with different content at a recycled address must not hit the conditioning cache.
    Frees the scene tensor, allocates another of the same shape (the caching allocator hands back the same block) and checks
    that the output follows the new content — for the per-step forward() path and for sampler_begin."""
    from models.functions import set_text_feature_provider
    B, N = 2, 1024
    model, diff = _mk(N, 500, "ddim5")
    txt = synth.text_features(B, seed=24)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        x = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(24)).to(DEV)
        t = torch.tensor([100, 100], device=DEV)
        xyz_a = synth.scene_points(B, N, seed=24).to(DEV)
        with torch.no_grad():
            out_a = model(x, t, c_text=["a"] * B, c_pc_xyz=xyz_a, c_pc_feat=None).clone()
            out_a2 = model(x, t, c_text=["a"] * B, c_pc_xyz=xyz_a, c_pc_feat=None)
            assert torch.equal(out_a, out_a2)  # same objects: cache hit, same answer
            host_b = synth.scene_points(B, N, seed=99)
            del xyz_a
            xyz_b = host_b.to(DEV)
            out_b = model(x, t, c_text=["a"] * B, c_pc_xyz=xyz_b, c_pc_feat=None)
            assert (out_b - out_a).abs().max().item() > 1e-3, "stale conditioning: output did not follow the new scene"
            # in-place edits of a cached tensor must also invalidate
            xyz_b.mul_(0.5)
            out_c = model(x, t, c_text=["a"] * B, c_pc_xyz=xyz_b, c_pc_feat=None)
            assert (out_c - out_b).abs().max().item() > 1e-3
            # sampler path: two jobs, same prompts, different scenes -> different samples
            xT = torch.randn(B, N, 6, device=DEV)
            s1 = diff.ddim_sample_loop(model, (B, N, 6), noise=xT, clip_denoised=False, eta=0.0,
                                       model_kwargs=dict(c_text=["a"] * B, c_pc_xyz=xyz_b, c_pc_feat=None)).clone()
            xyz_b = None
            xyz_c = synth.scene_points(B, N, seed=123).to(DEV)
            s2 = diff.ddim_sample_loop(model, (B, N, 6), noise=xT, clip_denoised=False, eta=0.0,
                                       model_kwargs=dict(c_text=["a"] * B, c_pc_xyz=xyz_c, c_pc_feat=None))
            assert (s1 - s2).abs().max().item() > 1e-3
    finally:
        set_text_feature_provider(None)


def test_cmdm_condition_cache_follows_contact_content():
    """Same prompts and scene, a NEW contact tensor (two_stage_generate allocates one per job): the CMDM must denoise against the
    new contact map, not a cached encoding of the previous one."""
    from amb200.config import cmdm_model_cfg
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    B, N, T, Dm = 2, 1024, 196, 263
    model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N), steps=4), device=DEV)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    model.to(DEV).eval()
    txt = synth.text_features(B, seed=25)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        xyz = synth.scene_points(B, N, seed=25).to(DEV)
        x_mask = synth.motion_mask(B, T, seed=25).to(DEV)
        x = synth.motion_noise(B, T, Dm, seed=25).to(DEV)
        t = torch.tensor([3, 3], device=DEV)
        outs = []
        with torch.no_grad():
            for seed in (25, 26):
                contact = synth.contact_map(B, N, seed=seed).to(DEV).clamp_(1e-20, 1.0)  # fresh tensor, _version 1, maybe same address
                outs.append(model(x, t, c_text=["a"] * B, c_pc_xyz=xyz, c_pc_contact=contact, x_mask=x_mask).clone())
                del contact
        valid = ~x_mask
        assert (outs[0] - outs[1])[valid].abs().max().item() > 1e-4
        samples = []
        for seed in (25, 26):
            contact = synth.contact_map(B, N, seed=seed).to(DEV).clamp_(1e-20, 1.0)
            samples.append(diff.p_sample_loop(model, (B, T, Dm), noise=x, clip_denoised=False,
                                              model_kwargs=dict(c_text=["a"] * B, c_pc_xyz=xyz, c_pc_contact=contact, x_mask=x_mask)).clone())
            del contact
        assert (samples[0] - samples[1])[valid].abs().max().item() > 1e-4
    finally:
        set_text_feature_provider(None)
