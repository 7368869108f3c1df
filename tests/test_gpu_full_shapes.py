"""GPU parity at the BASELINE configurations' own shapes (every oracle comparison ran on reduced batches):
  * one CMDM denoise step at config 2's batch (B=32, T=196, N=8192, mixed lengths) vs oracle/cmdm_ref.py,
  * the full 1000-step ancestral chain with injected noise (B=2, N=1024; conditioning hoisted on the CPU side exactly like the
    CUDA path hoists it) vs the oracle recursion — the end-to-end check BASELINE.md §4.4 asks for,
  * two-stage CDM -> CMDM generation vs an oracle two-stage run on the same injected noise.
Tolerance everywhere: 1e-3 max-abs on valid frames (BASELINE.json north_star)."""
import pytest
import torch

from amb200 import synth
from amb200.config import cdm_model_cfg, cmdm_model_cfg, full_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NET_TOL = 1e-3


def _mk(cfg, steps, respacing=""):
    from models.base import create_model_and_diffusion
    model, diff = create_model_and_diffusion(full_cfg(cfg, steps=steps, timestep_respacing=respacing), device=DEV)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    return model.to(DEV).eval(), diff


def _sd(model):
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}


def test_cmdm_step_at_config2_batch_vs_oracle():
    from models.functions import set_text_feature_provider
    from oracle import cmdm_ref
    B, N, T, Dm = 32, 8192, 196, 263
    model, _ = _mk(cmdm_model_cfg(N), 1000)
    xyz, contact = synth.scene_points(B, N, seed=31, dup_frac=0.05), synth.contact_map(B, N, seed=31)
    x, x_mask = synth.motion_noise(B, T, Dm, seed=31), synth.motion_mask(B, T, seed=31)
    txt = synth.text_features(B, seed=31)
    t = torch.tensor([(31 * i + 999) % 1000 for i in range(B)])
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        with torch.no_grad():
            out = model(x.to(DEV), t.to(DEV), c_text=["a"] * B, c_pc_xyz=xyz.to(DEV), c_pc_contact=contact.to(DEV), x_mask=x_mask.to(DEV))
            ref = cmdm_ref.cmdm_forward(_sd(model), x, t, txt, xyz, contact, x_mask)
        valid = ~x_mask
        err = (out.cpu() - ref)[valid].abs().max().item()
        assert err < NET_TOL, err
    finally:
        set_text_feature_provider(None)


def test_cmdm_1000_step_chain_vs_oracle():
    from diffusion.gaussian_diffusion import _draw_seed
    from models.functions import set_text_feature_provider
    from oracle import cmdm_ref, diffusion_ref as D
    B, N, T, Dm, STEPS = 2, 1024, 196, 263, 1000
    model, diff = _mk(cmdm_model_cfg(N), STEPS)
    xyz, contact = synth.scene_points(B, N, seed=32), synth.contact_map(B, N, seed=32)
    x, x_mask = synth.motion_noise(B, T, Dm, seed=32), synth.motion_mask(B, T, seed=32)
    txt = synth.text_features(B, seed=32)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        kw = dict(c_text=["a"] * B, c_pc_xyz=xyz.to(DEV), c_pc_contact=contact.to(DEV), x_mask=x_mask.to(DEV))
        img = x.to(DEV).clone()
        outs = list(diff._fast_loop("ddpm", model, img, kw, 0.0, _draw_seed(), False, True,
                                    step_noise=lambda k: synth.step_noise((B, T, Dm), k).to(DEV)))
        got = outs[-1]["sample"].cpu()
        sd = _sd(model)
        tab = D.make_tables(D.respaced(D.cosine_betas(STEPS), range(STEPS))[0])
        cont = cmdm_ref.contact_tokens(sd, xyz, contact)
        ref = x.clone()
        with torch.no_grad():
            for k in range(STEPS):
                t = torch.full((B,), STEPS - 1 - k, dtype=torch.long)
                x0 = cmdm_ref.cmdm_forward(sd, ref, t, txt, xyz, contact, x_mask, cont_emb=cont)
                ref = D.p_sample_step(tab, x0, ref, t, synth.step_noise((B, T, Dm), k))
        valid = ~x_mask
        err = (got - ref)[valid].abs().max().item()
        assert err < NET_TOL, err
    finally:
        set_text_feature_provider(None)


def test_two_stage_vs_oracle_two_stage():
    """Config 5 semantics at reduced step counts: CDM ddim10 (eta = 0) -> contact hand-off -> CMDM 12 ancestral steps with injected
    noise, vs the oracle running the same two stages on the CPU (utils/evaluate.py:55-66 + datasets/humanml3d.py:773-774 in between)."""
    from amb200.pipeline import contact_from_cdm_sample
    from diffusion.gaussian_diffusion import _draw_seed
    from models.functions import set_text_feature_provider
    from oracle import cdm_ref, cmdm_ref, diffusion_ref as D
    B, N, T, Dm = 2, 1024, 196, 263
    cdm, cdiff = _mk(cdm_model_cfg(N), 500, "ddim10")
    cmdm, mdiff = _mk(cmdm_model_cfg(N), 12)
    xyz = synth.scene_points(B, N, seed=33)
    txt = synth.text_features(B, seed=33)
    x_mask = synth.motion_mask(B, T, seed=33)
    xT = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(33))
    mT = synth.motion_noise(B, T, Dm, seed=33)
    mean, std, sigma = 0.2, 0.3, 0.8
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        # ---- B200
        s = cdiff.ddim_sample_loop(cdm, (B, N, 6), noise=xT.to(DEV), clip_denoised=False, eta=0.0,
                                   model_kwargs=dict(c_text=["a"] * B, c_pc_xyz=xyz.to(DEV), c_pc_feat=None))
        contact = contact_from_cdm_sample(s, mean, std)
        kw = dict(c_text=["a"] * B, c_pc_xyz=xyz.to(DEV), c_pc_contact=contact, x_mask=x_mask.to(DEV))
        outs = list(mdiff._fast_loop("ddpm", cmdm, mT.to(DEV).clone(), kw, 0.0, _draw_seed(), False, True,
                                     step_noise=lambda k: synth.step_noise((B, T, Dm), k).to(DEV)))
        got = outs[-1]["sample"].cpu()
        # ---- oracle
        sdc, sdm = _sd(cdm), _sd(cmdm)
        nb, tmap = D.respaced(D.cosine_betas(500), D.space_timesteps(500, "ddim10"))
        tabc = D.make_tables(nb)
        img = xT.clone()
        with torch.no_grad():
            for i in range(len(tmap) - 1, -1, -1):
                x0 = cdm_ref.cdm_forward(sdc, img, torch.tensor([tmap[i]] * B), txt, xyz)
                img = D.ddim_step(tabc, x0, img, torch.full((B,), i, dtype=torch.long), torch.zeros_like(img), eta=0.0)
            c = (img * std + mean).clamp(1e-20, 1.0)
            dist = torch.sqrt(-2 * torch.log(c) * sigma ** 2)
            c_ref = torch.exp(-0.5 * dist ** 2 / sigma ** 2)
            assert (contact.cpu() - c_ref).abs().max().item() < 1e-3
            tabm = D.make_tables(D.respaced(D.cosine_betas(12), range(12))[0])
            ref = mT.clone()
            for k in range(12):
                t = torch.full((B,), 11 - k, dtype=torch.long)
                x0 = cmdm_ref.cmdm_forward(sdm, ref, t, txt, xyz, c_ref, x_mask)
                ref = D.p_sample_step(tabm, x0, ref, t, synth.step_noise((B, T, Dm), k))
        valid = ~x_mask
        err = (got - ref)[valid].abs().max().item()
        assert err < NET_TOL, err
    finally:
        set_text_feature_provider(None)
