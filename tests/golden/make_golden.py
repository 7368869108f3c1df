"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE's own Python modules.

Run once in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What is pinned by these fixtures: the reference's diffusion tables / sampler / loss math, CDM-Perceiver
forward, CMDM trans_enc forward and the PointTransformer blocks (eval mode) — all executed from
/root/reference with four stub modules (omegaconf, clip, pointops_cuda, smplkit).
What is NOT pinned (source absent offline): `pointops_cuda` FPS/kNN (the oracle's C restatement is
injected in its place, lowest-index tie rule) and CLIP (a seeded [B,512] feature is injected through the
`encode_text_clip` hook).  Weights are NOT stored: both sides rebuild them with
amb200.synth.fill_state_dict(shapes, seed) from the key/shape list saved in state_keys.json.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.path.insert(2, os.path.join(ROOT, "afford-motion_b200"))


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return DictConfig(v) if isinstance(v, dict) else v


def install_stubs():
    om = types.ModuleType("omegaconf")
    om.DictConfig = DictConfig
    sys.modules["omegaconf"] = om
    sys.modules["clip"] = types.ModuleType("clip")
    sys.modules["pointops_cuda"] = types.ModuleType("pointops_cuda")
    sk = types.ModuleType("smplkit")
    sk.SMPLXLayer = lambda **k: None
    sys.modules["smplkit"] = sk


CDM_CFG = dict(
    name="CDM", input_feats=6, data_repr="contact_cont_joints", time_emb_dim=128,
    text_model=dict(version="ViT-B/32", max_length=20),
    scene_model=dict(use_scene_model=False, name="PointTransformerSeg", use_color=False, use_openscene=False,
                     num_points=1024, point_feat_dim=32, pretrained_weight=None, freeze=True),
    arch="Perceiver",
    arch_perceiver=dict(last_dim=256, point_pos_emb=True, encoder_q_input_channels=512, encoder_kv_input_channels=256,
                        encoder_num_heads=8, encoder_widening_factor=1, encoder_dropout=0.1,
                        encoder_residual_dropout=0.0, encoder_self_attn_num_layers=2, decoder_q_input_channels=256,
                        decoder_kv_input_channels=512, decoder_num_heads=8, decoder_widening_factor=1,
                        decoder_dropout=0.1, decoder_residual_dropout=0.0),
)


def cmdm_cfg(num_points):
    return dict(
        name="CMDM", input_feats=263, data_repr="h3d", time_emb_dim=512,
        contact_model=dict(contact_type="contact_cont_joints", contact_joints=[0, 10, 11, 12, 20, 21],
                           planes=[32, 64, 128, 256], num_points=num_points, blocks=[2, 2, 2, 2]),
        text_model=dict(version="ViT-B/32", max_length=20),
        arch="trans_enc", latent_dim=512, mask_motion=True, num_layers=[1, 1, 1, 1, 1], num_heads=8, dropout=0.1,
        dim_feedforward=1024,
    )


ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]  # e.g. `make_golden.py scene_seg` rewrites only matching fixtures


def _save(name, arrs):
    if ONLY and not any(o in name for o in ONLY):
        return
    np.savez_compressed(os.path.join(HERE, name), **arrs)
    print("wrote", name)


def main():
    install_stubs()
    from amb200 import synth
    from oracle import pointops_ref
    # The reference's `diffusion/` has no __init__.py (namespace package), so a regular package of the same name anywhere on
    # sys.path would shadow it: drop this repo's drop-in packages from the path before importing the reference.
    sys.path[:] = [p for p in sys.path if os.path.abspath(p) != os.path.join(ROOT, "afford-motion_b200")]
    for mod in [k for k in sys.modules if k == "diffusion" or k.startswith("diffusion.") or k == "models" or k.startswith("models.")]:
        del sys.modules[mod]

    import models.cdm as rcdm
    import models.cmdm as rcmdm
    import models.scene_models.pointops as rpo
    from diffusion import gaussian_diffusion as gd
    from diffusion.respace import SpacedDiffusion, space_timesteps

    # CPU stand-ins at the unpinned boundaries
    rpo.furthestsampling = pointops_ref.furthestsampling
    rpo.knnquery = pointops_ref.knnquery
    torch.cuda.IntTensor = lambda x: torch.IntTensor(x)  # pointtransformer.py:60

    text_holder = {}

    def fake_encode(model, raw_text, max_length=32, device="cpu"):
        return text_holder["feat"][: len(raw_text)].clone()

    for m in (rcdm, rcmdm):
        m.load_and_freeze_clip_model = lambda v: torch.nn.Module()
        m.encode_text_clip = fake_encode

    keys = {}

    # ---------------------------------------------------------------- diffusion tables + steps
    out = {}
    for T in (1000, 500):
        d = SpacedDiffusion(use_timesteps=space_timesteps(T, [T]), betas=gd.get_named_beta_schedule("cosine", T),
                            model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                            loss_type=gd.LossType.MSE, rescale_timesteps=False)
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                  "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                  "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"):
            out[f"T{T}_{k}"] = getattr(d, k)
        dd = SpacedDiffusion(use_timesteps=space_timesteps(T, "ddim100"), betas=gd.get_named_beta_schedule("cosine", T),
                             model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                             loss_type=gd.LossType.MSE, rescale_timesteps=False)
        out[f"T{T}_ddim100_map"] = np.array(dd.timestep_map)
        out[f"T{T}_ddim100_betas"] = dd.betas
        out[f"T{T}_ddim100_alphas_cumprod"] = dd.alphas_cumprod
    out["linear1000_betas"] = gd.get_named_beta_schedule("linear", 1000)
    out["space_300_10_15_20"] = np.array(sorted(space_timesteps(300, [10, 15, 20])))
    out["space_1000_ddim50"] = np.array(sorted(space_timesteps(1000, "ddim50")))
    _save("diffusion_tables.npz", out)

    # sampler / loss steps with a dummy model that returns a fixed x0_hat
    T = 1000
    d = SpacedDiffusion(use_timesteps=space_timesteps(T, [T]), betas=gd.get_named_beta_schedule("cosine", T),
                        model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                        loss_type=gd.LossType.MSE, rescale_timesteps=False)
    dd = SpacedDiffusion(use_timesteps=space_timesteps(T, "ddim100"), betas=gd.get_named_beta_schedule("cosine", T),
                         model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                         loss_type=gd.LossType.MSE, rescale_timesteps=False)
    B, Tm, D = 4, 48, 67  # elementwise math is shape-agnostic; keep the fixture small
    g = torch.Generator().manual_seed(7)
    x_t = torch.randn(B, Tm, D, generator=g)
    x0h = torch.randn(B, Tm, D, generator=g)
    noise = torch.randn(B, Tm, D, generator=g)
    x_mask = synth.motion_mask(B, Tm, seed=7)
    steps = {"x_t": x_t.numpy(), "x0h": x0h.numpy(), "noise": noise.numpy(), "x_mask": x_mask.numpy()}
    seen_t = {}

    def dummy(x, t, **kw):
        seen_t["t"] = t.clone()
        return x0h

    real_randn_like = torch.randn_like
    torch.randn_like = lambda x: noise  # inject the SAME noise (gaussian_diffusion.py:431,577,766)
    try:
        for tv in (999, 500, 1, 0):
            t = torch.tensor([tv] * B)
            steps[f"p_sample_t{tv}"] = d.p_sample(dummy, x_t, t, clip_denoised=False)["sample"].numpy()
        tmix = torch.tensor([999, 321, 1, 0])
        steps["t_mixed"] = tmix.numpy()
        steps["p_sample_mixed"] = d.p_sample(dummy, x_t, tmix, clip_denoised=False)["sample"].numpy()
        for tv in (99, 50, 1, 0):
            t = torch.tensor([tv] * B)
            steps[f"ddim_t{tv}"] = dd.ddim_sample(dummy, x_t, t, clip_denoised=False, eta=0.0)["sample"].numpy()
            steps[f"ddim_model_t{tv}"] = seen_t["t"].numpy()
        steps["ddim_eta05_t50"] = dd.ddim_sample(dummy, x_t, torch.tensor([50] * B), clip_denoised=False, eta=0.5)["sample"].numpy()
        steps["q_sample_mixed"] = d.q_sample(x0h, tmix, noise=noise).numpy()
        terms = d.training_losses(dummy, x_t, tmix, model_kwargs={"x_mask": x_mask}, noise=noise)
        steps["loss_mixed"] = terms["loss"].numpy()
        steps["mse_mixed"] = terms["mse"].numpy()
    finally:
        torch.randn_like = real_randn_like
    _save("diffusion_steps.npz", steps)

    # ---------------------------------------------------------------- CDM (config 1: B=2, N=1024)
    torch.manual_seed(0)
    cdm = rcdm.CDM(DictConfig(CDM_CFG), device="cpu").eval()
    shapes = {k: tuple(v.shape) for k, v in cdm.state_dict().items()}
    keys["CDM"] = {k: list(v) for k, v in shapes.items()}
    cdm.load_state_dict(synth.fill_state_dict(shapes, seed=0), strict=False)
    B, N = 2, 1024
    xyz = synth.scene_points(B, N, seed=11)
    x = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(11))
    text_holder["feat"] = synth.text_features(B, seed=11)
    res = {}
    with torch.no_grad():
        for tag, tv in (("a", [999, 3]), ("b", [500, 0])):
            t = torch.tensor(tv)
            res[f"t_{tag}"] = t.numpy()
            res[f"out_{tag}"] = cdm(x, t, c_text=["a"] * B, c_pc_xyz=xyz, c_pc_feat=None).numpy()
    _save("cdm_b2_n1024.npz", res)

    # ---------------------------------------------------------------- CMDM (B=3; N=1024 and N=8192)
    for N in (1024, 8192):
        torch.manual_seed(0)
        cm = rcmdm.CMDM(DictConfig(cmdm_cfg(N)), device="cpu").eval()
        shapes = {k: tuple(v.shape) for k, v in cm.state_dict().items()}
        keys["CMDM"] = {k: list(v) for k, v in shapes.items()}
        cm.load_state_dict(synth.fill_state_dict(shapes, seed=0), strict=False)
        B, Tm, D = 3, 196, 263
        xyz = synth.scene_points(B, N, seed=21, dup_frac=0.05)
        contact = synth.contact_map(B, N, seed=21)
        x = synth.motion_noise(B, Tm, D, seed=21)
        x_mask = synth.motion_mask(B, Tm, seed=21)
        text_holder["feat"] = synth.text_features(B, seed=21)
        res = {}
        with torch.no_grad():
            cont = cm.contact_encoder(xyz, contact)
            res["contact_tokens"] = cont.numpy()
            for tag, tv in (("a", [999, 500, 0]), ("b", [7, 7, 7])):
                t = torch.tensor(tv)
                res[f"t_{tag}"] = t.numpy()
                res[f"out_{tag}"] = cm(x, t, c_text=["a"] * B, c_pc_xyz=xyz, c_pc_contact=contact, x_mask=x_mask).numpy()
            if N == 1024:  # erase / mask conditioning switches (cmdm.py:142-155)
                er = torch.tensor([[True], [False], [True]])
                mk = torch.tensor([[False], [True], [True]])
                res["out_erase"] = cm(x, torch.tensor([10, 20, 30]), c_text=["a"] * B, c_pc_xyz=xyz, c_pc_contact=contact,
                                      x_mask=x_mask, c_text_erase=er, c_pc_erase=mk, c_text_mask=mk, c_pc_mask=er).numpy()
                # pointops boundary artefacts (oracle-defined, UNPINNED): stage-2 FPS + kNN of the packed batch
                p0 = xyz.reshape(B * N, 3)
                o = torch.tensor([N, 2 * N, 3 * N], dtype=torch.int32)
                no = torch.tensor([N // 4, 2 * (N // 4), 3 * (N // 4)], dtype=torch.int32)
                fidx = pointops_ref.furthestsampling(p0, o, no)
                kidx, kd = pointops_ref.knnquery(16, p0, p0[fidx.long()], o, no)
                res["fps_idx"] = fidx.numpy()
                res["knn_idx"] = kidx.numpy()
                res["knn_dist"] = kd.numpy()
                # 6-step ancestral chain with injected noise (gaussian_diffusion.py:488-536)
                T = 1000
                d = SpacedDiffusion(use_timesteps=space_timesteps(T, [T]), betas=gd.get_named_beta_schedule("cosine", T),
                                    model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                                    loss_type=gd.LossType.MSE, rescale_timesteps=False)
                img = x.clone()
                kw = dict(c_text=["a"] * B, c_pc_xyz=xyz, c_pc_contact=contact, x_mask=x_mask)
                chain_t = [999, 998, 997, 2, 1, 0]
                for si, tv in enumerate(chain_t):
                    nz = synth.step_noise(img.shape, si)
                    torch.randn_like = lambda a, _n=nz: _n
                    try:
                        img = d.p_sample(cm, img, torch.tensor([tv] * B), clip_denoised=False, model_kwargs=kw)["sample"]
                    finally:
                        torch.randn_like = real_randn_like
                res["chain_t"] = np.array(chain_t)
                res["chain_out"] = img.numpy()
        res["x_mask"] = x_mask.numpy()
        _save(f"cmdm_b3_n{N}.npz", res)

    # ---------------------------------------------------------------- CMDM training step (train mode, dropout p=0), B=2, N=1024
    torch.manual_seed(0)
    cm = rcmdm.CMDM(DictConfig(cmdm_cfg(1024)), device="cpu")
    shapes = {k: tuple(v.shape) for k, v in cm.state_dict().items()}
    cm.load_state_dict(synth.fill_state_dict(shapes, seed=0), strict=False)
    cm.train()
    for mod in cm.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    B, Tm, D = 2, 196, 263
    xyz = synth.scene_points(B, 1024, seed=31, dup_frac=0.05)
    contact = synth.contact_map(B, 1024, seed=31)
    x0 = synth.motion_noise(B, Tm, D, seed=31)
    x_mask = synth.motion_mask(B, Tm, seed=31)
    x_mask[1, 100:] = True
    text_holder["feat"] = synth.text_features(B, seed=31)
    noise = synth.step_noise((B, Tm, D), 77)
    tt = torch.tensor([700, 23])
    T = 1000
    d = SpacedDiffusion(use_timesteps=space_timesteps(T, [T]), betas=gd.get_named_beta_schedule("cosine", T),
                        model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                        loss_type=gd.LossType.MSE, rescale_timesteps=False)
    terms = d.training_losses(cm, x0, tt, model_kwargs=dict(c_text=["a"] * B, c_pc_xyz=xyz, c_pc_contact=contact, x_mask=x_mask), noise=noise)
    loss = terms["loss"].mean()
    loss.backward()
    tr = {"loss": terms["loss"].detach().numpy(), "x_mask": x_mask.numpy(), "t": tt.numpy()}
    norms = {}
    for n_, p_ in cm.named_parameters():
        if p_.grad is not None:
            norms[n_] = float(p_.grad.norm())
    tr["grad_names"] = np.array(sorted(norms))
    tr["grad_norms"] = np.array([norms[k] for k in sorted(norms)])
    for k in ("language_adapter.bias", "motion_layer.bias", "contact_encoder.enc1.0.linear.weight", "contact_encoder.enc2.0.bn.weight",
              "contact_encoder.enc4.1.transformer2.linear_w.5.weight", "contact_encoder.enc1.1.transformer2.linear_p.0.weight",
              "timestep_embedder.time_embed.0.bias", "self_attn_layer.layers.0.norm1.weight", "self_attn_layer.layers.4.self_attn.in_proj_bias",
              "contact_adapter.bias"):
        tr["grad::" + k] = dict(cm.named_parameters())[k].grad.numpy()
    sdt = cm.state_dict()
    for k in ("contact_encoder.enc1.0.bn.running_mean", "contact_encoder.enc1.0.bn.running_var",
              "contact_encoder.enc3.1.transformer2.linear_w.3.running_var"):
        tr["buf::" + k] = sdt[k].numpy()
    _save("cmdm_train_b2_n1024.npz", tr)

    # ---------------------------------------------------------------- CDM training step (train mode, dropout p=0), B=2, N=1024
    torch.manual_seed(0)
    cdm = rcdm.CDM(DictConfig(CDM_CFG), device="cpu")
    shapes = {k: tuple(v.shape) for k, v in cdm.state_dict().items()}
    cdm.load_state_dict(synth.fill_state_dict(shapes, seed=0), strict=False)
    cdm.train()
    for mod in cdm.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    B, N = 2, 1024
    xyz = synth.scene_points(B, N, seed=41)
    x0 = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(41))
    noise = synth.step_noise((B, N, 6), 78)
    text_holder["feat"] = synth.text_features(B, seed=41)
    tt = torch.tensor([400, 7])
    d5 = SpacedDiffusion(use_timesteps=space_timesteps(500, [500]), betas=gd.get_named_beta_schedule("cosine", 500),
                         model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                         loss_type=gd.LossType.MSE, rescale_timesteps=False)
    terms = d5.training_losses(cdm, x0, tt, model_kwargs=dict(c_text=["a"] * B, c_pc_xyz=xyz, c_pc_feat=None), noise=noise)
    terms["loss"].mean().backward()
    tr = {"loss": terms["loss"].detach().numpy(), "t": tt.numpy()}
    norms = {n_: float(p_.grad.norm()) for n_, p_ in cdm.named_parameters() if p_.grad is not None}
    tr["grad_names"] = np.array(sorted(norms))
    tr["grad_norms"] = np.array([norms[k] for k in sorted(norms)])
    for k in ("contact_layer.weight", "contact_model.encoder_adapter.weight", "contact_model.decoder_cross_attn.0.module.attention.q_proj.bias",
              "contact_model.encoder_cross_attn.0.module.kv_norm.weight", "timestep_embedder.time_embed.2.bias"):
        tr["grad::" + k] = dict(cdm.named_parameters())[k].grad.numpy()
    _save("cdm_train_b2_n1024.npz", tr)

    # ---------------------------------------------------------------- frozen PointTransformerSeg scene model (§8 f3) + CDM using it
    import models.scene_models.pointtransformer as rpt
    torch.cuda.FloatTensor = lambda *shape: torch.FloatTensor(*shape)  # pointops.py:175
    B, N = 2, 1024
    xyz = synth.scene_points(B, N, seed=51, dup_frac=0.05)
    color = torch.rand(B, N, 3, generator=torch.Generator().manual_seed(51))
    seg_res = {}
    for cdim in (3, 6):
        torch.manual_seed(0)
        seg = rpt.pointtransformer_seg_repro(c=cdim, num_points=N).eval()
        shapes = {k: tuple(v.shape) for k, v in seg.state_dict().items()}
        keys[f"PointTransformerSeg_c{cdim}"] = {k: list(v) for k, v in shapes.items()}
        seg.load_state_dict(synth.fill_state_dict(shapes, seed=0), strict=False)
        with torch.no_grad():
            seg_res[f"feat_c{cdim}"] = seg((xyz, color)).numpy()
    cfg_s = json.loads(json.dumps(CDM_CFG))
    cfg_s["scene_model"].update(use_scene_model=True, use_color=False, num_points=N)
    torch.manual_seed(0)
    cdm_s = rcdm.CDM(DictConfig(cfg_s), device="cpu").eval()
    shapes = {k: tuple(v.shape) for k, v in cdm_s.state_dict().items()}
    keys["CDM_scene"] = {k: list(v) for k, v in shapes.items()}
    cdm_s.load_state_dict(synth.fill_state_dict(shapes, seed=0), strict=False)
    text_holder["feat"] = synth.text_features(B, seed=51)
    x = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(52))
    tt = torch.tensor([450, 2])
    with torch.no_grad():
        seg_res["cdm_scene_out"] = cdm_s(x, tt, c_text=["a"] * B, c_pc_xyz=xyz, c_pc_feat=color).numpy()
    seg_res["t"] = tt.numpy()
    _save("scene_seg_b2_n1024.npz", seg_res)

    kp = os.path.join(HERE, "state_keys.json")
    if ONLY and os.path.exists(kp):
        keys = {**json.load(open(kp)), **keys}
    with open(kp, "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
