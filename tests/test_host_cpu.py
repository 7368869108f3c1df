"""CPU-only tests: C-ABI library loads and exports the declared boundary, state_dict compatibility with the
reference's key/shape list, diffusion host tables/respacing vs reference fixtures, loud failure without CUDA."""
import json
import os

import numpy as np
import pytest
import torch

from amb200 import lib, synth
from amb200.config import cdm_model_cfg, cmdm_model_cfg, full_cfg


def test_library_exports_declared_symbols():
    L = lib.load()
    declared = lib.declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert set(lib._SIGS) == set(declared)  # every declared entry point has a typed binding
    assert L.am_version() >= 100


def test_state_dict_keys_match_reference(golden_dir):
    from models.base import Model
    import models  # noqa: F401
    keys = json.load(open(os.path.join(golden_dir, "state_keys.json")))
    for name, cfg, kname in (("CDM", cdm_model_cfg(1024), "CDM"), ("CMDM", cmdm_model_cfg(8192), "CMDM"),
                             ("CDM", cdm_model_cfg(1024, use_scene_model=True), "CDM_scene")):
        m = Model.get(name)(cfg, device="cpu")
        mine = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert mine == keys[kname], kname
        name = kname
        # buffers the reference computes deterministically are bit-identical
        sd = synth.fill_state_dict({k: tuple(v) for k, v in keys[name].items()}, seed=0)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith(".pe") for k in missing)


def test_registry_semantics():
    from models.base import Model, Registry
    assert "CDM" in Model and "CMDM" in Model
    with pytest.raises(KeyError):
        Model.get("nope")
    r = Registry("x")

    @r.register()
    class A:  # noqa
        pass
    with pytest.raises(AssertionError):
        r.register(A)


def test_diffusion_tables_and_respacing(golden_dir):
    from models.base import create_gaussian_diffusion
    from diffusion.respace import space_timesteps
    g = np.load(os.path.join(golden_dir, "diffusion_tables.npz"))
    for T in (1000, 500):
        d = create_gaussian_diffusion(full_cfg(cmdm_model_cfg(), steps=T))
        assert d.num_timesteps == T and d.timestep_map == list(range(T))
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                  "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                  "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"):
            np.testing.assert_array_equal(getattr(d, k), g[f"T{T}_{k}"], err_msg=k)
        dd = create_gaussian_diffusion(full_cfg(cmdm_model_cfg(), steps=T, timestep_respacing="ddim100"))
        assert dd.num_timesteps == 100
        np.testing.assert_array_equal(np.array(dd.timestep_map), g[f"T{T}_ddim100_map"])
        np.testing.assert_array_equal(dd.betas, g[f"T{T}_ddim100_betas"])
    np.testing.assert_array_equal(np.array(sorted(space_timesteps(300, [10, 15, 20]))), g["space_300_10_15_20"])
    np.testing.assert_array_equal(np.array(sorted(space_timesteps(1000, "ddim50"))), g["space_1000_ddim50"])
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim333")


def test_uniform_sampling_matches_reference_stream():
    from diffusion.resample import uniform_sampling
    np.random.seed(2023)
    a = uniform_sampling(16, "cpu", 1000)
    np.random.seed(2023)
    w = np.ones([1000])
    b = np.random.choice(1000, size=(16,), p=w / np.sum(w))
    assert (a.numpy() == b).all() and a.dtype == torch.int64


def test_cpu_forward_fails_loudly():
    from models.base import create_model_and_diffusion
    model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(1024)), device="cpu")
    model.eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 196, 263), torch.zeros(1, dtype=torch.long), c_text=["a"], c_pc_xyz=torch.zeros(1, 1024, 3),
              c_pc_contact=torch.zeros(1, 1024, 6))
    from amb200 import ops
    with pytest.raises(lib.AmbError):
        ops.randn_(torch.zeros(8), 8, 1, 0, 0, 0)


def test_text_provider_hook_and_missing_clip():
    from models import functions as F
    F.set_text_feature_provider(None)
    m = F.load_and_freeze_clip_model("ViT-B/32")
    if isinstance(m, F._NoTextModel):
        with pytest.raises(RuntimeError, match="provider"):
            F.encode_text_clip(m, ["a"])
    feats = torch.arange(6.0).view(2, 3)
    F.set_text_feature_provider(lambda raw: feats[: len(raw)])
    assert torch.equal(F.encode_text_clip(m, ["a", "b"]), feats)
    F.set_text_feature_provider(None)
    assert F.get_lang_feat_dim_type("ViT-B/32") == (512, "clip")


def test_adamw_oracle_matches_torch_optim():
    """Pins oracle/optim_ref.py against torch.optim.AdamW (what utils/training.py:48 constructs) on CPU, incl. weight decay."""
    from oracle.optim_ref import adamw_step
    g = torch.Generator().manual_seed(0)
    p0 = [torch.randn(33, 7, generator=g), torch.randn(5, generator=g)]
    params = [torch.nn.Parameter(t.clone()) for t in p0]
    opt = torch.optim.AdamW(params, lr=3e-3, weight_decay=0.02)
    mine = [t.clone() for t in p0]
    m = [torch.zeros_like(t) for t in p0]
    v = [torch.zeros_like(t) for t in p0]
    for step in range(1, 6):
        grads = [torch.randn(t.shape, generator=g) for t in p0]
        for q, gr in zip(params, grads):
            q.grad = gr.clone()
        opt.step()
        for i in range(2):
            adamw_step(mine[i], grads[i], m[i], v[i], step, lr=3e-3, weight_decay=0.02)
            assert torch.allclose(mine[i], params[i].data, rtol=0, atol=1e-7), (step, i)


def test_fused_adamw_and_scene_model_host_behaviour():
    """Host-side contracts that need no GPU: the fused optimiser refuses CPU parameters loudly (no CPU fallback), the scene-model
    factory builds the reference's PointTransformerSeg tree (functions.py:96-126) frozen, and rejects the ablation-only encoder."""
    from amb200.lib import AmbError
    from amb200.optim import FusedAdamW
    from amb200.dist import allreduce_flat_
    from models.functions import load_scene_model
    with pytest.raises(AmbError):
        FusedAdamW([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
    assert allreduce_flat_([torch.ones(3)]) == 1  # no process group: no-op, world size 1
    seg = load_scene_model("PointTransformerSeg", 3, 1024, None, freeze=True)
    keys = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_keys.json")))["PointTransformerSeg_c3"]
    assert {k: list(v.shape) for k, v in seg.state_dict().items()} == keys
    assert not seg.training and not any(p.requires_grad for p in seg.parameters()) and seg.num_groups == 4
    with pytest.raises(NotImplementedError):
        load_scene_model("PointTransformerEnc", 3, 1024)
    with pytest.raises(RuntimeError):  # executing it needs the CUDA library + a GPU: CPU tensors are refused, not emulated
        seg((torch.zeros(1, 1024, 3), None))


def test_gemm_tail_plan_covers_every_tile_once_and_never_lengthens_the_last_round():
    """csrc/gemm_tc.cu::tc_tail_plan (host code of the CTA-pair GEMM): the remainder tiles of the last partial round are cut into
    2^shift column pieces; every tile is covered exactly once and the estimated makespan never exceeds the uncut schedule's."""
    import ctypes
    from amb200 import lib
    L = lib.load()
    f = L.am_tc_tail_plan_
    f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    cost = {0: 1.0, 1: 0.54, 2: 0.30}
    for clusters in (74, 50, 8):
        for pairs in range(1, 400):
            wide, shift = ctypes.c_int(-1), ctypes.c_int(-1)
            assert f(pairs, clusters, ctypes.byref(wide), ctypes.byref(shift)) == 0
            w, sh = wide.value, shift.value
            assert 0 <= w <= pairs and sh in (1, 2)
            assert w == pairs or w % clusters == 0          # whole rounds first
            rem = pairs - w
            items = w + (rem << sh)
            assert items >= pairs
            uncut = -(-pairs // clusters)                   # rounds of whole tiles
            planned = w // clusters + (-(-(rem << sh) // clusters)) * cost[sh] if rem else w / clusters
            assert planned <= uncut + 1e-9, (pairs, clusters, w, sh)
    # the trunk shapes of the headline: 82 tiles (N = 512) -> quarter pieces, 246 tiles (N = 1536) -> halves
    for pairs, want in ((82, 2), (164, 2), (246, 1)):
        wide, shift = ctypes.c_int(), ctypes.c_int()
        f(pairs, 74, ctypes.byref(wide), ctypes.byref(shift))
        assert (wide.value, shift.value) == ((pairs // 74) * 74, want)
