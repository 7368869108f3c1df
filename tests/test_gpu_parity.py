"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every comparison is CUDA path (through the C-ABI) vs
the CPU oracle / reference-generated golden fixtures on the same seeded inputs.
Tolerances: bit-exact for index work (FPS / kNN); 1e-3 max-abs fp32 (north_star) for network outputs, with much
tighter bounds where the arithmetic is elementwise."""
import json
import os

import numpy as np
import pytest
import torch

from amb200 import lib, ops, synth
from amb200.config import cdm_model_cfg, cmdm_model_cfg, full_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NET_TOL = 1e-3  # BASELINE.json north_star: within 1e-3 max-abs on fp32


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _cuda(t):
    return t.to(DEV)


@pytest.fixture(scope="module")
def cmdm_models():
    from models.base import create_model_and_diffusion
    out = {}
    for N in (1024, 8192):
        model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N)), device=DEV)
        sd = synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0)
        model.load_state_dict(sd, strict=False)
        out[N] = (model.to(DEV).eval(), diff)
    return out


def test_device_is_sm100():
    lib.check(lib.load().am_check_device(), "am_check_device")


# ------------------------------------------------------------------ sampler / loss kernels vs reference fixtures
def test_sampler_kernels_match_reference(golden_dir):
    from models.base import create_gaussian_diffusion
    g = _g(golden_dir, "diffusion_steps.npz")
    x_t, x0h, noise = (_cuda(torch.from_numpy(g[k])) for k in ("x_t", "x0h", "noise"))
    x_mask = _cuda(torch.from_numpy(g["x_mask"]))
    B = x_t.shape[0]
    d = create_gaussian_diffusion(full_cfg(cmdm_model_cfg(), steps=1000))
    dd = create_gaussian_diffusion(full_cfg(cmdm_model_cfg(), steps=1000, timestep_respacing="ddim100"))
    dummy = lambda x, t, **kw: x0h
    for tv in (999, 500, 1, 0):
        out = d.p_sample(dummy, x_t, torch.full((B,), tv, device=DEV), clip_denoised=False, noise=noise)["sample"]
        assert np.abs(out.cpu().numpy() - g[f"p_sample_t{tv}"]).max() < 2e-6
    tmix = _cuda(torch.from_numpy(g["t_mixed"]))
    out = d.p_sample(dummy, x_t, tmix, clip_denoised=False, noise=noise)["sample"]
    assert np.abs(out.cpu().numpy() - g["p_sample_mixed"]).max() < 2e-6
    seen = {}

    def dummy2(x, t, **kw):
        seen["t"] = t.clone()
        return x0h
    for tv in (99, 50, 1, 0):
        out = dd.ddim_sample(dummy2, x_t, torch.full((B,), tv, device=DEV), clip_denoised=False, eta=0.0, noise=noise)["sample"]
        assert np.abs(out.cpu().numpy() - g[f"ddim_t{tv}"]).max() < 5e-5
        assert (seen["t"].cpu().numpy() == g[f"ddim_model_t{tv}"]).all()
    out = dd.ddim_sample(dummy, x_t, torch.full((B,), 50, device=DEV), clip_denoised=False, eta=0.5, noise=noise)["sample"]
    assert np.abs(out.cpu().numpy() - g["ddim_eta05_t50"]).max() < 5e-5
    out = d.q_sample(x0h, tmix, noise=noise)
    assert np.abs(out.cpu().numpy() - g["q_sample_mixed"]).max() < 2e-6
    terms = d.training_losses(dummy, x_t, tmix, model_kwargs={"x_mask": x_mask}, noise=noise)
    np.testing.assert_allclose(terms["loss"].cpu().numpy(), g["loss_mixed"], rtol=2e-5)
    np.testing.assert_allclose(terms["mse"].cpu().numpy(), g["mse_mixed"], rtol=2e-5)
    # posterior mean/variance API
    pm = d.p_mean_variance(dummy, x_t, tmix, clip_denoised=False)
    assert pm["mean"].shape == x_t.shape and pm["log_variance"].shape == x_t.shape


def test_philox_normal_statistics_and_rank_invariance():
    n = 196 * 263
    a = torch.empty(8, n, device=DEV)
    ops.randn_(a, n, 8, 0, 1234, 7)
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1.0) < 5e-3
    assert abs((a ** 3).mean().item()) < 2e-2 and abs((a ** 4).mean().item() - 3.0) < 5e-2
    # global sample index addressing: samples 4..7 generated alone equal rows 4..7 of the full batch
    b = torch.empty(4, n, device=DEV)
    ops.randn_(b, n, 4, 4, 1234, 7)
    assert torch.equal(a[4:], b)
    c = torch.empty(8, n, device=DEV)
    ops.randn_(c, n, 8, 0, 1234, 8)
    assert not torch.equal(a, c)


# ------------------------------------------------------------------ pointops: bit-exact index parity
@pytest.mark.parametrize("B,N,stride", [(3, 1024, 4), (2, 8192, 4), (2, 2048, 4), (1, 20000, 4)])
def test_fps_bit_exact(B, N, stride):
    from oracle import pointops_ref
    xyz = synth.scene_points(B, N, seed=31, dup_frac=0.05).reshape(B * N, 3)
    o = torch.tensor([N * (i + 1) for i in range(B)], dtype=torch.int32)
    no = torch.tensor([(N // stride) * (i + 1) for i in range(B)], dtype=torch.int32)
    ref = pointops_ref.furthestsampling(xyz, o, no)
    got = ops.furthestsampling(_cuda(xyz), _cuda(o), _cuda(no), n_max=N, m_total=B * (N // stride))
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("k", [3, 8, 16])
def test_knn_bit_exact_and_properties(k):
    from oracle import pointops_ref
    B, N = 3, 2048
    xyz = synth.scene_points(B, N, seed=32, dup_frac=0.05).reshape(B * N, 3)
    o = torch.tensor([N, 2 * N, 3 * N], dtype=torch.int32)
    no = torch.tensor([N // 4, 2 * (N // 4), 3 * (N // 4)], dtype=torch.int32)
    fidx = pointops_ref.furthestsampling(xyz, o, no)
    q = xyz[fidx.long()].contiguous()
    ri, rd = pointops_ref.knnquery(k, xyz, q, o, no)
    gi, gd2 = ops.knnquery(k, _cuda(xyz), _cuda(q), _cuda(o), _cuda(no))
    assert torch.equal(gi.cpu(), ri)
    assert torch.equal(torch.sqrt(gd2.cpu()), rd)  # distances bit-equal (sqrt taken on the same side)
    assert (np.diff(gd2.cpu().numpy(), axis=1) >= 0).all()
    # self-kNN: nearest neighbour of a point is at distance 0
    si, sd = ops.knnquery(k, _cuda(xyz), _cuda(xyz), _cuda(o), _cuda(o))
    assert (sd[:, 0] == 0).all()
    r2, _ = pointops_ref.knnquery(k, xyz, xyz, o, o)
    assert torch.equal(si.cpu(), r2)


def test_knn_ragged_and_short_segments():
    from oracle import pointops_ref
    xyz = synth.scene_points(1, 300, seed=33).reshape(300, 3)
    o = torch.tensor([5, 40, 300], dtype=torch.int32)
    ri, rd = pointops_ref.knnquery(8, xyz, xyz, o, o)
    gi, gd2 = ops.knnquery(8, _cuda(xyz), _cuda(xyz), _cuda(o), _cuda(o))
    assert torch.equal(gi.cpu(), ri) and torch.equal(torch.sqrt(gd2.cpu()), rd)
    no = torch.tensor([2, 10, 70], dtype=torch.int32)
    rf = pointops_ref.furthestsampling(xyz, o, no)
    gf = ops.furthestsampling(_cuda(xyz), _cuda(o), _cuda(no), n_max=260, m_total=70)
    assert torch.equal(gf.cpu(), rf)


def test_pointops_module_api_matches_golden(golden_dir):
    from models.scene_models import pointops
    g = _g(golden_dir, "cmdm_b3_n1024.npz")
    B, N = 3, 1024
    xyz = _cuda(synth.scene_points(B, N, seed=21, dup_frac=0.05).reshape(B * N, 3))
    o = _cuda(torch.tensor([N, 2 * N, 3 * N], dtype=torch.int32))
    no = _cuda(torch.tensor([N // 4, 2 * (N // 4), 3 * (N // 4)], dtype=torch.int32))
    fidx = pointops.furthestsampling(xyz, o, no)
    assert (fidx.cpu().numpy() == g["fps_idx"]).all()
    kidx, kd = pointops.knnquery(16, xyz, xyz[fidx.long()].contiguous(), o, no)
    assert (kidx.cpu().numpy() == g["knn_idx"]).all()
    assert np.abs(kd.cpu().numpy() - g["knn_dist"]).max() < 1e-6  # sqrt taken on the GPU here
    grp = pointops.queryandgroup(16, xyz, xyz[fidx.long()].contiguous(), xyz, None, o, no, use_xyz=True)
    assert grp.shape == (B * N // 4, 16, 6)


# ------------------------------------------------------------------ dense building blocks vs torch fp32 reference
@pytest.mark.parametrize("M,N,K", [(10432, 1536, 512), (392, 263, 512), (5, 7, 3), (392, 512, 263), (130, 1024, 512),
                                   (16, 512, 512), (16, 257, 64), (128, 512, 512), (33, 96, 130), (2, 1536, 512)])  # last five: small-M kernel
def test_linear_f32_vs_torch(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    for act, ref_act in ((None, lambda v: v), ("gelu", torch.nn.functional.gelu), ("silu", torch.nn.functional.silu),
                         ("relu", torch.relu)):
        y = torch.empty(M, N, device=DEV)
        ops.linear(_cuda(x), _cuda(w), y, M, N, K, bias=_cuda(b), act=act, residual=_cuda(r))
        ref = ref_act(x.double() @ w.double().T + b.double()) + r.double()
        assert (y.cpu().double() - ref).abs().max() < 2e-5
    y = torch.empty(M, N, device=DEV)
    ops.linear(_cuda(x), _cuda(w), y, M, N, K, bias=_cuda(b), act="relu_after_res", residual=_cuda(r))
    assert (y.cpu().double() - torch.relu(x.double() @ w.double().T + b.double() + r.double())).abs().max() < 2e-5


@pytest.mark.parametrize("M,N,K", [(10432, 1536, 512), (10432, 512, 1024), (392, 263, 512), (300, 512, 263), (130, 96, 32), (64, 32, 64)])
def test_linear_tc_vs_fp64(M, N, K):
    """tcgen05 3-term bf16-split GEMM: fp32-class accuracy (error budget 2^-16 relative per product)."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    Kp = ops.pad32(K)
    a2, w2 = ops.split_bf16(_cuda(x), M, K), ops.split_bf16(_cuda(w), N, K)
    # split round trip: hi + lo reproduces the fp32 value to ~2^-17
    a2c = a2.float().cpu()
    assert (a2c[:, :K] + a2c[:, Kp:Kp + K] - x).abs().max() < 2e-5 * x.abs().max()
    for act, ref_act in ((None, lambda v: v), ("gelu", torch.nn.functional.gelu)):
        y = torch.full((M, N), float("nan"), device=DEV)
        y2 = torch.full((M, 2 * ops.pad32(N)), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.linear_tc(a2, w2, M, N, Kp, y=y, y2=y2, bias=_cuda(b), act=act, residual=_cuda(r))
        ref = ref_act(x.double() @ w.double().T + b.double()) + r.double()
        err = (y.cpu().double() - ref).abs().max().item()
        assert err < 1e-4, err
        Np = ops.pad32(N)
        y2c = y2.float().cpu()
        assert (y2c[:, :N] + y2c[:, Np:Np + N] - y.cpu()).abs().max() < 1e-4
        assert (y2c[:, N:Np] == 0).all() and (y2c[:, Np + N:] == 0).all()
    # agreement with the fp32 SIMT kernel (same contract)
    y1 = torch.empty(M, N, device=DEV)
    ops.linear(_cuda(x), _cuda(w), y1, M, N, K, bias=_cuda(b), act="relu_after_res", residual=_cuda(r))
    y3 = torch.empty(M, N, device=DEV)
    ops.linear_tc(a2, w2, M, N, Kp, y=y3, bias=_cuda(b), act="relu_after_res", residual=_cuda(r))
    assert (y1 - y3).abs().max() < 1e-4


def test_linear_small_m_strided_views_and_maps():
    """The CDM latent-side calls: strided x / w views (per-head column slices), unaligned weight rows, token row maps on both
    sides — all through the small-M kernel (M = 2B rows)."""
    g = torch.Generator().manual_seed(9)
    B, R, C, hd, DL = 8, 16, 256, 64, 512
    M2 = 2 * B
    Q = torch.randn(M2, DL, generator=g)
    kfold = torch.randn(8, C + 1, hd, generator=g)
    QF = torch.zeros(B, R, C + 4)
    want = QF.clone()
    Qd, kd, QFd = _cuda(Q), _cuda(kfold), _cuda(QF)
    for h in range(8):  # amb200.cdm_engine: qf[b, 2h+l, :C+1] = kfold[h] q_{h,l}
        ops.linear(Qd[:, h * hd:], kd[h], QFd, M2, C + 1, hd, ldx=DL, ldy=C + 4, ymap=(2, R, 2 * h))
        y = Q[:, h * hd:(h + 1) * hd].double() @ kfold[h].double().T          # [M2, C+1]
        want.view(B, R, C + 4)[:, 2 * h:2 * h + 2, :C + 1] = y.view(B, 2, C + 1).float()
    assert (QFd.cpu() - want).abs().max() < 2e-5
    # xmap + column-offset output + unaligned (odd ldw) weight rows
    Z = torch.randn(B, R, C, generator=g)
    Wv = torch.randn(DL, C + 1, generator=g)[:, :C]                              # ldw = C + 1: rows not 16-byte aligned
    AO = torch.zeros(M2, DL)
    Zd, Wd, AOd = _cuda(Z), _cuda(Wv.contiguous()), _cuda(AO)
    Wodd = _cuda(torch.randn(DL, C + 1, generator=torch.Generator().manual_seed(10)))
    for h in range(8):
        ops.linear(Zd, Wodd[h * hd:(h + 1) * hd], AOd[:, h * hd:], M2, hd, C, ldw=C + 1, ldy=DL, xmap=(2, R, 2 * h))
        zin = Z[:, 2 * h:2 * h + 2, :].reshape(M2, C).double()
        AO[:, h * hd:(h + 1) * hd] = (zin @ Wodd[h * hd:(h + 1) * hd, :C].cpu().double().T).float()
    assert (AOd.cpu() - AO).abs().max() < 2e-5


def test_linear_tc_row_maps():
    B, T, S, off, K, N = 3, 196, 326, 130, 64, 96
    g = torch.Generator().manual_seed(0)
    x, w = torch.randn(B * T, K, generator=g), torch.randn(N, K, generator=g)
    pe = torch.randn(T, N, generator=g)
    a2, w2 = ops.split_bf16(_cuda(x), B * T, K), ops.split_bf16(_cuda(w), N, K)
    y = torch.zeros(B, S, N, device=DEV)
    ops.linear_tc(a2, w2, B * T, N, 64, y=y, residual=_cuda(pe), ldr=N, res_mod=T, ymap=(T, S, off))
    ref = torch.zeros(B, S, N)
    ref[:, off:] = (x.double() @ w.double().T).float().view(B, T, N) + pe
    assert (y.cpu() - ref).abs().max() < 2e-5 * ref.abs().max()  # |y| ~ 10 here (unnormalised weights)
    # inverse (skip) map: GEMM over all [B,S] rows, only rows s >= off are written to a [B,T] output
    xs = torch.randn(B * S, K, generator=g)
    a3 = ops.split_bf16(_cuda(xs), B * S, K)
    out = torch.full((B, T, N), 7.0, device=DEV)
    ops.linear_tc(a3, w2, B * S, N, 64, y=out, ymap=(S, T, -off))
    ref2 = (xs.double() @ w.double().T).float().view(B, S, N)[:, off:]
    assert (out.cpu() - ref2).abs().max() < 2e-5 * ref2.abs().max()


def test_linear_row_maps():
    B, T, S, off, K, N = 3, 5, 9, 2, 16, 8
    g = torch.Generator().manual_seed(0)
    x, w = torch.randn(B, S, K, generator=g), torch.randn(N, K, generator=g)
    pe = torch.randn(T, N, generator=g)
    y = torch.zeros(B, S, N, device=DEV)
    ops.linear(_cuda(x), _cuda(w), y, B * T, N, K, residual=_cuda(pe), ldr=N, res_mod=T, xmap=(T, S, off), ymap=(T, S, off))
    ref = torch.zeros(B, S, N)
    ref[:, off:off + T] = x[:, off:off + T] @ w.T + pe
    assert (y.cpu() - ref).abs().max() < 1e-5


def test_layernorm_and_attention_vs_torch():
    g = torch.Generator().manual_seed(1)
    M, D = 777, 512
    x, r = torch.randn(M, D, generator=g) * 3, torch.randn(M, D, generator=g)
    gam, bet = torch.randn(D, generator=g), torch.randn(D, generator=g)
    y = torch.empty(M, D, device=DEV)
    ops.layernorm(_cuda(x), _cuda(gam), _cuda(bet), y, M, D, residual=_cuda(r))
    ref = torch.nn.functional.layer_norm((x + r).double(), (D,), gam.double(), bet.double(), 1e-5)
    assert (y.cpu().double() - ref).abs().max() < 2e-5
    for B, S in ((3, 326), (2, 2), (1, 212)):
        H, hd = 8, 64
        qkv = torch.randn(B, S, 3 * H * hd, generator=g)
        pad = torch.zeros(B, S, dtype=torch.bool)
        if S > 50:
            pad[0, S - 37:] = True
            pad[-1, 1] = True
        out = torch.empty(B, S, H * hd, device=DEV)
        ops.mha_fwd(_cuda(qkv), out, _cuda(pad.to(torch.uint8)), B, S, H, hd, hd ** -0.5)
        q, k, v = (t.view(B, S, H, hd).transpose(1, 2).double() for t in qkv.split(H * hd, -1))
        sc = (q @ k.transpose(-1, -2)) * hd ** -0.5
        sc = sc.masked_fill(pad[:, None, None, :], float("-inf"))
        ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B, S, H * hd)
        assert (out.cpu().double() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("B,S", [(3, 326), (2, 212), (1, 130), (2, 384), (1, 40), (2, 300), (2, 256), (1, 128), (40, 326)])
def test_mha_tc_vs_fp64(B, S):
    """tcgen05 attention (P in TMEM, 3-term bf16 split) vs an fp64 softmax(QK^T/8 + mask)V reference."""
    g = torch.Generator().manual_seed(S)
    H, hd = 8, 64
    qkv = torch.randn(B * S, 3 * H * hd, generator=g)
    pad = torch.zeros(B, S, dtype=torch.bool)
    if S > 50:
        pad[0, S - 37:] = True
        pad[-1, 1] = True
        pad[-1, 5:9] = True
    qkv2 = ops.split_bf16(_cuda(qkv), B * S, 3 * H * hd)
    out = torch.full((B * S, H * hd), float("nan"), device=DEV)
    out2 = torch.zeros(B * S, 2 * H * hd, dtype=torch.bfloat16, device=DEV)
    ops.mha_tc_fwd(qkv2, out, out2, _cuda(pad.to(torch.uint8)), B, S, H, hd, hd ** -0.5)
    q, k, v = (t.view(B, S, H, hd).transpose(1, 2).double() for t in qkv.view(B, S, -1).split(H * hd, -1))
    sc = (q @ k.transpose(-1, -2)) * hd ** -0.5
    sc = sc.masked_fill(pad[:, None, None, :], float("-inf"))
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, H * hd)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 5e-5, err
    o2 = out2.float().cpu()
    assert (o2[:, :H * hd] + o2[:, H * hd:] - out.cpu()).abs().max() < 1e-4
    # agreement with the fp32 SIMT attention kernel (S <= 352 there)
    if S <= 352:
        o1 = torch.empty(B, S, H * hd, device=DEV)
        ops.mha_fwd(_cuda(qkv.view(B, S, -1)), o1, _cuda(pad.to(torch.uint8)), B, S, H, hd, hd ** -0.5)
        assert (o1.view(B * S, -1) - out).abs().max() < 5e-5


@pytest.mark.parametrize("B,S,q0", [(3, 326, 130), (2, 212, 16), (2, 384, 255), (1, 130, 129), (33, 326, 130)])
def test_mha_tc_query_row_window(B, S, q0):
    """am_mha_tc_fwd_rows (query rows [q0, S) only, compact output) == the same rows of the full kernel, bit for bit."""
    g = torch.Generator().manual_seed(S + q0)
    H, hd = 8, 64
    qkv = torch.randn(B * S, 3 * H * hd, generator=g)
    pad = torch.zeros(B, S, dtype=torch.bool)
    pad[0, S - 37:] = True
    pad[-1, 5:9] = True
    pad_d = _cuda(pad.to(torch.uint8))
    qkv2 = ops.split_bf16(_cuda(qkv), B * S, 3 * H * hd)
    full32 = torch.empty(B * S, H * hd, device=DEV)
    full2 = torch.zeros(B * S, 2 * H * hd, dtype=torch.bfloat16, device=DEV)
    ops.mha_tc_fwd(qkv2, full32, full2, pad_d, B, S, H, hd, hd ** -0.5)
    So = S - q0
    win32 = torch.full((B * So, H * hd), float("nan"), device=DEV)
    win2 = torch.full((B * So, 2 * H * hd), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.mha_tc_fwd(qkv2, win32, win2, pad_d, B, S, H, hd, hd ** -0.5, q_row0=q0)
    assert torch.equal(win32.view(B, So, -1), full32.view(B, S, -1)[:, q0:])
    assert torch.equal(win2.view(B, So, -1), full2.view(B, S, -1)[:, q0:])


# ------------------------------------------------------------------ network parity vs reference-generated goldens
@pytest.mark.parametrize("N", [1024, 8192])
def test_cmdm_forward_matches_reference(golden_dir, cmdm_models, N):
    from models.functions import set_text_feature_provider
    g = _g(golden_dir, f"cmdm_b3_n{N}.npz")
    model, diff = cmdm_models[N]
    B, T, Dm = 3, 196, 263
    xyz = synth.scene_points(B, N, seed=21, dup_frac=0.05)
    contact = synth.contact_map(B, N, seed=21)
    x = synth.motion_noise(B, T, Dm, seed=21)
    x_mask = synth.motion_mask(B, T, seed=21)
    txt = synth.text_features(B, seed=21)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        cont = model.engine.scene.forward(_cuda(xyz), _cuda(contact))
        assert np.abs(cont.cpu().numpy() - g["contact_tokens"]).max() < 1e-4
        kw = dict(c_text=["a"] * B, c_pc_xyz=_cuda(xyz), c_pc_contact=_cuda(contact), x_mask=_cuda(x_mask))
        valid = (~x_mask).numpy()
        with torch.no_grad():
            for tag in ("a", "b"):
                out = model(_cuda(x), _cuda(torch.from_numpy(g[f"t_{tag}"])), **kw)
                err = np.abs(out.cpu().numpy() - g[f"out_{tag}"])[valid].max()
                assert err < NET_TOL, err
            if N == 1024:
                er = torch.tensor([[True], [False], [True]])
                mk = torch.tensor([[False], [True], [True]])
                out = model(_cuda(x), _cuda(torch.tensor([10, 20, 30])), c_text_erase=_cuda(er), c_pc_erase=_cuda(mk),
                            c_text_mask=_cuda(mk), c_pc_mask=_cuda(er), **kw)
                assert np.abs(out.cpu().numpy() - g["out_erase"])[valid].max() < NET_TOL
                # 6-step ancestral chain with injected noise, through diffusion.p_sample
                img = _cuda(x)
                for si, tv in enumerate(g["chain_t"].tolist()):
                    nz = _cuda(synth.step_noise(tuple(img.shape), si))
                    img = diff.p_sample(model, img, torch.full((B,), tv, device=DEV), clip_denoised=False, model_kwargs=kw,
                                        noise=nz)["sample"]
                assert np.abs(img.cpu().numpy() - g["chain_out"])[valid].max() < NET_TOL
    finally:
        set_text_feature_provider(None)


def test_cdm_forward_matches_reference(golden_dir):
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    g = _g(golden_dir, "cdm_b2_n1024.npz")
    B, N = 2, 1024
    model, diff = create_model_and_diffusion(full_cfg(cdm_model_cfg(N), steps=500), device=DEV)
    sd = synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0)
    model.load_state_dict(sd, strict=False)
    model.to(DEV).eval()
    xyz = synth.scene_points(B, N, seed=11)
    x = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(11))
    txt = synth.text_features(B, seed=11)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        with torch.no_grad():
            for tag in ("a", "b"):
                out = model(_cuda(x), _cuda(torch.from_numpy(g[f"t_{tag}"])), c_text=["a"] * B, c_pc_xyz=_cuda(xyz), c_pc_feat=None)
                err = np.abs(out.cpu().numpy() - g[f"out_{tag}"]).max()
                assert err < NET_TOL, err
    finally:
        set_text_feature_provider(None)


# ------------------------------------------------------------------ full-size properties (BASELINE shapes)
def test_cmdm_full_size_sampling_properties(cmdm_models):
    """B=32, T=196, N=8192 (BASELINE config 2 shapes, 12 steps): graph replay == eager, determinism under a fixed
    seed, batch-shard invariance of conditioning (sample b depends only on sample b), finite outputs."""
    from models.functions import set_text_feature_provider
    from models.base import create_gaussian_diffusion
    model, _ = cmdm_models[8192]
    diff = create_gaussian_diffusion(full_cfg(cmdm_model_cfg(8192), steps=12))
    B, T, Dm, N = 32, 196, 263, 8192
    xyz, contact = _cuda(synth.scene_points(B, N, seed=41)), _cuda(synth.contact_map(B, N, seed=41))
    x_mask = _cuda(synth.motion_mask(B, T, seed=41))
    txt = synth.text_features(B, seed=41)
    texts = [f"t{i}" for i in range(B)]
    set_text_feature_provider(lambda raw: torch.stack([txt[int(s[1:])] for s in raw]))
    try:
        kw = dict(c_text=texts, c_pc_xyz=xyz, c_pc_contact=contact, x_mask=x_mask)
        torch.manual_seed(7)
        a = diff.p_sample_loop(model, (B, T, Dm), clip_denoised=False, model_kwargs=kw)
        torch.manual_seed(7)
        b = diff.p_sample_loop(model, (B, T, Dm), clip_denoised=False, model_kwargs=kw)
        assert torch.isfinite(a).all() and torch.equal(a, b)
        # eager (no graph) must equal the graph-replayed loop bit for bit
        torch.manual_seed(7)
        from diffusion.gaussian_diffusion import _draw_seed
        seed = _draw_seed()
        img = torch.empty(B, T, Dm, device=DEV)
        ops.randn_(img, T * Dm, B, 0, seed, 0xFFFFFFFF)
        outs = list(diff._fast_loop("ddpm", model, img, kw, 0.0, seed, False, True, use_graph=False))
        assert torch.equal(outs[-1]["sample"], a)
        # shard invariance: the first 4 samples sampled alone (as a rank owning samples 0..3 would)
        kw4 = dict(c_text=texts[:4], c_pc_xyz=xyz[:4].contiguous(), c_pc_contact=contact[:4].contiguous(), x_mask=x_mask[:4].contiguous())
        torch.manual_seed(7)
        c = diff.p_sample_loop(model, (4, T, Dm), clip_denoised=False, model_kwargs=kw4)
        valid = ~x_mask[:4]
        assert (c - a[:4])[valid].abs().max() < 1e-4
    finally:
        set_text_feature_provider(None)
