"""GPU tests of the training path: every autograd Function (forward + backward kernels) against torch CPU autograd of the
same op, and one full CMDM training step (diffusion.training_losses -> loss.mean().backward()) against the step the
reference's own modules executed (golden) and the oracle's gradients for EVERY parameter."""
import os

import numpy as np
import pytest
import torch

from amb200 import synth
from amb200.config import cmdm_model_cfg, full_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, tol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1e-6) + 1e-7


def test_linear_layernorm_gelu_silu_functions():
    from amb200 import autograd_ops as A
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 5, 48, generator=g)
    w, b = torch.randn(29, 48, generator=g) / 7, torch.randn(29, generator=g)
    gam, bet = torch.randn(29, generator=g), torch.randn(29, generator=g)
    up = torch.randn(37, 5, 29, generator=g)

    def run(dev, lin, ln, gelu, silu):
        xs, ws, bs, gs, bts = (t.clone().to(dev).requires_grad_(True) for t in (x, w, b, gam, bet))
        y = silu(gelu(ln(lin(xs, ws, bs), gs, bts)))
        (y * up.to(dev)).sum().backward()
        return [y] + [t.grad for t in (xs, ws, bs, gs, bts)]
    ref = run("cpu", torch.nn.functional.linear, lambda v, g_, b_: torch.nn.functional.layer_norm(v, (29,), g_, b_, 1e-5),
              torch.nn.functional.gelu, torch.nn.functional.silu)
    got = run(DEV, A.linear, lambda v, g_, b_: A.LayerNormFn.apply(v, g_, b_, 1e-5), A.GeluFn.apply, A.SiluFn.apply)
    for r, o in zip(ref, got):
        assert _close(o, r, 2e-5)


def test_attention_function_with_mask():
    from amb200 import autograd_ops as A
    g = torch.Generator().manual_seed(1)
    B, S, H, hd = 2, 70, 8, 64
    qkv = torch.randn(B, S, 3 * H * hd, generator=g)
    pad = torch.zeros(B, S, dtype=torch.bool)
    pad[0, 50:] = True
    pad[1, 3] = True
    up = torch.randn(B, S, H * hd, generator=g)
    q0 = qkv.clone().requires_grad_(True)
    q, k, v = (t.view(B, S, H, hd).transpose(1, 2) for t in q0.split(H * hd, -1))
    sc = (q @ k.transpose(-1, -2)) * hd ** -0.5
    sc = sc.masked_fill(pad[:, None, None, :], float("-inf"))
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B, S, H * hd)
    (ref * up).sum().backward()
    q1 = qkv.clone().to(DEV).requires_grad_(True)
    out = A.AttentionFn.apply(q1, pad.to(torch.uint8).to(DEV), H, 0.0, 0, 0)
    (out * up.to(DEV)).sum().backward()
    assert _close(out, ref, 2e-5) and _close(q1.grad, q0.grad, 5e-5)
    # dropout: deterministic mask given (seed, site); backward uses the same mask (grad is zero where the output did not depend on p)
    q2 = qkv.clone().to(DEV).requires_grad_(True)
    o1 = A.AttentionFn.apply(q2, None, H, 0.3, 123, 5)
    o2 = A.AttentionFn.apply(q2.detach(), None, H, 0.3, 123, 5)
    o3 = A.AttentionFn.apply(q2.detach(), None, H, 0.3, 124, 5)
    assert torch.equal(o1, o2) and not torch.equal(o1, o3)


def test_dropout_function_statistics_and_backward():
    from amb200 import autograd_ops as A
    x = torch.ones(1 << 20, device=DEV, requires_grad=True)
    y = A.DropoutFn.apply(x, 0.1, 7, 3)
    keep = (y > 0).float().mean().item()
    assert abs(keep - 0.9) < 2e-3 and abs(y.mean().item() - 1.0) < 5e-3
    y.sum().backward()
    assert torch.equal(x.grad, y.detach())  # same mask, same 1/(1-p) scale


@pytest.mark.parametrize("M,C,relu", [(4096, 32, True), (777, 3, True), (50000, 256, False), (300, 260, True)])
def test_batchnorm_train_function(M, C, relu):
    from amb200 import autograd_ops as A
    g = torch.Generator().manual_seed(M + C)
    x = torch.randn(M, C, generator=g) * 2 + 0.5
    up = torch.randn(M, C, generator=g)
    bn_ref = torch.nn.BatchNorm1d(C)
    with torch.no_grad():
        bn_ref.weight.copy_(torch.randn(C, generator=g)); bn_ref.bias.copy_(torch.randn(C, generator=g))
    import copy
    bn_gpu = copy.deepcopy(bn_ref).to(DEV)
    bn_ref.train(); bn_gpu.train()
    x0 = x.clone().requires_grad_(True)
    y0 = bn_ref(x0)
    y0 = torch.relu(y0) if relu else y0
    (y0 * up).sum().backward()
    x1 = x.clone().to(DEV).requires_grad_(True)
    y1 = A.bn_train(x1, bn_gpu, relu=relu)
    (y1 * up.to(DEV)).sum().backward()
    assert _close(y1, y0, 5e-5) and _close(x1.grad, x0.grad, 2e-4)
    assert _close(bn_gpu.weight.grad, bn_ref.weight.grad, 2e-4) and _close(bn_gpu.bias.grad, bn_ref.bias.grad, 2e-4)
    assert _close(bn_gpu.running_mean, bn_ref.running_mean, 1e-5) and _close(bn_gpu.running_var, bn_ref.running_var, 1e-5)
    assert int(bn_gpu.num_batches_tracked) == 1


def test_point_group_functions():
    from amb200 import autograd_ops as A
    g = torch.Generator().manual_seed(3)
    n, k, c = 200, 8, 32
    c8 = c // 8
    qkv = torch.randn(n, 3 * c, generator=g)
    pr = torch.randn(n * k, c, generator=g)
    wl = torch.randn(n * k, c8, generator=g)
    idx = torch.randint(0, n, (n, k), generator=g, dtype=torch.int32)
    up = torch.randn(n, c, generator=g)
    # reference (pointtransformer.py:33-37 semantics)
    q0, p0, w0 = (t.clone().requires_grad_(True) for t in (qkv, pr, wl))
    kf, vf, qf = q0[:, c:2 * c], q0[:, 2 * c:], q0[:, :c]
    il = idx.long()
    wfull = kf[il] - qf[:, None, :] + p0.view(n, k, c)
    ws = torch.softmax(w0.view(n, k, c8), dim=1)
    out = ((vf[il] + p0.view(n, k, c)).view(n, k, 8, c8) * ws.unsqueeze(2)).sum(1).view(n, c)
    ((out * up).sum() + (wfull ** 2).sum()).backward()
    q1, p1, w1 = (t.clone().to(DEV).requires_grad_(True) for t in (qkv, pr, wl))
    idd = idx.to(DEV)
    wf = A.PtWFn.apply(q1, p1, idd, k)
    wsm = A.SoftmaxKFn.apply(w1, n, k)
    o1 = A.PtAggFn.apply(q1, p1, wsm, idd, k)
    ((o1 * up.to(DEV)).sum() + (wf ** 2).sum()).backward()
    assert _close(o1, out, 2e-5) and _close(wf.view(n, k, c), wfull, 2e-5)
    assert _close(q1.grad, q0.grad, 5e-5) and _close(p1.grad, p0.grad, 5e-5) and _close(w1.grad, w0.grad, 5e-5)
    # TransitionDown grouping + max-pool
    m, kk, cc = 50, 16, 12
    x = torch.randn(n, cc, generator=g)
    rel = torch.randn(m * kk, 3, generator=g)
    idx2 = torch.randint(0, n, (m, kk), generator=g, dtype=torch.int32)
    W = torch.randn(20, 3 + cc, generator=g)
    x0 = x.clone().requires_grad_(True)
    G0 = torch.cat([rel, x0[idx2.long().view(-1)]], 1)
    Z0 = (G0 @ W.T).view(m, kk, 20)
    o0 = Z0.max(dim=1).values
    (o0 ** 2).sum().backward()
    x1 = x.clone().to(DEV).requires_grad_(True)
    G1 = A.GroupCatFn.apply(rel.to(DEV), x1, idx2.to(DEV))
    o1 = A.MaxPoolKFn.apply(A.linear(G1, W.to(DEV)), m, kk)
    (o1 ** 2).sum().backward()
    assert _close(o1, o0, 2e-5) and _close(x1.grad, x0.grad, 5e-5)


def test_cmdm_training_step_matches_reference_and_oracle(golden_dir):
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    from tests.test_train_oracle_cpu import oracle_train_step
    g = np.load(os.path.join(golden_dir, "cmdm_train_b2_n1024.npz"))
    oloss, ograds, inp = oracle_train_step(golden_dir)
    model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(1024)), device=DEV)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    model.to(DEV).train()
    for mod in model.modules():  # the golden step ran with every dropout probability 0
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    set_text_feature_provider(lambda raw: inp["txt"][: len(raw)])
    try:
        kw = dict(c_text=["a"] * 2, c_pc_xyz=inp["xyz"].to(DEV), c_pc_contact=inp["contact"].to(DEV), x_mask=inp["x_mask"].to(DEV))
        terms = diff.training_losses(model, inp["x0"].to(DEV), inp["t"].to(DEV), model_kwargs=kw, noise=inp["noise"].to(DEV))
        assert set(terms) == {"loss", "mse"}
        terms["loss"].mean().backward()
    finally:
        set_text_feature_provider(None)
    np.testing.assert_allclose(terms["loss"].detach().cpu().numpy(), g["loss"], rtol=5e-5)
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(grads)
    mine = np.array([float(grads[n].norm()) for n in names])
    np.testing.assert_allclose(mine, g["grad_norms"], rtol=3e-3, atol=1e-7)
    gscale = max(float(v.norm()) for v in ograds.values())

    def rel_l2(a, b):
        # ||a-b|| relative to ||b||, with a floor at 1e-6 of the largest gradient norm: e.g. linear_w.5.bias has an exactly
        # zero true gradient (softmax over k is shift-invariant), what both sides compute there is rounding noise
        a, b = a.detach().cpu().double(), b.detach().cpu().double()
        return ((a - b).norm() / (b.norm() + 1e-6 * gscale)).item()
    # every parameter's gradient against the oracle's (relative L2: some encoder gradients are ~1e-6 sums of cancelling
    # terms, where an elementwise max-abs criterion only measures fp32 summation-order noise)
    worst = max((rel_l2(grads[n], ograds[n]), n) for n in names)
    assert worst[0] < 2e-2, worst
    for k in g.files:
        if k.startswith("grad::"):
            assert rel_l2(grads[k[6:]], torch.from_numpy(g[k])) < 2e-2, k
        if k.startswith("buf::"):
            assert _close(dict(model.named_buffers())[k[5:]], torch.from_numpy(g[k]), 1e-4), k
    # one optimiser step on the GPU parameters runs (utils/training.py:48-50,154)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.0)
    opt.step()
    assert all(torch.isfinite(p).all() for p in model.parameters())


def test_cdm_training_step_matches_reference_and_oracle(golden_dir):
    from amb200.config import cdm_model_cfg
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    from tests.test_train_oracle_cpu import oracle_cdm_train_step
    g = np.load(os.path.join(golden_dir, "cdm_train_b2_n1024.npz"))
    oloss, ograds, inp = oracle_cdm_train_step(golden_dir)
    model, diff = create_model_and_diffusion(full_cfg(cdm_model_cfg(1024), steps=500), device=DEV)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    model.to(DEV).train()
    model.arch_cfg["encoder_dropout"] = 0.0  # the golden step ran with attention dropout 0
    model.arch_cfg["decoder_dropout"] = 0.0
    set_text_feature_provider(lambda raw: inp["txt"][: len(raw)])
    try:
        kw = dict(c_text=["a"] * 2, c_pc_xyz=inp["xyz"].to(DEV), c_pc_feat=None)
        terms = diff.training_losses(model, inp["x0"].to(DEV), inp["t"].to(DEV), model_kwargs=kw, noise=inp["noise"].to(DEV))
        terms["loss"].mean().backward()
    finally:
        set_text_feature_provider(None)
    np.testing.assert_allclose(terms["loss"].detach().cpu().numpy(), g["loss"], rtol=5e-5)
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(grads)
    np.testing.assert_allclose(np.array([float(grads[n].norm()) for n in names]), g["grad_norms"], rtol=3e-3, atol=1e-8)
    gscale = max(float(v.norm()) for v in ograds.values())
    for n in names:
        a, b = grads[n].detach().cpu().double(), ograds[n].double()
        assert ((a - b).norm() / (b.norm() + 1e-6 * gscale)).item() < 2e-2, n


@pytest.mark.parametrize("M,K,N", [(2048, 32, 32), (2048, 64, 32), (4100, 96, 160), (2500, 512, 1024), (2048, 36, 44)])
def test_linear_fn_tensor_core_path_vs_fp64(M, K, N):
    """LinearFn's tcgen05 forward / dX / dW (3-term bf16 split) against fp64."""
    from amb200 import autograd_ops as A
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    up = torch.randn(M, N, generator=g)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    yd = xd @ wd.T + bd
    (yd * up.double()).sum().backward()
    x1, w1, b1 = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    assert A._use_tc(M, N, K)
    y1 = A.linear(x1, w1, b1, tc=True)
    (y1 * up.to(DEV)).sum().backward()

    def rel(a, ref):
        return float((a.detach().cpu().double() - ref).norm() / ref.norm())
    assert rel(y1, yd.detach()) < 2e-5
    assert rel(x1.grad, xd.grad) < 2e-5
    assert rel(w1.grad, wd.grad) < 2e-5
    assert rel(b1.grad, bd.grad) < 2e-5


@pytest.mark.parametrize("M,K,N", [(9000, 3, 3), (8500, 3, 64), (8200, 32, 4), (8200, 256, 32), (8300, 8, 8), (8192, 35, 64), (20000, 9, 32),
                                   (8200, 64, 8), (8200, 3, 130)])
def test_linear_fn_tall_skinny_paths(M, K, N):
    """Per-neighbour MLP shapes of the Point-Transformer encoder (n*k rows, 3..32 features on one side): forward, dX and dW go
    through csrc/rowgemm.cu (one thread per row / register-resident narrow side); checked against fp64."""
    from amb200 import autograd_ops as A
    g = torch.Generator().manual_seed(M + K + N)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    up = torch.randn(M, N, generator=g)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    ref = torch.nn.functional.linear(xd, wd, bd)
    (ref * up.double()).sum().backward()
    xs, ws_, bs = (t.clone().to(DEV).requires_grad_(True) for t in (x, w, b))
    y = A.linear(xs, ws_, bs)
    (y * up.to(DEV)).sum().backward()
    assert _close(y, ref, 1e-5)
    assert _close(xs.grad, xd.grad, 1e-5)
    assert _close(ws_.grad, wd.grad, 3e-5)   # fp32 atomics over ~300 row chunks
    assert _close(bs.grad, bd.grad, 3e-5)


def test_training_gradients_run_to_run_reproducibility():
    """The weight-gradient kernels accumulate with fp32 atomics (split-K dW, rowgemm dW, BatchNorm sums): the summation ORDER is not
    fixed, so two runs of the same step are not bit-identical.  This pins how far apart they can be: loss equal to 1e-6 relative (the
    forward's only atomics are the fp64 BatchNorm sums), every gradient within 2e-2 relative L2 of the other run (measured worst case 6.3e-3, on
    a bias whose gradient is a sum of cancelling terms) — i.e. the 2e-2 gate of the oracle comparison is set by this noise, not by a
    formulation difference."""
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    B, N, T, Dm = 2, 1024, 196, 263
    txt = synth.text_features(B, seed=81)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    runs = []
    try:
        for _ in range(2):
            model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(N)), device=DEV)
            model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
            model.to(DEV).train()
            for mod in model.modules():
                if isinstance(mod, torch.nn.Dropout):
                    mod.p = 0.0
                if isinstance(mod, torch.nn.MultiheadAttention):
                    mod.dropout = 0.0
            kw = dict(c_text=["a"] * B, c_pc_xyz=synth.scene_points(B, N, seed=81).to(DEV), c_pc_contact=synth.contact_map(B, N, seed=81).to(DEV),
                      x_mask=synth.motion_mask(B, T, seed=81).to(DEV))
            terms = diff.training_losses(model, synth.motion_noise(B, T, Dm, seed=81).to(DEV), torch.tensor([700, 20], device=DEV), model_kwargs=kw,
                                         noise=synth.motion_noise(B, T, Dm, seed=82).to(DEV))
            terms["loss"].mean().backward()
            runs.append((terms["loss"].detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}))
    finally:
        set_text_feature_provider(None)
    (l0, g0), (l1, g1) = runs
    assert torch.allclose(l0, l1, rtol=1e-6, atol=0)  # BatchNorm batch sums are fp64 atomics: order-dependent only far below fp32
    gscale = max(float(v.norm()) for v in g0.values())
    worst = max((((g0[n] - g1[n]).double().norm() / (g0[n].double().norm() + 1e-6 * gscale)).item(), n) for n in g0)
    print(f"run-to-run gradient difference (relative L2, worst parameter): {worst[0]:.3e} at {worst[1]}")
    assert worst[0] < 2e-2, worst  # measured 6.3e-3 at contact_encoder.enc1.1.transformer2.linear_p.0.bias (a sum of cancelling terms)
