"""CPU check of the rank-collapsed CDM Perceiver algebra (amb200/cdm_fold.py): a torch fp32 emulation of what the CUDA
kernels compute (csrc/perceiver_tc.cu: same constants, same operation order at the granularity that matters) against the
straight, unfolded oracle (oracle/cdm_ref.py, pinned to the reference by tests/golden/cdm_b2_n1024.npz).  This is host
logic only — the product path has no CPU route; the GPU parity tests compare the kernels themselves with the oracle."""
import torch
import torch.nn.functional as F

from amb200 import synth
from amb200.cdm_fold import R, fold_constants
from amb200.config import cdm_model_cfg
from oracle import cdm_ref
from oracle.nn_ref import lin, ln, timestep_embed


def _unpack_upper(p, n):
    T = torch.zeros(n, n)
    iu = torch.triu_indices(n, n)
    T[iu[0], iu[1]] = p
    return T


def emulate(m, K, x, t, text, xyz):
    """fp32 emulation of CDMEngine.forward (collapsed path)."""
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cm = m.contact_model
    d = K["dims"]
    C, KU, KZ, He, hd, Hd, hdd, J = d["C"], d["KU"], d["KZ"], d["He"], d["hd"], d["Hd"], d["hdd"], d["J"]
    B, N, _ = x.shape
    u = torch.cat([x, xyz], -1)
    ut = torch.cat([u, torch.ones(B, N, 1)], -1)                       # [B,N,KU]
    te = timestep_embed(sd, "timestep_embedder", t)[:, 0]
    cmn = "contact_model"
    L = torch.stack([lin(sd, cmn + ".language_adapter", text), lin(sd, cmn + ".time_embedding_adapter", te)], 1)  # [B,2,DL]
    # ---- encoder
    ca = cm.encoder_cross_attn[0].module
    pe = cmn + ".encoder_cross_attn"
    q = lin(sd, pe + ".0.module.attention.q_proj", ln(sd, pe + ".0.module.q_norm", L)) * hd ** -0.5   # [B,2,DL]
    qh = q.view(B, 2, He, hd)
    AEc = torch.einsum("blhk,hnk->bhln", qh, K["e_kfold"]).reshape(B, R, KU + 2)  # row 2h+l
    Te = _unpack_upper(K["e_chol"], KU)
    rstd = torch.rsqrt(((ut @ Te.T) ** 2).sum(-1) + 1e-5)              # [B,N]
    wr = rstd[..., None] * ut                                          # [B,N,KU]
    s = torch.einsum("bnk,brk->brn", wr, AEc[:, :, :KU]) + AEc[:, :, KU:KU + 1]
    p = torch.softmax(s, -1)                                           # [B,R,N]
    w = torch.einsum("brn,bnk->brk", p, wr)                            # [B,R,KU]
    Z = torch.einsum("ck,brk->brc", K["e_ecg"], w) + K["e_beta"]       # [B,R,C]
    att = ca.attention
    Wv = att.v_proj.weight.view(He, hd, C)
    o = torch.einsum("bhlc,hdc->blhd", Z.view(B, He, 2, C), Wv).reshape(B, 2, He * hd) + att.v_proj.bias
    La = lin(sd, pe + ".0.module.attention.o_proj", o) + L
    L = La + cdm_ref._mlp(sd, pe + ".1.module", La)
    for i in range(len(cm.encoder_self_attn)):
        L = cdm_ref.self_layer(sd, f"{cmn}.encoder_self_attn.{i}", L, He)
    # ---- decoder latent side
    dc = cm.decoder_cross_attn[0].module
    pd = cmn + ".decoder_cross_attn.0.module"
    kvn = ln(sd, pd + ".kv_norm", L)
    k_tok, v_tok = lin(sd, pd + ".attention.k_proj", kvn), lin(sd, pd + ".attention.v_proj", kvn)  # [B,2,C]
    AQc = torch.einsum("blhk,hnk->bhln", k_tok.view(B, 2, Hd, hdd), K["d_qfold"]).reshape(B, R, KU + 2)
    NS = d["NS"]
    UU = torch.einsum("blhk,nhk->bhln", v_tok.view(B, 2, Hd, hdd), K["d_ostack"].view(NS, Hd, hdd)).reshape(B, R, NS)
    Uc, MPt, G1up, HPt = UU[..., :C], UU[..., C:2 * C], UU[..., 2 * C:2 * C + KU], UU[..., 2 * C + KU:2 * C + KU + J]
    G1pp = Uc @ Uc.transpose(1, 2) / C
    G1 = torch.zeros(B, KZ, KZ)
    G1[:, :KU, :KU] = K["d_g1uu"]
    G1[:, :KU, KU:] = G1up.transpose(1, 2)
    G1[:, KU:, :KU] = G1up
    G1[:, KU:, KU:] = G1pp
    Mm = torch.cat([K["d_mu"].unsqueeze(0).expand(B, -1, -1), MPt.transpose(1, 2)], 2)   # [B,C,KZ]
    HP = torch.cat([K["d_hu"].unsqueeze(0).expand(B, -1, -1), HPt.transpose(1, 2)], 2)   # [B,J,KZ]
    # ---- decoder point side
    Tq = _unpack_upper(K["d_chol"], KU)
    rq = torch.rsqrt(((ut @ Tq.T) ** 2).sum(-1) + 1e-5)
    sc = rq[..., None] * torch.einsum("bnk,brk->bnr", ut, AQc[:, :, :KU]) + AQc[:, :, KU].unsqueeze(1)   # [B,N,R]
    pr = torch.softmax(sc.view(B, N, Hd, 2), -1).reshape(B, N, R)
    z = torch.cat([ut, pr], -1)                                        # [B,N,KZ]
    var1 = torch.einsum("bni,bij,bnj->bn", z, G1, z)
    r1 = torch.rsqrt(var1 + 1e-5)
    # bf16 (hi|lo) split of both GEMM operands, 3-term product with fp32 accumulation (what tcgen05 computes)
    def split(a):
        hi = a.bfloat16().float()
        return hi, (a - hi).bfloat16().float()
    zh, zl = split(z)
    mh, ml = split(Mm)
    acc = torch.einsum("bnk,bck->bnc", zl, mh) + torch.einsum("bnk,bck->bnc", zh, ml) + torch.einsum("bnk,bck->bnc", zh, mh)
    pre = r1[..., None] * acc + K["d_c1"]
    g = F.gelu(pre)
    return torch.einsum("bnk,bjk->bnj", z, HP) + g @ K["d_wg"][:, :J]


def _model(N):
    from models.base import Model
    import models  # noqa: F401
    m = Model.get("CDM")(cdm_model_cfg(N), device="cpu")
    sd = synth.fill_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=0)
    m.load_state_dict(sd, strict=False)
    return m.eval()


def test_collapsed_cdm_matches_unfolded_oracle():
    B, N = 2, 1024
    m = _model(N)
    K = fold_constants(m)
    assert K["dims"]["cin"] == 9 and K["dims"]["KZ"] == 26
    text = synth.text_features(B, seed=5)
    xyz = synth.scene_points(B, N, seed=5, dup_frac=0.05)
    x = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(3))
    t = torch.tensor([499, 3])
    with torch.no_grad():
        got = emulate(m, K, x, t, text, xyz)
        ref = cdm_ref.cdm_forward({k: v.detach() for k, v in m.state_dict().items()}, x, t, text, xyz)
    err = (got - ref).abs().max().item()
    assert err < 2e-4, err  # budget 1e-3 (BASELINE.json north_star); the fold + bf16-split noise floor is ~1e-5


def test_collapsed_cdm_large_inputs():
    """x_t at t = T-1 is N(0,1) but intermediate DDIM states and scaled scenes reach |u| ~ 5: the LayerNorm quadratic
    forms must stay accurate there."""
    B, N = 2, 512
    m = _model(N)
    K = fold_constants(m)
    text = synth.text_features(B, seed=7)
    xyz = synth.scene_points(B, N, seed=7) * 2.5
    x = 3.0 * torch.randn(B, N, 6, generator=torch.Generator().manual_seed(11))
    t = torch.tensor([250, 0])
    with torch.no_grad():
        got = emulate(m, K, x, t, text, xyz)
        ref = cdm_ref.cdm_forward({k: v.detach() for k, v in m.state_dict().items()}, x, t, text, xyz)
    assert (got - ref).abs().max().item() < 5e-4
