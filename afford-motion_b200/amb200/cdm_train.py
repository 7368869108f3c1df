"""Training-mode forward of CDM / ContactPerceiver (models/cdm.py:155-188,474-513; Perceiver-IO blocks
models/modules.py:222-661) as an autograd graph of libamb200 kernels.  Unlike the sampling engine nothing is folded:
K/V [B,N,512] are materialised so gradients reach k_proj / v_proj / the adapters exactly as in the reference."""
import torch

from . import autograd_ops as A


def _seed() -> int:
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _mlp(mlp, x):
    """modules.py:651-661: LN -> Linear -> GELU -> Linear."""
    h = A.LayerNormFn.apply(x, mlp[0].weight, mlp[0].bias, mlp[0].eps)
    h = A.GeluFn.apply(A.linear(h, mlp[1].weight, mlp[1].bias, tc=True))
    return A.linear(h, mlp[3].weight, mlp[3].bias, tc=True)


def _mha(att, xq, xkv, p_drop, seed, site):
    """modules.py:324-381: q *= c_head^-0.5 (folded into the score scale), no mask / rotary / cache."""
    H = att.num_heads
    q = A.linear(xq, att.q_proj.weight, att.q_proj.bias, tc=True)
    k = A.linear(xkv, att.k_proj.weight, att.k_proj.bias, tc=True)
    v = A.linear(xkv, att.v_proj.weight, att.v_proj.bias, tc=True)
    c = q.shape[-1] // H
    o = A.MHAFn.apply(q, k, v, H, c ** -0.5, p_drop, seed, site)
    return A.linear(o, att.o_proj.weight, att.o_proj.bias, tc=True)


def _cross_layer(layer, xq, xkv, p_drop, seed, site):
    """CrossAttentionLayer (modules.py:504-541) with Residual (:222-231): residual adds the UN-normalised x_q."""
    ca = layer[0].module
    qn = A.LayerNormFn.apply(xq, ca.q_norm.weight, ca.q_norm.bias, ca.q_norm.eps)
    kvn = A.LayerNormFn.apply(xkv, ca.kv_norm.weight, ca.kv_norm.bias, ca.kv_norm.eps)
    x = A.AddFn.apply(_mha(ca.attention, qn, kvn, p_drop, seed, site), xq, False)
    return A.AddFn.apply(_mlp(layer[1].module, x), x, False)


def _self_layer(layer, x, p_drop, seed, site):
    sa = layer[0].module
    n = A.LayerNormFn.apply(x, sa.norm.weight, sa.norm.bias, sa.norm.eps)
    x = A.AddFn.apply(_mha(sa.attention, n, n, p_drop, seed, site), x, False)
    return A.AddFn.apply(_mlp(layer[1].module, x), x, False)


def cdm_forward_train(m, x, timesteps, text_feat, kwargs):
    cm = m.contact_model
    cfg = m.arch_cfg
    seed = _seed()
    te = m.timestep_embedder
    h = te.pe[timesteps.long()]
    h = A.SiluFn.apply(A.linear(h, te.time_embed[0].weight, te.time_embed[0].bias, tc=True))
    time_emb = A.linear(h, te.time_embed[2].weight, te.time_embed[2].bias, tc=True)          # [B,1,128]
    text = text_feat.unsqueeze(1).float()
    u = x
    if m.point_feat_dim > 0:
        u = torch.cat([u, kwargs["c_pc_feat"].float()], dim=-1)
    u = torch.cat([u, kwargs["c_pc_xyz"].float()], dim=-1).contiguous()               # cdm.py:167-171
    enc_kv = A.linear(u, cm.encoder_adapter.weight, cm.encoder_adapter.bias, tc=True)         # [B,N,256]
    L = torch.cat([A.linear(text, cm.language_adapter.weight, cm.language_adapter.bias, tc=True),
                   A.linear(time_emb, cm.time_embedding_adapter.weight, cm.time_embedding_adapter.bias, tc=True)], dim=1)
    L = _cross_layer(cm.encoder_cross_attn, L, enc_kv, float(cfg.encoder_dropout), seed, 1)
    for i, layer in enumerate(cm.encoder_self_attn):
        L = _self_layer(layer, L, float(cfg.encoder_dropout), seed, 2 + i)
    dq = A.linear(enc_kv, cm.decoder_adapter.weight, cm.decoder_adapter.bias, tc=True)
    dq = _cross_layer(cm.decoder_cross_attn, dq, L, float(cfg.decoder_dropout), seed, 8)
    return A.linear(dq, m.contact_layer.weight, m.contact_layer.bias, tc=True)
