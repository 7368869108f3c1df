"""ORACLE (test infrastructure): CDM (arch='Perceiver', no scene model) forward, eval mode, fp32.

Restates models/cdm.py:155-188,474-513 and the Perceiver-IO stack of models/modules.py:222-661
(SURVEY Appendix A).  Straight (unfolded) formulation: K/V are materialised exactly as the
reference does, so the algebraic folds used by the CUDA path are checked against it, not assumed.
"""
import torch

from .nn_ref import lin, ln, gelu, timestep_embed


def _mha(sd, pre, xq, xkv, heads):
    """modules.py:324-381 (no mask, no rotary, dropout=identity)."""
    q, k, v = lin(sd, pre + ".q_proj", xq), lin(sd, pre + ".k_proj", xkv), lin(sd, pre + ".v_proj", xkv)
    B, Nq, C = q.shape
    c = C // heads
    q = q.view(B, Nq, heads, c).transpose(1, 2) * (c ** -0.5)
    k = k.view(B, -1, heads, c).transpose(1, 2)
    v = v.view(B, -1, heads, v.shape[-1] // heads).transpose(1, 2)
    a = torch.softmax(torch.einsum("bhic,bhjc->bhij", q, k), dim=-1)
    o = torch.einsum("bhij,bhjc->bhic", a, v).transpose(1, 2).reshape(B, Nq, -1)
    return lin(sd, pre + ".o_proj", o)


def _mlp(sd, pre, x):
    """modules.py:651-661."""
    return lin(sd, pre + ".3", gelu(lin(sd, pre + ".1", ln(sd, pre + ".0", x))))


def cross_layer(sd, pre, xq, xkv, heads):
    """modules.py:504-541 + Residual :222-231 (residual adds the UN-normalised x_q)."""
    a = _mha(sd, pre + ".0.module.attention", ln(sd, pre + ".0.module.q_norm", xq), ln(sd, pre + ".0.module.kv_norm", xkv), heads)
    x = a + xq
    return _mlp(sd, pre + ".1.module", x) + x


def self_layer(sd, pre, x, heads):
    """modules.py:544-578."""
    n = ln(sd, pre + ".0.module.norm", x)
    x = _mha(sd, pre + ".0.module.attention", n, n, heads) + x
    return _mlp(sd, pre + ".1.module", x) + x


def cdm_forward(sd, x, t, text_feat, xyz, point_feat=None, enc_heads=8, dec_heads=8, n_self=2):
    time_emb = timestep_embed(sd, "timestep_embedder", t)  # [B,1,128]
    text_emb = text_feat.unsqueeze(1).float()
    u = x if point_feat is None else torch.cat([x, point_feat], dim=-1)
    u = torch.cat([u, xyz], dim=-1)
    cm = "contact_model"
    enc_kv = lin(sd, cm + ".encoder_adapter", u)
    L = torch.cat([lin(sd, cm + ".language_adapter", text_emb), lin(sd, cm + ".time_embedding_adapter", time_emb)], dim=1)
    L = cross_layer(sd, cm + ".encoder_cross_attn", L, enc_kv, enc_heads)
    for i in range(n_self):
        L = self_layer(sd, f"{cm}.encoder_self_attn.{i}", L, enc_heads)
    dq = lin(sd, cm + ".decoder_adapter", enc_kv)
    dq = cross_layer(sd, cm + ".decoder_cross_attn", dq, L, dec_heads)
    return lin(sd, "contact_layer", dq)
