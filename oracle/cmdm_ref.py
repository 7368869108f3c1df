"""ORACLE (test infrastructure): CMDM (arch='trans_enc') forward, eval mode, fp32.

Restates models/cmdm.py:118-170,195-196 and the arithmetic of torch.nn.TransformerEncoderLayer
(post-LN, gelu, batch_first; cmdm.py:66-77) with explicit matmuls.  `text_feat` is the [B,512]
output of the `encode_text_clip` hook (CLIP itself is out of scope / parity unpinned).
"""
import math
import torch

from .nn_ref import lin, ln, gelu, timestep_embed
from .scene_ref import scene_map_encoder


def encoder_layer(sd, pre, x, key_pad, nhead=8):
    B, S, D = x.shape
    hd = D // nhead
    qkv = torch.nn.functional.linear(x, sd[pre + ".self_attn.in_proj_weight"], sd[pre + ".self_attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, nhead, hd).transpose(1, 2)
    k = k.view(B, S, nhead, hd).transpose(1, 2)
    v = v.view(B, S, nhead, hd).transpose(1, 2)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if key_pad is not None:
        sc = sc.masked_fill(key_pad[:, None, None, :], float("-inf"))
    a = torch.softmax(sc, dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, S, D)
    a = lin(sd, pre + ".self_attn.out_proj", a)
    x = ln(sd, pre + ".norm1", x + a)
    f = lin(sd, pre + ".linear2", gelu(lin(sd, pre + ".linear1", x)))
    return ln(sd, pre + ".norm2", x + f)


def contact_tokens(sd, xyz, contact, blocks=(2, 2, 2, 2)):
    return scene_map_encoder(sd, "contact_encoder", xyz, contact, blocks)


def cmdm_forward(sd, x, t, text_feat, xyz, contact, x_mask, nlayers=5, nhead=8, cont_emb=None,
                 c_text_mask=None, c_text_erase=None, c_pc_mask=None, c_pc_erase=None, mask_motion=True, train=False):
    """train=True: model.train() semantics with every dropout probability 0 (batch-statistics BatchNorm)."""
    from . import nn_ref
    nn_ref.TRAIN_BN = bool(train)
    try:
        return _cmdm_forward(sd, x, t, text_feat, xyz, contact, x_mask, nlayers, nhead, cont_emb, c_text_mask, c_text_erase, c_pc_mask,
                             c_pc_erase, mask_motion)
    finally:
        nn_ref.TRAIN_BN = False


def _cmdm_forward(sd, x, t, text_feat, xyz, contact, x_mask, nlayers, nhead, cont_emb, c_text_mask, c_text_erase, c_pc_mask, c_pc_erase,
                  mask_motion):
    B, T, _ = x.shape
    time_emb = timestep_embed(sd, "timestep_embedder", t)  # [B,1,512]
    text_emb = text_feat.unsqueeze(1).float()
    text_mask = torch.zeros(B, 1, dtype=torch.bool)
    if c_text_mask is not None:
        text_mask = torch.logical_or(text_mask, c_text_mask.repeat(1, 1))
    if c_text_erase is not None:
        text_emb = text_emb * (1.0 - c_text_erase.unsqueeze(-1).float())
    text_emb = lin(sd, "language_adapter", text_emb)
    if cont_emb is None:
        cont_emb = contact_tokens(sd, xyz, contact)
    G = cont_emb.shape[1]
    cont_mask = torch.zeros(B, G, dtype=torch.bool)
    if c_pc_mask is not None:
        cont_mask = torch.logical_or(cont_mask, c_pc_mask.repeat(1, G))
    if c_pc_erase is not None:
        cont_emb = cont_emb * (1.0 - c_pc_erase.unsqueeze(-1).float())
    cont_tok = lin(sd, "contact_adapter", cont_emb)
    mot = lin(sd, "motion_adapter", x)
    h = torch.cat([time_emb, text_emb, cont_tok, mot], dim=1)
    S = h.shape[1]
    h = h + sd["positional_encoder.pe"][:S, 0, :].unsqueeze(0)
    key_pad = None
    if mask_motion:
        key_pad = torch.cat([torch.zeros(B, 1, dtype=torch.bool), text_mask, cont_mask, x_mask], dim=1)
    for l in range(nlayers):
        h = encoder_layer(sd, f"self_attn_layer.layers.{l}", h, key_pad, nhead)
    h = h[:, 2 + G:, :]
    return lin(sd, "motion_layer", h)
