"""Training-mode forward of CMDM (models/cmdm.py:118-196) as an autograd graph of libamb200 kernels.

Differences from the sampling engine: nothing is cached or folded — conditioning is re-encoded every step (weights
change), BatchNorm uses batch statistics and updates the running buffers, dropout masks are Philox streams addressed by
(seed, site, element) so backward regenerates them, and every intermediate the backward needs is kept by autograd."""
import torch

from . import autograd_ops as A
from . import ops

STRIDES = [1, 4, 4, 4]
NSAMPLE = [8, 16, 16, 16]


def _seed() -> int:
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def pt_layer_train(layer, p, x, o, k):
    """PointTransformerLayer.forward (pointtransformer.py:26-38), train-mode BN."""
    n, c = x.shape
    wqkv = torch.cat([layer.linear_q.weight, layer.linear_k.weight, layer.linear_v.weight], 0)
    bqkv = torch.cat([layer.linear_q.bias, layer.linear_k.bias, layer.linear_v.bias], 0)
    qkv = A.linear(x, wqkv, bqkv)                                     # [n, 3c]  (q | k | v)
    idx, _ = ops.knnquery(k, p, p, o, o)                             # computed twice in the reference (:29-30)
    rel = torch.empty(n * k, 3, device=x.device)
    A._l.check(A.L().am_group_rel(p.data_ptr(), p.data_ptr(), idx.data_ptr(), rel.data_ptr(), n, k, A._st()), "am_group_rel")
    a = A.linear(rel, layer.linear_p[0].weight, layer.linear_p[0].bias)          # [n*k, 3]
    a = A.bn_train(a, layer.linear_p[1], relu=True)
    pr = A.linear(a, layer.linear_p[3].weight, layer.linear_p[3].bias)           # [n*k, c]
    w = A.PtWFn.apply(qkv, pr, idx, k)                                # k_nbr - q + pr
    w = A.bn_train(w, layer.linear_w[0], relu=True)
    w = A.linear(w, layer.linear_w[2].weight, layer.linear_w[2].bias)            # [n*k, c/8]
    w = A.bn_train(w, layer.linear_w[3], relu=True)
    w = A.linear(w, layer.linear_w[5].weight, layer.linear_w[5].bias)
    ws = A.SoftmaxKFn.apply(w, n, k)
    return A.PtAggFn.apply(qkv, pr, ws, idx, k)                        # [n, c]


def pt_block_train(blk, p, x, o, k):
    """PointTransformerBlock.forward (pointtransformer.py:115-123)."""
    y = A.bn_train(A.linear(x, blk.linear1.weight), blk.bn1, relu=True)
    y = A.bn_train(pt_layer_train(blk.transformer2, p, y, o, k), blk.bn2, relu=True)
    y = A.bn_train(A.linear(y, blk.linear3.weight), blk.bn3, relu=False)
    return A.AddFn.apply(y, x, True)


def scene_encoder_train(enc, xyz, feat):
    """SceneMapEncoder.forward (modules.py:152-167) in training mode -> [B, N/64, planes[-1]]."""
    B, N, _ = xyz.shape
    dev = xyz.device
    p = xyz.reshape(B * N, 3).float().contiguous()
    x = torch.cat((p, feat.reshape(B * N, -1).float()), 1).contiguous()
    n_seg = N
    o = (torch.arange(1, B + 1, device=dev, dtype=torch.int32) * n_seg).contiguous()
    for s in range(4):
        stage = getattr(enc, f"enc{s + 1}")
        td = stage[0]
        if STRIDES[s] == 1:
            x = A.bn_train(A.linear(x, td.linear.weight), td.bn, relu=True)
        else:
            m_seg = n_seg // STRIDES[s]
            n_o = (torch.arange(1, B + 1, device=dev, dtype=torch.int32) * m_seg).contiguous()
            m = B * m_seg
            k = NSAMPLE[s]
            fidx = ops.furthestsampling(p, o, n_o, n_max=n_seg, m_total=m)
            n_p = torch.empty(m, 3, device=dev)
            ops.gather_rows(p, fidx, n_p, m, 3)
            kidx, _ = ops.knnquery(k, p, n_p, o, n_o)
            rel = torch.empty(m * k, 3, device=dev)
            A._l.check(A.L().am_group_rel(p.data_ptr(), n_p.data_ptr(), kidx.data_ptr(), rel.data_ptr(), m, k, A._st()), "am_group_rel")
            G = A.GroupCatFn.apply(rel, x, kidx)                     # [m*k, 3+c]
            Z = A.bn_train(A.linear(G, td.linear.weight), td.bn, relu=True)
            x = A.MaxPoolKFn.apply(Z, m, k)
            p, o, n_seg = n_p, n_o, m_seg
        for blk in list(stage)[1:]:
            x = pt_block_train(blk, p, x, o, NSAMPLE[s])
    return x.view(B, n_seg, x.shape[1])


def cmdm_forward_train(m, x, timesteps, text_feat, kwargs):
    """CMDM.forward (cmdm.py:118-196, trans_enc) with dropout and batch-statistics BatchNorm."""
    dev = x.device
    B, T, _ = x.shape
    D = m.latent_dim
    seed = _seed()
    te = m.timestep_embedder
    h = te.pe[timesteps.long()]                                       # [B,1,temb] buffer gather (no grad)
    h = A.linear(h, te.time_embed[0].weight, te.time_embed[0].bias)
    h = A.SiluFn.apply(h)
    time_emb = A.linear(h, te.time_embed[2].weight, te.time_embed[2].bias)        # [B,1,D]
    text = text_feat.unsqueeze(1).float()
    if "c_text_erase" in kwargs:
        text = text * (1.0 - kwargs["c_text_erase"].unsqueeze(-1).float())
    text_emb = A.linear(text, m.language_adapter.weight, m.language_adapter.bias)
    cont = scene_encoder_train(m.contact_encoder, kwargs["c_pc_xyz"], kwargs["c_pc_contact"])
    G = cont.shape[1]
    if "c_pc_erase" in kwargs:
        cont = cont * (1.0 - kwargs["c_pc_erase"].unsqueeze(-1).float())
    cont_emb = A.linear(cont, m.contact_adapter.weight, m.contact_adapter.bias)
    mot = A.linear(x, m.motion_adapter.weight, m.motion_adapter.bias)
    X = torch.cat([time_emb, text_emb, cont_emb, mot], dim=1)        # [B,S,D] (data movement)
    S = X.shape[1]
    pe = m.positional_encoder.pe[:S, 0, :].unsqueeze(0).expand(B, S, D).contiguous()
    X = A.AddFn.apply(X, pe, False)
    X = A.DropoutFn.apply(X, m.positional_encoder.dropout.p, seed, 1)
    key_pad = None
    if m.mask_motion:
        kp = torch.zeros(B, S, dtype=torch.bool, device=dev)
        if "c_text_mask" in kwargs:
            kp[:, 1] = kwargs["c_text_mask"].view(B).bool()
        if "c_pc_mask" in kwargs:
            kp[:, 2:2 + G] = kwargs["c_pc_mask"].view(B, 1).bool()
        kp[:, 2 + G:] = kwargs["x_mask"].bool()
        key_pad = kp.to(torch.uint8).contiguous()
    site = 16
    for layer in m.self_attn_layer.layers:
        sa = layer.self_attn
        qkv = A.linear(X, sa.in_proj_weight, sa.in_proj_bias, tc=True)
        a = A.AttentionFn.apply(qkv, key_pad, sa.num_heads, float(sa.dropout), seed, site)
        a = A.linear(a, sa.out_proj.weight, sa.out_proj.bias, tc=True)
        a = A.DropoutFn.apply(a, layer.dropout1.p, seed, site + 1)
        X = A.LayerNormFn.apply(A.AddFn.apply(X, a, False), layer.norm1.weight, layer.norm1.bias, layer.norm1.eps)
        f = A.GeluFn.apply(A.linear(X, layer.linear1.weight, layer.linear1.bias, tc=True))
        f = A.DropoutFn.apply(f, layer.dropout.p, seed, site + 2)
        f = A.linear(f, layer.linear2.weight, layer.linear2.bias, tc=True)
        f = A.DropoutFn.apply(f, layer.dropout2.p, seed, site + 3)
        X = A.LayerNormFn.apply(A.AddFn.apply(X, f, False), layer.norm2.weight, layer.norm2.bias, layer.norm2.eps)
        site += 8
    Xm = X[:, 2 + G:, :].contiguous()
    return A.linear(Xm, m.motion_layer.weight, m.motion_layer.bias)
