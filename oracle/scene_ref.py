"""ORACLE (test infrastructure): SceneMapEncoder / PointTransformer blocks in eval mode.

Restates models/modules.py:124-167 and models/scene_models/pointtransformer.py:9-123 on packed
(p [n,3], x [n,c], o [b]) batches, using the oracle FPS/kNN (pointops_ref).  Eval-mode BatchNorm.
"""
import torch
import torch.nn.functional as F

from . import pointops_ref as P
from .nn_ref import lin, bn_eval


def pt_layer(sd, pre, p, x, o, nsample, share=8):
    """pointtransformer.py:26-38."""
    xq, xk, xv = lin(sd, pre + ".linear_q", x), lin(sd, pre + ".linear_k", x), lin(sd, pre + ".linear_v", x)
    idx, _ = P.knnquery(nsample, p, p, o, o)  # the reference computes this twice with identical args (:29,:30)
    gk = P.queryandgroup(nsample, p, p, xk, idx, o, o, use_xyz=True)
    gv = P.queryandgroup(nsample, p, p, xv, idx, o, o, use_xyz=False)
    pr, gk = gk[:, :, 0:3], gk[:, :, 3:]
    pr = lin(sd, pre + ".linear_p.0", pr)
    pr = F.relu(bn_eval(sd, pre + ".linear_p.1", pr))
    pr = lin(sd, pre + ".linear_p.3", pr)  # [n,k,c]
    w = gk - xq.unsqueeze(1) + pr  # the view(...).sum(2) at :33 is the identity (out==mid)
    w = F.relu(bn_eval(sd, pre + ".linear_w.0", w))
    w = lin(sd, pre + ".linear_w.2", w)
    w = F.relu(bn_eval(sd, pre + ".linear_w.3", w))
    w = lin(sd, pre + ".linear_w.5", w)
    w = torch.softmax(w, dim=1)
    n, k, c = gv.shape
    return ((gv + pr).view(n, k, share, c // share) * w.unsqueeze(2)).sum(1).view(n, c)


def transition_down(sd, pre, p, x, o, stride, nsample):
    """pointtransformer.py:53-69."""
    if stride != 1:
        ol = o.tolist()
        n_o, cnt, prev = [], 0, 0
        for e in ol:
            cnt += (e - prev) // stride
            n_o.append(cnt)
            prev = e
        n_o = torch.tensor(n_o, dtype=torch.int32)
        idx = P.furthestsampling(p, o, n_o)
        n_p = p[idx.long(), :]
        g = P.queryandgroup(nsample, p, n_p, x, None, o, n_o, use_xyz=True)  # [m,k,3+c]
        h = F.linear(g, sd[pre + ".linear.weight"])  # bias=False
        h = F.relu(bn_eval(sd, pre + ".bn", h))
        h = h.max(dim=1).values
        return n_p, h, n_o
    h = F.relu(bn_eval(sd, pre + ".bn", F.linear(x, sd[pre + ".linear.weight"])))
    return p, h, o


def pt_block(sd, pre, p, x, o, nsample):
    """pointtransformer.py:115-123."""
    idn = x
    y = F.relu(bn_eval(sd, pre + ".bn1", F.linear(x, sd[pre + ".linear1.weight"])))
    y = F.relu(bn_eval(sd, pre + ".bn2", pt_layer(sd, pre + ".transformer2", p, y, o, nsample)))
    y = bn_eval(sd, pre + ".bn3", F.linear(y, sd[pre + ".linear3.weight"]))
    return F.relu(y + idn)


def scene_map_encoder(sd, pre, xyz, feat, blocks=(2, 2, 2, 2)):
    """models/modules.py:152-167 -> [B, N/64, planes[-1]]."""
    B, N, _ = xyz.shape
    p = xyz.reshape(B * N, 3).contiguous()
    x = torch.cat((p, feat.reshape(B * N, -1)), 1)
    o = torch.tensor([N * (i + 1) for i in range(B)], dtype=torch.int32)
    strides, ns = [1, 4, 4, 4], [8, 16, 16, 16]
    for s in range(4):
        e = f"{pre}.enc{s + 1}"
        p, x, o = transition_down(sd, e + ".0", p, x, o, strides[s], ns[s])
        for bi in range(1, blocks[s]):
            x = pt_block(sd, f"{e}.{bi}", p, x, o, ns[s])
    return x.view(B, -1, x.shape[-1])


def transition_up(sd, pre, pxo1, pxo2=None):
    """pointtransformer.py:82-99.  Head form (pxo2 None): per-segment mean -> linear2+ReLU, repeated and concatenated to
    every row -> linear1+BN+ReLU.  Fusion form: linear1+BN+ReLU(x1) + 3-NN interpolation of linear2+BN+ReLU(x2)."""
    if pxo2 is None:
        _, x, o = pxo1
        rows, s = [], 0
        for e in o.tolist():
            xb = x[s:e]
            g = F.relu(lin(sd, pre + ".linear2.0", xb.sum(0, True) / (e - s)))
            rows.append(torch.cat((xb, g.repeat(e - s, 1)), 1))
            s = e
        h = lin(sd, pre + ".linear1.0", torch.cat(rows, 0))
        return F.relu(bn_eval(sd, pre + ".linear1.1", h))
    p1, x1, o1 = pxo1
    p2, x2, o2 = pxo2
    a = F.relu(bn_eval(sd, pre + ".linear1.1", lin(sd, pre + ".linear1.0", x1)))
    b = F.relu(bn_eval(sd, pre + ".linear2.1", lin(sd, pre + ".linear2.0", x2)))
    return a + P.interpolation(p2, p1, b, o2, o1)


def point_transformer_seg(sd, pre, xyz, feat, blocks=(2, 3, 4, 6, 3), c=6):
    """PointTransformerSeg.forward, (p, x) input form (pointtransformer.py:166-201) -> [B, N, 32].
    `pre` is the state_dict prefix of the scene model ('' or 'scene_model')."""
    B, N, _ = xyz.shape
    dot = (pre + ".") if pre else ""
    p = xyz.reshape(B * N, 3).contiguous()
    x = p if c == 3 else torch.cat((p, feat.reshape(B * N, -1)), 1)
    o = torch.tensor([N * (i + 1) for i in range(B)], dtype=torch.int32)
    strides, ns = [1, 4, 4, 4, 4], [8, 16, 16, 16, 16]
    lv = []
    for s in range(5):
        e = f"{dot}enc{s + 1}"
        p, x, o = transition_down(sd, e + ".0", p, x, o, strides[s], ns[s])
        for bi in range(1, blocks[s]):
            x = pt_block(sd, f"{e}.{bi}", p, x, o, ns[s])
        lv.append([p, x, o])
    for s in range(4, -1, -1):
        d = f"{dot}dec{s + 1}"
        p, x, o = lv[s]
        x = transition_up(sd, d + ".0", lv[s]) if s == 4 else transition_up(sd, d + ".0", lv[s], lv[s + 1])
        x = pt_block(sd, d + ".1", p, x, o, ns[s])  # _make_dec(blocks=2): one block after the TransitionUp
        lv[s] = [p, x, o]
    return lv[0][1].view(B, N, -1)
