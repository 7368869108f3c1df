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
