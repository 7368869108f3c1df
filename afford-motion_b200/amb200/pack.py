"""Weight packing for the CUDA engines: eval-mode BatchNorm folding, fused QKV / adapter matrices,
Perceiver fold matrices.  Pure host-side setup (runs once per weight version, not on the hot path)."""
from typing import Dict

import torch


def bn_fold(bn: torch.nn.BatchNorm1d):
    """Eval-mode BatchNorm1d == per-channel affine: y = x*s + t."""
    s = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    t = bn.bias.detach() - bn.running_mean.detach() * s
    return s.float().contiguous(), t.float().contiguous()


def c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


def pack_transition_down(td) -> Dict[str, torch.Tensor]:
    """models/scene_models/pointtransformer.py:41-69: linear (no bias) -> bn -> relu [-> max over k]."""
    s, t = bn_fold(td.bn)
    return {"W": c(td.linear.weight.detach() * s[:, None]), "shift": t}


def pack_pt_layer(layer, bn2) -> Dict[str, torch.Tensor]:
    """pointtransformer.py:9-38 (+ bn2 of the enclosing block :119)."""
    w = {}
    w["wqkv"] = c(torch.cat([layer.linear_q.weight, layer.linear_k.weight, layer.linear_v.weight], 0))
    w["bqkv"] = c(torch.cat([layer.linear_q.bias, layer.linear_k.bias, layer.linear_v.bias], 0))
    s, t = bn_fold(layer.linear_p[1])
    w["wp1"] = c(layer.linear_p[0].weight.detach() * s[:, None])
    w["bp1"] = c(layer.linear_p[0].bias.detach() * s + t)
    w["wp2"] = c(layer.linear_p[3].weight)
    w["bp2"] = c(layer.linear_p[3].bias)
    w["bnw_s"], w["bnw_t"] = bn_fold(layer.linear_w[0])
    s, t = bn_fold(layer.linear_w[3])
    w["ww1"] = c(layer.linear_w[2].weight.detach() * s[:, None])
    w["bw1"] = c(layer.linear_w[2].bias.detach() * s + t)
    w["ww2"] = c(layer.linear_w[5].weight)
    w["bw2"] = c(layer.linear_w[5].bias)
    w["post_s"], w["post_t"] = bn_fold(bn2)
    return w


def pack_pt_block(blk) -> Dict[str, torch.Tensor]:
    """pointtransformer.py:102-123."""
    w = pack_pt_layer(blk.transformer2, blk.bn2)
    s1, t1 = bn_fold(blk.bn1)
    w["w1"] = c(blk.linear1.weight.detach() * s1[:, None])
    w["b1"] = t1
    s3, t3 = bn_fold(blk.bn3)
    w["w3"] = c(blk.linear3.weight.detach() * s3[:, None])
    w["b3"] = t3
    return w


def params_version(module: torch.nn.Module) -> int:
    """Cheap change detector: in-place updates (optimizer.step, load_state_dict) bump tensor._version."""
    v = 0
    for p in module.parameters():
        v += p._version + (p.data_ptr() & 0xFFFF)
    for b in module.buffers():
        v += b._version
    return v


def pack_transition_up(tu) -> Dict[str, torch.Tensor]:
    """pointtransformer.py:72-99 with eval-mode BN folded.  Head form: linear1 over cat(x, g) is used as two column blocks of
    the same [c, 2c] matrix (w1a = first c columns, w1b = last c); fusion form: linear1+BN (w1, b1) on the fine level,
    linear2+BN (w2, b2) on the coarse one."""
    s1, t1 = bn_fold(tu.linear1[1])
    W1 = c(tu.linear1[0].weight.detach() * s1[:, None])
    b1 = c(tu.linear1[0].bias.detach() * s1 + t1)
    if tu.is_head:
        cc = W1.shape[0]
        return {"w1": W1, "w1a": W1, "w1b": c(W1[:, cc:]), "b1": b1, "w2": c(tu.linear2[0].weight), "b2": c(tu.linear2[0].bias)}
    s2, t2 = bn_fold(tu.linear2[1])
    return {"w1": W1, "b1": b1, "w2": c(tu.linear2[0].weight.detach() * s2[:, None]), "b2": c(tu.linear2[0].bias.detach() * s2 + t2)}
