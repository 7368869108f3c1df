"""`pointops` API of the reference (/root/reference/models/scene_models/pointops.py) on top of libamb200.

Only the four entry points the hot path reaches are provided: furthestsampling, knnquery,
queryandgroup, interpolation.  Unlike the reference there is no `pointops_cuda` pybind module: the native code is
the C-ABI library (include/amb200.h: am_furthestsampling, am_knnquery).
"""
import torch

from amb200 import ops


def furthestsampling(xyz, offset, new_offset):
    """pointops.py:10-27.  Host reads of offsets are kept here for API compatibility with arbitrary (ragged)
    offsets; the engines call amb200.ops.furthestsampling directly with static sizes (no sync)."""
    assert xyz.is_contiguous()
    o = offset.tolist()
    n_max = max(b - a for a, b in zip([0] + o[:-1], o))
    return ops.furthestsampling(xyz, offset.int().contiguous(), new_offset.int().contiguous(), n_max, int(new_offset[-1].item()))


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """pointops.py:30-45 -> (idx, sqrt(dist2))."""
    if new_xyz is None:
        new_xyz = xyz
    assert xyz.is_contiguous() and new_xyz.is_contiguous()
    idx, d2 = ops.knnquery(nsample, xyz, new_xyz, offset.int().contiguous(), new_offset.int().contiguous())
    return idx, torch.sqrt(d2)


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True):
    """pointops.py:79-100 (materialising variant kept for API completeness; the engines use the fused kernels)."""
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        idx, _ = knnquery(nsample, xyz, new_xyz, offset, new_offset)
    m, c = new_xyz.shape[0], feat.shape[1]
    flat = idx.reshape(-1)
    gx = torch.empty(m * nsample, 3, device=xyz.device)
    ops.gather_rows(xyz.contiguous(), flat, gx, m * nsample, 3)
    gx = gx.view(m, nsample, 3) - new_xyz.unsqueeze(1)
    gf = torch.empty(m * nsample, c, device=xyz.device)
    ops.gather_rows(feat.contiguous(), flat, gf, m * nsample, c)
    gf = gf.view(m, nsample, c)
    return torch.cat((gx, gf), -1) if use_xyz else gf


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """pointops.py:164-178."""
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=1, keepdim=True)
    out = torch.zeros(new_xyz.shape[0], feat.shape[1], device=feat.device)
    g = torch.empty(new_xyz.shape[0], feat.shape[1], device=feat.device)
    for i in range(k):
        ops.gather_rows(feat.contiguous(), idx[:, i].contiguous(), g, new_xyz.shape[0], feat.shape[1])
        out += g * weight[:, i].unsqueeze(-1)
    return out
