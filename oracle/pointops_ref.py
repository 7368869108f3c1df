"""ORACLE (test infrastructure): python face of pointops_ref.c plus the torch-side helpers of
/root/reference/models/scene_models/pointops.py (queryandgroup :79-100, interpolation :164-178)."""
import ctypes
import numpy as np
import torch

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.oracle_fps.restype = ctypes.c_int
        _lib.oracle_knn.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def furthestsampling(xyz: torch.Tensor, offset: torch.Tensor, new_offset: torch.Tensor) -> torch.Tensor:
    """pointops.py:10-27 -> idx int32 [m] (global row indices)."""
    lib = _load()
    x = np.ascontiguousarray(xyz.detach().cpu().numpy(), dtype=np.float32)
    o = np.ascontiguousarray(offset.cpu().numpy(), dtype=np.int32)
    no = np.ascontiguousarray(new_offset.cpu().numpy(), dtype=np.int32)
    idx = np.zeros(int(no[-1]), dtype=np.int32)
    lib.oracle_fps(ctypes.c_int(len(o)), _p(x), _p(o), _p(no), _p(idx))
    return torch.from_numpy(idx)


def knnquery(nsample: int, xyz: torch.Tensor, new_xyz, offset: torch.Tensor, new_offset: torch.Tensor):
    """pointops.py:30-45 -> (idx int32 [m,k], sqrt(dist2) fp32 [m,k])."""
    lib = _load()
    if new_xyz is None:
        new_xyz = xyz
    x = np.ascontiguousarray(xyz.detach().cpu().numpy(), dtype=np.float32)
    q = np.ascontiguousarray(new_xyz.detach().cpu().numpy(), dtype=np.float32)
    o = np.ascontiguousarray(offset.cpu().numpy(), dtype=np.int32)
    no = np.ascontiguousarray(new_offset.cpu().numpy(), dtype=np.int32)
    m = q.shape[0]
    idx = np.zeros((m, nsample), dtype=np.int32)
    d2 = np.zeros((m, nsample), dtype=np.float32)
    lib.oracle_knn(ctypes.c_int(len(o)), ctypes.c_int(m), ctypes.c_int(nsample), _p(x), _p(q), _p(o), _p(no),
                   _p(idx), _p(d2))
    return torch.from_numpy(idx), torch.sqrt(torch.from_numpy(d2))


def fps_numpy(xyz: np.ndarray, offset, new_offset) -> np.ndarray:
    """Independent numpy fp32 statement of the same rule (small cases; cross-checks the C code)."""
    xyz = xyz.astype(np.float32)
    out = []
    s_n = s_m = 0
    for e_n, e_m in zip(offset, new_offset):
        seg = xyz[s_n:e_n]
        m = e_m - s_m
        tmp = np.full(len(seg), 1e10, dtype=np.float32)
        last = 0
        sel = [0]
        for _ in range(1, m):
            d = seg - seg[last]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            tmp = np.minimum(tmp, d2.astype(np.float32))
            last = int(np.argmax(tmp))  # first maximal index == lowest index on ties
            sel.append(last)
        out.extend([s_n + i for i in sel])
        s_n, s_m = e_n, e_m
    return np.asarray(out, dtype=np.int32)


def knn_numpy(k, xyz: np.ndarray, new_xyz: np.ndarray, offset, new_offset):
    xyz = xyz.astype(np.float32)
    new_xyz = new_xyz.astype(np.float32)
    idx = np.zeros((len(new_xyz), k), dtype=np.int32)
    d2o = np.zeros((len(new_xyz), k), dtype=np.float32)
    s_n = s_m = 0
    for e_n, e_m in zip(offset, new_offset):
        seg = xyz[s_n:e_n]
        for q in range(s_m, e_m):
            d = seg - new_xyz[q]
            d2 = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
            order = np.argsort(d2, kind="stable")[:k]
            idx[q, :len(order)] = order + s_n
            d2o[q, :len(order)] = d2[order]
        s_n, s_m = e_n, e_m
    return idx, d2o


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True):
    """pointops.py:79-100."""
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        idx, _ = knnquery(nsample, xyz, new_xyz, offset, new_offset)
    m, c = new_xyz.shape[0], feat.shape[1]
    flat = idx.reshape(-1).long()
    grouped_xyz = xyz[flat, :].view(m, nsample, 3) - new_xyz.unsqueeze(1)
    grouped_feat = feat[flat, :].view(m, nsample, c)
    if use_xyz:
        return torch.cat((grouped_xyz, grouped_feat), -1)
    return grouped_feat


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """pointops.py:164-178."""
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=1, keepdim=True)
    weight = dist_recip / norm
    new_feat = torch.zeros(new_xyz.shape[0], feat.shape[1], dtype=feat.dtype)
    for i in range(k):
        new_feat += feat[idx[:, i].long(), :] * weight[:, i].unsqueeze(-1)
    return new_feat
