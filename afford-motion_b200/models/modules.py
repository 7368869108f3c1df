"""Building blocks with the reference's parameter names (/root/reference/models/modules.py).

These classes are parameter containers + host-side glue; the arithmetic runs in the CUDA engines
(amb200.cmdm_engine / amb200.cdm_engine / amb200.scene_engine)."""
from typing import List

import numpy as np
import torch
import torch.nn as nn

from models.scene_models.pointtransformer import PointTransformerBlock, TransitionDown


def get_positional_encoding(max_len: int, time_emb_dim: int) -> torch.Tensor:
    """modules.py:10-26 -> [max_len, 1, d] (same fp32 op order, so the buffer is bit-identical)."""
    pe = torch.zeros(max_len, time_emb_dim)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, time_emb_dim, 2).float() * (-np.log(10000.0) / time_emb_dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1)


class PositionalEncoding(nn.Module):
    """modules.py:28-36 (buffer `pe`); the add is fused into the motion-adapter GEMM epilogue."""

    def __init__(self, time_emb_dim, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.register_buffer("pe", get_positional_encoding(max_len, time_emb_dim))


class TimestepEmbedder(nn.Module):
    """modules.py:38-53: time_embed(pe[t]).  The engines precompute the table for every t once per weight version."""

    def __init__(self, d_model, time_embed_dim, max_len=5000):
        super().__init__()
        self.register_buffer("pe", get_positional_encoding(max_len, time_embed_dim))
        self.d_model, self.time_embed_dim = d_model, time_embed_dim
        self.time_embed = nn.Sequential(nn.Linear(time_embed_dim, d_model), nn.SiLU(), nn.Linear(d_model, d_model))


class SceneMapEncoder(nn.Module):
    """modules.py:124-167: enc1..enc4 = TransitionDown + (blocks-1) PointTransformerBlocks, planes [32,64,128,256]."""

    def __init__(self, point_feat_dim: int, planes: List, blocks: List, num_points: int = 8192) -> None:
        super().__init__()
        self.num_points = num_points
        self.c = point_feat_dim + 3
        self.in_planes = self.c
        stride, nsample = [1, 4, 4, 4], [8, 16, 16, 16]
        for i in range(4):
            setattr(self, f"enc{i + 1}", self._make_enc(planes[i], blocks[i], 8, stride[i], nsample[i]))

    @property
    def num_groups(self):
        return self.num_points // 64

    def _make_enc(self, planes, blocks, share_planes, stride, nsample):
        layers = [TransitionDown(self.in_planes, planes, stride, nsample)]
        self.in_planes = planes
        for _ in range(1, blocks):
            layers.append(PointTransformerBlock(planes, planes, share_planes, nsample=nsample))
        return nn.Sequential(*layers)


# ----------------------------------------------------------------------------- Perceiver-IO parameter tree
class _Wrapped(nn.Module):
    """`Residual` of modules.py:222-231: child is named `module`."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module


class MultiHeadAttention(nn.Module):
    """modules.py:234-323 parameters (q/k/v/o projections)."""

    def __init__(self, num_heads, num_q_input_channels, num_kv_input_channels):
        super().__init__()
        qk = num_q_input_channels
        self.num_heads = num_heads
        self.q_proj = nn.Linear(num_q_input_channels, qk)
        self.k_proj = nn.Linear(num_kv_input_channels, qk)
        self.v_proj = nn.Linear(num_kv_input_channels, qk)
        self.o_proj = nn.Linear(qk, num_q_input_channels)


class CrossAttention(nn.Module):
    def __init__(self, num_heads, num_q_input_channels, num_kv_input_channels):
        super().__init__()
        self.q_norm = nn.LayerNorm(num_q_input_channels)
        self.kv_norm = nn.LayerNorm(num_kv_input_channels)
        self.attention = MultiHeadAttention(num_heads, num_q_input_channels, num_kv_input_channels)


class SelfAttention(nn.Module):
    def __init__(self, num_heads, num_channels):
        super().__init__()
        self.norm = nn.LayerNorm(num_channels)
        self.attention = MultiHeadAttention(num_heads, num_channels, num_channels)


def MLP(num_channels: int, widening_factor: int) -> nn.Sequential:
    """modules.py:651-661: indices 0 (LN), 1 (Linear), 3 (Linear)."""
    return nn.Sequential(nn.LayerNorm(num_channels), nn.Linear(num_channels, widening_factor * num_channels), nn.GELU(),
                         nn.Linear(widening_factor * num_channels, num_channels))


def CrossAttentionLayer(num_heads, num_q_input_channels, num_kv_input_channels, widening_factor=1) -> nn.Sequential:
    """modules.py:504-541 -> keys `0.module.{q_norm,kv_norm,attention.*}`, `1.module.{0,1,3}`."""
    return nn.Sequential(_Wrapped(CrossAttention(num_heads, num_q_input_channels, num_kv_input_channels)),
                         _Wrapped(MLP(num_q_input_channels, widening_factor)))


def SelfAttentionBlock(num_layers, num_heads, num_channels, widening_factor=1) -> nn.Sequential:
    """modules.py:544-648 -> keys `{i}.0.module.{norm,attention.*}`, `{i}.1.module.{0,1,3}`."""
    return nn.Sequential(*[nn.Sequential(_Wrapped(SelfAttention(num_heads, num_channels)), _Wrapped(MLP(num_channels, widening_factor)))
                           for _ in range(num_layers)])
