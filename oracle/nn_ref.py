"""ORACLE (test infrastructure): small fp32 building blocks shared by the model restatements."""
import math
import numpy as np
import torch
import torch.nn.functional as F


def lin(sd, name, x):
    b = sd.get(name + ".bias")
    return F.linear(x, sd[name + ".weight"], b)


def ln(sd, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


TRAIN_BN = False  # set by oracle.cmdm_ref.cmdm_forward(train=True): BatchNorm1d uses batch statistics (model.train())


def bn_eval(sd, name, x, eps=1e-5):
    """BatchNorm1d on the channel (last) dim.  Eval: per-channel affine from running stats.  Train (TRAIN_BN): batch
    statistics over every other dim, biased variance (torch.nn.functional.batch_norm(training=True))."""
    if TRAIN_BN:
        dims = tuple(range(x.dim() - 1))
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        return (x - mean) / torch.sqrt(var + eps) * sd[name + ".weight"] + sd[name + ".bias"]
    s = sd[name + ".weight"] / torch.sqrt(sd[name + ".running_var"] + eps)
    return (x - sd[name + ".running_mean"]) * s + sd[name + ".bias"]


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def positional_table(max_len: int, d: int) -> torch.Tensor:
    """models/modules.py:10-26 -> [max_len, d]."""
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2).float() * (-np.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def timestep_embed(sd, prefix, t):
    """models/modules.py:38-53: time_embed(pe[t]) -> [B,1,d]."""
    pe = sd[prefix + ".pe"]  # [max_len,1,temb]
    h = pe[t]  # [B,1,temb]
    h = lin(sd, prefix + ".time_embed.0", h)
    h = F.silu(h)
    return lin(sd, prefix + ".time_embed.2", h)
