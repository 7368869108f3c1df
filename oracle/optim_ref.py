"""ORACLE (test infrastructure): AdamW update exactly as `torch.optim.AdamW` (the optimiser utils/training.py:48-50 builds)
applies it — decoupled weight decay, bias-corrected moments, single-tensor formulation — restated on plain tensors so the
fused CUDA kernel (am_adamw_flat) can be checked without going through torch.optim.  Pinned by tests/test_host_cpu.py
against torch.optim.AdamW itself on CPU."""
import math

import torch


def adamw_step(p, g, m, v, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """In-place AdamW update of (p, m, v) with gradient g; `step` is the 1-based update count."""
    b1, b2 = betas
    p.mul_(1.0 - lr * weight_decay)
    m.lerp_(g, 1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
    return p
