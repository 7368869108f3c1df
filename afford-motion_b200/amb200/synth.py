"""Synthetic weights and inputs of the named shapes (no datasets / checkpoints / CLIP weights are
available offline).  Shared by bench.py, the tests and the golden-vector generator so that the
reference modules, the oracle and the CUDA path all see bit-identical tensors.

Input distributions follow SURVEY.md §8(d): 4 m scene crops with the floor at z≈0
(prepare/generate_contact_data.py:361-435), contact = exp(-d²/(2·0.8²)) (datasets/humanml3d.py:773-774),
motion lengths in {40,44,…,196} (humanml3d.py:777-783).
"""
import zlib
from typing import Dict, Iterable, Tuple

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def fill_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, skip: Iterable[str] = ("pe",)) -> Dict[str, torch.Tensor]:
    """Deterministic, order-independent values for every parameter/buffer name in `shapes`.

    Linear weights ~ U(±1/sqrt(fan_in)) (PyTorch default scale), biases ~ 0.05·N(0,1), norm scales
    1+0.1·N(0,1), BatchNorm running stats randomised so eval-mode BN is a non-trivial affine map.
    Sinusoidal `pe` buffers are left to the module (they are deterministic).
    """
    out = {}
    for name, shape in shapes.items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf in skip or name.startswith("text_model."):
            continue
        g = _gen(name, seed)
        shape = tuple(shape)
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros(shape, dtype=torch.long)
        elif leaf == "running_mean":
            out[name] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            out[name] = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) >= 2:
            a = 1.0 / (shape[-1] ** 0.5)
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif leaf in ("weight",):
            out[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:  # bias, in_proj_bias
            out[name] = 0.05 * torch.randn(shape, generator=g)
    return out


def scene_points(B: int, N: int, seed: int = 2023, dup_frac: float = 0.0) -> torch.Tensor:
    g = _gen("xyz", seed)
    xy = torch.rand(B, N, 2, generator=g) * 4 - 2
    z = torch.rand(B, N, 1, generator=g) * 2.5
    p = torch.cat([xy, z], -1)
    if dup_frac > 0:  # crops with < N points duplicate rows (generate_contact_data.py:418-423)
        k = int(N * dup_frac)
        src = torch.randint(0, N, (B, k), generator=g)
        dst = torch.randint(0, N, (B, k), generator=g)
        for b in range(B):
            p[b, dst[b]] = p[b, src[b]]
    return p.contiguous()


def contact_map(B: int, N: int, J: int = 6, seed: int = 2023) -> torch.Tensor:
    g = _gen("contact", seed)
    d = torch.rand(B, N, J, generator=g) * 3.0
    return torch.exp(-0.5 * d * d / (0.8 ** 2)).contiguous()


def text_features(B: int, dim: int = 512, seed: int = 2023) -> torch.Tensor:
    return 0.4 * torch.randn(B, dim, generator=_gen("text", seed))


def motion_noise(B: int, T: int = 196, D: int = 263, seed: int = 2023) -> torch.Tensor:
    return torch.randn(B, T, D, generator=_gen("motion", seed))


def motion_mask(B: int, T: int = 196, seed: int = 2023, all_valid: bool = False) -> torch.Tensor:
    """x_mask [B,T] bool, True = padded frame."""
    if all_valid:
        return torch.zeros(B, T, dtype=torch.bool)
    g = _gen("mask", seed)
    lens = 40 + 4 * torch.randint(0, (T - 40) // 4 + 1, (B,), generator=g)
    lens[0] = T
    return torch.arange(T)[None, :] >= lens[:, None]


def step_noise(shape, step: int, seed: int = 1234) -> torch.Tensor:
    """Injected per-step normal noise for parity runs (CPU mt19937 vs CUDA Philox streams differ,
    gaussian_diffusion.py:431 — so parity harnesses upload the same tensors to both sides)."""
    return torch.randn(*shape, generator=_gen(f"step{step}", seed))
