"""Two-stage generation (BASELINE config 5): CDM affordance map -> CMDM motion, with the hand-off kept on the device.

The reference writes the CDM samples to disk as distances (`utils/evaluate.py:55-66,170-223`: denormalise, clip to
[1e-20, 1], d = sqrt(-2 ln c * sigma^2), save pred_contact/*.npy) and the CMDM dataset re-exponentiates them
(`datasets/humanml3d.py:766,773-774`: c = exp(-d^2 / (2 sigma^2))).  exp(-(sqrt(-2 ln c s^2))^2 / (2 s^2)) == c, so the
round trip is the identity on the clipped value: `contact_from_cdm_sample` applies just the denormalise + clip."""
from typing import List, Sequence, Tuple

import torch


def contact_from_cdm_sample(sample: torch.Tensor, mean, std) -> torch.Tensor:
    """CDM sample (normalised contact) -> contact map in (0, 1] as CMDM's `c_pc_contact` expects."""
    mean = torch.as_tensor(mean, dtype=sample.dtype, device=sample.device)
    std = torch.as_tensor(std, dtype=sample.dtype, device=sample.device)
    return (sample * std + mean).clamp_(1e-20, 1.0).contiguous()


def contact_to_distance(contact: torch.Tensor, sigma: float = 0.8) -> torch.Tensor:
    """Exporter for the reference's `.npy` wire format (distances), utils/evaluate.py:55-66."""
    return torch.sqrt(-2.0 * torch.log(contact) * sigma ** 2)


@torch.no_grad()
def two_stage_generate(cdm, cdm_diffusion, cmdm, cmdm_diffusion, texts: List[str], xyz: torch.Tensor, x_mask: torch.Tensor,
                       motion_shape: Sequence[int], contact_mean=0.0, contact_std=1.0, ddim: bool = True,
                       eta: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """texts [B], xyz [B,N,3] (device), x_mask [B,T] -> (motion [B,T,D], contact [B,N,J]).  Stays on the owning rank /
    device; each stage runs its own device-resident CUDA-graph loop."""
    B, N, _ = xyz.shape
    J = cdm.contact_dim
    kw1 = dict(c_text=texts, c_pc_xyz=xyz, c_pc_feat=None)
    loop = cdm_diffusion.ddim_sample_loop if ddim else cdm_diffusion.p_sample_loop
    extra = dict(eta=eta) if ddim else {}
    sample = loop(cdm, (B, N, J), clip_denoised=False, model_kwargs=kw1, **extra)
    contact = contact_from_cdm_sample(sample, contact_mean, contact_std)
    kw2 = dict(c_text=texts, c_pc_xyz=xyz, c_pc_contact=contact, x_mask=x_mask)
    motion = cmdm_diffusion.p_sample_loop(cmdm, (B,) + tuple(motion_shape), clip_denoised=False, model_kwargs=kw2)
    return motion, contact
