"""ORACLE (test infrastructure): diffusion schedules, tables and sampler/loss math.

Restates /root/reference/diffusion/gaussian_diffusion.py and respace.py for the configured
diffusion (START_X, FIXED_SMALL, MSE, rescale_timesteps=False; configs/default.yaml:31-40).
Tables are float64 numpy exactly as the reference builds them; per-step math is torch fp32.
"""
import math
import numpy as np
import torch


def cosine_betas(T: int, max_beta: float = 0.999) -> np.ndarray:
    """gaussian_diffusion.py:19-63 ('cosine' + betas_for_alpha_bar)."""
    ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - ab((i + 1) / T) / ab(i / T), max_beta) for i in range(T)], dtype=np.float64)


def linear_betas(T: int) -> np.ndarray:
    """gaussian_diffusion.py:28-36."""
    scale = 1000 / T
    return np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)


def space_timesteps(num_timesteps: int, section_counts) -> set:
    """respace.py:8-61."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[4:])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start = 0
    steps = []
    for i, cnt in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < cnt:
            raise ValueError(f"cannot divide section of {size} steps into {cnt}")
        stride = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


def make_tables(betas: np.ndarray) -> dict:
    """gaussian_diffusion.py:119-170."""
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    return dict(
        betas=betas,
        alphas_cumprod=ac,
        alphas_cumprod_prev=ac_prev,
        alphas_cumprod_next=np.append(ac[1:], 0.0),
        sqrt_alphas_cumprod=np.sqrt(ac),
        sqrt_one_minus_alphas_cumprod=np.sqrt(1.0 - ac),
        log_one_minus_alphas_cumprod=np.log(1.0 - ac),
        sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
        sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
        posterior_variance=pv,
        posterior_log_variance_clipped=np.log(np.append(pv[1], pv[1:])),
        posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    )


def respaced(base_betas: np.ndarray, use_timesteps) -> tuple:
    """respace.py:73-87 -> (new_betas, timestep_map)."""
    use = set(use_timesteps)
    ac = np.cumprod(1.0 - np.asarray(base_betas, dtype=np.float64))
    last = 1.0
    nb, tmap = [], []
    for i, a in enumerate(ac):
        if i in use:
            nb.append(1 - a / last)
            last = a
            tmap.append(i)
    return np.array(nb), tmap


def _ext(arr: np.ndarray, t: torch.Tensor, ndim: int) -> torch.Tensor:
    """gaussian_diffusion.py:829-842 (fp64 gather, THEN .float())."""
    r = torch.from_numpy(arr)[t.cpu()].float()
    while r.dim() < ndim:
        r = r[..., None]
    return r


def q_sample(tab, x0, t, noise):
    """gaussian_diffusion.py:189-207."""
    return _ext(tab["sqrt_alphas_cumprod"], t, x0.dim()) * x0 + _ext(tab["sqrt_one_minus_alphas_cumprod"], t, x0.dim()) * noise


def p_sample_step(tab, x0_hat, x_t, t, noise):
    """gaussian_diffusion.py:209-231, 306-315, 396-440 (START_X, FIXED_SMALL, clip_denoised=False)."""
    nd = x_t.dim()
    mean = _ext(tab["posterior_mean_coef1"], t, nd) * x0_hat + _ext(tab["posterior_mean_coef2"], t, nd) * x_t
    logvar = _ext(tab["posterior_log_variance_clipped"], t, nd)
    nz = (t != 0).float().view(-1, *([1] * (nd - 1)))
    return mean + nz * torch.exp(0.5 * logvar) * noise


def ddim_step(tab, x0_hat, x_t, t, noise, eta=0.0):
    """gaussian_diffusion.py:346-350, 538-586."""
    nd = x_t.dim()
    eps = (_ext(tab["sqrt_recip_alphas_cumprod"], t, nd) * x_t - x0_hat) / _ext(tab["sqrt_recipm1_alphas_cumprod"], t, nd)
    ab = _ext(tab["alphas_cumprod"], t, nd)
    abp = _ext(tab["alphas_cumprod_prev"], t, nd)
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    mean_pred = x0_hat * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
    nz = (t != 0).float().view(-1, *([1] * (nd - 1)))
    return mean_pred + nz * sigma * noise


def masked_mse(x0, pred, x_mask):
    """gaussian_diffusion.py:815-818 with nn.py:93-97; x_mask [B,T] bool True=pad -> loss [B]."""
    d = x0.shape[-1]
    keep = (~x_mask).float().unsqueeze(-1)
    se = (x0 - pred) ** 2
    return (se * keep).flatten(1).sum(1) / (keep.flatten(1).sum(1) * d)
