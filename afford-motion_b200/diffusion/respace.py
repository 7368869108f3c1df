"""Timestep respacing — drop-in for /root/reference/diffusion/respace.py.

Differences in mechanism, not behaviour: the spaced->original timestep map lives on the device once (the reference
rebuilds `th.tensor(self.timestep_map)` and copies it H2D on every model call, respace.py:124-126)."""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """respace.py:8-61: 'ddimN' striding or per-section counts -> set of retained original timesteps."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    start, kept = 0, []
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += frac
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """respace.py:64-114: keep `use_timesteps` of a base process; betas re-derived from the retained alpha-bars."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        alphas_cumprod = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64), axis=0)
        last, new_betas, self.timestep_map = 1.0, [], []
        for i, ac in enumerate(alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def training_losses(self, model, *args, **kwargs):
        return super().training_losses(self._wrap_model(model), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t  # done by the wrapped model


class _WrappedModel:
    """respace.py:117-129."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._maps = {}

    def map_tensor(self, device, dtype):
        key = (str(device), dtype)
        if key not in self._maps:
            self._maps[key] = th.tensor(self.timestep_map, device=device, dtype=dtype)
        return self._maps[key]

    def __call__(self, x, ts, **kwargs):
        new_ts = self.map_tensor(ts.device, ts.dtype)[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)
