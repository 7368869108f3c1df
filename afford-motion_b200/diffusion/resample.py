"""diffusion/resample.py:7-12 — uniform timestep sampling used by TrainLoop (utils/training.py:141).
The loss-aware resampler (resample.py:76-110) is never constructed by the reference and is out of scope."""
import numpy as np
import torch as th


def uniform_sampling(batch_size: int, device, ddpm_steps: int):
    w = np.ones([ddpm_steps])
    p = w / np.sum(w)
    indices_np = np.random.choice(len(p), size=(batch_size,), p=p)  # same host RNG stream as the reference
    return th.from_numpy(indices_np).long().to(device)
