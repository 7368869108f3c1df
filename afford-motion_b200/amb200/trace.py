"""Optional NVTX ranges around the phases of a sampling / training job (AMB200_NVTX=1): conditioning encode, graph capture,
graph replay, sampler update.  Off by default — a push/pop pair costs host time on a loop that enqueues a 1000-step job in 2 ms.
Use with `ncu --nvtx --nvtx-include "denoise_steps/"` (nsys is not installed in this image) or any NVTX-aware tool."""
import contextlib
import os

import torch

ON = os.environ.get("AMB200_NVTX", "0") == "1"


@contextlib.contextmanager
def rng(name: str):
    if not ON:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()
