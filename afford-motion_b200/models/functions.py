"""Text-encoder hooks with the reference's signatures (/root/reference/models/functions.py:46-94).

CLIP itself is out of scope (third-party frozen tower, no weights offline): the call site is kept,
results are cached per unique string (the reference re-encodes the same prompts on every denoise step,
cmdm.py:133-135), and a feature provider can be registered where CLIP is not installed (bench / tests).
"""
from typing import Callable, List, Optional

import torch

_PROVIDER: Optional[Callable[[List[str]], torch.Tensor]] = None
_CACHE = {}


def set_text_feature_provider(fn: Optional[Callable[[List[str]], torch.Tensor]]) -> None:
    """fn(list[str]) -> [B, feat_dim] tensor (any device).  Used in place of CLIP.encode_text."""
    global _PROVIDER
    _PROVIDER = fn
    _CACHE.clear()


class _NoTextModel(torch.nn.Module):
    """Placeholder when the `clip` package is absent: holds no parameters (text_model.* keys are dropped from
    checkpoints anyway, utils/training.py:97)."""

    def __init__(self, version: str):
        super().__init__()
        self.version = version


def load_and_freeze_clip_model(version: str) -> torch.nn.Module:
    try:
        import clip  # type: ignore
    except ImportError:
        return _NoTextModel(version)
    clip_model, _ = clip.load(version, device="cpu", jit=False)
    clip_model.eval()
    for p in clip_model.parameters():
        p.requires_grad = False
    return clip_model


def encode_text_clip(clip_model: torch.nn.Module, raw_text: List[str], max_length: int = 32, device="cpu") -> torch.Tensor:
    if _PROVIDER is not None:
        return _PROVIDER(list(raw_text)).to(device).detach()
    if isinstance(clip_model, _NoTextModel):
        raise RuntimeError("CLIP is not installed: register a feature provider with "
                           "models.functions.set_text_feature_provider(fn) (fn(list[str]) -> [B,512]).")
    import clip  # type: ignore
    miss = [s for s in dict.fromkeys(raw_text) if (s, max_length) not in _CACHE]
    if miss:
        if max_length is not None:
            ctx = max_length + 2
            assert ctx < 77
            toks = clip.tokenize(miss, context_length=ctx, truncate=True).to(device)
            toks = torch.cat([toks, torch.zeros([toks.shape[0], 77 - ctx], dtype=toks.dtype, device=toks.device)], dim=1)
        else:
            toks = clip.tokenize(miss, truncate=True).to(device)
        with torch.no_grad():
            enc = clip_model.encode_text(toks).detach()
        for s, e in zip(miss, enc):
            _CACHE[(s, max_length)] = e
    return torch.stack([_CACHE[(s, max_length)] for s in raw_text]).to(device)


def get_lang_feat_dim_type(model_name: str):
    if model_name == "bert-base-uncased":
        return 768, "bert"
    if model_name == "ViT-B/32":
        return 512, "clip"
    if model_name == "ViT-L/14@336px":
        return 768, "clip"
    raise NotImplementedError(model_name)
