"""Text-encoder hooks with the reference's signatures (/root/reference/models/functions.py:46-94).

CLIP itself is out of scope (third-party frozen tower, no weights offline): the call site is kept,
results are cached per unique string (the reference re-encodes the same prompts on every denoise step,
cmdm.py:133-135), and a feature provider can be registered where CLIP is not installed (bench / tests).
"""
from collections import OrderedDict
from typing import Callable, List, Optional

import torch

_PROVIDER: Optional[Callable[[List[str]], torch.Tensor]] = None
_CACHE = OrderedDict()  # (id(text model), prompt, max_length) -> feature; LRU, bounded
_CACHE_MAX = 4096


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
    mk = id(clip_model)  # per text tower: two models with different CLIP versions in one process must not share entries
    miss = [s for s in dict.fromkeys(raw_text) if (mk, s, max_length) not in _CACHE]
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
            _CACHE[(mk, s, max_length)] = e
    out = torch.stack([_CACHE[(mk, s, max_length)] for s in raw_text]).to(device)
    for s in raw_text:  # LRU order; bounded so a training run over many captions does not pin every feature on the GPU
        _CACHE.move_to_end((mk, s, max_length))
    while len(_CACHE) > _CACHE_MAX:
        _CACHE.popitem(last=False)
    return out


def get_lang_feat_dim_type(model_name: str):
    if model_name == "bert-base-uncased":
        return 768, "bert"
    if model_name == "ViT-B/32":
        return 512, "clip"
    if model_name == "ViT-L/14@336px":
        return 768, "clip"
    raise NotImplementedError(model_name)


def load_scene_model(model_name: str, model_dim: int, num_points: int, pretrained_weight: str = None, freeze: bool = True) -> torch.nn.Module:
    """functions.py:96-126.  PointTransformerSeg is the scene model every CDM config names (configs/model/cdm.yaml:18);
    PointTransformerEnc is only used by the ContactPointTrans ablations, which are out of scope."""
    from models.scene_models.pointtransformer import pointtransformer_seg_repro
    if model_name != "PointTransformerSeg":
        raise NotImplementedError(model_name)
    scene_model = pointtransformer_seg_repro(c=model_dim, num_points=num_points)
    if pretrained_weight is not None:
        scene_model.load_pretrained_weight(weight_path=pretrained_weight)
    if freeze:
        scene_model.eval()
        for p in scene_model.parameters():
            p.requires_grad = False
    return scene_model
