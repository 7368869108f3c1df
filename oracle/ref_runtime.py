"""ORACLE (test infrastructure): runs the reference's OWN modules (staged under oracle/_ref by oracle/build_ref.py) on the CPU.

Two uses, never from the product path:
  * `reference_models()`  — import the reference's `models` / `diffusion` packages (for `bench.py --impl reference`, the
    `cpu_baseline` legs and oracle cross-checks).  Must run in a process that has NOT imported the drop-in packages of the same
    names (afford-motion_b200/models, /diffusion): bench.py's reference arm is such a process.
  * `load_reference_drivers()` — only the third-party stubs, so that the reference's DRIVERS (utils/training.py::TrainLoop,
    test.py) import and run against the drop-in `models` / `diffusion` (tests/test_gpu_dropin_drivers.py).

Stubs (none of these packages is installed offline): omegaconf (dict with attribute access), hydra (no-op
`main` decorator), clip, smplkit, natsort, and `pointops_cuda`, whose two live entry points (models/scene_models/pointops.py:23,42)
are served by the C restatement oracle/pointops_ref.c.  CLIP is replaced by a feature provider behind `encode_text_clip`.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


class DictConfig(dict):
    """Attribute-access dict standing in for omegaconf.DictConfig (the reference only reads `cfg.key`)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return DictConfig(v) if isinstance(v, dict) and not isinstance(v, DictConfig) else v

    def __setattr__(self, k, v):
        self[k] = v


def available() -> bool:
    from . import build_ref
    return build_ref.available()


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_third_party_stubs(cpu_pointops: bool = True):
    class _OmegaConf:
        @staticmethod
        def to_yaml(cfg):
            return repr(dict(cfg))

    _mod("omegaconf", DictConfig=DictConfig, OmegaConf=_OmegaConf)

    def _hydra_main(*a, **k):
        return lambda fn: fn
    _mod("hydra", main=_hydra_main)
    if cpu_pointops and "clip" not in sys.modules:
        _mod("clip")  # only for the reference's models/functions.py; the drop-in detects a missing `clip` by ImportError
    _mod("smplkit", SMPLXLayer=lambda **k: None)
    _mod("natsort", natsorted=sorted)
    try:
        import loguru  # noqa: F401
    except ImportError:
        import logging
        _mod("loguru", logger=logging.getLogger("reference"))
    if not cpu_pointops:
        return
    # pointops_cuda: the two live entry points, in-place fill like the pybind module (pointops.py:23,42)
    from . import pointops_ref

    def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
        idx.copy_(pointops_ref.furthestsampling(xyz, offset, new_offset))

    def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
        i, d = pointops_ref.knnquery(nsample, xyz, new_xyz, offset, new_offset)
        idx.copy_(i)
        dist2.copy_(d * d)  # the wrapper takes the sqrt itself (pointops.py:43)
    _mod("pointops_cuda", furthestsampling_cuda=furthestsampling_cuda, knnquery_cuda=knnquery_cuda)
    # the reference allocates with torch.cuda.{Int,Float}Tensor (pointops.py:21-22,40-41; pointtransformer.py:60): CPU stand-ins

    def _int(*a):
        return torch.IntTensor(*a)

    def _float(*a):
        return torch.FloatTensor(*a)
    torch.cuda.IntTensor = _int
    torch.cuda.FloatTensor = _float


def load_reference_drivers(datasets_base, evaluate_mod):
    """The reference's DRIVERS against the drop-in packages: returns (utils.training module, test.py module).
    `models` / `diffusion` must resolve to afford-motion_b200/ (put it first on sys.path); `utils.*` and test.py come from
    oracle/_ref.  `datasets_base` / `evaluate_mod` are caller-provided stand-ins for `datasets.base` (needs real datasets) and
    `utils.evaluate` (needs SMPL-X and evaluator checkpoints); `datasets.misc` is the reference's own collate code."""
    import importlib.util
    import random

    import numpy as np
    assert available(), "oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists"
    install_third_party_stubs(cpu_pointops=False)
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)  # AFTER the drop-in: only `utils` (absent from the drop-in) resolves here

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m
    sys.modules["datasets"] = types.ModuleType("datasets")
    sys.modules["datasets.base"] = datasets_base
    _load("datasets.misc", "datasets/misc.py")
    sys.modules["utils.evaluate"] = evaluate_mod
    import utils.training as rtrain  # the reference's TrainLoop / load_ckpt
    rtest = _load("reference_test_py", "test.py")
    # test.py imports torch / numpy / random only under `if __name__ == '__main__'` (test.py:132-136): provide its script globals
    rtest.torch, rtest.np, rtest.random = torch, np, random
    return rtrain, rtest


def reference_models(text_provider):
    """-> (models.base module, diffusion.gaussian_diffusion module) of the REFERENCE, with CLIP replaced by `text_provider`
    (fn(list[str]) -> [B,512])."""
    assert available(), "oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists"
    for k in list(sys.modules):
        if k in ("models", "diffusion", "utils") or k.startswith(("models.", "diffusion.", "utils.")):
            if not getattr(sys.modules[k], "__file__", "") or REF_ROOT not in (sys.modules[k].__file__ or ""):
                raise RuntimeError(f"oracle.ref_runtime.reference_models: the drop-in package '{k}' is already imported in this "
                                   "process; the reference's same-named packages need a fresh process")
    install_third_party_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import models.base as rbase  # noqa: E402  (the reference's)
    import models.cdm as rcdm
    import models.cmdm as rcmdm
    from diffusion import gaussian_diffusion as rgd

    def fake_encode(model, raw_text, max_length=32, device="cpu"):
        return text_provider(list(raw_text)).to(device)

    for m in (rcdm, rcmdm):
        m.load_and_freeze_clip_model = lambda v: torch.nn.Module()
        m.encode_text_clip = fake_encode
    return rbase, rgd


def full_cfg(model_cfg: dict, steps: int = 1000, timestep_respacing: str = "") -> DictConfig:
    """configs/default.yaml:31-40 + a model section, as the reference's create_model_and_diffusion expects."""
    return DictConfig(dict(model=dict(model_cfg), diffusion=dict(predict_xstart=True, steps=steps, noise_schedule="cosine",
                                                                  timestep_respacing=timestep_respacing, rescale_timesteps=False,
                                                                  loss_type="MSE", learn_sigma=False, sigma_small=True)))
