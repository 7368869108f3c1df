"""Model registry + factories — the drop-in boundary of /root/reference/models/base.py:7-83.

`Model` has the registry semantics of utils/registry.py:10-91 (register by class name, duplicate names assert,
`get` raises KeyError).  `create_model_and_diffusion(cfg, device=...)` reads the same `cfg.model.*` /
`cfg.diffusion.*` keys and returns `(nn.Module, SpacedDiffusion)`.
"""
from typing import Any, Dict, Iterator, Tuple

import torch.nn as nn


class Registry:
    def __init__(self, name: str) -> None:
        self._name = name
        self._obj_map: Dict[str, Any] = {}

    def _do_register(self, name: str, obj: Any) -> None:
        assert name not in self._obj_map, f"An object named '{name}' was already registered in '{self._name}' registry!"
        self._obj_map[name] = obj

    def register(self, obj: Any = None) -> Any:
        if obj is None:
            def deco(func_or_class: Any) -> Any:
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)

    def get(self, name: str) -> Any:
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name: str) -> bool:
        return name in self._obj_map

    def __iter__(self) -> Iterator[Tuple[str, Any]]:
        return iter(self._obj_map.items())

    def __repr__(self) -> str:
        return f"Registry of {self._name}: " + ", ".join(sorted(self._obj_map))


Model = Registry("model")


def create_model(cfg, *args, **kwargs) -> nn.Module:
    """base.py:9-18."""
    return Model.get(cfg.model.name)(cfg.model, *args, **kwargs)


def create_gaussian_diffusion(cfg, *args, **kwargs):
    """base.py:20-70: same config keys, same enum choices."""
    from diffusion import gaussian_diffusion as gd
    from diffusion.respace import SpacedDiffusion, space_timesteps

    dcfg = cfg.diffusion
    steps = dcfg.steps
    respacing = dcfg.timestep_respacing if dcfg.timestep_respacing else [steps]
    betas = gd.get_named_beta_schedule(dcfg.noise_schedule, steps)
    mean_type = gd.ModelMeanType.START_X if dcfg.predict_xstart else gd.ModelMeanType.EPSILON
    loss_type = {"MSE": gd.LossType.MSE, "RESCALED_MSE": gd.LossType.RESCALED_MSE, "KL": gd.LossType.KL,
                 "RESCALED_KL": gd.LossType.RESCALED_KL}[dcfg.loss_type]
    if dcfg.learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if dcfg.sigma_small else gd.ModelVarType.FIXED_LARGE
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, respacing), betas=betas, model_mean_type=mean_type,
                           model_var_type=var_type, loss_type=loss_type, rescale_timesteps=dcfg.rescale_timesteps)


def create_model_and_diffusion(cfg, *args, **kwargs):
    """base.py:72-83."""
    return create_model(cfg, *args, **kwargs), create_gaussian_diffusion(cfg, *args, **kwargs)
