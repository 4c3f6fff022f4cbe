"""Env-model factory (reference: gops/create_pkg/create_env_model.py:28-147): registry filled by a
directory scan of gops_b200/env/env_*/env_model/, same `create_env_model` signature and wrapper
assembly order (innermost first: MaskAtDone, ActionRepeat, ShapingReward, ScaleObservation,
ClipObservation, ClipAction, ScaleAction)."""
import importlib
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Union

import numpy as np

from gops_b200.env.wrapper.action_repeat import ActionRepeatModel
from gops_b200.env.wrapper.clip_action import ClipActionModel
from gops_b200.env.wrapper.clip_observation import ClipObservationModel
from gops_b200.env.wrapper.mask_at_done import MaskAtDoneModel
from gops_b200.env.wrapper.scale_action import ScaleActionModel
from gops_b200.env.wrapper.scale_observation import ScaleObservationModel
from gops_b200.env.wrapper.shaping_reward import ShapingRewardModel
from gops_b200.utils.gops_path import env_path, underline2camel


@dataclass
class Spec:
    env_id: str
    entry_point: Callable
    kwargs: dict = field(default_factory=dict)


registry: Dict[str, Spec] = {}


def register(env_id: str, entry_point: Union[Callable, str], **kwargs):
    registry[env_id] = Spec(env_id=env_id, entry_point=entry_point, kwargs=kwargs)


def create_env_model(
    env_id: str,
    *,
    reward_shift: Optional[float] = None,
    reward_scale: Optional[float] = None,
    obs_shift: Union[np.ndarray, float, list, None] = None,
    obs_scale: Union[np.ndarray, float, list, None] = None,
    clip_obs: bool = True,
    clip_action: bool = True,
    mask_at_done: bool = True,
    repeat_num: Optional[int] = None,
    sum_reward: bool = True,
    action_scale: bool = True,
    min_action: Union[float, int, np.ndarray, list] = -1.0,
    max_action: Union[float, int, np.ndarray, list] = 1.0,
    **kwargs,
) -> object:
    spec_ = registry.get(env_id + "_model")
    if spec_ is None:
        raise KeyError(f"No registered env with id: {env_id}_model")
    _kwargs = spec_.kwargs.copy()
    _kwargs.update(kwargs)
    _kwargs["device"] = "cuda"          # this package computes on the GPU only
    if not callable(spec_.entry_point):
        raise RuntimeError(f"{spec_.env_id} registered but entry_point is not specified")
    env_model = spec_.entry_point(**_kwargs)

    if mask_at_done:
        env_model = MaskAtDoneModel(env_model)
    if repeat_num is not None:
        env_model = ActionRepeatModel(env_model, repeat_num, sum_reward)
    if reward_scale is not None or reward_shift is not None:
        reward_scale = 1.0 if reward_scale is None else reward_scale
        reward_shift = 0.0 if reward_shift is None else reward_shift
        env_model = ShapingRewardModel(env_model, reward_shift, reward_scale)
    if obs_shift is not None or obs_scale is not None:
        obs_scale = 1.0 if obs_scale is None else obs_scale
        obs_shift = 0.0 if obs_shift is None else obs_shift
        env_model = ScaleObservationModel(env_model, obs_shift, obs_scale)
    if clip_obs:
        env_model = ClipObservationModel(env_model)
    if clip_action:
        env_model = ClipActionModel(env_model)
    if action_scale:
        env_model = ScaleActionModel(env_model, min_action, max_action)
    return env_model


def _scan():
    for env_dir_name in sorted(e for e in os.listdir(env_path) if e.startswith("env_")):
        env_model_path = os.path.join(env_path, env_dir_name, "env_model")
        if not os.path.isdir(env_model_path):
            continue
        for file in sorted(os.listdir(env_model_path)):
            if file.endswith(".py") and file[0] != "_" and "base" not in file:
                env_id = file[:-3]
                mdl = importlib.import_module(f"gops_b200.env.{env_dir_name}.env_model.{env_id}")
                camel = underline2camel(env_id)
                if hasattr(mdl, "env_model_creator"):
                    register(env_id=env_id, entry_point=getattr(mdl, "env_model_creator"))
                elif hasattr(mdl, camel):
                    register(env_id=env_id, entry_point=getattr(mdl, camel))
                else:
                    print(f"env {env_id} has no env_model_creator or {camel} in {env_dir_name}")


_scan()
