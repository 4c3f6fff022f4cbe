"""ScaleObservationModel (reference: gops/env/wrapper/scale_observation.py:72-116): the inner model sees
obs / scale - shift, the caller sees (next_obs + shift) * scale.  Fused into the rollout kernel."""
from typing import Union

import numpy as np
import torch

from gops_b200.env.wrapper.base import ModelWrapper


class ScaleObservationModel(ModelWrapper):
    def __init__(self, model, shift: Union[np.ndarray, float, list] = 0.0,
                 scale: Union[np.ndarray, float, list] = 1.0):
        super().__init__(model)
        dev = model.obs_lower_bound.device
        as_t = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev) if isinstance(v, (np.ndarray, list)) else v
        self.shift, self.scale = as_t(shift), as_t(scale)

    def describe(self, cfg):
        n = self.unwrapped.obs_dim
        full = lambda v: (torch.zeros(n) + (v.detach().cpu() if torch.is_tensor(v) else float(v))).numpy().astype(np.float32)
        cfg["obs_scaling"] = 1
        cfg["obs_scale"], cfg["obs_shift"] = full(self.scale), full(self.shift)
