"""ScaleActionModel (reference: gops/env/wrapper/scale_action.py:44-83): affine map of
[min_action, max_action] onto the model's action bounds, clipped on both sides."""
from typing import Union

import numpy as np
import torch

from gops_b200.env.wrapper.base import ActionModelWrapper


class ScaleActionModel(ActionModelWrapper):
    def __init__(self, model, min_action: Union[float, int, np.ndarray, list],
                 max_action: Union[float, int, np.ndarray, list]):
        super().__init__(model)
        lo = model.action_lower_bound
        as_t = lambda v: torch.as_tensor(v, dtype=lo.dtype, device=lo.device) if isinstance(v, (np.ndarray, list)) else v
        self.min_action = torch.zeros_like(lo) + as_t(min_action)
        self.max_action = torch.zeros_like(lo) + as_t(max_action)
        self.action_lower_bound = self.min_action
        self.action_upper_bound = self.max_action

    def describe(self, cfg):
        cfg["action_scale"] = 1
        cfg["min_action"] = self.min_action.detach().cpu().numpy()
        cfg["max_action"] = self.max_action.detach().cpu().numpy()

    def action(self, action: torch.Tensor) -> torch.Tensor:
        # host-side helper kept for API parity (tiny, not on the hot path)
        low, high = self.model.action_lower_bound, self.model.action_upper_bound
        action = torch.clip(action, self.min_action, self.max_action)
        action = low + (high - low) * ((action - self.min_action) / (self.max_action - self.min_action))
        return torch.clip(action, low, high)
