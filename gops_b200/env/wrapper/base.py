"""Base wrappers for model-type environments (reference: gops/env/wrapper/base.py:23-97).

A wrapper here records its constants (`describe`) for the fused kernels; `forward` on ANY level of
the chain evaluates that level and everything below it in one `gops_b200_model_step` launch."""
from typing import Tuple

import torch

from gops_b200.utils.gops_typing import InfoDict


class ModelWrapper:
    def __init__(self, model):
        self.model = model
        self.forward_flag = False

    def describe(self, cfg: dict):
        """Write this wrapper's constants into the chain description (see gops_b200/env/fused.py)."""

    def forward(self, obs: torch.Tensor, action: torch.Tensor, done: torch.Tensor, info: InfoDict
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, InfoDict]:
        from gops_b200.env.fused import fused_forward
        return fused_forward(self, obs, action, done, info)

    def __getattr__(self, name):
        if name in ("model", "__setstate__", "__getstate__"):
            raise AttributeError(name)
        return getattr(self.model, name)

    @property
    def unwrapped(self):
        return self.model.unwrapped


class ActionModelWrapper(ModelWrapper):
    def action(self, action: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError
