"""FHADP with a learned Lagrange multiplier (reference gops/algorithm/fhadp_lagrangian.py:25-85):
loss = -mean(v_r) + softplus(multiplier_param) * mean(sum_k gamma^k sum_i max(c_i, 0)); every `multiplier_delay` updates
the multiplier parameter takes an Adam ascent step on the constraint term.  Fused kernel: csrc/kernel.cuh, cstr_mode 2."""
__all__ = ["FHADPLagrangian"]

import math
from typing import Tuple

import torch
import torch.nn as nn

from gops_b200.algorithm.fhadp import ApproxContainer   # noqa: F401  (registry contract)
from gops_b200.algorithm.fhadp_exterior import MODE_LAGRANGIAN, _ConstrainedFHADP
from gops_b200.utils.flat_params import ScalarAdam


class FHADPLagrangian(_ConstrainedFHADP):
    _mode = MODE_LAGRANGIAN

    def __init__(self, *, pre_horizon: int, gamma: float = 1.0, multiplier: float = 1.0, multiplier_lr: float = 1e-3,
                 multiplier_delay: int = 10, index: int = 0, **kwargs):
        super().__init__(pre_horizon=pre_horizon, gamma=gamma, index=index, **kwargs)
        # inverse of softplus
        self.multiplier_param = nn.Parameter(torch.tensor(math.log(math.exp(multiplier) - 1), dtype=torch.float32))
        self.multiplier_optim = ScalarAdam(self.multiplier_param, lr=multiplier_lr)
        self.multiplier_delay = multiplier_delay
        self.update_step = 0

    @property
    def adjustable_parameters(self) -> Tuple[str]:
        return (*super().adjustable_parameters, "multiplier", "multiplier_lr", "multiplier_delay")

    def _coef(self) -> float:
        return torch.nn.functional.softplus(self.multiplier_param).item()

    def _after_update(self, loss_constraint: float):
        self.tb_info["Loss/Lagrange multiplier-RL iter"] = self._coef_used
        self.update_step += 1
        if self.update_step % self.multiplier_delay == 0:
            self.multiplier_optim.grad = -loss_constraint        # d(-param * loss_constraint) / d param
            self.multiplier_optim.step()
