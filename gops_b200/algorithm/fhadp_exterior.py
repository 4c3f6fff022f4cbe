"""FHADP with an exterior penalty on constraint violations (reference gops/algorithm/fhadp_exterior.py:20-78).

loss = -mean(v_r) + penalty * mean(sum_k gamma^k sum_i max(c_i, 0)^2); the penalty grows by `penalty_increase` every
`penalty_delay` updates up to `max_penalty`.  The constraint term, its gradient through the observation model and the
MaskAtDone-frozen observations are part of the fused rollout kernel (csrc/kernel.cuh, cstr_mode 1)."""
__all__ = ["FHADPExterior"]

from typing import Tuple

import torch

from gops_b200.algorithm.fhadp import ApproxContainer, FHADP   # noqa: F401  (ApproxContainer: registry contract)
from gops_b200.utils.tensorboard_setup import tb_tags

MODE_EXTERIOR, MODE_LAGRANGIAN, MODE_INTERIOR = 1, 2, 3


class _ConstrainedFHADP(FHADP):
    """Shared plumbing: hand mode + coefficient of THIS update to the plan, split the scalar tail into the tb tags."""

    _mode = MODE_EXTERIOR

    def _coef(self) -> float:
        raise NotImplementedError

    def _after_update(self, loss_constraint: float):
        pass

    def _launch_gradient(self, data) -> torch.Tensor:
        from gops_b200 import _lib
        pol = self.networks.policy
        plan = self._plan(_lib.ALG_FHADP, pol, None, self.pre_horizon, self.gamma)
        self._coef_used = float(self._coef())
        _lib.check(_lib.lib().gops_b200_plan_set_constraint(plan.handle, self._mode, self._coef_used))
        self._batch = data["obs"].shape[0] * self._world()[1]
        return self._rollout_grad(plan, data, pol.flat_params, pol.flat_params, None, None)

    def _publish(self, tail: torch.Tensor, start_time: float):
        import time
        loss, c_ext, third, n_feasible = self._tail_to_host(tail)
        coef = self._coef_used
        if self._mode == MODE_INTERIOR:
            loss_reward = loss - third - coef * c_ext          # third = mean(v_int * feasible) / penalty
        else:
            loss_reward = loss - coef * c_ext
        self.tb_info[tb_tags["loss_actor"]] = loss
        self.tb_info[tb_tags["loss_actor_reward"]] = loss_reward
        self.tb_info[tb_tags["loss_actor_constraint"]] = c_ext
        self._extra_tb(n_feasible / self._batch)
        self._after_update(c_ext)
        self.tb_info[tb_tags["alg_time"]] = (time.time() - start_time) * 1000  # ms

    def _extra_tb(self, feasible_ratio: float):
        pass


class FHADPExterior(_ConstrainedFHADP):
    def __init__(self, *, pre_horizon: int, gamma: float = 1.0, penalty: float = 1.0, penalty_increase: float = 1.1,
                 penalty_delay: float = 100, max_penalty: float = 1e3, index: int = 0, **kwargs):
        super().__init__(pre_horizon=pre_horizon, gamma=gamma, index=index, **kwargs)
        self.penalty, self.penalty_increase = penalty, penalty_increase
        self.penalty_delay, self.max_penalty = penalty_delay, max_penalty
        self.update_step = 0

    @property
    def adjustable_parameters(self) -> Tuple[str]:
        return (*super().adjustable_parameters, "penalty", "penalty_increase", "penalty_delay")

    def _coef(self) -> float:
        return self.penalty

    def _after_update(self, loss_constraint: float):
        self.update_step += 1
        if self.update_step % self.penalty_delay == 0:
            self.penalty = min(self.penalty * self.penalty_increase, self.max_penalty)
        self.tb_info["Loss/Penalty coefficient-RL iter"] = self.penalty    # the reference logs it AFTER the increase
