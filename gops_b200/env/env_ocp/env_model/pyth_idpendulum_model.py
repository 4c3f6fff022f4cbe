"""Inverted double pendulum, model type (reference: gops/env/env_ocp/env_model/pyth_idpendulum_model.py).
The dynamics (Dynamics.f_xu :31-124, rewards :126-149, done :151-172) and their adjoint are
implemented in gops_b200/csrc/models.cuh (ModelIdp)."""
from typing import Union

import numpy as np
import torch

from gops_b200 import _lib
from gops_b200.env.env_ocp.env_model.pyth_base_model import PythBaseModel


class PythInvertedpendulum(PythBaseModel):
    MODEL_KIND = _lib.MODEL_IDPENDULUM

    def __init__(self, device: Union[torch.device, str, None] = None):
        self.discrete_num = 5
        super().__init__(obs_dim=6, action_dim=1, dt=0.01, obs_lower_bound=[-np.inf] * 6,
                         obs_upper_bound=[np.inf] * 6, action_lower_bound=[-1.0], action_upper_bound=[1.0],
                         device=device)


def env_model_creator(**kwargs):
    """make env model `pyth_idpendulum`"""
    return PythInvertedpendulum(kwargs.get("device", None))
