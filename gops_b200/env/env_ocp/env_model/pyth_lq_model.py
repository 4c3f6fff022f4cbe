"""Linear-quadratic control, model type (reference: gops/env/env_ocp/resources/lq_base.py LqModel
:317-357, LQDynamics :35-141; creator gops/env/env_ocp/env_model/pyth_lq_model.py:18-34).
x' = inv(I - A dt) (B u dt + x);  r = rs * (rsh - (sum Q x^2 + sum R u^2)) -- kernel: ModelLq."""
from typing import Union

import numpy as np
import torch

from gops_b200 import _lib
from gops_b200.env.env_ocp.env_model.pyth_base_model import PythBaseModel
from gops_b200.env.env_ocp.resources import lq_configs


class LqModel(PythBaseModel):
    MODEL_KIND = _lib.MODEL_LQ

    def __init__(self, config: dict, device: Union[torch.device, str, None] = None):
        lb_state, hb_state = np.array(config["state_low"]), np.array(config["state_high"])
        lb_action, hb_action = np.array(config["action_low"]), np.array(config["action_high"])
        super().__init__(obs_dim=lb_state.shape[0], action_dim=lb_action.shape[0], dt=config["dt"],
                         obs_lower_bound=lb_state, obs_upper_bound=hb_state, action_lower_bound=lb_action,
                         action_upper_bound=hb_action, device=device)
        if self.obs_dim > _lib.MAX_LQ_N or self.action_dim > _lib.MAX_ACT:
            raise NotImplementedError("pyth_lq: fused kernels support n <= 8, m <= 4")
        self.config = config
        # fp32 pseudo-inverse of (I - A dt) exactly as LQDynamics.__init__ (:55-57); init-time only
        A = torch.as_tensor(config["A"], dtype=torch.float32)
        IA = torch.eye(self.obs_dim) - A * config["dt"]
        self.inv_IA = torch.linalg.pinv(IA)
        self.B = torch.as_tensor(config["B"], dtype=torch.float32)
        self.Q = torch.as_tensor(config["Q"], dtype=torch.float32)
        self.R = torch.as_tensor(config["R"], dtype=torch.float32)
        self.reward_scale, self.reward_shift = config["reward_scale"], config["reward_shift"]

    def fill_plan_desc(self, desc):
        super().fill_plan_desc(desc)
        n, m = self.obs_dim, self.action_dim
        desc.lq_n, desc.lq_m = n, m
        for i, v in enumerate(self.inv_IA.reshape(-1).tolist()):
            desc.lq_inv_IA[i] = v
        for i, v in enumerate(self.B.reshape(-1).tolist()):
            desc.lq_B[i] = v
        for i, v in enumerate(self.Q.tolist()):
            desc.lq_Q[i] = v
        for i, v in enumerate(self.R.tolist()):
            desc.lq_R[i] = v
        desc.lq_dt = float(self.dt)
        desc.lq_reward_scale, desc.lq_reward_shift = float(self.reward_scale), float(self.reward_shift)


def env_model_creator(**kwargs):
    """make env model `pyth_lq`"""
    lqc = kwargs.get("lq_config", None)
    if lqc is None:
        config = lq_configs.config_s3a1
    elif isinstance(lqc, str):
        assert hasattr(lq_configs, "config_" + lqc)
        config = getattr(lq_configs, "config_" + lqc)
    elif isinstance(lqc, dict):
        config = lqc
    else:
        raise RuntimeError("lq_config invalid")
    return LqModel(config, kwargs.get("device", None))
