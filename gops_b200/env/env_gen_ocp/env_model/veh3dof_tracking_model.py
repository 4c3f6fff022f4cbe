"""env_gen_ocp 3-DoF vehicle tracking, model type (reference:
gops/env/env_gen_ocp/env_model/veh3dof_tracking_model.py:11-102, EnvModel.forward
env_model/pyth_base_model.py:109-119, robot step robot/veh3dof_model.py:24-58).  The reference
trajectory is a tensor [B, 2P+1, 4] indexed by a shared integer t; kernel: ModelVehTrack."""
import math
from typing import Union

import torch

from gops_b200 import _lib
from gops_b200.env.env_gen_ocp.pyth_base import ContextState, State
from gops_b200.env.env_ocp.env_model.pyth_base_model import PythBaseModel


class Veh3DoFTrackingModel(PythBaseModel):
    MODEL_KIND = _lib.MODEL_VEH3DOF_TRACKING

    def __init__(self, pre_horizon: int = 10, max_acc: float = 3.0, max_steer: float = math.pi / 6,
                 device: Union[torch.device, str, None] = None, **kwargs):
        self.pre_horizon = pre_horizon
        super().__init__(obs_dim=6 + 4 * pre_horizon, action_dim=2, dt=0.1,
                         action_lower_bound=[-max_steer, -max_acc], action_upper_bound=[max_steer, max_acc],
                         device=device)

    def fill_plan_desc(self, desc):
        desc.model = self.MODEL_KIND
        desc.veh_pre_horizon = self.pre_horizon

    def fill_batch(self, batch, info, f32, keep):
        state = info.get("state")
        if not isinstance(state, State):
            raise KeyError("veh3dof_tracking: info['state'] must be a gops_b200 State(robot_state, ContextState)")
        robot, ref = f32(state.robot_state), f32(state.context_state.reference)
        keep += [robot, ref]
        batch.state, batch.reference = robot.data_ptr(), ref.data_ptr()
        batch.ref_t = int(state.context_state.t)
        batch.ref_len = int(ref.shape[1])


    def alloc_next_info(self, B, dev):
        return {"state": torch.empty((B, 6), dtype=torch.float32, device=dev)}

    def make_next_info(self, info, extra):
        st = info["state"]
        ctx = st.context_state
        return {"state": State(robot_state=extra["state"].to(st.robot_state.device),
                               context_state=ContextState(reference=ctx.reference, constraint=ctx.constraint,
                                                          t=ctx.t + 1))}


def env_model_creator(**kwargs) -> Veh3DoFTrackingModel:
    return Veh3DoFTrackingModel(**kwargs)
