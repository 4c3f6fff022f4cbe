"""3-DoF vehicle trajectory tracking, model type (reference:
gops/env/env_ocp/env_model/pyth_veh3dofconti_model.py:65-203 with the analytic reference generator
gops/env/env_ocp/resources/ref_traj_model.py).  Dynamics, reference generation, ego-frame observation
and their adjoints run in gops_b200/csrc/models_veh.cuh (ModelVehConti)."""
from copy import deepcopy
from typing import Dict, Optional, Union

import numpy as np
import torch

from gops_b200 import _lib
from gops_b200.env.env_ocp.env_model.pyth_base_model import PythBaseModel
from gops_b200.env.env_ocp.resources.ref_traj_data import DEFAULT_PATH_PARAM, DEFAULT_SPEED_PARAM


def fill_reftraj(rt: _lib.RefTraj, path_param: dict, speed_param: dict):
    s, d, tr, c = path_param["sine"], path_param["double_lane"], path_param["triangle"], path_param["circle"]
    rt.sine_A, rt.sine_omega, rt.sine_phi = s["A"], s["omega"], s["phi"]
    rt.dl_t1, rt.dl_t2, rt.dl_t3, rt.dl_t4, rt.dl_y1, rt.dl_y2 = d["t1"], d["t2"], d["t3"], d["t4"], d["y1"], d["y2"]
    rt.tri_A, rt.tri_T, rt.circ_r = tr["A"], tr["T"], c["r"]
    sp, cs = speed_param["sine"], speed_param["constant"]
    rt.sp_A, rt.sp_omega, rt.sp_phi, rt.sp_b, rt.sp_const = sp["A"], sp["omega"], sp["phi"], sp["b"], cs["u"]


class Veh3dofcontiModel(PythBaseModel):
    MODEL_KIND = _lib.MODEL_VEH3DOFCONTI

    def __init__(self, pre_horizon: int = 10, device: Union[torch.device, str, None] = None,
                 path_para: Optional[Dict[str, Dict]] = None, u_para: Optional[Dict[str, Dict]] = None,
                 max_steer: float = np.pi / 6, **kwargs):
        self.pre_horizon = pre_horizon
        super().__init__(obs_dim=6 + 4 * pre_horizon, action_dim=2, dt=0.1,
                         action_lower_bound=[-max_steer, -3], action_upper_bound=[max_steer, 3], device=device)
        self.path_param = deepcopy(DEFAULT_PATH_PARAM)
        for k, v in (path_para or {}).items():
            self.path_param[k].update(v)
        self.speed_param = deepcopy(DEFAULT_SPEED_PARAM)
        for k, v in (u_para or {}).items():
            self.speed_param[k].update(v)

    def fill_plan_desc(self, desc):
        desc.model = self.MODEL_KIND
        desc.veh_pre_horizon = self.pre_horizon
        fill_reftraj(desc.reftraj, self.path_param, self.speed_param)

    def fill_batch(self, batch, info, f32, keep):
        for key, field in (("state", "state"), ("ref_points", "ref_points"), ("path_num", "path_num"),
                           ("u_num", "u_num"), ("ref_time", "ref_time")):
            if key not in info:
                raise KeyError(f"pyth_veh3dofconti: info['{key}'] is required")
            tns = f32(info[key])
            keep.append(tns)
            setattr(batch, field, tns.data_ptr())
        if keep[-4].shape[1:] != (self.pre_horizon + 1, 4):
            raise RuntimeError(f"ref_points must be [B, {self.pre_horizon + 1}, 4], got {tuple(keep[-4].shape)}")


    def alloc_next_info(self, B, dev):
        return {"state": torch.empty((B, 6), dtype=torch.float32, device=dev),
                "ref_points": torch.empty((B, self.pre_horizon + 1, 4), dtype=torch.float32, device=dev),
                "ref_time": torch.empty(B, dtype=torch.float32, device=dev)}

    def make_next_info(self, info, extra):
        next_info = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in info.items()}
        dev = info["state"].device
        next_info.update({"state": extra["state"].to(dev), "ref_points": extra["ref_points"].to(dev),
                          "path_num": info["path_num"], "u_num": info["u_num"], "ref_time": extra["ref_time"].to(dev)})
        return next_info


def env_model_creator(**kwargs):
    """make env model `pyth_veh3dofconti`"""
    return Veh3dofcontiModel(**kwargs)
