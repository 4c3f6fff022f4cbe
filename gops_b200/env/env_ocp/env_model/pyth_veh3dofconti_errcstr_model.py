"""3-DoF vehicle tracking with tracking-error constraints, model type (reference:
gops/env/env_ocp/env_model/pyth_veh3dofconti_errcstr_model.py:19-55): `pyth_veh3dofconti` whose `forward` also returns
info["constraint"] = (|y_err| - y_error_tol, |u_err| - u_error_tol) of the INCOMING observation -- the constraint
provider of the constrained FHADP variants (fhadp_exterior / fhadp_lagrangian / fhadp_interior)."""
from typing import Any, Dict, Optional, Union

import torch

from gops_b200.env.env_ocp.env_model.pyth_veh3dofconti_model import Veh3dofcontiModel


class Veh3dofcontiErrCstrModel(Veh3dofcontiModel):
    def __init__(self, pre_horizon: int, device: Union[torch.device, str, None] = None,
                 path_para: Optional[Dict[str, Dict]] = None, u_para: Optional[Dict[str, Dict]] = None,
                 y_error_tol: float = 0.2, u_error_tol: float = 2.0, **kwargs: Any):
        super().__init__(pre_horizon, device, path_para, u_para)
        self.y_error_tol, self.u_error_tol = y_error_tol, u_error_tol

    def fill_plan_desc(self, desc):
        super().fill_plan_desc(desc)
        desc.veh_errcstr = 1
        desc.veh_y_error_tol, desc.veh_u_error_tol = float(self.y_error_tol), float(self.u_error_tol)

    def get_constraint(self, obs: torch.Tensor, info=None) -> torch.Tensor:
        return torch.stack((obs[:, 1].abs() - self.y_error_tol, obs[:, 3].abs() - self.u_error_tol), dim=1)

    def make_next_info(self, info, extra):
        next_info = super().make_next_info(info, extra)
        if "_obs_in" in extra:
            next_info["constraint"] = self.get_constraint(extra["_obs_in"])
        return next_info


def env_model_creator(**kwargs):
    """make env model `pyth_veh3dofconti_errcstr`"""
    return Veh3dofcontiErrCstrModel(**kwargs)
