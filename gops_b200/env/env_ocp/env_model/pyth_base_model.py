"""Base class of model-type environments (reference: gops/env/env_ocp/env_model/pyth_base_model.py:21-85).

Objects are descriptors: they carry dimensions, bounds and physical constants and know how to
write themselves into a kernel plan; `forward` launches the fused single-step kernel."""
from abc import ABCMeta
from typing import Callable, Optional, Sequence, Tuple, Union

import torch

from gops_b200.utils.gops_typing import InfoDict


def _default_device(device):
    if device in (None, "cuda"):
        return "cuda" if torch.cuda.is_available() else "cpu"
    return device


class PythBaseModel(metaclass=ABCMeta):
    MODEL_KIND = -1

    def __init__(self, obs_dim: int, action_dim: int, dt: Optional[float] = None,
                 obs_lower_bound: Optional[Sequence] = None, obs_upper_bound: Optional[Sequence] = None,
                 action_lower_bound: Optional[Sequence] = None, action_upper_bound: Optional[Sequence] = None,
                 device: Union[torch.device, str, None] = None):
        self.obs_dim, self.action_dim, self.dt = obs_dim, action_dim, dt
        inf = float("inf")
        device = _default_device(device)
        mk = lambda v, fill, n: torch.tensor(list(v) if v is not None else [fill] * n, dtype=torch.float32, device=device)
        self.obs_lower_bound = mk(obs_lower_bound, -inf, obs_dim)
        self.obs_upper_bound = mk(obs_upper_bound, inf, obs_dim)
        self.action_lower_bound = mk(action_lower_bound, -inf, action_dim)
        self.action_upper_bound = mk(action_upper_bound, inf, action_dim)
        self.device = device

    def forward(self, obs: torch.Tensor, action: torch.Tensor, done: torch.Tensor, info: InfoDict
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, InfoDict]:
        from gops_b200.env.fused import fused_forward
        return fused_forward(self, obs, action, done, info)

    get_constraint: Callable[[torch.Tensor, InfoDict], torch.Tensor] = None
    get_terminal_cost: Callable[[torch.Tensor], torch.Tensor] = None

    @property
    def unwrapped(self):
        return self

    # ---- kernel-plan plumbing -----------------------------------------------------------
    def fill_plan_desc(self, desc):
        desc.model = self.MODEL_KIND
        lo = self.obs_lower_bound.detach().cpu().tolist()
        hi = self.obs_upper_bound.detach().cpu().tolist()
        for i in range(min(len(lo), len(desc.obs_low))):
            desc.obs_low[i], desc.obs_high[i] = lo[i], hi[i]

    def fill_batch(self, batch, info, f32, keep):
        pass

    def alloc_next_info(self, B, dev) -> dict:
        return {}

    def make_next_info(self, info, extra) -> InfoDict:
        return {"constraint": None}
