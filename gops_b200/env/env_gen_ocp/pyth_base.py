"""State containers of the env_gen_ocp family (reference: gops/env/env_gen_ocp/pyth_base.py:14-141):
`State(robot_state, ContextState(reference, constraint, t))`; only what the model path needs."""
from dataclasses import dataclass, fields
from typing import Generic, Optional, TypeVar, Union

import numpy as np
import torch

stateType = TypeVar("stateType", np.ndarray, torch.Tensor)


def _map(obj, fn):
    vals = []
    for f in fields(obj):
        v = getattr(obj, f.name)
        vals.append(fn(v) if isinstance(v, (np.ndarray, torch.Tensor)) else v)
    return obj.__class__(*vals)


@dataclass
class ContextState(Generic[stateType]):
    reference: stateType
    constraint: Optional[stateType] = None
    t: Union[int, stateType] = 0

    def array2tensor(self):
        return _map(self, lambda v: torch.from_numpy(v) if isinstance(v, np.ndarray) else v)

    def tensor2array(self):
        return _map(self, lambda v: v.numpy() if isinstance(v, torch.Tensor) else v)

    def cuda(self):
        return _map(self, lambda v: v.cuda() if isinstance(v, torch.Tensor) else v)

    def __getitem__(self, index):
        return _map(self, lambda v: v[index])

    def index_by_t(self):
        vals = []
        for f in fields(self):
            v = getattr(self, f.name)
            if f.name == "t":
                vals.append(0)
            elif isinstance(v, (np.ndarray, torch.Tensor)) and v.ndim > 2:
                vals.append(v[np.arange(v.shape[0]), self.t])
            else:
                vals.append(v)
        return self.__class__(*vals)


@dataclass
class State(Generic[stateType]):
    robot_state: stateType
    context_state: ContextState[stateType]

    def array2tensor(self):
        return State(torch.from_numpy(self.robot_state), self.context_state.array2tensor())

    def tensor2array(self):
        return State(self.robot_state.numpy(), self.context_state.tensor2array())

    def cuda(self):
        return State(self.robot_state.cuda(), self.context_state.cuda())

    def __getitem__(self, index):
        return State(robot_state=self.robot_state[index], context_state=self.context_state[index])

    def __len__(self):
        return 1 if self.robot_state.ndim == 1 else self.robot_state.shape[0]
