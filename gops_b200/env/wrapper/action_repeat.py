"""ActionRepeatModel (reference: gops/env/wrapper/action_repeat.py:54-87).  Not yet supported by the
fused kernels: constructing it raises instead of silently computing elsewhere."""
from gops_b200.env.wrapper.base import ModelWrapper


class ActionRepeatModel(ModelWrapper):
    def __init__(self, model, repeat_num=1, sum_reward=True):
        raise NotImplementedError("gops_b200: repeat_num is not supported by the fused kernels yet")
