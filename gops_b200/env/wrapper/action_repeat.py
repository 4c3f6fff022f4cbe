"""ActionRepeatModel (reference: gops/env/wrapper/action_repeat.py:54-87): the masked model step is applied
`repeat_num` times with the same action; rewards are summed (or only the last one kept).  The reference does not
advance `done` / `info` between repeats (line 84) -- for the state==obs models that is exactly "repeat the
dynamics", which the fused kernel implements; for the vehicle models (state in `info`) it is not built."""
from gops_b200.env.wrapper.base import ModelWrapper


class ActionRepeatModel(ModelWrapper):
    def __init__(self, model, repeat_num: int = 1, sum_reward: bool = True):
        super().__init__(model)
        self.repeat_num, self.sum_reward = int(repeat_num), bool(sum_reward)

    def describe(self, cfg):
        cfg["repeat_num"], cfg["sum_reward"] = self.repeat_num, int(self.sum_reward)
