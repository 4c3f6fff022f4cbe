"""ShapingRewardModel (reference: gops/env/wrapper/shaping_reward.py:53-88): r <- (r + shift) * scale."""
from gops_b200.env.wrapper.base import ModelWrapper


class ShapingRewardModel(ModelWrapper):
    def __init__(self, model, reward_shift=0.0, reward_scale=1.0):
        super().__init__(model)
        self.shift, self.scale = reward_shift, reward_scale

    def describe(self, cfg):
        cfg["reward_shaping"] = 1
        cfg["reward_shift"], cfg["reward_scale"] = float(self.shift), float(self.scale)
