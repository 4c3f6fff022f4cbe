"""Mixin giving apprfuncs `get_act_dist` (reference: gops/utils/act_distribution_cls.py)."""


class Action_Distribution:
    def __init__(self):
        super().__init__()

    def get_act_dist(self, logits):
        act_dist = getattr(self, "action_distribution_cls")(logits)
        if hasattr(self, "act_high_lim"):
            act_dist.act_high_lim = getattr(self, "act_high_lim")
            act_dist.act_low_lim = getattr(self, "act_low_lim")
        return act_dist
