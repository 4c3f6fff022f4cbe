"""MaskAtDoneModel (reference: gops/env/wrapper/mask_at_done.py:21-40): next_obs/reward are masked by
the incoming done flag, next_done |= done."""
from gops_b200.env.wrapper.base import ModelWrapper


class MaskAtDoneModel(ModelWrapper):
    def describe(self, cfg):
        cfg["mask_at_done"] = 1
