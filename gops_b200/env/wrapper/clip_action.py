"""ClipActionModel (reference: gops/env/wrapper/clip_action.py:22-40): clip action to the model bounds."""
from gops_b200.env.wrapper.base import ModelWrapper


class ClipActionModel(ModelWrapper):
    def describe(self, cfg):
        cfg["clip_action"] = 1
