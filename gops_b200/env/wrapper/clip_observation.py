"""ClipObservationModel (reference: gops/env/wrapper/clip_observation.py:22-44): clip next_obs to the
model's observation bounds."""
from gops_b200.env.wrapper.base import ModelWrapper


class ClipObservationModel(ModelWrapper):
    def describe(self, cfg):
        cfg["clip_obs"] = 1
