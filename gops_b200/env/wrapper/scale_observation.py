"""ScaleObservationModel (reference: gops/env/wrapper/scale_observation.py:72-116).  Not yet
supported by the fused kernels: constructing it raises instead of silently computing elsewhere."""
from gops_b200.env.wrapper.base import ModelWrapper


class ScaleObservationModel(ModelWrapper):
    def __init__(self, model, shift=0.0, scale=1.0):
        raise NotImplementedError("gops_b200: obs_shift / obs_scale are not supported by the fused kernels yet")
