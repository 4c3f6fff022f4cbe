"""FHADP with an interior-point (log-barrier) treatment of the constraints (reference
gops/algorithm/fhadp_interior.py:22-99): samples whose whole rollout is feasible pay the log barrier / penalty, the
others the exterior penalty * penalty.  Fused kernel: csrc/kernel.cuh, cstr_mode 3."""
__all__ = ["FHADPInterior"]

from gops_b200.algorithm.fhadp import ApproxContainer   # noqa: F401  (registry contract)
from gops_b200.algorithm.fhadp_exterior import MODE_INTERIOR, FHADPExterior


class FHADPInterior(FHADPExterior):
    _mode = MODE_INTERIOR

    def _extra_tb(self, feasible_ratio: float):
        self.tb_info["Loss/Feasible ratio-RL iter"] = feasible_ratio
