"""Default reference-path / reference-speed parameters (numerical data as shipped by the reference,
gops/env/env_ocp/resources/ref_traj_data.py:19-37)."""
import math

DEFAULT_PATH_PARAM = {
    "sine": {"A": 1.5, "omega": 2 * math.pi / 10, "phi": 0.0},
    "double_lane": {"t1": 5.0, "t2": 9.0, "t3": 14.0, "t4": 18.0, "y1": 0.0, "y2": 3.5},
    "triangle": {"A": 3.0, "T": 10.0},
    "circle": {"r": 100.0},
}
DEFAULT_SPEED_PARAM = {
    "sine": {"A": 1.0, "omega": 2 * math.pi / 10, "phi": 0.0, "b": 5.0},
    "constant": {"u": 5.0},
}
