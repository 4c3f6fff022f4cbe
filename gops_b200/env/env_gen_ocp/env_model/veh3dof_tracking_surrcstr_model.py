"""env_gen_ocp 3-DoF vehicle tracking with a surrounding-vehicle collision constraint, model type (reference:
gops/env/env_gen_ocp/env_model/veh3dof_tracking_surrcstr_model.py:13-181).  Same structure as veh3dof_tracking_detour
(one surrounding vehicle in ContextState.constraint, four extra observation entries, bicircle constraint of the incoming
state) with the tracking model's reward / termination bound and circle radius sqrt(2)/2 * veh_width (:88)."""
from gops_b200.env.env_gen_ocp.env_model.veh3dof_tracking_detour_model import Veh3DoFTrackingDetourModel


class Veh3DoFTrackingSurrCstrModel(Veh3DoFTrackingDetourModel):
    VARIANT = 2
    RADIUS_FACTOR = 2 ** 0.5 / 2      # r = np.sqrt(2) / 2 * veh_width (veh3dof_tracking_surrcstr_model.py:88)


def env_model_creator(**kwargs) -> Veh3DoFTrackingSurrCstrModel:
    return Veh3DoFTrackingSurrCstrModel(**kwargs)
