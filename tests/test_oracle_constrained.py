"""CPU: the oracle's constrained-FHADP restatement (fhadp_constrained_loss) on pyth_veh3dofconti_errcstr and on the
env_gen_ocp veh3dof_tracking_detour model against the unmodified reference (tests/golden/cstr_*.npz, detour_*.npz:
fhadp / fhadp_exterior / fhadp_lagrangian / fhadp_interior, two updates each)."""
import numpy as np
import pytest
import torch

from golden_util import inputs_from, load, net_from, rel_l2
from oracle import gops_oracle as orc

MODES = {"cstr_fhadpexterior": ("exterior", (2.0, 3.0)), "cstr_fhadpinterior": ("interior", (2.0, 3.0)),
         "cstr_fhadplagrangian": ("lagrangian", None),
         "detour_fhadpexterior": ("exterior", (2.0, 3.0)), "detour_fhadpinterior": ("interior", (2.0, 3.0)),
         "detour_fhadplagrangian": ("lagrangian", None), "surrcstr_fhadpexterior": ("exterior", (2.0, 3.0))}


@pytest.mark.parametrize("name", sorted(MODES))
def test_constrained_losses_and_gradients(name):
    torch.set_num_threads(4)
    mode, coefs = MODES[name]
    rec = load(name)
    if name.startswith("detour") or name.startswith("surrcstr"):
        env = orc.create_env_model("veh3dof_tracking_" + name.split("_")[0], pre_horizon=10)
        data = inputs_from(rec, "veh3dof_tracking_detour")
    else:
        env = orc.create_env_model("pyth_veh3dofconti_errcstr", pre_horizon=10, y_error_tol=1.2, u_error_tol=2.2)
        data = inputs_from(rec, "pyth_veh3dofconti")
    for it in (0, 1):
        prefix = "init/" if it == 0 else "it0/post/"
        pol = net_from(rec, prefix, "policy", "elu", requires_grad=True)
        pol.time_input = True
        coef = coefs[it] if coefs else float(rec[f"it{it}/tb/Loss/Lagrange multiplier-RL iter"])
        loss, l_r, l_c, feas = orc.fhadp_constrained_loss(mode, pol, env, data, 10, 0.97, coef)
        loss.backward()
        tb = {k.split("/tb/")[1]: float(v) for k, v in rec.items() if k.startswith(f"it{it}/tb/")}
        assert abs(loss.item() - tb["Loss/Actor loss-RL iter"]) <= 2e-6 * max(1.0, abs(loss.item()))
        assert abs(l_r.item() - tb["Loss/Actor reward loss-RL iter"]) <= 2e-6
        assert abs(l_c.item() - tb["Loss/Actor constraint loss-RL iter"]) <= 2e-6
        if mode == "interior":
            assert abs(float(feas) - tb["Loss/Feasible ratio-RL iter"]) < 1e-6
        keys = [f"it{it}/grad/policy.pi.{2 * j}.{w}" for j in range(3) for w in ("weight", "bias")]
        got = [t.grad.numpy() for pair in pol.layers for t in pair]
        assert rel_l2(got, [rec[k] for k in keys]) < 1e-5


def test_detour_plain_fhadp():
    """FHADP on the detour model: the constraint is provided but unused; reward weights, termination bound and the four
    surrounding-vehicle observation entries differ from veh3dof_tracking."""
    torch.set_num_threads(4)
    rec = load("detour_fhadp")
    env = orc.create_env_model("veh3dof_tracking_detour", pre_horizon=10)
    data = inputs_from(rec, "veh3dof_tracking_detour")
    for it in (0, 1):
        pol = net_from(rec, "init/" if it == 0 else "it0/post/", "policy", "elu", requires_grad=True)
        pol.time_input = True
        loss = orc.fhadp_loss(pol, env, data, 10, 0.97)
        loss.backward()
        ref = float(rec[f"it{it}/tb/Loss/Actor loss-RL iter"])
        assert abs(loss.item() - ref) <= 2e-6 * max(1.0, abs(ref))
        keys = [f"it{it}/grad/policy.pi.{2 * j}.{w}" for j in range(3) for w in ("weight", "bias")]
        assert rel_l2([t.grad.numpy() for pair in pol.layers for t in pair], [rec[k] for k in keys]) < 1e-5
