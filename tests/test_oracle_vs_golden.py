"""Pins the CPU oracle (oracle/gops_oracle.py) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  Tolerances: loss 2e-6 relative (both sides are fp32 CPU torch
with the same op order, so they agree to round-off), gradients 1e-5 relative L2."""
import numpy as np
import pytest
import torch

from golden_util import CASES, DEFAULT_LR, load, oracle_eval, rel_l2, net_from, inputs_from
from oracle import gops_oracle as orc

FAST = [n for n in CASES if n not in ("fhadp_veh3dof_tracking_p60_w256", "fhadp_idp_h80")]


def _loss_key(rec, it):
    keys = [k for k in rec if k.startswith(f"it{it}/tb/")]
    if it % 2 == 0 and any("Critic loss" in k for k in keys):
        return next(k for k in keys if "Critic loss" in k)
    return next(k for k in keys if "Actor loss" in k)


@pytest.mark.parametrize("name", FAST + ["fhadp_idp_h80"])
def test_loss_and_grads(name):
    torch.set_num_threads(4)
    its = [0, 1] if CASES[name][1] == "INFADP" else [0]
    for it in its:
        out = oracle_eval(name, it)
        rec = out["rec"]
        ref_loss = float(rec[_loss_key(rec, it)])
        assert abs(out["loss"] - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss)), (name, it, out["loss"], ref_loss)
        keys = sorted(out["grads"])
        err = rel_l2([out["grads"][k].numpy() for k in keys], [rec[f"it{it}/grad/{k}"] for k in keys])
        assert err < 1e-5, (name, it, err)


def test_wide_net_tracking_case():
    out = oracle_eval("fhadp_veh3dof_tracking_p60_w256")
    rec = out["rec"]
    ref_loss = float(rec[_loss_key(rec, 0)])
    assert abs(out["loss"] - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss))
    keys = sorted(out["grads"])
    assert rel_l2([out["grads"][k].numpy() for k in keys], [rec[f"it0/grad/{k}"] for k in keys]) < 1e-5


def test_trace_matches_reference_rollout():
    name = "fhadp_idp_h30"
    env_id, _, act, mk, wk, ak = CASES[name]
    rec = load(name)
    env = orc.create_env_model(env_id, **mk, **wk)
    pol = net_from(rec, "init/", "policy", act)
    pol.time_input = True
    trace = []
    with torch.no_grad():
        orc.fhadp_loss(pol, env, inputs_from(rec, env_id), ak["pre_horizon"], 1.0, trace=trace)
    n = rec["trace_obs"].shape[1]
    for k, (o, a, r, d) in enumerate(trace):
        np.testing.assert_allclose(o[:n].numpy(), rec["trace_obs"][k], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(a[:n].numpy(), rec["trace_act"][k], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(r[:n].numpy(), rec["trace_rew"][k], rtol=1e-5, atol=1e-4)
        assert (d[:n].numpy() == rec["trace_done"][k]).all()


@pytest.mark.parametrize("name", ["fhadp_idp_h30", "infadp_lq_s4a2", "infadp_veh3dofconti"])
def test_adam_and_polyak(name):
    env_id, alg, act, mk, wk, ak = CASES[name]
    its = [0, 1] if alg == "INFADP" else [0]
    states = {}
    for it in its:
        out = oracle_eval(name, it)
        rec = out["rec"]
        net = "policy" if (alg == "FHADP" or it % 2 == 1) else "v"
        seq = "pi" if net == "policy" else "v"
        prefix = "init/" if it == 0 else f"it{it - 1}/post/"
        lr = ak.get("policy_lr" if net == "policy" else "value_lr", DEFAULT_LR.get(name, 1e-3))
        keys = [k for k in sorted(out["grads"])]
        params = [torch.tensor(rec[prefix + k]) for k in keys]
        new = orc.adam_step(params, [out["grads"][k] for k in keys], states.setdefault(net, {}), lr)
        for k, p in zip(keys, new):
            np.testing.assert_allclose(p.numpy(), rec[f"it{it}/post/{k}"], rtol=0, atol=2e-7)
        if alg == "INFADP":
            tgt_keys = [k.replace(net + ".", net + "_target.", 1) for k in keys]
            tgt = orc.polyak([torch.tensor(rec[prefix + k]) for k in tgt_keys], new, ak.get("tau", 0.005))
            for k, p in zip(tgt_keys, tgt):
                np.testing.assert_allclose(p.numpy(), rec[f"it{it}/post/{k}"], rtol=0, atol=2e-7)


def test_checkpoint_known_answer():
    rec = load("ckpt_fhadp_idp")
    pol = net_from(rec, "sd/", "policy", "gelu")
    pol.time_input = True
    env = orc.create_env_model("pyth_idpendulum", reward_scale=1.0)
    o, d = torch.tensor(rec["obs0"]), torch.zeros(1)
    acts = []
    with torch.no_grad():
        for _ in range(5):
            a = pol.act(o, 1)
            acts.append(a.numpy()[0])
            o, r, d, _ = env.forward(o, a, d, {})
        loss = orc.fhadp_loss(pol, env, {"obs": torch.tensor(rec["obs0"]), "done": torch.zeros(1)}, 80)
    np.testing.assert_allclose(np.stack(acts), rec["closed_loop_actions"], rtol=1e-5, atol=1e-6)
    assert abs(loss.item() - float(rec["loss_h80"])) < 1e-5 * abs(float(rec["loss_h80"]))
