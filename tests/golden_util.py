"""Helpers shared by the parity tests: load golden cases written by oracle/make_golden.py and
re-evaluate them with the CPU oracle (oracle/gops_oracle.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import gops_oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

_SC46 = (1.0 + 0.5 * torch.rand(46, generator=torch.Generator().manual_seed(7))).tolist()
_SH46 = (0.2 * torch.rand(46, generator=torch.Generator().manual_seed(8)) - 0.1).tolist()

# name -> (env_id, algorithm, hidden_act, model kwargs, wrapper kwargs, alg kwargs)
CASES = {
    "fhadp_idp_h30": ("pyth_idpendulum", "FHADP", "gelu", {}, dict(reward_scale=1.0), dict(pre_horizon=30)),
    "fhadp_idp_h80": ("pyth_idpendulum", "FHADP", "gelu", {}, dict(reward_scale=1.0), dict(pre_horizon=80)),
    "fhadp_idp_h80_nomask": ("pyth_idpendulum", "FHADP", "gelu", {}, dict(reward_scale=1.0, mask_at_done=False),
                             dict(pre_horizon=80)),
    "fhadp_idp_trained_h80": ("pyth_idpendulum", "FHADP", "gelu", {}, dict(reward_scale=1.0), dict(pre_horizon=80)),
    "fhadp_idp_relu_done": ("pyth_idpendulum", "FHADP", "relu", {}, dict(reward_scale=0.5, reward_shift=1.0),
                            dict(pre_horizon=12, gamma=0.97)),
    "infadp_lq_s4a2": ("pyth_lq", "INFADP", "gelu", dict(lq_config="s4a2"), dict(reward_scale=1.0, reward_shift=0.0),
                       dict(policy_lr=8e-4, value_lr=3e-4)),
    "infadp_lq_s4a2_n40": ("pyth_lq", "INFADP", "gelu", dict(lq_config="s4a2"), dict(reward_scale=1.0, reward_shift=0.0),
                           dict(forward_step=40, tau=0.2, gamma=0.97, policy_lr=8e-4, value_lr=3e-4)),
    "fhadp_lq_s3a1_clip": ("pyth_lq", "FHADP", "elu", dict(lq_config="s3a1"), dict(reward_scale=0.1), dict(pre_horizon=20)),
    "infadp_idp": ("pyth_idpendulum", "INFADP", "relu", {}, dict(reward_scale=1.0), {}),
    "infadp_veh3dofconti": ("pyth_veh3dofconti", "INFADP", "relu", dict(pre_horizon=10), {}, {}),
    "fhadp_veh3dofconti_p12": ("pyth_veh3dofconti", "FHADP", "elu", dict(pre_horizon=12), {}, dict(pre_horizon=12)),
    "fhadp_veh3dof_tracking_p10": ("veh3dof_tracking", "FHADP", "elu", dict(pre_horizon=10), {}, dict(pre_horizon=10)),
    "fhadp_veh3dof_tracking_p60_w256": ("veh3dof_tracking", "FHADP", "elu", dict(pre_horizon=60), {}, dict(pre_horizon=60)),
    "fhadp_idp_obsscale": ("pyth_idpendulum", "FHADP", "gelu", {},
                           dict(reward_scale=0.2, obs_scale=[0.5, 2.0, 2.0, 1.0, 0.25, 0.5],
                                obs_shift=[0.1, 0.0, -0.05, 0.0, 0.2, 0.0]), dict(pre_horizon=15)),
    "infadp_lq_obsscale_repeat3": ("pyth_lq", "INFADP", "tanh", dict(lq_config="s4a2"),
                                   dict(reward_scale=0.1, obs_scale=[2.0, 0.5, 1.5, 1.0], repeat_num=3), {}),
    "fhadp_idp_repeat2_last": ("pyth_idpendulum", "FHADP", "elu", {},
                               dict(reward_scale=1.0, repeat_num=2, sum_reward=False), dict(pre_horizon=6)),
    "infadp_veh3dofconti_obsscale": ("pyth_veh3dofconti", "INFADP", "relu", dict(pre_horizon=10),
                                     dict(obs_scale=_SC46, obs_shift=_SH46), {}),
}
DEFAULT_LR = {"fhadp_idp_h30": 1e-4, "fhadp_idp_h80": 1e-4, "fhadp_idp_trained_h80": 1e-4, "fhadp_idp_h80_nomask": 1e-4}


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def inputs_from(rec, env_id, dtype=torch.float32):
    data = {}
    for k, v in rec.items():
        if not k.startswith("in_"):
            continue
        data[k[3:]] = torch.from_numpy(np.asarray(v))
    if env_id == "veh3dof_tracking":
        data["state"] = (data.pop("robot_state"), data.pop("reference"), int(data.pop("t")))
    if env_id == "veh3dof_tracking_detour":
        data["state"] = (data.pop("robot_state"), data.pop("reference"), int(data.pop("t")), data.pop("surr"))
    if dtype != torch.float32:
        def cast(x):
            if isinstance(x, tuple):
                return tuple(cast(e) for e in x)
            return x.to(dtype) if torch.is_tensor(x) and x.is_floating_point() else x
        data = {k: cast(v) for k, v in data.items()}
    return data


def net_from(rec, prefix, net, hidden_act, dtype=torch.float32, requires_grad=False):
    """prefix e.g. 'init/' ; net in {'policy','v','v_target','policy_target'}."""
    seq = "pi" if net.startswith("policy") else "v"
    layers = []
    j = 0
    while f"{prefix}{net}.{seq}.{j}.weight" in rec:
        w = torch.tensor(rec[f"{prefix}{net}.{seq}.{j}.weight"], dtype=dtype).requires_grad_(requires_grad)
        b = torch.tensor(rec[f"{prefix}{net}.{seq}.{j}.bias"], dtype=dtype).requires_grad_(requires_grad)
        layers.append((w, b))
        j += 2
    hi = lo = None
    if net.startswith("policy"):
        hi = torch.tensor(rec[f"{prefix}{net}.act_high_lim"], dtype=dtype)
        lo = torch.tensor(rec[f"{prefix}{net}.act_low_lim"], dtype=dtype)
    return orc.NetSpec(layers, hidden_act, "linear", hi, lo, time_input=False)


def oracle_eval(name, iteration=0, dtype=torch.float32, prefix=None):
    """Re-evaluate a golden case with the oracle. Returns dict(loss, grads{key:tensor}, extra)."""
    env_id, alg, act, mk, wk, ak = CASES[name]
    rec = load(name)
    prefix = prefix or ("init/" if iteration == 0 else f"it{iteration - 1}/post/")
    env = orc.create_env_model(env_id, dtype=dtype, **mk, **wk)
    data = inputs_from(rec, env_id, dtype)
    pol = net_from(rec, prefix, "policy", act, dtype, requires_grad=True)
    if alg == "FHADP":
        pol.time_input = True
        loss = orc.fhadp_loss(pol, env, data, ak["pre_horizon"], ak.get("gamma", 1.0))
        loss.backward()
        return dict(loss=loss.item(), grads=_grads("policy", "pi", pol), rec=rec)
    n, gamma = ak.get("forward_step", 10), ak.get("gamma", 0.99)
    v = net_from(rec, prefix, "v", act, dtype, requires_grad=True)
    vt = net_from(rec, prefix, "v_target", act, dtype)
    if iteration % 2 == 0:
        loss, vmean = orc.infadp_loss_value(v, pol, vt, env, data, n, gamma)
        loss.backward()
        return dict(loss=loss.item(), vmean=vmean.item(), grads=_grads("v", "v", v), rec=rec)
    loss = orc.infadp_loss_policy(pol, vt, env, data, n, gamma)
    loss.backward()
    return dict(loss=loss.item(), grads=_grads("policy", "pi", pol), rec=rec)


def _grads(net, seq, spec):
    g = {}
    for j, (w, b) in enumerate(spec.layers):
        g[f"{net}.{seq}.{2 * j}.weight"] = w.grad
        g[f"{net}.{seq}.{2 * j}.bias"] = b.grad
    return g


def rel_l2(a, b):
    a = np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in a])
    b = np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in b])
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def slice_data(data, lo, hi):
    """Rows [lo, hi) of an oracle input dict (veh3dof_tracking carries (robot_state, reference, t) under 'state')."""
    out = {}
    for k, v in data.items():
        if isinstance(v, tuple):
            out[k] = (v[0][lo:hi], v[1][lo:hi], v[2])
        elif torch.is_tensor(v) and v.dim() > 0:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def oracle_chunked(loss_of_chunk, data, params, chunk=32768):
    """Batch-mean loss and gradient of the CPU oracle evaluated in chunks (the loss is a mean over independent samples,
    so it is the size-weighted sum of chunk means): keeps autograd memory bounded at BASELINE batch sizes.
    `loss_of_chunk(d)` returns the chunk-mean loss tensor or a tuple (loss, *extras); extras are size-weighted too."""
    B = data["obs"].shape[0]
    for p in params:
        p.grad = None
    total, extras = 0.0, None
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        out = loss_of_chunk(slice_data(data, lo, hi))
        loss = out[0] if isinstance(out, tuple) else out
        w = (hi - lo) / B
        (loss * w).backward()
        total += float(loss) * w
        if isinstance(out, tuple):
            ex = [float(e) * w for e in out[1:]]
            extras = ex if extras is None else [a + b for a, b in zip(extras, ex)]
    return total, [p.grad.detach().clone() for p in params], extras
