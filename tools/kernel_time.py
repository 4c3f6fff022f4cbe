"""Kernel-only time of the fused rollout update for a few (env, activation, batch, horizon) points: one line each.
usage: python tools/kernel_time.py [idp|lq] act B H [path]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gops_b200 import _lib  # noqa: E402
from gops_b200.create_pkg.create_alg import create_alg  # noqa: E402
from gops_b200.trainer import device_sampler as ds  # noqa: E402


def run(env, act, B, H, path="auto", iters=5):
    lq = env == "lq"
    obs_dim, act_dim = (4, 2) if lq else (6, 1)
    kw = dict(env_id="pyth_lq" if lq else "pyth_idpendulum", algorithm="FHADP", seed=0, trainer="off_serial_trainer",
              use_gpu=True, action_type="continu", obsv_dim=obs_dim, action_dim=act_dim,
              action_high_limit=np.ones(act_dim, np.float32), action_low_limit=-np.ones(act_dim, np.float32),
              policy_func_name="FiniteHorizonPolicy", policy_func_type="MLP", policy_hidden_sizes=[64, 64],
              policy_hidden_activation=act, policy_act_distribution="default", policy_learning_rate=1e-4,
              value_func_type="MLP", reward_scale=1.0, pre_horizon=H)
    if lq:
        kw["lq_config"] = "s4a2"
    torch.manual_seed(0)
    alg = create_alg(**kw)
    alg.kernel_path = path
    data = ds.sample_lq(B, "s4a2", "cuda", 1) if lq else ds.sample_idpendulum(B, "cuda", 1)
    for _ in range(2):
        alg._compute_gradient(data)
    plan = next(iter(alg._plans.values()))
    _lib.check(_lib.lib().gops_b200_plan_enable_timing(plan.handle, 1))
    ts = []
    for _ in range(iters):
        alg._compute_gradient(data)
        ms = C.c_float()
        _lib.check(_lib.lib().gops_b200_plan_last_kernel_ms(plan.handle, C.byref(ms)))
        ts.append(ms.value)
    print(f"{env} {act} B={B} H={H} path={alg.last_kernel_path()} kernel_ms={min(ts):.3f} "
          f"env-steps/s={B * H / (min(ts) * 1e-3):.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] if len(sys.argv) > 5 else "auto")
    else:
        for act in ("gelu", "relu", "tanh", "elu"):
            run("idp", act, 1 << 18, 30)
        for B in (1 << 15, 1 << 16, 1 << 12, 256):
            run("idp", "gelu", B, 30)
        run("idp", "gelu", 1 << 18, 30, "mma")
        run("lq", "gelu", 1 << 18, 30)
