#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs other than the headline bench line (C2, C3, C5),
timed through the plugin API (`alg.local_update`, CUDA-resident inputs, CUDA events, median of N).
Writes one JSON object per line; run on the GPU box:  python tools/bench_configs.py > gpurun_out/configs.jsonl"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import usable_cores
torch.set_num_threads(usable_cores())
from gops_b200.create_pkg.create_alg import create_alg
from gops_b200.trainer import device_sampler as ds


def kwargs(env_id, alg, obs_dim, act_dim, hid, act, **extra):
    kw = dict(env_id=env_id, algorithm=alg, seed=0, trainer="off_serial_trainer", use_gpu=True, action_type="continu",
              obsv_dim=obs_dim, action_dim=act_dim, action_high_limit=np.ones(act_dim, np.float32),
              action_low_limit=-np.ones(act_dim, np.float32),
              policy_func_name="FiniteHorizonPolicy" if alg == "FHADP" else "DetermPolicy", policy_func_type="MLP",
              policy_hidden_sizes=[hid, hid], policy_hidden_activation=act, policy_act_distribution="default",
              policy_learning_rate=1e-3, value_func_name="StateValue", value_func_type="MLP",
              value_hidden_sizes=[hid, hid], value_hidden_activation=act, value_learning_rate=1e-3)
    kw.update(extra)
    return kw


def time_updates(alg, data, iters, n=12, warm=3):
    """median ms of local_update for each iteration parity in `iters` (INFADP alternates PEV / PIM)."""
    out = {}
    for it in iters:
        ts = []
        for i in range(warm + n):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            alg.local_update(data, it)
            e1.record()
            torch.cuda.synchronize()
            if i >= warm:
                ts.append(e0.elapsed_time(e1))
        out[it] = statistics.median(ts)
    return out


def sample(env_id, batch, seed, pre_horizon=10, lq_config="s4a2"):
    """Synthetic update inputs drawn on the device (the reset laws of gops_b200/trainer/device_sampler.py)."""
    if env_id == "pyth_idpendulum":
        return ds.sample_idpendulum(batch, "cuda", seed)
    if env_id == "pyth_lq":
        return ds.sample_lq(batch, lq_config, "cuda", seed)
    if env_id in ("pyth_veh3dofconti", "pyth_veh3dofconti_errcstr"):
        return ds.sample_veh3dofconti(batch, pre_horizon, "cuda", seed)
    if env_id == "veh3dof_tracking":
        return ds.sample_veh3dof_tracking(batch, pre_horizon, "cuda", seed)
    if env_id == "veh3dof_tracking_detour":
        return ds.sample_veh3dof_tracking_detour(batch, pre_horizon, "cuda", seed)
    raise KeyError(env_id)


def main():
    torch.manual_seed(0)
    res = []
    # C2: INFADP veh3dofconti B=4096, P=10, forward_step=10, [64,64] relu
    alg = create_alg(**kwargs("pyth_veh3dofconti", "INFADP", 46, 2, 64, "relu", pre_horizon=10))
    data = sample("pyth_veh3dofconti", 4096, 1, pre_horizon=10)
    ms = time_updates(alg, data, [0, 1])
    res.append({"config": "C2 INFADP veh3dofconti B=4096 n=10 [64,64] relu", "ms_pev": ms[0], "ms_pim": ms[1],
                "env_steps_per_s_pev": 4096 * 10 / ms[0] * 1e3, "env_steps_per_s_pim": 4096 * 10 / ms[1] * 1e3})
    # C3: FHADP veh3dof_tracking H=P=60, [256,256] elu, 8192 samples per GPU (65536 over 8 GPUs)
    alg = create_alg(**kwargs("veh3dof_tracking", "FHADP", 246, 2, 256, "elu", pre_horizon=60))
    data = sample("veh3dof_tracking", 8192, 2, pre_horizon=60)
    ms = time_updates(alg, data, [0], n=5, warm=2)
    res.append({"config": "C3 FHADP veh3dof_tracking H=60 B=8192/GPU [256,256] elu", "ms": ms[0],
                "env_steps_per_s": 8192 * 60 / ms[0] * 1e3})
    # C5: INFADP LQ s4a2 batch sweep, forward_step 10, [64,64] gelu
    for logb in (10, 12, 14, 16, 18, 20):
        B = 1 << logb
        alg = create_alg(**kwargs("pyth_lq", "INFADP", 4, 2, 64, "gelu", lq_config="s4a2", reward_scale=1.0))
        data = sample("pyth_lq", B, 3)
        ms = time_updates(alg, data, [0, 1], n=8)
        res.append({"config": f"C5 INFADP LQ s4a2 B=2^{logb} n=10 [64,64] gelu", "ms_pev": ms[0], "ms_pim": ms[1],
                    "env_steps_per_s_pev": B * 10 / ms[0] * 1e3, "env_steps_per_s_pim": B * 10 / ms[1] * 1e3})
    # C1 small-batch point: the reference's own CPU-runnable case (B=256, H=30)
    alg = create_alg(**kwargs("pyth_idpendulum", "FHADP", 6, 1, 64, "gelu", pre_horizon=30, reward_scale=1.0))
    data = sample("pyth_idpendulum", 256, 4)
    ms = time_updates(alg, data, [0])
    res.append({"config": "C1 FHADP idpendulum B=256 H=30 (reference's CPU config)", "ms": ms[0],
                "env_steps_per_s": 256 * 30 / ms[0] * 1e3})
    for r in res:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
