#!/usr/bin/env python
"""FHADP on the inverted double pendulum, trained entirely on the GPU (counterpart of the reference's
example_train/fhadp/fhadp_mlp_idpendulum_serial.py: same hyper-parameters, default horizon 80)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gops_b200.create_pkg.create_alg import create_alg
from gops_b200.trainer.device_trainer import DeviceStateSampler, OnDeviceSerialTrainer

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pre_horizon", type=int, default=80)
    ap.add_argument("--replay_batch_size", type=int, default=4096)
    ap.add_argument("--max_iteration", type=int, default=5000)
    ap.add_argument("--policy_learning_rate", type=float, default=1e-4)
    ap.add_argument("--save_folder", type=str, default=None)
    ap.add_argument("--seed", type=int, default=12345)
    args = vars(ap.parse_args())
    torch.manual_seed(args["seed"])
    kw = dict(env_id="pyth_idpendulum", algorithm="FHADP", trainer="off_serial_trainer", use_gpu=True,
              action_type="continu", obsv_dim=6, action_dim=1, action_high_limit=np.ones(1, np.float32),
              action_low_limit=-np.ones(1, np.float32), policy_func_name="FiniteHorizonPolicy", policy_func_type="MLP",
              policy_hidden_sizes=[64, 64], policy_hidden_activation="gelu", policy_act_distribution="default",
              value_func_type="MLP", reward_scale=1.0, **args)
    alg = create_alg(**kw)
    trainer = OnDeviceSerialTrainer(alg, DeviceStateSampler("pyth_idpendulum", "cuda", args["seed"]), log_save_interval=250,
                                    **args)
    trainer.train()
    for it, tb in trainer.history:
        print(it, {k: round(v, 4) for k, v in tb.items()})
