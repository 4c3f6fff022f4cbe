#!/usr/bin/env python
"""Convergence probes on the GPU (system-level sanity of the fused update):
  * FHADP idpendulum H=30: actor loss must fall from ~+230 (pendulum drops) towards -10*H*...;
  * INFADP LQ s4a2: the learned policy approaches the discounted-LQR gain the reference computes in
    LQDynamics.compute_control_matrix (gops/env/env_ocp/resources/lq_base.py:61-71)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from gops_b200.create_pkg.create_alg import create_alg
from gops_b200.trainer.device_trainer import DeviceStateSampler, OnDeviceSerialTrainer

sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_configs import kwargs


def lqr_gain(cfg, gamma=0.99):
    from scipy.linalg import solve_discrete_are
    A0, B0 = np.array(cfg["A"], dtype=np.float64), np.array(cfg["B"], dtype=np.float64)
    dt = cfg["dt"]
    A = np.linalg.pinv(np.eye(A0.shape[0]) - A0 * dt) * np.sqrt(gamma)
    B = A @ B0 * dt
    Q, R = np.diag(np.array(cfg["Q"], dtype=np.float64)), np.diag(np.array(cfg["R"], dtype=np.float64))
    P = solve_discrete_are(A, B, Q, R)
    return np.linalg.pinv(R + B.T @ P @ B) @ B.T @ P @ A


def main():
    torch.manual_seed(0)
    t0 = time.time()
    alg = create_alg(**kwargs("pyth_idpendulum", "FHADP", 6, 1, 64, "gelu", pre_horizon=30, reward_scale=1.0,
                              policy_learning_rate=3e-4))
    tr = OnDeviceSerialTrainer(alg, DeviceStateSampler("pyth_idpendulum", "cuda", 1), replay_batch_size=4096,
                               max_iteration=int(os.environ.get("IDP_ITERS", 3000)), log_save_interval=250)
    tr.train()
    print("FHADP idpendulum H=30:", [(it, round(tb["Loss/Actor loss-RL iter"], 2)) for it, tb in tr.history],
          f"{time.time() - t0:.1f}s")

    from gops_b200.env.env_ocp.resources import lq_configs
    t0 = time.time()
    kw = kwargs("pyth_lq", "INFADP", 4, 2, 64, "gelu", lq_config="s4a2", reward_scale=1.0)
    kw.update(policy_learning_rate=8e-4, value_learning_rate=3e-4)
    alg = create_alg(**kw)
    alg.set_parameters({"forward_step": int(os.environ.get("LQ_N", 50)), "tau": 0.2, "gamma": 0.99})
    tr = OnDeviceSerialTrainer(alg, DeviceStateSampler("pyth_lq", "cuda", 2, lq_config="s4a2"), replay_batch_size=1024,
                               max_iteration=int(os.environ.get("LQ_ITERS", 8000)), log_save_interval=1000)
    tr.train()
    K = lqr_gain(lq_configs.config_s4a2)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(512, 4, generator=g) * torch.tensor([0.7, 0.3, 0.7, 0.3]) * 0.5
    u = alg.networks.policy(x.cuda()).cpu().numpy() * 8.0          # ScaleAction: [-1,1] -> [-8,8]
    u_star = -(x.numpy() @ K.T)
    cos = np.sum(u * u_star) / (np.linalg.norm(u) * np.linalg.norm(u_star))
    rel = np.linalg.norm(u - u_star) / np.linalg.norm(u_star)
    print("INFADP LQ s4a2:", [(it, {k.split('/')[1][:12]: round(v, 3) for k, v in tb.items() if 'Time' not in k})
                              for it, tb in tr.history][-3:])
    print(f"policy vs LQR gain: cosine {cos:.4f}, rel.err {rel:.3f}, {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
