"""Run a few updates of one secondary configuration (for ncu launch lists): python tools/run_config.py c3|dsac [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gops_b200.create_pkg.create_alg import create_alg  # noqa: E402
from gops_b200.trainer import device_sampler as ds  # noqa: E402

which, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
dev = torch.device("cuda")
if which == "c3":
    alg = create_alg(**bench.alg_kwargs("veh3dof_tracking", "FHADP", 256, "elu", 6 + 4 * 60, 2, pre_horizon=60,
                                        policy_learning_rate=1e-3))
    data = ds.sample_veh3dof_tracking(8192, 60, dev, seed=4)
    step = lambda i: alg.local_update(data, i)
else:
    from gops_b200.trainer.device_buffer import DeviceReplayBuffer
    alg = create_alg(**bench.alg_kwargs("pyth_idpendulum", "DSAC", 256, "gelu", 6, 1, policy_func_name="StochaPolicy",
                                        policy_hidden_sizes=[256, 256, 256], value_hidden_sizes=[256, 256, 256],
                                        policy_act_distribution="TanhGaussDistribution", policy_min_log_std=-20,
                                        policy_max_log_std=1, value_func_name="ActionValueDistri", value_learning_rate=3e-4,
                                        policy_learning_rate=3e-4, alpha_learning_rate=5e-5, gamma=0.99, tau=0.005,
                                        auto_alpha=True, alpha=0.2, delay_update=2, TD_bound=10, bound=True))
    buf = DeviceReplayBuffer(6, 1, 1 << 18, device=dev, seed=1)
    o = ds.sample_idpendulum(1 << 16, dev, seed=2)["obs"]
    buf.add_batch({"obs": o, "act": torch.rand(1 << 16, 1, device=dev) * 2 - 1, "rew": torch.randn(1 << 16, device=dev),
                   "obs2": o + 0.01 * torch.randn_like(o), "done": torch.zeros(1 << 16, device=dev)})
    step = lambda i: alg.local_update(buf.sample_batch(8192), i)
for i in range(n):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    step(i)
e1.record()
torch.cuda.synchronize()
print(which, "ms/update", e0.elapsed_time(e1) / n)
