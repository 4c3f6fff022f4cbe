#!/usr/bin/env python
"""Small end-to-end updates for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_configs import kwargs, to_dev
from gops_b200.create_pkg.create_alg import create_alg
from oracle import gops_oracle as orc

torch.manual_seed(0)
which = sys.argv[1:] or ["idp", "lq", "veh", "wide", "tc"]
if "idp" in which:
    alg = create_alg(**kwargs("pyth_idpendulum", "FHADP", 6, 1, 64, "gelu", pre_horizon=3, reward_scale=1.0))
    for B in (700, 130):     # cfg1/cfg2 tiles incl. ragged tails
        alg.local_update(to_dev(orc.sample_inputs("pyth_idpendulum", B, 1)), 0)
if "lq" in which:
    alg = create_alg(**kwargs("pyth_lq", "INFADP", 4, 2, 64, "relu", lq_config="s4a2", reward_scale=1.0))
    alg.set_parameters({"forward_step": 3})
    d = to_dev(orc.sample_inputs("pyth_lq", 300, 2, lq_config="s4a2"))
    alg.local_update(d, 0)
    alg.local_update(d, 1)
if "veh" in which:
    alg = create_alg(**kwargs("pyth_veh3dofconti", "INFADP", 46, 2, 64, "relu", pre_horizon=10))
    alg.set_parameters({"forward_step": 2})
    d = to_dev(orc.sample_inputs("pyth_veh3dofconti", 200, 3, pre_horizon=10))
    alg.local_update(d, 0)
    alg.local_update(d, 1)
if "wide" in which:
    alg = create_alg(**kwargs("veh3dof_tracking", "FHADP", 46, 2, 256, "elu", pre_horizon=10))
    alg.set_parameters({"pre_horizon": 2})
    alg.local_update(to_dev(orc.sample_inputs("veh3dof_tracking", 70, 4, pre_horizon=10)), 0)
if "tc" in which:          # tcgen05 / TMEM kernels: hybrid rollout (forced) and batched inference incl. a ragged tail
    os.environ["GOPS_B200_ROLLOUT"] = "tc"
    alg = create_alg(**kwargs("pyth_idpendulum", "FHADP", 6, 1, 64, "gelu", pre_horizon=3, reward_scale=1.0))
    for B in (700, 130):
        alg.local_update(to_dev(orc.sample_inputs("pyth_idpendulum", B, 1)), 0)
    alg = create_alg(**kwargs("pyth_lq", "INFADP", 4, 2, 64, "relu", lq_config="s4a2", reward_scale=1.0))
    alg.set_parameters({"forward_step": 3})
    d = to_dev(orc.sample_inputs("pyth_lq", 300, 2, lq_config="s4a2"))
    alg.local_update(d, 0)
    alg.local_update(d, 1)
    os.environ.pop("GOPS_B200_ROLLOUT")
    os.environ["GOPS_B200_INFER"] = "tc"
    alg.networks.policy(torch.randn(4321, 4, device="cuda"))
    alg.networks.v(torch.randn(129, 4, device="cuda"))
    os.environ.pop("GOPS_B200_INFER")
torch.cuda.synchronize()
print("sanitize_case done")
