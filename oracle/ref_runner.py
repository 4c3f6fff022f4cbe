"""Run the UNMODIFIED reference (through oracle/ref_shim.py) for the bench's reference arm and the integration tests.

TEST / MEASUREMENT INFRASTRUCTURE: only tests/, bench.py's reference / cpu_baseline / gpu_eager legs and smoke() may
import this.  The reference is found at /root/reference (build container) or oracle/_ref (GPU box, oracle/build_ref.py).
"""
import time

import numpy as np
import torch

from oracle import ref_shim


def available() -> bool:
    return ref_shim.available()


def c1_kwargs(H=30, use_gpu=False, algorithm="FHADP"):
    """example_train/fhadp/fhadp_mlp_idpendulum_serial.py:34-75 with pre_horizon = H (the BASELINE C1 config)."""
    return dict(env_id="pyth_idpendulum", algorithm=algorithm, pre_horizon=H, seed=0, trainer="off_serial_trainer",
                cnn_shared=False, use_gpu=use_gpu, action_type="continu", obsv_dim=6, action_dim=1,
                action_high_limit=np.ones(1, dtype=np.float32), action_low_limit=-np.ones(1, dtype=np.float32),
                policy_func_name="FiniteHorizonPolicy", policy_func_type="MLP", policy_hidden_sizes=[64, 64],
                policy_hidden_activation="gelu", policy_act_distribution="default", policy_learning_rate=1e-4,
                value_func_type="MLP", reward_scale=1.0)


def create_reference_alg(kwargs, device="cpu"):
    """gops/create_pkg/create_alg.py:60-97 of the reference; `use_gpu` places env-model constants
    (create_env_model.py:95) and the caller moves the networks (off_serial_trainer.py:52-55 does `networks.cuda()`)."""
    ref_shim.install()
    from gops.create_pkg.create_alg import create_alg
    alg = create_alg(**kwargs)
    if device != "cpu":
        alg.networks.to(device)
    return alg


def time_reference_updates(batch, steps, warmup, threads=None, H=30, device="cpu", seed=1):
    """env-steps/s of the reference's own FHADP.local_update (fhadp.py:87-125 + base.py:94-98) on synthetic C1
    inputs.  device='cpu': wall clock on `threads` torch threads; device='cuda': PyTorch eager on the GPU
    (the 'existing Blackwell path' of SURVEY 8(d)), timed with a device synchronize on both sides."""
    if threads:
        torch.set_num_threads(threads)
    cuda = device != "cpu"
    torch.manual_seed(0)
    alg = create_reference_alg(c1_kwargs(H, use_gpu=cuda), device)
    g = torch.Generator().manual_seed(seed)
    h = torch.tensor([5, 0.1, 0.1, 0.3, 0.3, 0.3])
    data = {"obs": (torch.rand(batch, 6, generator=g) * 2 - 1) * h, "done": torch.zeros(batch)}
    if cuda:
        data = {k: v.to(device) for k, v in data.items()}
    times = []
    for i in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        alg.local_update(data, i)
        if cuda:
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * H * len(times) / total, total / len(times)
