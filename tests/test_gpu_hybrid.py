"""Hybrid rollout kernel (tcgen05 / TMEM forward sweep + mma.sync reverse sweep, GOPS_B200_ROLLOUT=hy; csrc/mlp_tc_fwd.cuh)
held to the same bars as the pure mma.sync kernel: golden vectors of the unmodified reference (loss, gradient, Adam
step), the fp64 oracle on ragged batches, degenerate shapes and the no-grad trace.  The cases are the ones of
test_gpu_parity.py whose policy the tcgen05 forward is built for (64-wide, <= 16 inputs, state == obs models)."""
import numpy as np
import pytest
import torch

import test_gpu_parity as base
from golden_util import CASES

pytestmark = pytest.mark.gpu

TC_GOLDEN = [n for n in base.GOLDEN
             if CASES[n][0] in ("pyth_idpendulum", "pyth_lq") and "w256" not in n]


@pytest.fixture(autouse=True)
def _force_hy(monkeypatch):
    monkeypatch.setenv("GOPS_B200_ROLLOUT", "hy")


@pytest.mark.parametrize("name", TC_GOLDEN)
def test_hy_golden_loss_grad_update(name):
    base.test_golden_loss_grad_update(name)


@pytest.mark.parametrize("env_id,algname,act,B,H", [
    ("pyth_idpendulum", "FHADP", "gelu", 3000, 30),
    ("pyth_idpendulum", "FHADP", "tanh", 777, 7),
    ("pyth_idpendulum", "INFADP", "elu", 2048, 10),
    ("pyth_lq", "INFADP", "gelu", 5000, 10),
    ("pyth_lq", "FHADP", "selu", 1000, 25),
    ("pyth_lq", "INFADP", "sigmoid", 130, 3),
])
def test_hy_against_oracle_fp64(env_id, algname, act, B, H):
    base.test_against_oracle_fp64(env_id, algname, act, B, H)


@pytest.mark.parametrize("B,H", [(1, 1), (1, 5), (17, 1), (129, 2), (513, 3)])
def test_hy_edge_shapes(B, H):
    base.test_edge_shapes_against_oracle(B, H)


def test_hy_trace_matches_reference_rollout():
    base.test_trace_matches_reference_rollout()


def test_hy_path_is_taken_and_deterministic(monkeypatch):
    """Same inputs: hybrid and mma.sync gradients agree to tolerance but not bitwise (the tcgen05 forward rounds
    differently), and the hybrid path reproduces itself bit for bit."""
    alg, rec = base.build_alg("fhadp_idp_h30")
    data = base.data_from(rec, "pyth_idpendulum")

    def grads():
        alg._compute_gradient(data)
        torch.cuda.synchronize()
        return np.concatenate([p.grad.detach().cpu().numpy().ravel() for p in alg.networks.policy.parameters()])

    g_hy, g_hy2 = grads(), grads()
    monkeypatch.setenv("GOPS_B200_ROLLOUT", "mma")
    g_mma = grads()
    assert np.array_equal(g_hy, g_hy2)
    assert not np.array_equal(g_hy, g_mma)
    assert np.linalg.norm(g_hy - g_mma) <= 2e-4 * np.linalg.norm(g_mma)
