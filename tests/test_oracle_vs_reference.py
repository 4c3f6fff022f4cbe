"""CPU: the oracle restatement against the LIVE unmodified reference on fresh inputs (beyond the committed golden
vectors).  Runs wherever the reference tree is reachable: /root/reference in the build container, oracle/_ref (the
copy made by oracle/build_ref.py) on the GPU box; skipped otherwise."""
import numpy as np
import pytest
import torch

from golden_util import rel_l2
from oracle import gops_oracle as orc
from oracle import ref_runner, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not reachable")


@pytest.mark.parametrize("H,B,seed", [(30, 64, 101), (7, 33, 102)])
def test_fhadp_idpendulum_fresh_inputs(H, B, seed):
    torch.set_num_threads(4)
    torch.manual_seed(seed)
    alg = ref_runner.create_reference_alg(ref_runner.c1_kwargs(H))
    data = orc.sample_inputs("pyth_idpendulum", B, seed=seed)
    data["done"][::9] = 1.0
    pi = alg.networks.policy.pi
    layers = [(pi[j].weight.detach().clone().requires_grad_(True), pi[j].bias.detach().clone().requires_grad_(True))
              for j in (0, 2, 4)]
    pol = orc.NetSpec(layers, "gelu", "linear", torch.ones(1), -torch.ones(1), time_input=True)
    loss = orc.fhadp_loss(pol, orc.create_env_model("pyth_idpendulum", reward_scale=1.0), data, H)
    loss.backward()
    tb = alg.local_update({k: v.clone() for k, v in data.items()}, 0)
    ref_loss = tb["Loss/Actor loss-RL iter"]
    assert abs(loss.item() - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss))
    ref_g = [p.grad.numpy() for p in alg.networks.policy.parameters()]
    assert rel_l2([p.grad.numpy() for p in pol.params()], ref_g) < 1e-5
