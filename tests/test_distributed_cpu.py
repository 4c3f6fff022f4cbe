"""N>1 host logic on CPU (gloo, world_size 2): batch sharding + the single SUM all-reduce of the flat
[gradient | loss | ...] buffer reproduce the full-batch mean gradient.  The per-shard gradients come from the
CPU oracle here (the fused kernel needs a GPU); the code under test is gops_b200.algorithm.base
{shard_inv_batch, allreduce_flat} and the FlatParams gradient-buffer layout."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gops_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, H, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gops_b200.algorithm.base import allreduce_flat, shard_inv_batch
    from gops_b200.utils.flat_params import GRAD_TAIL
    gen = torch.Generator().manual_seed(0)                      # identical replicas
    layers = [(w.requires_grad_(True), b.requires_grad_(True)) for w, b in orc.init_mlp([7, 64, 64, 1], gen)]
    pol = orc.NetSpec(layers, "gelu", "linear", torch.ones(1), -torch.ones(1), time_input=True)
    env = orc.create_env_model("pyth_idpendulum", reward_scale=1.0)
    data = orc.sample_inputs("pyth_idpendulum", B, seed=5)
    shard = B // world
    mine = {k: v[rank * shard:(rank + 1) * shard] for k, v in data.items()}
    inv_B = shard_inv_batch(shard, world)
    # what the fused kernel returns per rank: sums scaled by 1 / B_global
    loss_local = orc.fhadp_loss(pol, env, mine, H) * (shard * inv_B)
    loss_local.backward()
    n = sum(p.numel() for p in pol.params())
    gbuf = torch.zeros(n + GRAD_TAIL)
    gbuf[:n] = torch.cat([p.grad.reshape(-1) for p in pol.params()])
    gbuf[n] = loss_local.detach()
    allreduce_flat(gbuf)
    torch.save(gbuf, os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_allreduce_equals_full_batch(tmp_path):
    B, H, world = 64, 6, 2
    mp.spawn(_worker, args=(world, _free_port(), B, H, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1), "all ranks must hold the identical reduced buffer"
    gen = torch.Generator().manual_seed(0)
    layers = [(w.requires_grad_(True), b.requires_grad_(True)) for w, b in orc.init_mlp([7, 64, 64, 1], gen)]
    pol = orc.NetSpec(layers, "gelu", "linear", torch.ones(1), -torch.ones(1), time_input=True)
    env = orc.create_env_model("pyth_idpendulum", reward_scale=1.0)
    loss = orc.fhadp_loss(pol, env, orc.sample_inputs("pyth_idpendulum", B, seed=5), H)
    loss.backward()
    full = torch.cat([p.grad.reshape(-1) for p in pol.params()])
    n = full.numel()
    assert abs(float(g0[n]) - loss.item()) <= 1e-5 * abs(loss.item())
    assert float((g0[:n] - full).norm() / full.norm()) < 1e-5


def test_flat_params_views_survive_module_moves():
    """state_dict keys stay those of the reference and every parameter / gradient is a view of one buffer."""
    import numpy as np
    from gops_b200.apprfunc.mlp import FiniteHorizonPolicy
    from gops_b200.utils.act_distribution_type import DiracDistribution
    pol = FiniteHorizonPolicy(obs_dim=6, act_dim=1, hidden_sizes=[64, 64], hidden_activation="gelu",
                              output_activation="linear", act_high_lim=np.ones(1, np.float32),
                              act_low_lim=-np.ones(1, np.float32), action_distribution_cls=DiracDistribution)
    assert list(pol.state_dict()) == ["act_high_lim", "act_low_lim", "pi.0.weight", "pi.0.bias", "pi.2.weight",
                                      "pi.2.bias", "pi.4.weight", "pi.4.bias"]
    fp = pol.flat_params
    flat = fp.sync()
    assert flat.numel() == 7 * 64 + 64 + 64 * 64 + 64 + 64 + 1
    fp.bind_grads()
    off = 0
    for p in pol.pi.parameters():
        assert p.data.data_ptr() == flat.data_ptr() + 4 * off and p.grad.data_ptr() == fp.gbuf.data_ptr() + 4 * off
        off += p.numel()
    sd = {k: v.clone() + 1 for k, v in pol.state_dict().items()}
    pol.load_state_dict(sd)                      # in-place copy keeps the aliasing
    assert torch.equal(fp.sync()[:7 * 64], sd["pi.0.weight"].reshape(-1))
    pol.double().float()                         # a dtype round trip re-allocates parameters ...
    flat2 = fp.sync()                            # ... and sync() re-flattens them
    assert flat2.data_ptr() == next(pol.pi.parameters()).data.data_ptr()


def test_peer_exchange_is_not_selected_without_nccl(monkeypatch):
    """utils/peer_reduce.group_peer: the peer-memory path needs an initialised NCCL group of >= 2 ranks on one host; in
    every other situation the caller falls through to dist.all_reduce (and a wrong mode string is an error)."""
    from gops_b200.utils import peer_reduce
    assert peer_reduce.group_peer(1000, torch.device("cpu")) is None            # no process group
    monkeypatch.setenv("GOPS_B200_ALLREDUCE", "nccl")
    assert peer_reduce.mode() == "nccl" and peer_reduce.group_peer(1000, torch.device("cpu")) is None
    monkeypatch.setenv("GOPS_B200_ALLREDUCE", "ring")
    with pytest.raises(RuntimeError, match="auto, p2p or nccl"):
        peer_reduce.mode()
    monkeypatch.setenv("GOPS_B200_ALLREDUCE", "auto")
    assert peer_reduce.group_peer(peer_reduce.MAX_P2P_FLOATS + 1, torch.device("cpu")) is None   # bandwidth regime: NCCL
