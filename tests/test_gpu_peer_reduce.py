"""Peer-memory gradient exchange fused with Adam (csrc/peer.cu, utils/peer_reduce.py).

The protocol (push to the peers' slots, sequence flags, parity double-buffering, fixed summation order) is exercised on
ONE device with the ranks wired in-process on separate streams; the cudaIpc wiring and the NVLink path run in the
two-process test, which needs two GPUs (skipped otherwise)."""
import ctypes as C
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ring(world, cap):
    from gops_b200.utils.peer_reduce import PeerReduce
    PeerReduce(1, 0, 16).allreduce(torch.zeros(8, device="cuda"))        # loads the kernel before anyone spins on a peer
    torch.cuda.synchronize()
    peers = [PeerReduce(world, r, cap) for r in range(world)]
    bases = [p.local_base() for p in peers]
    for p in peers:
        p.connect_local(bases)
    return peers, [torch.cuda.Stream() for _ in range(world)]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sum_is_rank_ordered_and_identical_on_all_ranks(world):
    peers, streams = _ring(world, 6000)
    gen = torch.Generator(device="cuda").manual_seed(world)
    for it, n in enumerate([4743, 4743, 5, 6000, 4743, 1, 333]):          # slot parity reuse, ragged chunks, n < #CTAs
        src = [torch.randn(n, device="cuda", generator=gen) * 10 ** (r % 3) for r in range(world)]
        bufs = [s.clone() for s in src]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                peers[r].allreduce(bufs[r])
        torch.cuda.synchronize()
        want = torch.zeros(n, device="cuda")
        for r in range(world):
            want = want + src[r]                                          # fp32, rank order
        for r in range(world):
            assert torch.equal(bufs[r], want), (it, n, r)
    assert not any(p.error() for p in peers)


def test_fused_adam_equals_adam_on_the_summed_gradient():
    from gops_b200 import _lib
    world, npar, tail = 4, 4739, 4
    peers, streams = _ring(world, npar + tail)
    gen = torch.Generator(device="cuda").manual_seed(7)
    p0 = torch.randn(npar, device="cuda", generator=gen)
    P = [p0.clone() for _ in range(world)]
    M = [torch.zeros(npar, device="cuda") for _ in range(world)]
    V = [torch.zeros(npar, device="cuda") for _ in range(world)]
    rp, rm, rv = p0.clone(), torch.zeros(npar, device="cuda"), torch.zeros(npar, device="cuda")
    lr, b1, b2, eps = 3e-4, 0.9, 0.999, 1e-8
    for step in range(1, 6):
        src = [torch.randn(npar + tail, device="cuda", generator=gen) for _ in range(world)]
        bufs = [s.clone() for s in src]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                _lib.check(_lib.lib().gops_b200_peer_allreduce(
                    peers[r].handle, _lib.ptr(bufs[r]), npar + tail, _lib.ptr(P[r]), _lib.ptr(M[r]), _lib.ptr(V[r]), npar,
                    step, lr, b1, b2, eps, C.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        want = torch.zeros(npar + tail, device="cuda")
        for r in range(world):
            want = want + src[r]
        _lib.check(_lib.lib().gops_b200_adam_step(_lib.ptr(rp), _lib.ptr(want), _lib.ptr(rm), _lib.ptr(rv), npar, step, lr, b1, b2,
                                                  eps, _lib.stream_ptr()))
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(bufs[r], want)
            assert torch.equal(P[r], rp) and torch.equal(M[r], rm) and torch.equal(V[r], rv), (step, r)


def test_argument_errors_are_reported():
    from gops_b200.utils.peer_reduce import PeerReduce
    p = PeerReduce(2, 0, 64)
    with pytest.raises(RuntimeError, match="not connected"):
        p.allreduce(torch.zeros(8, device="cuda"))
    q = PeerReduce(1, 0, 64)
    with pytest.raises(RuntimeError, match="exceed"):
        q.allreduce(torch.zeros(100, device="cuda"))
    with pytest.raises(RuntimeError, match="world must be"):
        PeerReduce(17, 0, 64)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, mode):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["GOPS_B200_ALLREDUCE"] = mode
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    from bench_configs import kwargs
    from gops_b200.create_pkg.create_alg import create_alg
    from oracle import gops_oracle as orc
    torch.manual_seed(0)                                                 # identical replicas
    alg = create_alg(**kwargs("pyth_idpendulum", "FHADP", 6, 1, 64, "gelu", pre_horizon=8, reward_scale=1.0))
    B = 1024
    data = orc.sample_inputs("pyth_idpendulum", B, 11)
    shard = B // world
    mine = {k: v[rank * shard:(rank + 1) * shard].cuda() for k, v in data.items()}
    losses = []
    for it in range(4):
        alg.local_update(mine, it)
        losses.append(alg.tb_info["Loss/Actor loss-RL iter"])
    torch.cuda.synchronize()
    from gops_b200.utils import peer_reduce
    used_p2p = peer_reduce._group_peer is not None
    if used_p2p:
        assert not peer_reduce._group_peer.error()
    torch.save({"flat": alg.networks.policy.flat_params.sync().cpu(), "losses": losses, "p2p": used_p2p},
               os.path.join(out_dir, f"{mode}{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_processes_ipc_matches_nccl(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    for mode in ("p2p", "nccl"):
        mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    r = {(m, k): torch.load(tmp_path / f"{m}{k}.pt") for m in ("p2p", "nccl") for k in range(world)}
    assert r[("p2p", 0)]["p2p"] and r[("p2p", 1)]["p2p"] and not r[("nccl", 0)]["p2p"]
    assert torch.equal(r[("p2p", 0)]["flat"], r[("p2p", 1)]["flat"]), "replicas must stay bit-identical"
    # two ranks: a + b is the same number in either order, so the NCCL run is reproduced exactly
    assert torch.equal(r[("p2p", 0)]["flat"], r[("nccl", 0)]["flat"])
    assert r[("p2p", 0)]["losses"] == r[("nccl", 0)]["losses"]
