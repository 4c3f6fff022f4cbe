"""Gradient exchange between data-parallel ranks over NVLink peer memory, fused with Adam (csrc/peer.cu).

The reference's synchronous trainer hands every replica's gradient to the learner, which applies the average
(gops/trainer/off_sync_trainer.py:97-120 -> `alg.remote_update`).  Here every rank owns one GPU and the flat vector
[gradient | loss | critic mean | #done] of its shard; `PeerReduce.allreduce` sums it over the ranks in ONE kernel per
rank (push into the peers' exchange regions, sequence flags, fixed summation order) and, when an optimizer is given,
applies Adam in the same kernel.  torch.distributed is used once, to hand the 64-byte cudaIpc handles around.
"""
import ctypes as C
import os
import socket
import warnings
from typing import List, Optional

import torch

from gops_b200 import _lib

MAX_P2P_FLOATS = 1 << 20          # beyond this the exchange is bandwidth- not latency-bound: NCCL's ring is the right tool


class PeerReduce:
    def __init__(self, world: int, rank: int, max_floats: int, device: Optional[torch.device] = None):
        self.world, self.rank = int(world), int(rank)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gops_b200_peer_create(self.world, self.rank, int(max_floats), C.byref(self.handle)))
        self.cap = (int(max_floats) + 3) // 4 * 4

    # -- wiring ------------------------------------------------------------------------------------------------------
    def export(self) -> bytes:
        buf = C.create_string_buffer(64)
        _lib.check(_lib.lib().gops_b200_peer_export(self.handle, buf))
        return buf.raw

    def connect(self, handles: List[bytes]):
        """handles[r] = export() of rank r (one process per GPU)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        _lib.check(_lib.lib().gops_b200_peer_connect(self.handle, C.c_char_p(blob)))

    def local_base(self) -> int:
        p = C.c_void_p()
        _lib.check(_lib.lib().gops_b200_peer_local_base(self.handle, C.byref(p)))
        return int(p.value)

    def connect_local(self, bases: List[int]):
        """bases[r] = local_base() of rank r; for ranks that live in ONE process (several streams / devices)."""
        arr = (C.c_void_p * self.world)(*[C.c_void_p(b) for b in bases])
        _lib.check(_lib.lib().gops_b200_peer_connect_local(self.handle, arr))

    # -- the collective ----------------------------------------------------------------------------------------------
    def allreduce(self, buf: torch.Tensor, optimizer=None):
        """buf <- sum over ranks, in place, on the current stream of buf's device.  `optimizer` (FusedAdam over the
        parameters whose flat gradient is the head of buf): its step is applied by the same kernel."""
        if not (buf.is_cuda and buf.dtype == torch.float32 and buf.is_contiguous()):
            raise RuntimeError("PeerReduce.allreduce: a contiguous fp32 CUDA vector is required")
        n = buf.numel()
        if n > self.cap:
            raise RuntimeError(f"PeerReduce.allreduce: {n} floats exceed the exchange slots ({self.cap})")
        with torch.cuda.device(buf.device):
            if optimizer is None:
                _lib.check(_lib.lib().gops_b200_peer_allreduce(self.handle, _lib.ptr(buf), n, None, None, None, 0, 0, 0.0, 0.0,
                                                               0.0, 0.0, _lib.stream_ptr()))
            else:
                flat, m, v, step, g = optimizer.fused_state()
                if flat.numel() > n or flat.device != buf.device:
                    raise RuntimeError("PeerReduce.allreduce: optimizer parameters do not match the gradient vector")
                _lib.check(_lib.lib().gops_b200_peer_allreduce(
                    self.handle, _lib.ptr(buf), n, _lib.ptr(flat), _lib.ptr(m), _lib.ptr(v), flat.numel(), step,
                    float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), _lib.stream_ptr()))
        return buf

    def error(self) -> bool:
        """True if a peer failed to arrive within the kernel's time limit at some point (results were NaN-poisoned)."""
        e = C.c_int32()
        _lib.check(_lib.lib().gops_b200_peer_error(self.handle, C.byref(e)))
        return bool(e.value)

    def close(self):
        if self.handle:
            _lib.lib().gops_b200_peer_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_group_peer: Optional[PeerReduce] = None
_group_failed = False


def mode() -> str:
    """GOPS_B200_ALLREDUCE = auto (default: peer memory when every rank can, else NCCL) | p2p (required) | nccl."""
    m = os.environ.get("GOPS_B200_ALLREDUCE", "auto")
    if m not in ("auto", "p2p", "nccl"):
        raise RuntimeError("GOPS_B200_ALLREDUCE must be auto, p2p or nccl")
    return m


def group_peer(n_floats: int, device: torch.device) -> Optional[PeerReduce]:
    """The PeerReduce of the default process group (created collectively on first use), or None when the exchange goes
    through NCCL: single rank, non-NCCL backend, ranks on several hosts, a vector in the bandwidth regime, or a rank
    that could not map its peers (decided by ALL ranks together, so they never disagree on the path)."""
    global _group_peer, _group_failed
    import torch.distributed as dist
    if mode() == "nccl" or _group_failed or n_floats > MAX_P2P_FLOATS:
        return None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2 or dist.get_backend() != "nccl":
        return None
    if _group_peer is not None and _group_peer.cap >= n_floats and _group_peer.device == device:
        return _group_peer
    world, rank = dist.get_world_size(), dist.get_rank()
    ok, peer, why = 1, None, ""
    try:
        if world > 16:
            raise RuntimeError("more than 16 ranks")
        if _group_peer is not None:
            _group_peer.close()
            _group_peer = None
        peer = PeerReduce(world, rank, max(n_floats, 1 << 16), device)
        mine = (socket.gethostname(), peer.export())
    except Exception as e:      # noqa: BLE001 -- reported below, and every rank must still reach the collectives
        ok, why, mine = 0, str(e), (socket.gethostname(), b"\0" * 64)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    if ok and len({h for h, _ in everyone}) != 1:
        ok, why = 0, "ranks on several hosts"
    if ok:
        try:
            peer.connect([h for _, h in everyone])
        except Exception as e:  # noqa: BLE001
            ok, why = 0, str(e)
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 1:
        _group_peer = peer
        return peer
    if peer is not None:
        peer.close()
    _group_failed = True
    if mode() == "p2p":
        raise RuntimeError("GOPS_B200_ALLREDUCE=p2p: peer-memory exchange unavailable" + (f" ({why})" if why else ""))
    if rank == 0:
        warnings.warn("gops_b200: peer-memory gradient exchange unavailable, using NCCL all-reduce" + (f" ({why})" if why else ""))
    return None
