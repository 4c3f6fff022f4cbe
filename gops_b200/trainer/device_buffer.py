"""Replay buffer resident in device memory (reference: gops/trainer/buffer/replay_buffer.py:27-108, a NumPy ring buffer on
the host whose sampled batch is copied to the GPU for every update).  Same ring semantics (`store` / `add_batch` /
`sample_batch` with uniform index sampling, float32 batch tensors incl. `done` and `logp`), but the storage, the index
draw and the gather stay on the device, so `alg.local_update(buffer.sample_batch(B), it)` involves no host memory.
Plumbing only: torch tensors and torch indexing."""
from typing import Dict

import torch


class DeviceReplayBuffer:
    def __init__(self, obsv_dim: int, action_dim: int, buffer_max_size: int, device="cuda", seed: int = 0, **kwargs):
        self.obsv_dim, self.act_dim, self.max_size = int(obsv_dim), int(action_dim), int(buffer_max_size)
        self.device = torch.device(device)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.buf = {"obs": z(self.max_size, self.obsv_dim), "obs2": z(self.max_size, self.obsv_dim),
                    "act": z(self.max_size, self.act_dim), "rew": z(self.max_size), "done": z(self.max_size),
                    "logp": z(self.max_size)}
        self.ptr, self.size = 0, 0
        self.gen = torch.Generator(device=self.device).manual_seed(int(seed) + 100)

    def __len__(self):
        return self.size

    def add_batch(self, samples: Dict[str, torch.Tensor]) -> None:
        """samples: dict of [n, ...] tensors with the buffer's keys (a batched `store`)."""
        n = samples["obs"].shape[0]
        idx = (self.ptr + torch.arange(n, device=self.device)) % self.max_size
        for k, v in self.buf.items():
            if k in samples:
                v[idx] = samples[k].to(self.device, torch.float32).reshape(n, *v.shape[1:])
        self.ptr = (self.ptr + n) % self.max_size
        self.size = min(self.size + n, self.max_size)

    def sample_batch(self, batch_size: int) -> Dict[str, torch.Tensor]:
        idx = torch.randint(0, self.size, (int(batch_size),), generator=self.gen, device=self.device)
        return {k: v[idx] for k, v in self.buf.items()}
