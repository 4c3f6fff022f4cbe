"""On-device serial trainer: the caller of the fused ADP update.

The reference's serial loop (gops/trainer/off_serial_trainer.py:79-173) steps a NumPy data env on the CPU, fills a
host replay buffer and copies a batch to the GPU for every update.  For the model-based ADP algorithms the replay
batch only provides INITIAL STATES, so here they are drawn directly on the device from the data envs' initial-state
distributions (idpendulum: pyth_idpendulum.py:36-38 uniform box; LQ: lq_base.py:151-155 Gaussian init_mean/init_std)
and handed to `alg.local_update` without any host round trip.  TensorBoard tags, the `apprfunc_{it}.pkl` checkpoint
format and the `config.json` dump follow the reference so that `example_run/*` keeps working on the results."""
import json
import os
import time
from typing import Dict, Optional

import torch

from gops_b200.utils.tensorboard_setup import tb_tags


class DeviceStateSampler:
    """Batched initial-state sampler living on the GPU (plumbing: torch RNG, no arithmetic of the hot path)."""

    def __init__(self, env_id: str, device, seed: int = 0, **kwargs):
        self.env_id, self.device = env_id, torch.device(device)
        self.gen = torch.Generator(device=self.device).manual_seed(int(seed))
        if env_id == "pyth_idpendulum":
            self.high = torch.tensor([5, 0.1, 0.1, 0.3, 0.3, 0.3], dtype=torch.float32, device=self.device)
        elif env_id == "pyth_lq":
            from gops_b200.env.env_ocp.resources import lq_configs
            cfg = kwargs.get("lq_config", "s3a1")
            cfg = getattr(lq_configs, "config_" + cfg) if isinstance(cfg, str) else cfg
            self.mean = torch.tensor(cfg["init_mean"], dtype=torch.float32, device=self.device)
            self.std = torch.tensor(cfg["init_std"], dtype=torch.float32, device=self.device)
        else:
            raise NotImplementedError(f"DeviceStateSampler: no on-device initial-state law for {env_id} yet")

    def sample(self, batch: int) -> Dict[str, torch.Tensor]:
        if self.env_id == "pyth_idpendulum":
            obs = (torch.rand(batch, 6, generator=self.gen, device=self.device) * 2 - 1) * self.high
        else:
            obs = self.mean + self.std * torch.randn(batch, self.mean.numel(), generator=self.gen, device=self.device)
        return {"obs": obs, "done": torch.zeros(batch, device=self.device)}


class OnDeviceSerialTrainer:
    def __init__(self, alg, sampler: DeviceStateSampler, *, replay_batch_size: int, max_iteration: int,
                 log_save_interval: int = 100, apprfunc_save_interval: int = 0, save_folder: Optional[str] = None,
                 ini_network_dir: Optional[str] = None, sample_interval: int = 1, **kwargs):
        self.alg, self.sampler, self.networks = alg, sampler, alg.networks
        if ini_network_dir is not None:
            self.networks.load_state_dict(torch.load(ini_network_dir))
        self.replay_batch_size, self.max_iteration = int(replay_batch_size), int(max_iteration)
        self.log_save_interval, self.apprfunc_save_interval = int(log_save_interval), int(apprfunc_save_interval)
        self.sample_interval = max(1, int(sample_interval))
        self.save_folder, self.iteration, self.history = save_folder, 0, []
        self.writer = None
        if save_folder is not None:
            os.makedirs(os.path.join(save_folder, "apprfunc"), exist_ok=True)
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=save_folder, flush_secs=20)
            except Exception:
                self.writer = None
            with open(os.path.join(save_folder, "config.json"), "w") as f:
                json.dump({k: v for k, v in kwargs.items() if isinstance(v, (int, float, str, bool, list))}, f, indent=1)
        self._batch = None
        self.start_time = time.time()

    def step(self):
        if self._batch is None or self.iteration % self.sample_interval == 0:
            self._batch = self.sampler.sample(self.replay_batch_size)
        self.networks.train()
        tb = self.alg.local_update(self._batch, self.iteration)
        self.networks.eval()
        if self.iteration % self.log_save_interval == 0:
            self.history.append((self.iteration, dict(tb)))
            if self.writer is not None:
                for tag, val in tb.items():
                    self.writer.add_scalar(tag, val, self.iteration)
        if self.apprfunc_save_interval and self.iteration % self.apprfunc_save_interval == 0:
            self.save_apprfunc()
        self.iteration += 1
        return tb

    def train(self):
        while self.iteration < self.max_iteration:
            self.step()
        self.save_apprfunc()
        if self.writer is not None:
            self.writer.flush()

    def save_apprfunc(self):
        if self.save_folder is None:
            return
        sd = {k: v.detach().cpu() for k, v in self.networks.state_dict().items()}
        torch.save(sd, os.path.join(self.save_folder, "apprfunc", f"apprfunc_{self.iteration}.pkl"))
