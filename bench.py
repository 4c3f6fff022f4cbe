#!/usr/bin/env python
"""bench.py -- batched env-steps/s of the FHADP rollout+update hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): FHADP on pyth_idpendulum, FiniteHorizonPolicy [64,64] gelu, horizon 30,
batch 2^18 PER GPU (weak scaling), synthetic initial states from the data env's reset box.
One step = one `alg.local_update(data, it)` through the plugin API: weight packing, fused rollout
forward+backward, partial reduction, (NCCL all-reduce of the flat gradient when N>1), fused Adam, and
the host read of the loss scalar.

value : inputs resident in HBM when the timed region starts (a rotating set of input batches whose total
        size exceeds L2, so no step finds its inputs in L2);
e2e   : same call with PINNED HOST tensors -- H2D copy of the batch and D2H read of the loss inside the
        timed region;
roofline : the fused rollout kernel alone, timed with CUDA events on its launch stream inside the library
        (gops_b200_plan_enable_timing).  The kernel is compute bound by design: `achieved` is
        algorithmic TFLOP/s, `peak` the tensor roof for FP32-accurate (BF16x3 / 3xTF32) GEMMs derived from the measured bf16
        peak; the FP32-FFMA, HBM and raw bf16 fractions are reported beside it.
cpu_baseline : the CPU oracle port (oracle/gops_oracle.py, the reference's algorithm in PyTorch-CPU) on a
        bounded sample of the same workload, all host threads.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = 30
OBS_DIM, ACT_DIM, HID = 6, 1, 64
MAC = (OBS_DIM + 1) * HID + HID * HID + HID * ACT_DIM          # 4608
FLOP_PER_ENV_STEP = 6 * MAC + 1500                             # MLP fwd+bwd + dynamics
BYTES_PER_ENV_STEP = (OBS_DIM * 4 + 4) / H                     # obs + done read once per sample
L2_BYTES = 126 * 1024 * 1024


def alg_kwargs():
    import numpy as np
    return dict(env_id="pyth_idpendulum", algorithm="FHADP", pre_horizon=H, seed=0, trainer="off_serial_trainer",
                use_gpu=True, action_type="continu", obsv_dim=OBS_DIM, action_dim=ACT_DIM,
                action_high_limit=np.ones(ACT_DIM, dtype=np.float32), action_low_limit=-np.ones(ACT_DIM, dtype=np.float32),
                policy_func_name="FiniteHorizonPolicy", policy_func_type="MLP", policy_hidden_sizes=[HID, HID],
                policy_hidden_activation="gelu", policy_act_distribution="default", policy_learning_rate=1e-4,
                value_func_type="MLP", reward_scale=1.0)


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line): one `nvidia-smi -lms 100`
    child process started before and killed after the region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.4)          # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[x.strip() for x in ln.split(",")] for ln in out.strip().splitlines() if ln.count(",") >= 6]
            except Exception:
                pass
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(rows)}


def usable_cores():
    """Host cores this process may really use: min(affinity, cgroup cpu quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def idp_init_batch(batch, seed):
    """Synthetic inputs of the GPU arm: the initial-state box of pyth_idpendulum (reference pyth_idpendulum.py:36-38,
    pyth_base_env.py:61-65): obs ~ U(-h, h), nobody done.  (The oracle is only imported by the CPU legs below.)"""
    import torch
    g = torch.Generator().manual_seed(seed)
    h = torch.tensor([5, 0.1, 0.1, 0.3, 0.3, 0.3])
    return {"obs": (torch.rand(batch, 6, generator=g, dtype=torch.float32) * 2 - 1) * h, "done": torch.zeros(batch)}


def cpu_update_rate(batch, steps, warmup, threads):
    """env-steps/s of the CPU oracle port: loss + autograd backward + Adam, as FHADP.local_update."""
    import torch
    torch.set_num_threads(threads)
    from oracle import gops_oracle as orc
    gen = torch.Generator().manual_seed(0)
    layers = [(w.requires_grad_(True), b.requires_grad_(True))
              for w, b in orc.init_mlp([OBS_DIM + 1, HID, HID, ACT_DIM], gen)]
    pol = orc.NetSpec(layers, "gelu", "linear", torch.ones(ACT_DIM), -torch.ones(ACT_DIM), time_input=True)
    env = orc.create_env_model("pyth_idpendulum", reward_scale=1.0)
    opt = torch.optim.Adam(pol.params(), lr=1e-4)
    data = orc.sample_inputs("pyth_idpendulum", batch, seed=1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = orc.fhadp_loss(pol, env, data, H)
        loss.backward()
        opt.step()
        loss.item()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * H * len(times) / total, total / len(times)


def best_cpu_threads(batch):
    """The reference's tiny-op eager path does not scale to many threads; give it the better of (all usable cores)
    and (16 threads) so the baseline is not handicapped by oversubscription."""
    cands = sorted({usable_cores(), min(usable_cores(), 16)})
    if len(cands) == 1:
        return cands[0]
    rates = {c: cpu_update_rate(batch, 1, 1, c)[0] for c in cands}
    return max(rates, key=rates.get)


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU algorithm (oracle port; the python reference cannot travel
    to the GPU box) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    sample_b = args.cpu_batch
    threads = best_cpu_threads(sample_b)
    rate, sec = cpu_update_rate(sample_b, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "batched env-steps/sec (FHADP rollout+update)", "value": rate,
        "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"FHADP pyth_idpendulum H={H} FiniteHorizonPolicy[64,64] gelu",
                   "sample": f"batch {sample_b} per step on CPU"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"B={sample_b}, H={H}, {args.steps} updates (oracle/gops_oracle.py)"},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=1 << 18)
    ap.add_argument("--cpu-batch", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    torch.set_num_threads(usable_cores())      # the box reports 128 cpus under a 16-core cgroup quota
    import torch.distributed as dist
    from gops_b200.create_pkg.create_alg import create_alg
    from gops_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W, K, Bg = max(args.warmup, 3), args.steps, args.batch_per_gpu

    torch.manual_seed(0)                       # identical replicas on every rank
    alg = create_alg(**alg_kwargs())
    batch_bytes = Bg * (OBS_DIM + 1) * 4
    n_sets = max(2, math.ceil(1.5 * L2_BYTES / batch_bytes))
    host_sets = []
    for i in range(n_sets):
        d = idp_init_batch(Bg, seed=1000 * rank + i)
        host_sets.append({k: v.pin_memory() for k, v in d.items()})
    dev_sets = [{k: v.to(dev) for k, v in d.items()} for d in host_sets]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(sets, steps, collect_kernel=False):
        plan = next(iter(alg._plans.values())) if alg._plans else None
        kms = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            alg.local_update(sets[i % len(sets)], i)
            if collect_kernel and plan is not None:
                import ctypes as C
                ms = C.c_float()
                _lib.check(_lib.lib().gops_b200_plan_last_kernel_ms(plan.handle, C.byref(ms)))
                kms.append(ms.value)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), kms

    timed(dev_sets, W)                                    # warm-up (creates the plan, compiles nothing)
    plan = next(iter(alg._plans.values()))
    _lib.check(_lib.lib().gops_b200_plan_enable_timing(plan.handle, 1))
    timed(dev_sets, 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total, kernel_ms = timed(dev_sets, K, collect_kernel=True)
    clocks = sampler.stop()
    _lib.check(_lib.lib().gops_b200_plan_enable_timing(plan.handle, 0))
    timed(host_sets, W)
    ms_e2e, _ = timed(host_sets, K)

    env_steps = Bg * world * H
    value = env_steps * K / (ms_total * 1e-3)
    e2e = env_steps * K / (ms_e2e * 1e-3)

    if rank == 0:
        import ctypes as C
        info = (C.c_int32 * 4)()
        _lib.check(_lib.lib().gops_b200_plan_launch_info(plan.handle, info))
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured" if peaks else "fallback"
        k_ms = statistics.mean(kernel_ms) if kernel_ms else ms_total / K
        sm_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        fp32_peak = n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12
        ach_tflops = Bg * H * FLOP_PER_ENV_STEP / (k_ms * 1e-3) / 1e12
        ach_gbs = Bg * H * BYTES_PER_ENV_STEP / (k_ms * 1e-3) / 1e9
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            try:
                traffic = json.load(open(tfile)).get("rollout_kernel_dram_bytes_per_launch")
            except Exception:
                pass
        # Tensor roof for FP32-accurate GEMMs: the MLP layers run as 3xTF32 mma.sync (three TF32 MMAs per product);
        # TF32 dense peak = measured bf16 peak / 2.  The FP32 FFMA roof is the one SURVEY 8(d) names for a CUDA-core
        # implementation; both are reported, plus HBM.
        tf32x3_peak = tens_peak / 2.0 / 3.0
        roof = {"bound": "tensor", "achieved": ach_tflops, "peak": tf32x3_peak, "unit": "TFLOP/s",
                "frac": ach_tflops / tf32x3_peak, "traffic": traffic,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src}) / 2 (TF32) / 3 (3xTF32 split)",
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms * K / ms_total,
                "algorithmic_flop_per_env_step": FLOP_PER_ENV_STEP, "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP,
                "fp32_ffma": {"achieved_tflops": ach_tflops, "peak_tflops": fp32_peak, "frac": ach_tflops / fp32_peak,
                              "of": f"{n_sm} SM x 128 lanes x 2 x {sm_mhz:.0f} MHz sampled under load"},
                "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": hbm_peak, "frac": ach_gbs / hbm_peak, "of": peak_src},
                "tensor_bf16": {"achieved_tflops": ach_tflops, "peak_tflops": tens_peak, "frac": ach_tflops / tens_peak,
                                "of": peak_src + " bf16 sustained"},
                "launch": {"grid": info[0], "block": info[1], "tile_samples": info[2], "smem_bytes": info[3]}}
        cpu = None
        if not args.no_cpu_baseline:
            threads = best_cpu_threads(args.cpu_batch)
            rate, sec = cpu_update_rate(args.cpu_batch, 3, 1, threads)
            cpu = {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
                   "sample": f"B={args.cpu_batch}, H={H}, 3 updates after 1 warm-up (oracle/gops_oracle.py)"}
        # kernels of this library per update: pack_params[_tcf], [pack_params_tc,] rollout_kernel, reduce_partials, adam
        sm_count = torch.cuda.get_device_properties(0).multi_processor_count
        forced = os.environ.get("GOPS_B200_ROLLOUT", "")
        path = forced if forced in ("tc", "hy", "mma") else ("tc" if Bg >= sm_count * 512 else "mma")
        launches_per_step = 5 if path == "hy" else 4
        roof["launch"]["kernel_path"] = {
            "tc": "full tcgen05: BF16x3 UMMA for every dense product, weight gradients accumulate in TMEM",
            "hy": "hybrid: tcgen05/TMEM (3xTF32) forward sweep + mma.sync reverse sweep",
            "mma": "mma.sync (3xTF32) forward and reverse sweeps"}[path]
        if path == "tc":   # six bf16 MMAs per FP32-accurate product: the same roof as three TF32 MMAs at half the bf16 rate
            roof["peak_source"] = (f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src}) / 6 "
                                   "(BF16x3: six bf16 products per FP32-accurate product)")
        line = {
            "metric": "batched env-steps/sec (FHADP rollout+update)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"FHADP pyth_idpendulum H={H} FiniteHorizonPolicy[64,64] gelu, "
                                   f"batch {Bg} per GPU (global {Bg * world})",
                       "parallelism": f"dp{world}", "l2_policy": f"{n_sets} rotating input batches "
                       f"({n_sets * batch_bytes / 2**20:.0f} MiB total > L2)"},
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": batch_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": launches_per_step * K,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
