#!/usr/bin/env python
"""Prints the measured parity margins of the fused CUDA path against the reference's golden vectors
(tests/golden, produced by the unmodified reference): loss relative error, gradient relative-L2 error,
per case and iteration.  Run on the GPU box:  python tools/parity_report.py > gpurun_out/parity.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from golden_util import CASES, rel_l2
import test_gpu_parity as T


def main():
    print(f"{'case':38s} it  {'loss(ref)':>14s} {'loss(gpu)':>14s} {'rel.err':>9s} {'grad relL2':>10s}")
    for name in T.GOLDEN:
        env_id, algname = CASES[name][0], CASES[name][1]
        alg, rec = T.build_alg(name)
        for it in ([0, 1] if algname == "INFADP" else [0]):
            if it > 0:
                alg.load_state_dict({k.split("/post/")[1]: torch.from_numpy(v) for k, v in rec.items()
                                     if k.startswith(f"it{it - 1}/post/")})
            try:
                tb = alg.get_remote_update_info(T.data_from(rec, env_id), it)[0]
            except RuntimeError as e:
                print(f"{name:38s} {it}   skipped: {e}")
                continue
            torch.cuda.synchronize()
            lk = T.loss_key(rec, it)
            ref = float(rec[lk]); got = tb[lk.split("/tb/")[1]]
            net = "v" if (algname == "INFADP" and it % 2 == 0) else "policy"
            gkeys = sorted(k for k in rec if k.startswith(f"it{it}/grad/{net}."))
            named = dict(getattr(alg.networks, net).named_parameters())
            g = [named[k.split(f"/grad/{net}.")[1]].grad.detach().cpu().numpy() for k in gkeys]
            print(f"{name:38s} {it}  {ref:14.6f} {got:14.6f} {abs(got - ref) / max(1.0, abs(ref)):9.1e} "
                  f"{rel_l2(g, [rec[k] for k in gkeys]):10.1e}")


if __name__ == "__main__":
    main()
