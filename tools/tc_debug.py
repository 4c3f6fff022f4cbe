"""Per-tensor comparison of the hybrid (tcgen05 forward) rollout kernel against the mma.sync kernel on one golden case."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as base  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fhadp_idp_h30"
MODE = sys.argv[2] if len(sys.argv) > 2 else "hy"
env_id, algname = base.CASES[name][0], base.CASES[name][1]
alg, rec = base.build_alg(name)
data = base.data_from(rec, env_id)
its = [0, 1] if algname == "INFADP" else [0]
for it in its:
    out = {}
    for mode in ("mma", MODE):
        os.environ["GOPS_B200_ROLLOUT"] = mode
        alg2, _ = base.build_alg(name)
        if it > 0:
            alg2.load_state_dict({k.split("/post/")[1]: torch.from_numpy(v) for k, v in rec.items()
                                  if k.startswith(f"it{it - 1}/post/")})
        alg2._INFADP__compute_gradient(data, it) if algname == "INFADP" else alg2._compute_gradient(data)
        torch.cuda.synchronize()
        net = "v" if (algname == "INFADP" and it % 2 == 0) else "policy"
        mod = getattr(alg2.networks, net)
        out[mode] = {k: p.grad.detach().cpu().numpy().copy() for k, p in mod.named_parameters()}
        print(mode, it, {k: float(v) for k, v in alg2.tb_info.items() if "oss" in k})
    for k in out["mma"]:
        a, b = out[MODE][k], out["mma"][k]
        err = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
        cos = float((a * b).sum() / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
        print(f"it{it} {k:12s} shape {str(a.shape):10s} rel {err:.3e} cos {cos:+.6f} |hy| {np.linalg.norm(a):.4e} |mma| {np.linalg.norm(b):.4e}")
        if err > 1e-3 and a.ndim == 2:
            r = np.abs(a - b)
            print("   row errs", np.round(r.sum(1)[:8], 5), "col errs", np.round(r.sum(0)[:8], 5))
            print("   hy[0,:6]", a[0, :6], "\n   mma[0,:6]", b[0, :6])
            # is hy a transposed / scaled version?
            if a.shape[0] == a.shape[1]:
                et = np.linalg.norm(a.T - b) / np.linalg.norm(b)
                print("   rel err of transpose", et)
            print("   ratio median", np.median(a[np.abs(b) > 1e-8] / b[np.abs(b) > 1e-8]))
