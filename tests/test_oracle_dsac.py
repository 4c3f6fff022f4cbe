"""CPU: pins oracle/dsac_oracle.py against the unmodified reference's DSAC.local_update (tests/golden/dsac_idp.npz, made by
oracle/make_golden.py run_dsac with the reference's noise recorded): scalars, gradients of q / policy / log_alpha for
four consecutive updates starting from the reference's own post-update weights."""
import numpy as np
import torch

from golden_util import load, rel_l2
from oracle import dsac_oracle as do


def _layers(rec, prefix, net, seq, grad):
    out, j = [], 0
    while f"{prefix}{net}.{seq}.{j}.weight" in rec:
        out.append((torch.tensor(rec[f"{prefix}{net}.{seq}.{j}.weight"]).requires_grad_(grad),
                    torch.tensor(rec[f"{prefix}{net}.{seq}.{j}.bias"]).requires_grad_(grad)))
        j += 2
    return out


def test_dsac_oracle_matches_reference():
    torch.set_num_threads(4)
    rec = load("dsac_idp")
    data = {k[3:]: torch.tensor(v) for k, v in rec.items() if k.startswith("in_")}
    n_it = 1 + max(int(k[2:k.index("/")]) for k in rec if k.startswith("it"))
    for it in range(n_it):
        prefix = "init/" if it == 0 else f"it{it - 1}/post/"
        pol, polT = _layers(rec, prefix, "policy", "policy", True), _layers(rec, prefix, "policy_target", "policy", False)
        q, qT = _layers(rec, prefix, "q", "q", True), _layers(rec, prefix, "q_target", "q", False)
        log_alpha = torch.tensor(rec[prefix + "log_alpha"]).requires_grad_(True)
        noise = {k: torch.tensor(rec[f"it{it}/{k}"]) for k in ("eps_new", "eps_next", "z_next")}
        lq, lp, la, info = do.dsac_losses(pol, polT, q, qT, log_alpha, data, noise, gamma=0.99)
        gq = torch.autograd.grad(lq, [t for pair in q for t in pair])
        gp = torch.autograd.grad(lp, [t for pair in pol for t in pair])
        ga = torch.autograd.grad(la, [log_alpha])[0]
        tb = {k.split("/tb/")[1]: float(v) for k, v in rec.items() if k.startswith(f"it{it}/tb/")}
        assert abs(lp.item() - tb["Loss/Actor loss-RL iter"]) <= 2e-6 * max(1.0, abs(lp.item()))
        assert abs(info["q"] - tb["DSAC/critic_avg_q-RL iter"]) < 1e-6 and abs(info["q_std"] - tb["DSAC/critic_avg_std-RL iter"]) < 1e-6
        assert abs(info["entropy"] - tb["DSAC/entropy-RL iter"]) < 2e-6 and abs(info["alpha"] - tb["DSAC/alpha-RL iter"]) < 1e-6
        kq = sorted(k for k in rec if k.startswith(f"it{it}/grad/q."))
        kp = sorted(k for k in rec if k.startswith(f"it{it}/grad/policy."))
        names_q = [f"it{it}/grad/q.q.{2 * j}.{w}" for j in range(len(q)) for w in ("weight", "bias")]
        names_p = [f"it{it}/grad/policy.policy.{2 * j}.{w}" for j in range(len(pol)) for w in ("weight", "bias")]
        assert sorted(names_q) == kq and sorted(names_p) == kp
        assert rel_l2([g.numpy() for g in gq], [rec[k] for k in names_q]) < 1e-5
        assert rel_l2([g.numpy() for g in gp], [rec[k] for k in names_p]) < 1e-5
        assert abs(float(ga) - float(rec[f"it{it}/grad/log_alpha"])) < 1e-5
