"""CPU: the oracle restatement replays the reference's K = 20 update trajectory (tests/golden/traj_fhadp_idp_k20.npz):
losses, LinearLR schedule, Adam, weights after 1 / 10 / 20 updates.  Pins oracle.adam_step + fhadp_loss over many
consecutive updates, not just one."""
import numpy as np
import torch

from golden_util import load, net_from
from oracle import gops_oracle as orc


def test_oracle_replays_reference_trajectory():
    torch.set_num_threads(4)
    rec = load("traj_fhadp_idp_k20")
    pol = net_from(rec, "init/", "policy", "gelu", requires_grad=True)
    pol.time_input = True
    env = orc.create_env_model("pyth_idpendulum", reward_scale=1.0)
    batches = [{"obs": torch.from_numpy(rec[f"b{j}/in_obs"]), "done": torch.from_numpy(rec[f"b{j}/in_done"])}
               for j in range(3)]
    state, lr0 = {}, 3e-4
    for it in range(len(rec["losses"])):
        # torch LinearLR(start 1.0, end 0.25, total 16): factor of update `it` (the scheduler steps after the update)
        lr = lr0 * (1.0 + (0.25 - 1.0) * min(it, 16) / 16)
        for p in pol.params():
            p.grad = None
        loss = orc.fhadp_loss(pol, env, batches[it % 3], 30)
        loss.backward()
        ref = float(rec["losses"][it])
        assert abs(loss.item() - ref) <= 2e-5 * max(1.0, abs(ref)), (it, loss.item(), ref)
        new = orc.adam_step([p.detach() for p in pol.params()], [p.grad for p in pol.params()], state, lr)
        with torch.no_grad():
            for p, q in zip(pol.params(), new):
                p.copy_(q)
        lr_next = lr0 * (1.0 + (0.25 - 1.0) * min(it + 1, 16) / 16)
        assert abs(lr_next - float(rec["lrs"][it])) < 1e-12
        if f"after{it + 1}/policy.pi.0.weight" in rec:
            for j, (w, b) in enumerate(pol.layers):
                np.testing.assert_allclose(w.detach().numpy(), rec[f"after{it + 1}/policy.pi.{2 * j}.weight"], rtol=0,
                                           atol=3e-6)
