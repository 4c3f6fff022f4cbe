"""K = 20 consecutive updates against the UNMODIFIED reference (tests/golden/traj_*.npz, oracle/make_golden.py
run_trajectory): per-update losses, the LinearLR schedule driven through `local_update` (base.py:94-98), every 5th
update through get_remote_update_info + remote_update (the Ray-replica entry points, base.py:100-104), the weights
after 1, 10 and 20 updates and the trained policy's actions.

Bars.  Loss of update k: 1e-4 relative for k = 0 (same weights), 5e-4 afterwards -- the weights then differ by the
fp32 noise Adam amplifies where a gradient element is zero within round-off (update 1 moves every weight by
lr * sign(g)).  Weights after K updates: at least 97 % of the elements within 2 % of the distance travelled, none
further than the distance itself.  Trained-policy actions: 2e-3 absolute (actions are in [-1, 1])."""
import numpy as np
import pytest
import torch

from golden_util import load

pytestmark = pytest.mark.gpu


def _mk(name):
    from gops_b200.create_pkg.create_alg import create_alg
    rec = load(name)
    lq = name.startswith("traj_infadp_lq")
    obs_dim, act_dim = (4, 2) if lq else (6, 1)
    kw = dict(env_id="pyth_lq" if lq else "pyth_idpendulum", algorithm="INFADP" if lq else "FHADP", seed=0,
              trainer="off_serial_trainer", cnn_shared=False, use_gpu=True, action_type="continu", obsv_dim=obs_dim,
              action_dim=act_dim, action_high_limit=np.ones(act_dim, dtype=np.float32),
              action_low_limit=-np.ones(act_dim, dtype=np.float32),
              policy_func_name="DetermPolicy" if lq else "FiniteHorizonPolicy", policy_func_type="MLP",
              policy_hidden_sizes=[64, 64], policy_hidden_activation="gelu", policy_act_distribution="default",
              value_func_name="StateValue", value_func_type="MLP", value_hidden_sizes=[64, 64],
              value_hidden_activation="gelu", reward_scale=1.0)
    if lq:
        kw.update(lq_config="s4a2", reward_shift=0.0, policy_learning_rate=8e-4, value_learning_rate=3e-4)
    else:
        kw.update(pre_horizon=30, policy_learning_rate=3e-4, value_learning_rate=1e-3,
                  policy_scheduler={"name": "LinearLR", "params": {"start_factor": 1.0, "end_factor": 0.25,
                                                                   "total_iters": 16}})
    alg = create_alg(**kw)
    alg.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("init/")})
    nb = 1 + max(int(k[1:k.index("/")]) for k in rec if k[0] == "b" and k[1].isdigit())
    batches = [{k.split("/in_")[1]: torch.from_numpy(v) for k, v in rec.items() if k.startswith(f"b{j}/in_")}
               for j in range(nb)]
    return alg, rec, batches


@pytest.mark.parametrize("name,remote_every", [("traj_fhadp_idp_k20", 5), ("traj_infadp_lq_k20", 0)])
def test_twenty_updates_follow_the_reference(name, remote_every):
    alg, rec, batches = _mk(name)
    K = len(rec["losses"])
    infadp = name.startswith("traj_infadp")
    travelled = 0.0
    for it in range(K):
        data = batches[it % len(batches)]
        lr_now = alg.networks.policy_optimizer.param_groups[0]["lr"]
        if remote_every and it % remote_every == remote_every - 1:
            tb, upd = alg.get_remote_update_info(data, it)
            alg.remote_update(upd)
        else:
            tb = alg.local_update(data, it)
        tag = "Loss/Critic loss-RL iter" if (infadp and it % 2 == 0) else "Loss/Actor loss-RL iter"
        ref = float(rec["losses"][it])
        rtol = 1e-4 if it == 0 else 5e-4
        assert abs(tb[tag] - ref) <= rtol * max(1.0, abs(ref)), (it, tb[tag], ref)
        # the scheduler stepped exactly like the reference's (lr AFTER the update is what the golden stores)
        assert abs(alg.networks.policy_optimizer.param_groups[0]["lr"] - float(rec["lrs"][it])) <= 1e-12, it
        travelled += lr_now
        if f"after{it + 1}/policy.pi.0.weight" in rec:
            sd = alg.state_dict()
            for k in (k for k in rec if k.startswith(f"after{it + 1}/") and ".pi." in k and "target" not in k):
                got, want = sd[k.split("/", 1)[1]].detach().cpu().numpy(), rec[k]
                delta = np.abs(got - want)
                dist = travelled if not infadp else travelled / 2 + 1e-12     # the policy steps every other update
                assert delta.max() <= 1.05 * dist + 1e-7, (it, k, delta.max(), dist)
                assert np.mean(delta <= 2e-2 * dist + 1e-7) >= 0.97, (it, k, float(np.mean(delta <= 2e-2 * dist + 1e-7)))
    obs = batches[0]["obs"][:64]
    act = alg.networks.policy(obs.cuda(), 1) if not infadp else alg.networks.policy(obs.cuda())
    np.testing.assert_allclose(act.cpu().numpy(), rec["final_actions"], rtol=0, atol=2e-3)
