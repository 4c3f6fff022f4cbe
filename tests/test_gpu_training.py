"""System-level checks of the fused update through the on-device serial trainer: training converges where the
reference's own theory says it should (no reference run needed):
  * FHADP idpendulum, H = 30: reward <= 10 per step, so the optimum of the loss is > -300; an untrained policy sits at
    about +230 (the pendulum falls within ~20 steps).  After 1 500 updates the loss must be below -280.
  * INFADP LQ s4a2: the learned policy must line up with the discounted LQR gain that the reference computes in
    LQDynamics.compute_control_matrix (gops/env/env_ocp/resources/lq_base.py:61-71)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _kwargs(env_id, alg, obs_dim, act_dim, act, **extra):
    kw = dict(env_id=env_id, algorithm=alg, seed=0, trainer="off_serial_trainer", use_gpu=True, action_type="continu",
              obsv_dim=obs_dim, action_dim=act_dim, action_high_limit=np.ones(act_dim, np.float32),
              action_low_limit=-np.ones(act_dim, np.float32),
              policy_func_name="FiniteHorizonPolicy" if alg == "FHADP" else "DetermPolicy", policy_func_type="MLP",
              policy_hidden_sizes=[64, 64], policy_hidden_activation=act, policy_act_distribution="default",
              policy_learning_rate=3e-4, value_func_name="StateValue", value_func_type="MLP",
              value_hidden_sizes=[64, 64], value_hidden_activation=act, value_learning_rate=3e-4)
    kw.update(extra)
    return kw


def test_fhadp_idpendulum_learns_to_balance():
    from gops_b200.create_pkg.create_alg import create_alg
    from gops_b200.trainer.device_trainer import DeviceStateSampler, OnDeviceSerialTrainer
    torch.manual_seed(0)
    alg = create_alg(**_kwargs("pyth_idpendulum", "FHADP", 6, 1, "gelu", pre_horizon=30, reward_scale=1.0))
    tr = OnDeviceSerialTrainer(alg, DeviceStateSampler("pyth_idpendulum", "cuda", 1), replay_batch_size=4096,
                               max_iteration=1500, log_save_interval=100)
    tr.train()
    losses = [tb["Loss/Actor loss-RL iter"] for _, tb in tr.history]
    assert losses[0] > -250 and losses[-1] < -280 and losses[-1] > -300, losses


def test_infadp_lq_approaches_lqr_gain(tmp_path):
    from scipy.linalg import solve_discrete_are
    from gops_b200.create_pkg.create_alg import create_alg
    from gops_b200.env.env_ocp.resources import lq_configs
    from gops_b200.trainer.device_trainer import DeviceStateSampler, OnDeviceSerialTrainer
    torch.manual_seed(0)
    kw = _kwargs("pyth_lq", "INFADP", 4, 2, "gelu", lq_config="s4a2", reward_scale=1.0)
    kw.update(policy_learning_rate=8e-4, value_learning_rate=3e-4)
    alg = create_alg(**kw)
    alg.set_parameters({"forward_step": 50, "tau": 0.2, "gamma": 0.99})
    tr = OnDeviceSerialTrainer(alg, DeviceStateSampler("pyth_lq", "cuda", 2, lq_config="s4a2"), replay_batch_size=1024,
                               max_iteration=8000, log_save_interval=1000, save_folder=str(tmp_path),
                               apprfunc_save_interval=4000)
    tr.train()
    cfg = lq_configs.config_s4a2
    A0, B0, dt, gamma = np.array(cfg["A"], float), np.array(cfg["B"], float), cfg["dt"], 0.99
    A = np.linalg.pinv(np.eye(4) - A0 * dt) * np.sqrt(gamma)
    B = A @ B0 * dt
    Q, R = np.diag(np.array(cfg["Q"], float)), np.diag(np.array(cfg["R"], float))
    P = solve_discrete_are(A, B, Q, R)
    K = np.linalg.pinv(R + B.T @ P @ B) @ B.T @ P @ A
    x = torch.randn(512, 4, generator=torch.Generator().manual_seed(3)) * torch.tensor([0.7, 0.3, 0.7, 0.3]) * 0.5
    u = alg.networks.policy(x.cuda()).cpu().numpy() * 8.0         # ScaleAction maps [-1, 1] to the +-8 action box
    u_star = -(x.numpy() @ K.T)
    cos = float(np.sum(u * u_star) / (np.linalg.norm(u) * np.linalg.norm(u_star)))
    assert cos > 0.98, cos
    # checkpoints carry the reference's state_dict keys
    sd = torch.load(tmp_path / "apprfunc" / "apprfunc_8000.pkl")
    assert {"policy.pi.0.weight", "v.v.4.bias", "v_target.v.0.weight", "policy_target.pi.2.bias",
            "policy.act_high_lim"} <= set(sd)


def test_device_samplers_match_the_oracle_laws():
    """On-device initial-state samplers (trainer/device_sampler.py) vs. the oracle's restatement of the data envs'
    reset laws: same reference points for the same (t0, path, speed), same ego-frame observation."""
    from gops_b200.trainer import device_sampler as ds
    from oracle import gops_oracle as orc
    d = ds.sample_veh3dofconti(2048, 10, "cuda", seed=1)
    ref = orc.RefTraj()
    t0, p, s = d["ref_time"].cpu(), d["path_num"].cpu(), d["u_num"].cpu()
    for i in (0, 4, 10):
        tt = t0 + i * 0.1
        want = torch.stack((ref.x(tt, p, s), ref.y(tt, p, s), ref.phi(tt, p, s), ref.u(tt, p, s)), 1)
        got = d["ref_points"][:, i].cpu()
        np.testing.assert_allclose(np.delete(got.numpy(), 2, 1), np.delete(want.numpy(), 2, 1), rtol=1e-5, atol=2e-4)
        np.testing.assert_allclose(got[:, 2].numpy(), want[:, 2].numpy(), rtol=0, atol=8e-3)   # fp32 finite-difference phi
    np.testing.assert_allclose(d["obs"].cpu().numpy(), orc.veh_obs(d["state"].cpu(), d["ref_points"].cpu()).numpy(),
                               rtol=1e-5, atol=1e-5)
    assert set(p.unique().tolist()) == {0.0, 1.0, 2.0, 3.0} and set(s.unique().tolist()) == {0.0, 1.0}
    t = ds.sample_veh3dof_tracking(64, 20, "cuda", seed=2)
    assert t["state"].context_state.reference.shape == (64, 41, 4) and t["obs"].shape == (64, 86)


def test_trainer_evaluator_and_best_checkpoint(tmp_path):
    """Vehicle sampler -> INFADP updates -> batched on-device evaluator -> `_opt.pkl` bookkeeping of
    off_serial_trainer.py:126-141 (only after max_iteration / 5, previous best removed)."""
    import os
    from gops_b200.create_pkg.create_alg import create_alg
    from gops_b200.trainer.device_trainer import DeviceEvaluator, DeviceStateSampler, OnDeviceSerialTrainer
    torch.manual_seed(0)
    kw = _kwargs("pyth_veh3dofconti", "INFADP", 46, 2, "relu", pre_horizon=10)
    kw.update(policy_learning_rate=1e-3, value_learning_rate=1e-3)
    alg = create_alg(**kw)
    sampler = DeviceStateSampler("pyth_veh3dofconti", "cuda", 3, pre_horizon=10)
    ev = DeviceEvaluator(alg, DeviceStateSampler("pyth_veh3dofconti", "cuda", 4, pre_horizon=10), num_eval_episode=64,
                         max_step=50)
    r0 = ev.run_evaluation(0)
    tr = OnDeviceSerialTrainer(alg, sampler, replay_batch_size=2048, max_iteration=600, log_save_interval=100,
                               save_folder=str(tmp_path), evaluator=ev, eval_interval=100)
    tr.train()
    tars = [tb["Evaluation/1. TAR-RL iter"] for _, tb in tr.history if "Evaluation/1. TAR-RL iter" in tb]
    assert len(tars) >= 5 and all(np.isfinite(tars))
    assert max(tars[2:]) > r0, (r0, tars)                    # tracking improves over the untrained policy
    opt = [f for f in os.listdir(tmp_path / "apprfunc") if f.endswith("_opt.pkl")]
    assert len(opt) == 1 and int(opt[0].split("_")[1]) >= 600 / 5, opt
