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
