"""Linear-quadratic benchmark systems (numerical data as shipped by the reference,
gops/env/env_ocp/resources/lq_configs.py:15-116)."""


def _cfg(A, B, Q, R, dt, init_mean, init_std, state_bound, action_bound, max_step, reward_scale=1.0, reward_shift=0):
    n = len(A)
    return dict(A=A, B=B, Q=Q, R=R, dt=dt, init_mean=init_mean, init_std=init_std,
                state_high=[state_bound] * n, state_low=[-state_bound] * n,
                action_high=[action_bound] * len(R), action_low=[-action_bound] * len(R),
                max_step=max_step, reward_scale=reward_scale, reward_shift=reward_shift)


config_s2a1 = _cfg([[0.0, 1.0], [0.0, 0.0]], [[0.0], [1.0]], [2, 1], [1.0], 0.05, [0.0, 0.0], [1.0, 1.0], 20.0, 5.0, 200)
config_s3a1 = _cfg([[-1.01887, 0.90506, -0.00215], [0.82225, -1.07741, -0.17555], [0.0, 0.0, -1.0]],
                   [[0.0], [0.0], [5.0]], [50.0, 1, 1], [1.0], 0.1, [0, 0, 0], [2, 2, 2], 20, 5.0, 200)
config_s4a2 = _cfg([[0, 1, 0, 0], [0, 1, 0, 0], [0.1, -0.2, 0, 0.5], [-0.2, 0.1, 0.1, 0]],
                   [[0, 0], [-2, -1], [0.0, 0], [1, 1.5]], [1, 2, 2, 1], [1.0, 1.0], 0.1, [0, 0, 0, 0],
                   [0.7, 0.3, 0.7, 0.3], 15, 8.0, 200)
config_s5a1 = _cfg([[1, 1, 0, 0, 0], [0, 0.2, 1, 0, 0], [0, 0, 0.3, 1, 0], [0, 0, 0, 0.4, 1], [0, 0, 0, 0, 0.5]],
                   [[1], [1], [1], [1], [1]], [50, 10, 20, 10, 10], [100], 0.05, [0] * 5, [0.1] * 5, 50, 10, 500)
config_s6a3 = _cfg([[0, 1, 0, 0, 0, 0], [3, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0], [2.5, 0, 0, 0, 0, 0],
                    [0, 0, 0, 0, 1, 0], [-2, 0, 0, 0, 0, 0]],
                   [[0, 0, 0], [1.5, 1.5, 0], [0.0, 0, 0], [0.5, 0.5, 0.5], [0, 0, 1], [2, 2, 2]],
                   [0, 2, 10, 10, 5, 5], [1.0, 1.0, 1.0], 0.05, [0] * 6, [0.1] * 6, 10, 10.0, 500)
