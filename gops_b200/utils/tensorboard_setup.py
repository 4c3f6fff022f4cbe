"""TensorBoard tag table consumed by the trainers (reference: gops/utils/tensorboard_setup.py:154-168)."""
tb_tags = {
    "TAR of RL iteration": "Evaluation/1. TAR-RL iter",
    "TAR of total time": "Evaluation/2. TAR-Total time [s]",
    "TAR of collected samples": "Evaluation/3. TAR-Collected samples",
    "TAR of replay samples": "Evaluation/4. TAR-Replay samples",
    "Buffer RAM of RL iteration": "RAM/RAM [MB]-RL iter",
    "loss_actor": "Loss/Actor loss-RL iter",
    "loss_actor_reward": "Loss/Actor reward loss-RL iter",
    "loss_actor_constraint": "Loss/Actor constraint loss-RL iter",
    "loss_critic": "Loss/Critic loss-RL iter",
    "alg_time": "Time/Algorithm time [ms]-RL iter",
    "sampler_time": "Time/Sampler time [ms]-RL iter",
    "critic_avg_value": "Train/Critic avg value-RL iter",
    "lips_value": "Lipschitz/Lipschitz value - RL iter",
}
