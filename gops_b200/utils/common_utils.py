"""Config helpers on the path (reference: gops/utils/common_utils.py:26-135, :195-237)."""
import random
import sys

import numpy as np
import torch
import torch.nn as nn

from gops_b200.utils.act_distribution_type import (DiracDistribution, TanhGaussDistribution,  # noqa: F401
                                                    ValueDiracDistribution)

_ACTIVATIONS = {"relu": nn.ReLU, "elu": nn.ELU, "gelu": nn.GELU, "selu": nn.SELU, "sigmoid": nn.Sigmoid,
                "tanh": nn.Tanh, "linear": nn.Identity}


def get_activation_func(key: str):
    assert isinstance(key, str)
    if key not in _ACTIVATIONS:
        print("input activation name:" + key)
        raise RuntimeError
    return _ACTIVATIONS[key]


def activation_name(cls) -> str:
    for k, v in _ACTIVATIONS.items():
        if v is cls:
            return k
    raise RuntimeError(f"unknown activation class {cls}")


def get_apprfunc_dict(key: str, **kwargs):
    """Build the constructor kwargs of an approximate function from the flat config dict."""
    var = dict()
    var["apprfunc"] = kwargs[key + "_func_type"]
    var["name"] = kwargs[key + "_func_name"]
    var["obs_dim"] = kwargs["obsv_dim"]
    var["min_log_std"] = kwargs.get(key + "_min_log_std", float("-20"))
    var["max_log_std"] = kwargs.get(key + "_max_log_std", float("2"))
    var["std_type"] = kwargs.get(key + "_std_type", "mlp_shared")
    var["norm_matrix"] = kwargs.get("norm_matrix", None)
    var["pre_horizon"] = kwargs.get("pre_horizon", None)
    apprfunc_type = kwargs[key + "_func_type"]
    if apprfunc_type != "MLP":
        raise NotImplementedError(f"gops_b200 implements the MLP apprfunc family only, got {apprfunc_type}")
    var["hidden_sizes"] = kwargs[key + "_hidden_sizes"]
    var["hidden_activation"] = kwargs[key + "_hidden_activation"]
    var["output_activation"] = kwargs.get(key + "_output_activation", "linear")
    if kwargs["action_type"] == "continu":
        var["act_high_lim"] = np.array(kwargs["action_high_limit"])
        var["act_low_lim"] = np.array(kwargs["action_low_limit"])
        var["act_dim"] = kwargs["action_dim"]
    else:
        var["act_num"] = kwargs["action_num"]
    if kwargs["policy_act_distribution"] == "default":
        var["action_distribution_cls"] = (DiracDistribution if kwargs["action_type"] == "continu"
                                          else ValueDiracDistribution)
    else:
        var["action_distribution_cls"] = getattr(sys.modules[__name__], kwargs["policy_act_distribution"])
    return var


def seed_everything(seed=None) -> int:
    seed = int(seed if seed is not None else np.random.randint(0, 2 ** 31))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    return seed


def set_seed(trainer_name, seed, offset, env=None):
    """Sub-process replicas (async / sync trainers) re-seed; serial trainers do not."""
    if trainer_name.split("_")[1] in ["async", "sync"]:
        seed_everything(seed + offset)
        if env is not None:
            env.seed(seed + offset)
        return seed + offset, env
    if env is not None:
        env.seed(seed)
    return None, env
