"""CPU restatement of the DSAC update (TEST INFRASTRUCTURE: only tests/ may import this).

Follows gops/algorithm/dsac.py:155-290 with the apprfuncs of gops/apprfunc/mlp.py:149-221 (StochaPolicy, std_type
"mlp_shared"), :271-296 (ActionValueDistri) and TanhGaussDistribution.rsample (gops/utils/act_distribution_type.py:37-50)
in plain PyTorch (autograd), with the Gaussian noise passed in explicitly (the reference draws it from torch's global
generator: eps_new for the actor's action, eps_next for the target action, z_next for the target critic sample).
Pinned against the unmodified reference by tests/test_oracle_dsac.py on tests/golden/dsac_idp.npz (noise recorded while
the reference ran)."""
import math

import torch
import torch.nn.functional as F

EPS = 1e-6


def mlp_apply(layers, x, act="gelu"):
    for j, (w, b) in enumerate(layers):
        x = F.linear(x, w, b)
        if j < len(layers) - 1:
            x = getattr(F, act)(x)
    return x


def policy_logits(layers, obs, min_log_std, max_log_std, act="gelu"):
    """StochaPolicy.forward, mlp.py:203-221: cat(mean, exp(clamp(log_std)))."""
    out = mlp_apply(layers, obs, act)
    mean, log_std = torch.chunk(out, 2, dim=-1)
    return torch.cat((mean, torch.clamp(log_std, min_log_std, max_log_std).exp()), dim=-1)


def rsample(logits, eps, hi, lo):
    """TanhGaussDistribution.rsample, act_distribution_type.py:37-50 with the noise given."""
    mean, std = torch.chunk(logits, 2, dim=-1)
    action = mean + std * eps
    limited = (hi - lo) / 2 * torch.tanh(action) + (hi + lo) / 2
    gauss = (-((action - mean) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)
    logp = gauss - torch.log(1 + EPS - torch.tanh(action) ** 2).sum(-1) - torch.log((hi - lo) / 2).sum(-1)
    return limited, logp


def q_head(layers, obs, act_, act="gelu"):
    """ActionValueDistri.forward, mlp.py:289-296 -> (mean, std)."""
    out = mlp_apply(layers, torch.cat([obs, act_], dim=-1), act)
    return out[..., 0], F.softplus(out[..., 1])


def dsac_losses(policy, policy_target, q, q_target, log_alpha, data, noise, gamma, bound=True, min_log_std=-20.0,
                max_log_std=1.0, target_entropy=-1.0, hi=None, lo=None, act="gelu"):
    """One DSAC.__compute_gradient (dsac.py:155-199): returns (loss_q, loss_policy, loss_alpha, info) whose autograd
    gradients are those the reference leaves in q / policy / log_alpha (q frozen for the actor loss)."""
    obs, a, rew, obs2, done = data["obs"], data["act"], data["rew"], data["obs2"], data["done"]
    hi = torch.ones(a.shape[-1], dtype=obs.dtype) if hi is None else hi
    lo = -torch.ones(a.shape[-1], dtype=obs.dtype) if lo is None else lo
    alpha = log_alpha.detach().exp().item()
    logits = policy_logits(policy, obs, min_log_std, max_log_std, act)
    new_act, new_logp = rsample(logits, noise["eps_new"], hi, lo)
    with torch.no_grad():
        logits2 = policy_logits(policy_target, obs2, min_log_std, max_log_std, act)
        act2, logp2 = rsample(logits2, noise["eps_next"], hi, lo)
        qn_mean, qn_std = q_head(q_target, obs2, act2, act)
        q_next = qn_mean + torch.clamp(noise["z_next"], -3, 3) * qn_std
    qm, qs = q_head(q, obs, a, act)
    target = (rew + (1 - done) * gamma * (q_next - alpha * logp2)).detach()
    td_bound = 3 * qs.detach().mean()
    target_b = (qm.detach() + torch.clamp(target - qm.detach(), -td_bound, td_bound)).detach()
    if bound:
        loss_q = torch.mean((qm - target) ** 2 / (2 * qs.detach() ** 2) + (qm.detach() - target_b) ** 2 / (2 * qs ** 2)
                            + torch.log(qs))
    else:
        loss_q = -torch.distributions.Normal(qm, qs).log_prob(target).mean()
    q_frozen = [(w.detach(), b.detach()) for w, b in q]
    qp, _ = q_head(q_frozen, obs, new_act, act)
    loss_policy = (alpha * new_logp - qp).mean()
    loss_alpha = -log_alpha * (new_logp.detach() + target_entropy).mean()
    info = dict(q=qm.detach().mean().item(), q_std=qs.detach().mean().item(), entropy=-new_logp.detach().mean().item(),
                policy_mean=torch.tanh(logits[..., 0]).mean().item(), policy_std=logits[..., 1].mean().item(), alpha=alpha)
    return loss_q, loss_policy, loss_alpha, info
