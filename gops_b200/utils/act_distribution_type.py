"""Action distributions used by the deterministic policies on this path
(reference: gops/utils/act_distribution_type.py:141-167)."""
import torch

EPS = 1e-6


class DiracDistribution:
    def __init__(self, logits):
        self.logits = logits

    def sample(self):
        return self.logits, torch.zeros_like(self.logits).sum(-1)

    def mode(self):
        return self.logits


class ValueDiracDistribution:
    def __init__(self, logits):
        self.logits = logits

    def sample(self):
        return torch.argmax(self.logits, dim=-1), torch.tensor([0.0])

    def mode(self):
        return torch.argmax(self.logits, dim=-1)


class TanhGaussDistribution:
    """Squashed Gaussian of the stochastic policies (reference act_distribution_type.py:18-76): the sampler / evaluator
    side (`sample`, `mode`, `log_prob`) on whatever device the logits live on.  Training never goes through this class:
    the DSAC update samples and differentiates inside the library (gops_b200_dsac_sample[_backward])."""

    def __init__(self, logits):
        self.logits = logits
        self.mean, self.std = torch.chunk(logits, chunks=2, dim=-1)
        self.act_high_lim = torch.tensor([1.0], device=logits.device)
        self.act_low_lim = torch.tensor([-1.0], device=logits.device)

    def _squash(self, action):
        return (self.act_high_lim - self.act_low_lim) / 2 * torch.tanh(action) + (self.act_high_lim + self.act_low_lim) / 2

    def _log_prob_pre(self, action):
        gauss = -((action - self.mean) ** 2) / (2 * self.std ** 2) - torch.log(self.std) - 0.9189385332046727
        return gauss.sum(-1) - torch.log(1 + EPS - torch.tanh(action) ** 2).sum(-1) \
            - torch.log((self.act_high_lim - self.act_low_lim) / 2).sum(-1)

    def sample(self):
        action = self.mean + self.std * torch.randn_like(self.mean)
        return self._squash(action), self._log_prob_pre(action)

    rsample = sample

    def log_prob(self, action_limited):
        action = torch.atanh((1 - EPS) * (2 * action_limited - (self.act_high_lim + self.act_low_lim))
                             / (self.act_high_lim - self.act_low_lim))
        return self._log_prob_pre(action)

    def mode(self):
        return self._squash(self.mean)
