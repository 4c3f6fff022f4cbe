"""Action distributions used by the deterministic policies on this path
(reference: gops/utils/act_distribution_type.py:141-167)."""
import torch


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
