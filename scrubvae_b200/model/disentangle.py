"""Gradient-reversal scrubber head (reference model/disentangle.py:541-660) as parameter containers
with the reference's names and construction order.  The four MLPs, the reversed gradient
(-alpha * g into mu) and the head losses run in libscv.so kernels sequenced by the engine."""
from __future__ import annotations

import torch
import torch.nn as nn

from .residual import _EngineOnly


class GradientReversalLayer(_EngineOnly):
    """Reference :559-565 — `alpha` is a plain tensor attribute (not a buffer, not in state_dict)."""

    def __init__(self, alpha):
        super().__init__()
        self.alpha = torch.tensor(alpha, requires_grad=False)


class MLPEnsemble(_EngineOnly):
    """Reference :583-632 — four ReLU MLPs on the same input."""

    def __init__(self, in_dim, out_dim, bound=False):
        super().__init__()
        self.mlp1 = nn.Sequential(nn.Linear(in_dim, in_dim), nn.ReLU(), nn.Linear(in_dim, in_dim), nn.ReLU(),
                                  nn.Linear(in_dim, out_dim))
        self.mlp2 = nn.Sequential(nn.Linear(in_dim, in_dim), nn.ReLU(), nn.Linear(in_dim, out_dim))
        self.mlp3 = nn.Sequential(nn.Linear(in_dim, in_dim), nn.ReLU(), nn.Linear(in_dim, in_dim // 2), nn.ReLU(),
                                  nn.Linear(in_dim // 2, out_dim))
        self.mlp4 = nn.Sequential(nn.Linear(in_dim, in_dim * 2), nn.ReLU(), nn.Linear(in_dim * 2, in_dim * 2),
                                  nn.ReLU(), nn.Linear(in_dim * 2, out_dim))

    def members(self):
        return [self.mlp1, self.mlp2, self.mlp3, self.mlp4]


class GRScrubber(_EngineOnly):
    """Reference :635-660."""

    def __init__(self, in_dim, out_dim, alpha=1.0, bound=False):
        super().__init__()
        self.reversal = nn.Sequential(GradientReversalLayer(alpha), MLPEnsemble(in_dim, out_dim, bound))

    def reset_parameters(self):
        # nn.Linear.reset_parameters writes in place, so the flat parameter buffer stays attached
        for mlp in self.reversal[1].members():
            for head in mlp:
                if isinstance(head, nn.Linear):
                    head.reset_parameters()
