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


class MutInfoEstimator(nn.Module):
    """Reference model/disentangle.py:234-317: the stored samples of the kernel mutual-information estimate behind the
    "mcmi" loss.  Same constructor; the pairwise log-sum-exps and their gradient run in scv_mi_loss (csrc/scv_mi.cu).
    The trainer rebuilds it after every optimizer step from the UPDATED encoder (train/trainer.py:184-199)."""

    def __init__(self, x_s, y_s, bandwidth, var_mode="sphere", model_var=None, device="cuda"):
        super().__init__()
        self.register_buffer("x_s", x_s.detach().to(torch.float32).contiguous())
        self.register_buffer("y_s", y_s.detach().to(torch.float32).contiguous())
        self.num_s = x_s.shape[0]
        assert y_s.shape[0] == self.num_s
        self.x_dim, self.y_dim = x_s.shape[1], y_s.shape[1]
        self.var_mode = var_mode
        self.gamma = float(bandwidth)
        log2pi = float(torch.log(torch.tensor(2 * torch.pi)))
        if var_mode == "sphere":
            self.register_buffer("var_s", torch.tensor([self.gamma], device=x_s.device))
            self.register_buffer("logA_x", self.x_dim * (log2pi + torch.log(self.var_s)))
        elif var_mode == "diagonal":
            self.register_buffer("var_s", (model_var.detach().diagonal(dim1=-2, dim2=-1) ** 2 + self.gamma).contiguous())
            self.register_buffer("logA_x", (self.x_dim * log2pi + torch.sum(torch.log(self.var_s), dim=-1)).contiguous())
        else:
            raise ValueError(var_mode)

    def forward(self, x, y):
        """Value of the estimate (no autograd: inside the training step the loss and its gradient come from the
        engine's launch lists, train/losses.get_batch_loss)."""
        from .._ops import get_ops
        ops = get_ops()
        out = torch.zeros(1, dtype=torch.double, device=x.device)
        x, y = x.detach().float().contiguous(), y.detach().float().contiguous()
        ops.mi_loss(x, y, y.shape[1], self.x_s, self.y_s, self.var_s if self.var_mode == "diagonal" else None, self.logA_x,
                    self.gamma, self.num_s, x.shape[0], self.x_dim, self.y_dim, loss=out)
        return out[0].float()
