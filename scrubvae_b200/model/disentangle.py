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


class MovingAvgLeastSquares(nn.Module):
    """Reference model/disentangle.py:393-538, polynomial order 1: two linear decoders of the scrubbed variable from the
    latent mean, fitted by exponentially weighted least squares with forgetting factors lam0 < lam1 that drift towards the
    better one.  Same constructor, buffers (Sxx0, Sxy0, Sxx1, Sxy1, lam0, lam1: the reference's state_dict keys) and
    methods; the solve, the predictions, the loss with its gradient into mu, the forgetting-factor rule and the covariance
    update run in csrc/scv_mals.cu, sequenced by the engine for the training step.  The standalone methods below use the
    same kernels (ops handed over by the engine that owns the model)."""

    def __init__(self, nx, ny, lamdiff=1e-1, delta=1e-4, bias=False, polynomial_order=1, l2_reg=0):
        super().__init__()
        if polynomial_order != 1:
            raise NotImplementedError("scrubvae_b200: moving_avg_lsq is built for polynomial order 1 "
                                      "(order 2 is a 2145 x 2145 solve per step: SURVEY.md §8(f) rank 4, not built)")
        self.bias = bool(bias)
        self.polynomial_order = polynomial_order
        self.z = int(nx)
        nx = int(nx) + int(self.bias)
        if nx > 136 or ny > 16:
            raise NotImplementedError("scrubvae_b200: moving_avg_lsq kernels hold z <= 135 (+ bias) and ny <= 16")
        self.l2_reg = 0 if l2_reg is None else l2_reg
        print("Moving Avg Least Squares Bias: {}".format(self.bias))
        self.register_buffer("Sxx0", torch.eye(nx))
        self.register_buffer("Sxy0", torch.zeros(nx, ny))
        self.register_buffer("Sxx1", torch.eye(nx))
        self.register_buffer("Sxy1", torch.zeros(nx, ny))
        self.register_buffer("lam0", torch.tensor([0.9]))
        self.register_buffer("lam1", torch.tensor([0.9]) + lamdiff)
        self.delta = delta
        self.lamdiff = lamdiff
        self._ops = None
        self._l01 = None

    def _need_ops(self):
        if self._ops is None:
            from .._ops import get_ops
            self._ops = get_ops()
        return self._ops

    def forward(self, x):
        ops = self._need_ops()
        B, ny = x.shape[0], self.Sxy0.shape[1]
        x = x.detach().to(torch.float32).contiguous()
        W0, W1 = torch.empty_like(self.Sxy0), torch.empty_like(self.Sxy1)
        ops.mals_solve(self.Sxx0, self.Sxy0, self.Sxx1, self.Sxy1, self.l2_reg, self.bias, self.Sxx0.shape[0], ny, W0, W1)
        y0, y1 = torch.zeros(B, ny, device=x.device), torch.zeros(B, ny, device=x.device)
        ops.mals_loss(x, x.shape[1], y0, ny, W0, W1, self.bias, B, self.z, ny, yhat0=y0, yhat1=y1)  # y0 doubles as a dummy y
        return [y0, y1]

    def update(self, x, y):
        ops = self._need_ops()
        x = x.detach().to(torch.float32).contiguous()
        y = y.detach().to(torch.float32).reshape(x.shape[0], -1).contiguous()
        ops.mals_update(x, x.shape[1], y, y.shape[1], self.bias, x.shape[0], self.z, y.shape[1], self.lam0, self.lam1,
                        self.Sxx0, self.Sxy0, self.Sxx1, self.Sxy1)
        return self

    def evaluate_loss(self, yhat0, yhat1, y):
        """Sum of squared errors of the two decoders, averaged; moves lam0 / lam1 (reference :505-538).  No gradient: the
        training path gets d loss / d mu from the engine's backward."""
        ops = self._need_ops()
        y = y.detach().to(torch.float32).reshape(yhat0.shape)
        l01 = torch.stack([((y - yhat0.detach()).double() ** 2).sum(), ((y - yhat1.detach()).double() ** 2).sum()])
        out = torch.zeros(1, dtype=torch.double, device=y.device)
        ops.mals_finalize(l01, self.lam0, self.lam1, self.delta, self.lamdiff, 1, loss=out)
        return out[0].to(torch.float32)


class QuadraticDiscriminantFilter(nn.Module):
    """Reference model/disentangle.py:90-232: per class two one-vs-rest quadratic (Gaussian) classifiers of the latent mean
    with streaming means / covariances and self-tuning forgetting factors lama < lamb.  Same constructor and buffers (m0a,
    S0a, m1a, S1a, m0b, S0b, m1b, S1b, lama, lamb: the reference's state_dict keys and registration order); the inverses and
    log-determinants, the log-likelihood-ratio loss with its gradient into mu, the forgetting-factor rule and the
    class-conditional update run in csrc/scv_qda.cu, sequenced by the engine."""

    def __init__(self, nx, classes, lamdiff=1e-2, delta=1e-3):
        super().__init__()
        self.classes = classes
        n_classes = len(classes)
        if nx > 128 or n_classes > 16:
            raise NotImplementedError("scrubvae_b200: the qda kernels hold z <= 128 and <= 16 classes")
        for name in ["0a", "1a", "0b", "1b"]:
            self.register_buffer("m{}".format(name), torch.zeros(n_classes, nx))
            self.register_buffer("S{}".format(name), torch.eye(nx)[None, :].repeat(n_classes, 1, 1))
        self.register_buffer("lama", torch.ones(n_classes) * 0.2)
        self.register_buffer("lamb", torch.ones(n_classes) * 0.2 + lamdiff)
        self.delta = delta
        self.lamdiff = lamdiff
        self.z = int(nx)
        self._ops = None

    def forward(self, *args, **kwargs):
        return

    def m4(self):
        return [self.m0a, self.m1a, self.m0b, self.m1b]

    def S4(self):
        return [self.S0a, self.S1a, self.S0b, self.S1b]


class MovingAverageFilter(nn.Module):
    """Reference model/disentangle.py:9-88: per class two running means of the latent mean with self-tuning forgetting
    factors; the loss is the norm of the pairwise differences between the classes' mean estimates.  Same constructor and
    buffers (m1, m2, lam1, lam2); kernels scv_ma_* (csrc/scv_qda.cu), sequenced by the engine."""

    def __init__(self, nx, classes, lamdiff=1e-2, delta=1e-3):
        super().__init__()
        self.classes = classes
        if nx > 128 or len(classes) > 16:
            raise NotImplementedError("scrubvae_b200: the moving_avg kernels hold z <= 128 and <= 16 classes")
        self.register_buffer("m1", torch.zeros(len(self.classes), nx))
        self.register_buffer("m2", torch.zeros(len(self.classes), nx))
        self.register_buffer("lam1", torch.ones(len(self.classes)) * 0.5)
        self.register_buffer("lam2", torch.ones(len(self.classes)) * 0.5 + lamdiff)
        self.delta = delta
        self.lamdiff = lamdiff
        self._ops = None

    def forward(self, *args, **kwargs):
        return
