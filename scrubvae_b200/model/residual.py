"""SC-VAE modules with the reference's class names, constructor signatures, attribute names,
construction order (hence identical RNG consumption / initial weights under the same seed) and
state_dict keys — mirror of the reference's model/residual.py.

The `nn.Conv1d` / `nn.BatchNorm1d` / ... children below are PARAMETER CONTAINERS only: no torch
kernel is ever run through them.  All arithmetic is done by the sm_100a kernels of libscv.so,
sequenced by `scrubvae_b200.engine.Engine` (ResVAE.forward / encode / decode call into it).
Calling `forward` on an inner block raises: there is no PyTorch fallback path.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def find_latent_dim(window_size: int, kernel: int, num_layers: int, dilation=None) -> int:
    """Reference model/residual.py:6-20 (stride-2 case)."""
    l_out = window_size
    for _ in range(num_layers):
        l_out = (l_out + 2 * (kernel // 2) - (kernel - 1) - 1) / 2 + 1
    return int(l_out)


def find_out_dim(latent_dim: int, kernel: int, num_layers: int, dilation=None) -> int:
    """Reference model/residual.py:23-36 (stride-2 case)."""
    l_out = latent_dim
    for _ in range(num_layers):
        l_out = (l_out - 1) * 2 - 2 * (kernel // 2) + (kernel - 1) + 1
    return int(l_out)


class _EngineOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container; the SC-VAE step runs through "
            "ResVAE.forward (sm_100a kernels) — there is no per-module PyTorch path.")


class CholeskyL(_EngineOnly):
    """Reference model/residual.py:39-68; fused into scv_reparam_fwd."""

    def __init__(self, z_dim: int, is_diag: bool):
        super().__init__()
        self.z_dim, self.is_diag = z_dim, is_diag


class ResidualBlock(_EngineOnly):
    """Reference model/residual.py:71-119."""

    def __init__(self, in_channels, out_channels, kernel=3, activation="prelu", dilation=1):
        super().__init__()
        _check_supported(activation, dilation)
        self.residual = nn.Sequential(
            nn.Conv1d(in_channels, out_channels // 2, kernel, 2, kernel // 2, bias=True),
            nn.BatchNorm1d(out_channels // 2, eps=1e-4),
            nn.PReLU(),
            nn.Conv1d(out_channels // 2, out_channels, kernel, 1, kernel // 2, bias=True),
        )
        self.skip = nn.Conv1d(in_channels, out_channels, kernel, 2, kernel // 2, bias=True)
        self.add = nn.Sequential(nn.BatchNorm1d(out_channels, eps=1e-4), nn.PReLU())


class ResidualBlockTranspose(_EngineOnly):
    """Reference model/residual.py:122-180."""

    def __init__(self, in_channels, out_channels, kernel=3, scale_factor=2, activation="prelu", dilation=1):
        super().__init__()
        _check_supported(activation, dilation)
        self.residual = nn.Sequential(
            nn.ConvTranspose1d(in_channels, in_channels // 2, kernel, 1, kernel // 2, bias=True),
            nn.BatchNorm1d(in_channels // 2, eps=1e-4),
            nn.PReLU(),
            nn.ConvTranspose1d(in_channels // 2, out_channels, kernel, 2, kernel // 2, bias=True),
        )
        self.skip = nn.Sequential(
            nn.Upsample(scale_factor=scale_factor, mode="linear", align_corners=False),
            nn.Conv1d(in_channels, out_channels, kernel + 1, 1, kernel // 2, bias=True),
        )
        self.add = nn.Sequential(nn.BatchNorm1d(out_channels, eps=1e-4), nn.PReLU())


def _check_supported(activation, dilation):
    if activation != "prelu":
        raise NotImplementedError("scrubvae_b200: only activation='prelu' (the reference default) is built")
    if dilation != 1:
        raise NotImplementedError("scrubvae_b200: init_dilation is not supported (reference default None)")


class ResidualEncoder(_EngineOnly):
    """Reference model/residual.py:183-240."""

    def __init__(self, in_channels, ch=[64, 128, 256, 512, 1024], kernel=5, z_dim=128, window=200,
                 activation="prelu", is_diag=False, prior="gaussian", init_dilation=None):
        super().__init__()
        if prior != "gaussian":
            raise NotImplementedError("scrubvae_b200: only the gaussian prior is built (reference default)")
        if init_dilation is not None:
            raise NotImplementedError("scrubvae_b200: init_dilation is not supported")
        self.prior = prior
        self.conv_in = nn.Conv1d(in_channels, ch[0], 7, 1, 3)
        self.activation = nn.PReLU()
        if activation != "prelu":
            _check_supported(activation, 1)
        self.res_layers = nn.Sequential(
            *[ResidualBlock(ch[i], ch[i + 1], kernel, activation, 1) for i in range(len(ch) - 1)])
        self.flatten = nn.Flatten()
        flatten_dim = find_latent_dim(window, kernel, len(ch) - 1) * ch[-1]
        sig_dim = z_dim if is_diag else z_dim * (z_dim + 1) // 2  # reference model/residual.py:216
        self.fc_mu = nn.Linear(flatten_dim, z_dim)
        self.fc_sigma = nn.Sequential(nn.Linear(flatten_dim, sig_dim), CholeskyL(z_dim, is_diag))


class ResidualDecoder(_EngineOnly):
    """Reference model/residual.py:243-292."""

    def __init__(self, out_channels, ch=[64, 128, 256, 512, 1024], kernel=5, z_dim=128, window=200,
                 activation="prelu", conditional_dim=0, init_dilation=None):
        super().__init__()
        if init_dilation is not None:
            raise NotImplementedError("scrubvae_b200: init_dilation is not supported")
        self.conditional_dim = conditional_dim
        flatten_dim = find_latent_dim(window, kernel, len(ch) - 1) * ch[-1]
        self.fc_in = nn.Linear(z_dim + conditional_dim, flatten_dim)
        self.unflatten = nn.Unflatten(1, (ch[-1], -1))
        self.res_layers = nn.Sequential(
            *[ResidualBlockTranspose(ch[-i], ch[-i - 1], kernel, activation=activation, dilation=1)
              for i in range(1, len(ch))])
        l_out = find_out_dim(find_latent_dim(window, kernel, len(ch) - 1), kernel, len(ch) - 1)
        final_kernel = window - l_out + 7
        self.conv_out = nn.ConvTranspose1d(ch[0], out_channels, final_kernel, 1, 3)


class VAE(nn.Module):
    """Reference model/residual.py:295-362."""

    def __init__(self, prior="gaussian"):
        super().__init__()
        self.prior = prior
        self.dist_params = ["mu", "L"]


class ResVAE(VAE):
    """Reference model/residual.py:365-491.  Same constructor; forward/encode/decode run on the
    sm_100a engine and return the reference's `data_o` dictionary."""

    def __init__(self, in_channels, ch=[64, 128, 256, 512, 1024], kernel=5, z_dim=128, window=200,
                 activation="prelu", is_diag=False, conditional_dim=0, init_dilation=None, disentangle=None,
                 kinematic_tree=None, arena_size=None, disentangle_keys=None, conditional_keys=None,
                 discrete_classes=None, prior="gaussian", precision="tf32"):
        super().__init__(prior=prior)
        if kernel % 2 != 1:
            raise NotImplementedError("scrubvae_b200: odd kernel sizes only")
        if arena_size is None:
            raise NotImplementedError("scrubvae_b200: arena_size is required (root channels are part of the input)")
        self.in_channels, self.ch, self.kernel, self.z_dim, self.window = in_channels, list(ch), kernel, z_dim, window
        self.is_diag, self.conditional_dim = is_diag, conditional_dim
        self.kinematic_tree = kinematic_tree
        self.register_buffer("arena_size", torch.as_tensor(arena_size, dtype=torch.float32))
        self.disentangle_keys = disentangle_keys
        self.conditional_keys = conditional_keys
        self.discrete_classes = discrete_classes if discrete_classes is not None else {}
        self.precision = precision
        self.encoder = ResidualEncoder(in_channels, ch=ch, kernel=kernel, z_dim=z_dim, window=window,
                                       activation=activation, is_diag=is_diag, prior=prior,
                                       init_dilation=init_dilation)
        self.decoder = ResidualDecoder(in_channels, ch=ch, kernel=kernel, z_dim=z_dim, window=window,
                                       activation=activation, conditional_dim=conditional_dim,
                                       init_dilation=init_dilation)
        if disentangle is not None:
            self.disentangle = nn.ModuleDict()
            for k, v in disentangle.items():
                self.disentangle[k] = nn.ModuleDict(v)
        else:
            self.disentangle = nn.ModuleDict()
        for method in self.disentangle.keys():
            if method not in ("grad_reversal", "moving_avg_lsq", "qda", "moving_avg"):
                raise NotImplementedError(
                    f"scrubvae_b200: scrubber method '{method}' is outside the built hot path "
                    "(conditional, grad_reversal, moving_avg_lsq, qda, moving_avg; SURVEY.md §8)")
        self.mi_estimator = None
        self._engine = None
        self._noise = None  # test hook: injected reparameterisation noise (B, z)

    # -- engine plumbing
    @property
    def engine(self):
        if self._engine is None:
            from ..engine import Engine
            self._engine = Engine(self)
        return self._engine

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if self._engine is not None:
            self._engine.invalidate()
        return out

    # parameters may temporarily live in the packed GEMM layout (engine.TrainStep(resident=True)): refresh the
    # reference-layout views before anything reads or replaces them through the nn.Module API
    def state_dict(self, *a, **k):
        if self._engine is not None:
            self._engine.ensure_flat()
        return super().state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        if self._engine is not None:
            self._engine.ensure_flat()
            self._engine.resident_valid = False
        return super().load_state_dict(*a, **k)

    def normalize_root(self, root):
        return 2 * (root - self.arena_size[0]) / (self.arena_size[1] - self.arena_size[0]) - 1

    def inv_normalize_root(self, norm_root):
        return 0.5 * (norm_root + 1) * (self.arena_size[1] - self.arena_size[0]) + self.arena_size[0]

    def forward(self, data):
        return self.engine.forward(data, training=self.training)

    def encode(self, data):
        return self.engine.encode(data, training=self.training)

    def decode(self, z, data):
        return self.engine.decode(z, data, training=self.training)
