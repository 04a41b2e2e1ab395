"""CPU oracle for the SC-VAE training step (TEST INFRASTRUCTURE — NOT PRODUCT CODE).

A functional, from-scratch restatement in plain PyTorch-on-CPU (fp32) / numpy of the
reference algorithm on the north-star hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module; the product package `scrubvae_b200` never does.

Parity pin: validated against the UNMODIFIED reference imported from /root/reference
(tests/golden/make_golden.py writes fixtures; tests/test_oracle.py checks the oracle
against them, and against the live reference when the tree is present).  The reference
itself ships no tests or golden vectors, so the pin is "outputs of the
reference itself run here".

Every function cites the reference file:line (relative to /root/reference/src/scrubvae)
it follows.  Weights are passed as a dict keyed by the reference's state_dict names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# configs/mouse_skeleton.yaml:86-92 and :95-112
KINEMATIC_TREE = [[0, 1, 2, 3, 4], [0, 5], [1, 6, 7, 8], [1, 9, 10, 11],
                  [5, 12, 13, 14], [5, 15, 16, 17]]
OFFSET = [[0, 0, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0], [-1, 0, 0],
          [0, 1, 0], [0, 1, 0], [0, 1, 0], [0, -1, 0], [0, -1, 0], [0, -1, 0],
          [0, 1, 0], [0, 1, 0], [0, 1, 0], [0, -1, 0], [0, -1, 0], [0, -1, 0]]
ARENA = [[-100.0, -100.0, 0.0], [100.0, 100.0, 50.0]]

FEAT_DIM = {"avg_speed": 1, "part_speed": 4, "avg_speed_3d": 3, "heading": 2,
            "heading_change": 1, "fluorescence": 1}  # get/model.py:19-27


# --------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------
def cont6d_to_matrix(c: torch.Tensor, eps: float = 0.0) -> torch.Tensor:
    """data/quaternion.py:337-353.  Columns [x y z]."""
    xr, yr = c[..., 0:3], c[..., 3:6]
    x = xr / (xr.norm(dim=-1, keepdim=True) + eps)
    z = torch.linalg.cross(x, yr, dim=-1)
    z = z / (z.norm(dim=-1, keepdim=True) + eps)
    y = torch.linalg.cross(z, x, dim=-1)
    return torch.stack([x, y, z], dim=-1)


def fwd_kin(c6d: torch.Tensor, offsets: torch.Tensor, root: torch.Tensor,
            tree=KINEMATIC_TREE, eps: float = 0.0) -> torch.Tensor:
    """data/dataset.py:83-116 (do_root_R=True).  c6d (F,J,6), offsets (F,J,3), root (F,3).
    Every chain restarts from the ROOT joint's rotation (joint 0), also chains whose first
    joint is not joint 0."""
    Fn, J = c6d.shape[0], c6d.shape[1]
    mats = cont6d_to_matrix(c6d, eps)  # (F,J,3,3)
    pose: List[Optional[torch.Tensor]] = [None] * J
    pose[0] = root
    for chain in tree:
        R = mats[:, 0]
        for i in range(1, len(chain)):
            j = chain[i]
            R = R @ mats[:, j]
            pose[j] = (R @ offsets[:, j].unsqueeze(-1)).squeeze(-1) + pose[chain[i - 1]]
    return torch.stack(pose, dim=1)


def quat_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """data/quaternion.py:291-317."""
    r, i, j, k = q.unbind(-1)
    s = 2.0 / (q * q).sum(-1)
    o = torch.stack([1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
                     s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
                     s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)], -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def quat_to_cont6d(q: torch.Tensor) -> torch.Tensor:
    """data/quaternion.py:325-334: first two COLUMNS of R(q), concatenated."""
    m = quat_to_matrix(q)
    return torch.cat([m[..., 0], m[..., 1]], dim=-1)


def qmul(q: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """data/quaternion.py:34-52 (Hamilton product q*r, real part first)."""
    qw, qx, qy, qz = q.unbind(-1)
    rw, rx, ry, rz = r.unbind(-1)
    return torch.stack([rw * qw - rx * qx - ry * qy - rz * qz,
                        rw * qx + rx * qw - ry * qz + rz * qy,
                        rw * qy + rx * qz + ry * qw - rz * qx,
                        rw * qz - rx * qy + ry * qx + rz * qw], dim=-1)


def qinv(q: torch.Tensor) -> torch.Tensor:
    """data/quaternion.py:17-21."""
    return q * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=q.dtype)


def qrot(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """data/quaternion.py:55-74."""
    qv = q[..., 1:]
    uv = torch.linalg.cross(qv, v, dim=-1)
    uuv = torch.linalg.cross(qv, uv, dim=-1)
    return v + 2 * (q[..., :1] * uv + uuv)


def qbetween(v0: torch.Tensor, v1: torch.Tensor) -> torch.Tensor:
    """data/quaternion.py:409-420."""
    v = torch.linalg.cross(v0, v1, dim=-1)
    w = torch.sqrt((v0 ** 2).sum(-1, keepdim=True) * (v1 ** 2).sum(-1, keepdim=True)) \
        + (v0 * v1).sum(-1, keepdim=True)
    q = torch.cat([w, v], dim=-1)
    return q / q.norm(dim=-1, keepdim=True)


# --------------------------------------------------------------------------------------
# preprocessing (numpy in, like the reference; fp32 after the *_np casts)
# --------------------------------------------------------------------------------------
def window_indices(ids: np.ndarray, stride: int, window: int) -> np.ndarray:
    """data/dataset.py:198-233.  int64 (N_w, window); bit-exact contract."""
    ids = np.asarray(ids)
    n = len(ids)
    change = np.concatenate([[0], np.where(np.diff(ids, prepend=ids[0]) != 0)[0], [n]])
    out = []
    for a, b in zip(change[:-1], change[1:]):
        if b - a >= window:
            starts = np.arange(a, b - window + 1, stride, dtype=np.int64)
            out.append(starts[:, None] + np.arange(window, dtype=np.int64)[None, :])
    if not out:
        raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")  # :231
    return np.concatenate(out, axis=0)


def speed_outliers(pose: np.ndarray, threshold: float = 2.25) -> np.ndarray:
    """data/dataset.py:299-309.  pose (N,W,J,3) float64."""
    d = np.diff(pose, n=1, axis=-3)
    spd = np.sqrt((d ** 2).sum(-1)).mean(axis=(-1, -2))
    return np.where(spd > threshold)[0]


def speed_parts(pose: np.ndarray) -> np.ndarray:
    """data/dataset.py:134-163 + :362-374 -> avg_speed_3d (N,3) float64."""
    parts = [[0, 1, 2, 3, 4, 5], [1, 6, 7, 8, 9, 10, 11], [5, 12, 13, 14, 15, 16, 17]]
    root_spd = np.sqrt((np.diff(pose[..., 0, :], n=1, axis=-2) ** 2).sum(-1)).mean(-1)
    out = np.zeros((len(root_spd), len(parts) + 1))
    out[:, 0] = root_spd
    cen = pose - pose[..., 0:1, :]
    for i, part in enumerate(parts):
        pp = cen if part[0] == 0 else cen - cen[:, part[0]:part[0] + 1, :]
        rel = (np.diff(pp[..., part[1:], :], n=1, axis=-3) ** 2).sum(-1)
        out[:, i + 1] = np.sqrt(rel).mean(axis=(-1, -2))
    return np.concatenate([out[:, :2], out[:, 2:].mean(axis=-1, keepdims=True)], axis=-1)


def frame_yaw(pose_mid: np.ndarray) -> np.ndarray:
    """data/dataset.py:236-243 with root_i=0, front_i=1."""
    f = pose_mid[:, 1, :] - pose_mid[:, 0, :]
    f = f / np.linalg.norm(f, axis=-1)[..., None]
    return -np.arctan2(f[:, 1], f[:, 0])


def inv_kin(pose: np.ndarray, tree=KINEMATIC_TREE, offset=OFFSET) -> np.ndarray:
    """data/dataset.py:11-46 with forward_indices=[1,0] (:402).  pose (F,J,3) float64 ->
    local quaternions (F,J,4) float64 holding fp32-rounded values (the *_np helpers cast
    to float and the results are stored into a float64 array)."""
    pose = np.asarray(pose, dtype=np.float64)
    off = torch.tensor(np.asarray(offset), dtype=torch.float32)
    fwd = pose[:, 0, :] - pose[:, 1, :]
    fwd = fwd / np.linalg.norm(fwd, axis=-1)[..., None]
    target = torch.tensor([[1.0, 0.0, 0.0]]).expand(len(fwd), 3)
    root_q = qbetween(torch.from_numpy(fwd).float(), target).clone()
    root_q[0] = torch.tensor([1.0, 0.0, 0.0, 0.0])  # quirk (iv): global frame 0 only
    local = np.zeros(pose.shape[:-1] + (4,))
    local[:, 0] = root_q.numpy()
    for chain in tree:
        R = root_q
        for i in range(len(chain) - 1):
            u = off[chain[i + 1]][None].expand(len(pose), 3)
            v = pose[:, chain[i + 1]] - pose[:, chain[i]]
            v = v / np.linalg.norm(v, axis=-1)[..., None]
            ruv = qbetween(u, torch.from_numpy(v).float())
            Rloc = qmul(qinv(R), ruv)
            local[:, chain[i + 1], :] = Rloc.numpy()
            R = qmul(R, Rloc)
    return local


def segment_len(pose: np.ndarray, tree=KINEMATIC_TREE, offset=OFFSET) -> np.ndarray:
    """data/dataset.py:279-296.  Quirk (iii): `offset` is an int64 array (YAML ints), so the
    scaled offsets are truncated toward zero when stored."""
    off = np.asarray(offset)
    parents = [0] * len(off)
    parents[0] = -1
    for chain in tree:
        for j in range(1, len(chain)):
            parents[chain[j]] = chain[j - 1]
    out = np.moveaxis(np.tile(off[..., None], pose.shape[0]), -1, 0).copy()  # int64
    for i in range(1, off.shape[0]):
        ln = np.linalg.norm(pose[:, i, :] - pose[:, parents[i], :], axis=1)[..., None]
        out[:, i] = ln * out[:, i]  # float -> int64 truncation
    return out


def preprocess(pose_frames: np.ndarray, ids_frames: np.ndarray, window: int = 51,
               stride: int = 2, speed_threshold: Optional[float] = 2.25,
               direction_process: str = "midfwd") -> Dict[str, torch.Tensor]:
    """data/dataset.py:313-454 with data_keys = x6d, root, offsets, target_pose, heading,
    avg_speed_3d, ids.  pose_frames (N,J,3) float64, ids_frames (N,) int."""
    winds = window_indices(ids_frames, stride, window)
    pose = pose_frames[winds]
    ids = ids_frames[winds][:, window // 2]
    if speed_threshold is not None:
        bad = speed_outliers(pose, speed_threshold)
        pose = np.delete(pose, bad, 0)
        ids = np.delete(ids, bad, 0)
    data = {"raw_pose": pose, "avg_speed_3d": speed_parts(pose)}
    yaw = frame_yaw(pose[:, window // 2])[..., None]
    data["heading"] = np.concatenate([np.sin(yaw), np.cos(yaw)], axis=-1)  # :260-267
    root = pose[..., 0, :].copy()
    if direction_process in ("midfwd", "x360"):
        c = np.zeros(root.shape)
        c[..., [0, 1]] = root[:, window // 2, [0, 1]][:, None, :]
        root -= c
    lq = inv_kin(pose.reshape((-1,) + pose.shape[-2:])).reshape(pose.shape[:-1] + (-1,))
    if direction_process == "midfwd":
        fq = np.zeros((len(yaw), 4))
        fq[:, 3] = np.sin(yaw / 2)[:, 0]
        fq[:, 0] = np.cos(yaw / 2)[:, 0]
        fq = np.repeat(fq[:, None, :], window, axis=1)
        fqt = torch.from_numpy(fq).float()
        lq[..., 0, :] = qmul(fqt, torch.from_numpy(lq[..., 0, :]).float()).numpy()
        root = qrot(fqt, torch.from_numpy(root).float()).numpy()
    data["x6d"] = quat_to_cont6d(torch.from_numpy(lq).float()).numpy()
    data["offsets"] = segment_len(pose.reshape((-1,) + pose.shape[-2:])).reshape(pose.shape)
    data["root"] = root
    out = {k: torch.tensor(v, dtype=torch.float32) for k, v in data.items()}
    out["ids"] = torch.tensor(ids, dtype=torch.int16)
    x = out["x6d"].reshape((-1,) + out["x6d"].shape[-2:])
    o = out["offsets"].reshape(x.shape[:2] + (-1,))
    out["target_pose"] = fwd_kin(x, o, torch.zeros(x.shape[0], 3), eps=1e-8).reshape(
        out["x6d"].shape[:-1] + (3,))
    out["window_inds"] = torch.from_numpy(winds)
    return out


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synth_batch(B: int, window: int = 51, seed: int = 0, n_ids: int = 4) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    J = 18
    q = torch.randn(B, window, J, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    x6d = quat_to_cont6d(q)
    offd = torch.tensor(OFFSET, dtype=torch.float32)
    ln = torch.trunc(5 + 25 * torch.rand(B, 1, J, 1, generator=g))
    offsets = (offd[None, None] * ln).expand(B, window, J, 3).contiguous()
    lo, hi = torch.tensor(ARENA[0]), torch.tensor(ARENA[1])
    root = lo + (hi - lo) * torch.rand(B, window, 3, generator=g)
    psi = (torch.rand(B, generator=g) * 2 - 1) * math.pi
    heading = torch.stack([torch.sin(psi), torch.cos(psi)], dim=-1)
    speed = torch.randn(B, 3, generator=g)
    ids = torch.randint(0, n_ids, (B, 1), generator=g).to(torch.int16)
    tp = fwd_kin(x6d.reshape(-1, J, 6), offsets.reshape(-1, J, 3),
                 torch.zeros(B * window, 3), eps=1e-8).reshape(B, window, J, 3)
    return {"x6d": x6d.contiguous(), "root": root, "offsets": offsets, "target_pose": tp,
            "heading": heading, "avg_speed_3d": speed, "ids": ids}


def synth_eps(B: int, z_dim: int, seed: int = 2) -> torch.Tensor:
    return torch.randn(B, z_dim, generator=torch.Generator().manual_seed(seed))


# --------------------------------------------------------------------------------------
# model forward (functional; sd = reference state_dict)
# --------------------------------------------------------------------------------------
class Cfg:
    def __init__(self, ch=(64, 128, 256, 512, 1024), kernel=5, z_dim=64, window=51,
                 conditional=("heading",), grad_reversal=("heading",), alpha=1.0,
                 n_keypts=18, discrete_classes=None, arena=ARENA, is_diag=False):
        self.ch, self.kernel, self.z_dim, self.window = list(ch), kernel, z_dim, window
        self.conditional, self.grad_reversal = list(conditional), list(grad_reversal)
        self.alpha, self.n_keypts = alpha, n_keypts
        self.discrete_classes = discrete_classes or {}
        self.arena, self.is_diag = arena, is_diag
        self.in_channels = n_keypts * 6 + 3  # get/model.py:33-35

    def feat_dim(self, k):
        return len(self.discrete_classes[k]) if k in self.discrete_classes else FEAT_DIM[k]


def _bn(x, sd, pre, training, new_stats, eps=1e-4, momentum=0.1):
    """nn.BatchNorm1d(eps=1e-4): model/residual.py:88,112,146,173; SURVEY App. B.  F.batch_norm is
    the functional form of the same torch module (train: batch mean / biased variance, running
    statistics updated with the unbiased variance)."""
    rm, rv = sd[pre + ".running_mean"].clone(), sd[pre + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, sd[pre + ".weight"], sd[pre + ".bias"], training, momentum, eps)
    if training and new_stats is not None:
        new_stats[pre + ".running_mean"] = rm
        new_stats[pre + ".running_var"] = rv
        new_stats[pre + ".num_batches_tracked"] = sd[pre + ".num_batches_tracked"] + 1
    return y


def _prelu(x, a):
    return F.prelu(x, a)


def upsample2_linear(x):
    """nn.Upsample(scale_factor=2, mode='linear', align_corners=False), residual.py:160."""
    L = x.shape[-1]
    left = torch.cat([x[..., :1], x[..., :-1]], dim=-1)
    right = torch.cat([x[..., 1:], x[..., -1:]], dim=-1)
    even = 0.25 * left + 0.75 * x
    odd = 0.75 * x + 0.25 * right
    return torch.stack([even, odd], dim=-1).reshape(x.shape[:-1] + (2 * L,))


def encoder(sd, x, cfg: Cfg, training, new_stats):
    """model/residual.py:183-240; blocks :71-119."""
    p = "encoder."
    h = _prelu(F.conv1d(x, sd[p + "conv_in.weight"], sd[p + "conv_in.bias"], 1, 3),
               sd[p + "activation.weight"])
    k = cfg.kernel
    for i in range(len(cfg.ch) - 1):
        b = f"{p}res_layers.{i}."
        skip = F.conv1d(h, sd[b + "skip.weight"], sd[b + "skip.bias"], 2, k // 2)
        r = F.conv1d(h, sd[b + "residual.0.weight"], sd[b + "residual.0.bias"], 2, k // 2)
        r = _prelu(_bn(r, sd, b + "residual.1", training, new_stats), sd[b + "residual.2.weight"])
        r = F.conv1d(r, sd[b + "residual.3.weight"], sd[b + "residual.3.bias"], 1, k // 2)
        h = _prelu(_bn(r + skip, sd, b + "add.0", training, new_stats), sd[b + "add.1.weight"])
    flat = h.flatten(1)
    mu = F.linear(flat, sd[p + "fc_mu.weight"], sd[p + "fc_mu.bias"])
    sig = F.linear(flat, sd[p + "fc_sigma.0.weight"], sd[p + "fc_sigma.0.bias"])
    return mu, cholesky_L(sig, cfg.z_dim, cfg.is_diag)


def cholesky_L(sig, z, is_diag=False):
    """model/residual.py:39-68: row-major tril scatter (is_diag: the z values go on the diagonal, :55-56),
    softplus on the diagonal."""
    idx = torch.arange(z)[None, :].repeat(2, 1) if is_diag else torch.tril_indices(z, z)
    L = torch.zeros(sig.shape[0], z, z, dtype=sig.dtype)
    L = L.index_put((torch.arange(sig.shape[0])[:, None], idx[0][None], idx[1][None]), sig)
    d = F.softplus(torch.diagonal(L, dim1=-2, dim2=-1))
    return L - torch.diag_embed(torch.diagonal(L, dim1=-2, dim2=-1)) + torch.diag_embed(d)


def decoder(sd, zc, cfg: Cfg, training, new_stats):
    """model/residual.py:243-292; blocks :122-180."""
    p = "decoder."
    h = F.linear(zc, sd[p + "fc_in.weight"], sd[p + "fc_in.bias"]).unflatten(1, (cfg.ch[-1], -1))
    k = cfg.kernel
    for i in range(len(cfg.ch) - 1):
        b = f"{p}res_layers.{i}."
        skip = F.conv1d(upsample2_linear(h), sd[b + "skip.1.weight"], sd[b + "skip.1.bias"], 1, k // 2)
        r = F.conv_transpose1d(h, sd[b + "residual.0.weight"], sd[b + "residual.0.bias"], 1, k // 2)
        r = _prelu(_bn(r, sd, b + "residual.1", training, new_stats), sd[b + "residual.2.weight"])
        r = F.conv_transpose1d(r, sd[b + "residual.3.weight"], sd[b + "residual.3.bias"], 2, k // 2)
        h = _prelu(_bn(r + skip, sd, b + "add.0", training, new_stats), sd[b + "add.1.weight"])
    return torch.tanh(F.conv_transpose1d(h, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"], 1, 3))


def mlp_ensemble(sd, pre, z):
    """model/disentangle.py:583-632: four ReLU MLPs on the same input, returned as a list."""
    outs = []
    for name in ("mlp1", "mlp2", "mlp3", "mlp4"):
        h = z
        idxs = sorted({int(k[len(pre) + len(name) + 1:].split(".")[0]) for k in sd
                       if k.startswith(pre + name + ".")})
        for n, li in enumerate(idxs):
            h = F.linear(h, sd[f"{pre}{name}.{li}.weight"], sd[f"{pre}{name}.{li}.bias"])
            if n < len(idxs) - 1:
                h = F.relu(h)
        outs.append(h)
    return outs


class _RevGrad(torch.autograd.Function):
    """model/disentangle.py:541-553."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return -ctx.alpha * g, None


def forward(sd: Dict[str, torch.Tensor], data: Dict[str, torch.Tensor], cfg: Cfg,
            eps_noise: Optional[torch.Tensor], training: bool = True,
            new_stats: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """VAE.forward model/residual.py:318-362 with ResVAE.encode :438-459 / decode :461-491."""
    arena = sd["arena_size"]
    B = data["x6d"].shape[0]
    nr = 2 * (data["root"] - arena[0]) / (arena[1] - arena[0]) - 1  # :428-431
    x_in = torch.cat([data["x6d"].reshape(B, cfg.window, -1), nr], dim=-1)
    mu, L = encoder(sd, x_in.moveaxis(1, -1), cfg, training, new_stats)
    if training:
        z = torch.matmul(L, eps_noise[..., None]).squeeze(-1) + mu  # :315-316
    else:
        z = mu
    out = {"mu": mu, "L": L, "z": z}
    zc = z
    if cfg.conditional:
        var = [F.one_hot(data[k].ravel().long(), cfg.feat_dim(k)).to(z.dtype)
               if k in cfg.discrete_classes else data[k] for k in cfg.conditional]
        out["var"] = torch.cat(var, dim=-1)
        zc = torch.cat([z, out["var"]], dim=-1)
    xh = decoder(sd, zc, cfg, training, new_stats).moveaxis(-1, 1)  # (B,W,C)
    nroot = xh[..., -3:]
    out["root"] = 0.5 * (nroot + 1) * (arena[1] - arena[0]) + arena[0]  # :433-436
    out["x6d"] = xh[..., :-3].reshape(B, cfg.window, -1, 6)
    out["disentangle"] = {"grad_reversal": {}}
    for k in cfg.grad_reversal:
        pre = f"disentangle.grad_reversal.{k}.reversal.1."
        out["disentangle"]["grad_reversal"][k] = mlp_ensemble(sd, pre, _RevGrad.apply(mu, cfg.alpha))
    return out


def batch_loss(data, out, cfg: Cfg, loss_scale: Dict[str, float]) -> Dict[str, torch.Tensor]:
    """train/losses.py:182-324 (prior :138-146, jpe :148-171, root :216-219, *_gr :267-284)."""
    B = data["x6d"].shape[0]
    res = {}
    if "prior" in loss_scale:
        mu, L = out["mu"], out["L"]
        var = torch.matmul(L, L.transpose(-1, -2))
        res["prior"] = -0.5 * torch.sum(1 + 2 * torch.log(torch.diagonal(L, dim1=-1, dim2=-2))
                                        - mu.pow(2) - torch.diagonal(var, dim1=-1, dim2=-2)) / B
    if "jpe" in loss_scale:
        tp = data["target_pose"]
        xh = out["x6d"]
        ph = fwd_kin(xh.reshape((-1,) + xh.shape[-2:]),
                     data["offsets"].reshape((-1,) + data["offsets"].shape[-2:]),
                     torch.zeros(tp.shape[0] * tp.shape[1], 3), eps=1e-8).reshape(tp.shape)
        res["jpe"] = torch.sum((tp - ph) ** 2) / (tp.shape[0] * tp.shape[-1] * tp.shape[-2])
    if "root" in loss_scale:
        res["root"] = torch.sum((out["root"] - data["root"]) ** 2) / B
    nk = len(cfg.grad_reversal)
    for key in cfg.grad_reversal:
        acc = 0
        ens = out["disentangle"]["grad_reversal"][key]
        for e in ens:
            if key == "ids":
                acc = acc + F.cross_entropy(e, data[key].ravel().long(), reduction="sum")
            else:
                acc = acc + torch.sum((e - data[key]) ** 2)
            acc = acc / len(ens) / nk / B  # quirk (i): normalisation INSIDE the loop
        res[key + "_gr"] = acc
    res["total"] = sum(loss_scale[k] * res[k] for k in list(res.keys()) if loss_scale[k] != 0)
    return res


# --------------------------------------------------------------------------------------
# step tail (train/trainer.py:158-167)
# --------------------------------------------------------------------------------------
def param_names(sd) -> List[str]:
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked") or k == "arena_size")]


def clip_coef(grads: Sequence[torch.Tensor], max_norm: float = 1e6):
    """torch.nn.utils.clip_grad_norm_ as used at train/trainer.py:164."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adam_update(p, g, m, v, step, lr, kind="adamw", b1=0.9, b2=0.999, eps=1e-8, wd=0.01):
    """torch.optim.Adam / AdamW defaults (train/trainer.py:60-65)."""
    p = p * (1 - lr * wd) if kind == "adamw" else p.clone()
    m = torch.lerp(m, g, 1 - b1)
    v = (v * b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    return p.addcdiv_(m, denom, value=-(lr / bc1)), m, v


def train_step(sd, data, cfg: Cfg, loss_scale, eps_noise, opt_state=None, lr=1e-4,
               optimizer="adamw", step=1):
    """One train_test_epoch(mode='train') iteration body, train/trainer.py:126-182.
    Returns (losses, grads, new_sd, new_opt_state)."""
    names = param_names(sd)
    work = {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    new_stats = {}
    out = forward(work, data, cfg, eps_noise, True, new_stats)
    losses = batch_loss(data, out, cfg, loss_scale)
    grads_l = torch.autograd.grad(losses["total"], [work[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads_l)}
    _, coef = clip_coef(list(grads.values()))
    new_sd = dict(sd)
    new_sd.update(new_stats)
    opt_state = opt_state or {k: (torch.zeros_like(sd[k]), torch.zeros_like(sd[k])) for k in names}
    new_opt = {}
    for k in names:
        g = grads[k] * coef
        p, m, v = adam_update(sd[k], g, opt_state[k][0], opt_state[k][1], step, lr, optimizer)
        new_sd[k], new_opt[k] = p, (m, v)
    return {k: v.detach() for k, v in losses.items()}, grads, new_sd, new_opt, out


# --------------------------------------------------------------------------------------
# deterministic weight init for tests / bench (NOT the reference's init: a fixed recipe
# both sides can regenerate from a seed without the reference being present)
# --------------------------------------------------------------------------------------
def state_dict_shapes(cfg: Cfg) -> Dict[str, tuple]:
    ch, k, z, W = cfg.ch, cfg.kernel, cfg.z_dim, cfg.window
    s: Dict[str, tuple] = {"arena_size": (2, 3)}

    def bn(pre, c):
        s[pre + ".weight"] = (c,); s[pre + ".bias"] = (c,)
        s[pre + ".running_mean"] = (c,); s[pre + ".running_var"] = (c,)
        s[pre + ".num_batches_tracked"] = ()

    s["encoder.conv_in.weight"] = (ch[0], cfg.in_channels, 7); s["encoder.conv_in.bias"] = (ch[0],)
    s["encoder.activation.weight"] = (1,)
    L = W
    for i in range(len(ch) - 1):
        b = f"encoder.res_layers.{i}."
        ci, co = ch[i], ch[i + 1]
        s[b + "residual.0.weight"] = (co // 2, ci, k); s[b + "residual.0.bias"] = (co // 2,)
        bn(b + "residual.1", co // 2); s[b + "residual.2.weight"] = (1,)
        s[b + "residual.3.weight"] = (co, co // 2, k); s[b + "residual.3.bias"] = (co,)
        s[b + "skip.weight"] = (co, ci, k); s[b + "skip.bias"] = (co,)
        bn(b + "add.0", co); s[b + "add.1.weight"] = (1,)
        L = (L + 2 * (k // 2) - (k - 1) - 1) // 2 + 1
    flat = L * ch[-1]
    s["encoder.fc_mu.weight"] = (z, flat); s["encoder.fc_mu.bias"] = (z,)
    sig = z if cfg.is_diag else z * (z + 1) // 2  # model/residual.py:216
    s["encoder.fc_sigma.0.weight"] = (sig, flat); s["encoder.fc_sigma.0.bias"] = (sig,)
    cd = sum(cfg.feat_dim(c) for c in cfg.conditional)
    s["decoder.fc_in.weight"] = (flat, z + cd); s["decoder.fc_in.bias"] = (flat,)
    Ld = L
    for i in range(len(ch) - 1):
        b = f"decoder.res_layers.{i}."
        ci, co = ch[-1 - i], ch[-2 - i]
        s[b + "residual.0.weight"] = (ci, ci // 2, k); s[b + "residual.0.bias"] = (ci // 2,)
        bn(b + "residual.1", ci // 2); s[b + "residual.2.weight"] = (1,)
        s[b + "residual.3.weight"] = (ci // 2, co, k); s[b + "residual.3.bias"] = (co,)
        s[b + "skip.1.weight"] = (co, ci, k + 1); s[b + "skip.1.bias"] = (co,)
        bn(b + "add.0", co); s[b + "add.1.weight"] = (1,)
        Ld = 2 * Ld - 1
    kf = W - Ld + 7
    s["decoder.conv_out.weight"] = (ch[0], cfg.in_channels, kf); s["decoder.conv_out.bias"] = (cfg.in_channels,)
    for key in cfg.grad_reversal:
        pre = f"disentangle.grad_reversal.{key}.reversal.1."
        d = cfg.feat_dim(key)
        dims = {"mlp1": [z, z, z, d], "mlp2": [z, z, d], "mlp3": [z, z, z // 2, d],
                "mlp4": [z, 2 * z, 2 * z, d]}
        for name, dd in dims.items():
            for n in range(len(dd) - 1):
                s[f"{pre}{name}.{2 * n}.weight"] = (dd[n + 1], dd[n])
                s[f"{pre}{name}.{2 * n}.bias"] = (dd[n + 1],)
    return s


def synth_state_dict(cfg: Cfg, seed: int = 1) -> Dict[str, torch.Tensor]:
    """Seeded fan-in-uniform weights with the reference's state_dict keys/shapes."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(cfg).items():
        if k == "arena_size":
            sd[k] = torch.tensor(cfg.arena, dtype=torch.float32)
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(shp)
        elif shp == (1,):
            sd[k] = torch.full((1,), 0.25)  # PReLU init
        elif ".residual.1." in k or ".add.0." in k:  # BN affine: perturbed so tests see them
            base = 1.0 if k.endswith("weight") else 0.0
            sd[k] = base + 0.1 * (torch.rand(shp, generator=g) * 2 - 1)
        else:
            if len(shp) == 1:
                bound = 0.05
            else:
                bound = 1.0 / math.sqrt(int(np.prod(shp[1:])))
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    return sd


# --------------------------------------------------------------------------------------------------------------------
# moving_avg_lsq scrubber (MovingAvgLeastSquares, reference model/disentangle.py:393-538), polynomial order 1
# --------------------------------------------------------------------------------------------------------------------
def mals_init(z: int, ny: int, bias: bool = False, lamdiff: float = 1e-1) -> Dict[str, torch.Tensor]:
    """buffers as registered by the reference constructor (:417-433)"""
    nx = z + int(bias)
    return {"Sxx0": torch.eye(nx), "Sxy0": torch.zeros(nx, ny), "Sxx1": torch.eye(nx), "Sxy1": torch.zeros(nx, ny),
            "lam0": torch.tensor([0.9]), "lam1": torch.tensor([0.9]) + lamdiff}


def _mals_x(mu, bias):
    return torch.column_stack((mu, torch.ones(mu.shape[0], 1))) if bias else mu


def mals_forward(st, mu, bias=False, l2_reg=0.0):
    """forward (:466-487): decoder weights from the normal equations, predictions of both decoders"""
    x = _mals_x(mu, bias)
    l2 = torch.ones(x.shape[1]) * l2_reg
    if bias:
        l2[-1] = 0
    W0 = torch.linalg.solve(st["Sxx0"].diagonal_scatter(st["Sxx0"].diagonal() + l2), st["Sxy0"])
    W1 = torch.linalg.solve(st["Sxx1"].diagonal_scatter(st["Sxx1"].diagonal() + l2), st["Sxy1"])
    return x @ W0, x @ W1, W0, W1


def mals_evaluate(st, yhat0, yhat1, y, delta=1e-4, lamdiff=1e-1):
    """evaluate_loss (:505-538): mean of the two summed squared errors; the forgetting factors move towards the better one"""
    l0, l1 = ((y - yhat0) ** 2).sum(), ((y - yhat1) ** 2).sum()
    if l0 < l1:
        st["lam0"] = torch.clamp(st["lam0"] - delta, 0.0, 1.0)
        st["lam1"] = st["lam0"] + lamdiff
    else:
        st["lam1"] = torch.clamp(st["lam1"] + delta, 0.0, 1.0)
        st["lam0"] = st["lam1"] - lamdiff
    return (l0 + l1) * 0.5


def mals_update(st, mu, y, bias=False):
    """update (:489-503): exponentially weighted running covariances"""
    x = _mals_x(mu, bias)
    xx, xy = x.T @ x, x.T @ y
    st["Sxx0"] = st["lam0"] * st["Sxx0"] + xx
    st["Sxy0"] = st["lam0"] * st["Sxy0"] + xy
    st["Sxx1"] = st["lam1"] * st["Sxx1"] + xx
    st["Sxy1"] = st["lam1"] * st["Sxy1"] + xy
    return st


# --------------------------------------------------------------------------------------------------------------------
# qda scrubber (QuadraticDiscriminantFilter, reference model/disentangle.py:90-232)
# --------------------------------------------------------------------------------------------------------------------
def qda_init(z: int, n_classes: int, lamdiff: float = 1e-2) -> Dict[str, torch.Tensor]:
    """buffers as registered by the reference constructor (:104-125)"""
    st = {}
    for name in ("0a", "1a", "0b", "1b"):
        st["m" + name] = torch.zeros(n_classes, z)
        st["S" + name] = torch.eye(z)[None, :].repeat(n_classes, 1, 1)
    st["lama"] = torch.ones(n_classes) * 0.2
    st["lamb"] = st["lama"] + lamdiff
    return st


def _cgll(x, m, S):
    """Gaussian log likelihood up to a constant (:130-135)"""
    r = x - m
    return -0.5 * (torch.logdet(S) + torch.sum(r * torch.linalg.solve(S, r.T).T, axis=1))


def qda_evaluate(st, x, y, classes, delta=1e-3, lamdiff=1e-2, update=True):
    """evaluate_loss (:173-232): average log-likelihood ratio of the two classifiers; forgetting factors drift"""
    loss = 0
    for i, label in enumerate(classes):
        i0, i1 = (y != label).ravel(), (y == label).ravel()
        lla0, lla1 = _cgll(x, st["m0a"][i:i + 1], st["S0a"][i]), _cgll(x, st["m1a"][i:i + 1], st["S1a"][i])
        llb0, llb1 = _cgll(x, st["m0b"][i:i + 1], st["S0b"][i]), _cgll(x, st["m1b"][i:i + 1], st["S1b"][i])
        lla, llb = torch.sum(i0 * lla0 + i1 * lla1), torch.sum(i0 * llb0 + i1 * llb1)
        if update and lla > llb:
            st["lama"][i] = torch.clamp(st["lama"][i] - delta, 0.0, 1.0)
            st["lamb"][i] = st["lama"][i] + lamdiff
        elif update:
            st["lamb"][i] = torch.clamp(st["lamb"][i] + delta, 0.0, 1.0)
            st["lama"][i] = st["lamb"][i] - lamdiff
        s = (i1 * 2 - 1).float()
        loss = loss + (s @ (lla1 - lla0) + s @ (llb1 - llb0)) * 0.5
    return loss / len(classes)


def qda_update(st, x, y, classes):
    """update (:137-171): class-conditional batch mean / covariance (correction 0) blended with lama (A) and lamb (B)"""
    for i, label in enumerate(classes):
        for side, sel in ((0, (y != label).ravel()), (1, (y == label).ravel())):
            mean, cov = torch.mean(x[sel], axis=0), torch.cov(x[sel].T, correction=0)
            for lam, ab in ((st["lama"][i], "a"), (st["lamb"][i], "b")):
                st[f"m{side}{ab}"][i] = (1 - lam) * st[f"m{side}{ab}"][i] + lam * mean
                st[f"S{side}{ab}"][i] = (1 - lam) * st[f"S{side}{ab}"][i] + lam * cov
    return st


# --------------------------------------------------------------------------------------------------------------------
# moving_avg scrubber (MovingAverageFilter, reference model/disentangle.py:9-88)
# --------------------------------------------------------------------------------------------------------------------
def ma_init(z: int, n_classes: int, lamdiff: float = 1e-2) -> Dict[str, torch.Tensor]:
    lam1 = torch.ones(n_classes) * 0.5
    return {"m1": torch.zeros(n_classes, z), "m2": torch.zeros(n_classes, z), "lam1": lam1, "lam2": lam1 + lamdiff}


def ma_evaluate(st, x, y, classes, delta=1e-3, lamdiff=1e-2):
    """evaluate_loss (:32-74): forgetting factors drift towards the running mean closer to the batch mean; the loss is the
    norm of the upper-triangular pairwise differences between the classes' mean estimates"""
    m1, m2 = torch.zeros_like(st["m1"]), torch.zeros_like(st["m2"])
    for i, label in enumerate(classes):
        xbar = torch.mean(x[(y == label).ravel(), :], axis=0)
        if torch.linalg.norm(xbar - st["m1"][i]) < torch.linalg.norm(xbar - st["m2"][i]):
            st["lam1"][i] = torch.clamp(st["lam1"][i] - delta, 0.0, 1.0)
            st["lam2"][i] = st["lam1"][i] + lamdiff
        else:
            st["lam2"][i] = torch.clamp(st["lam2"][i] + delta, 0.0, 1.0)
            st["lam1"][i] = st["lam2"][i] - lamdiff
        m1[i] = (1 - st["lam1"][i]) * xbar + st["lam1"][i] * st["m1"][i]
        m2[i] = (1 - st["lam2"][i]) * xbar + st["lam2"][i] * st["m2"][i]
    est = 0.5 * (m1 + m2)
    return torch.linalg.norm(torch.triu(est.T[..., None] - est.T[..., None, :], diagonal=1))


def ma_update(st, x, y, classes):
    """update (:76-88)"""
    for i, label in enumerate(classes):
        xbar = torch.mean(x[(y == label).ravel(), :], axis=0)
        st["m1"][i] = (1 - st["lam1"][i]) * xbar + st["lam1"][i] * st["m1"][i]
        st["m2"][i] = (1 - st["lam2"][i]) * xbar + st["lam2"][i] * st["m2"][i]
    return st
