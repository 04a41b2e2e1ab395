"""Pose-window preprocessing with the reference's function names and meaning
(reference data/dataset.py: get_window_indices :198-233, preprocess_save_data :313-454), computed by the
libscv.so preprocessing kernels on frames resident in HBM.  Disk I/O (read.pose_h5, the *.h5 caches) is the
caller's business and out of scope: these functions take the arrays the reference reads.

There is no CPU path: `pose` is moved to the CUDA device and every array product comes from a kernel.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

# reference data/dataset.py:362-374: the three body parts of get_speed_parts for the mouse skeleton
MOUSE_SPEED_PARTS = [[0, 1, 2, 3, 4, 5], [1, 6, 7, 8, 9, 10, 11], [5, 12, 13, 14, 15, 16, 17]]
MODES = {None: 0, "midfwd": 1, "x360": 2}


def window_starts(ids: np.ndarray, stride: int, window: int) -> np.ndarray:
    """Start frame of every window: frames are split where the animal id changes and each run of at least
    `window` frames yields starts run_begin, run_begin + stride, ... (reference :198-233; int64, host metadata)."""
    ids = np.asarray(ids).reshape(-1)
    n = len(ids)
    if n == 0:
        raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
    change = np.concatenate([[0], np.flatnonzero(ids[1:] != ids[:-1]) + 1, [n]])
    out = [np.arange(a, b - window + 1, stride, dtype=np.int64) for a, b in zip(change[:-1], change[1:]) if b - a >= window]
    if not out:
        raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")  # what the reference raises (:231)
    return np.concatenate(out)


def _flat_tree(tree: Sequence[Sequence[int]]) -> torch.Tensor:
    t = [len(tree)]
    for chain in tree:
        t += [len(chain)] + [int(j) for j in chain]
    return torch.tensor(t, dtype=torch.int32)


def get_window_indices(ids, stride: int, window: int, device="cuda", ops=None) -> torch.Tensor:
    """(N_w, window) int64 index matrix on `device`, bit-identical to the reference's get_window_indices."""
    if ops is None:
        from .._ops import get_ops
        ops = get_ops()
    starts = torch.from_numpy(window_starts(ids, stride, window)).to(device)
    winds = torch.empty(starts.numel(), window, dtype=torch.int64, device=device)
    ops.window_indices(starts, starts.numel(), window, winds)
    return winds


def preprocess_windows(pose, ids, kinematic_tree, offset, window: int = 51, stride: int = 2,
                       speed_threshold: Optional[float] = 2.25, direction_process: Optional[str] = "midfwd",
                       speed_parts=MOUSE_SPEED_PARTS, device="cuda", ops=None) -> Dict[str, torch.Tensor]:
    """The in-memory part of preprocess_save_data (reference :349-449) for data_keys
    x6d, root, offsets, target_pose, heading, avg_speed_3d, ids (+ window_inds of ALL windows, kept = the
    indices that survived the speed-outlier filter).  pose (N, J, 3) float64, ids (N,) int.  Tensors on `device`."""
    if direction_process not in MODES:
        raise ValueError(f"direction_process {direction_process!r} is not one of {list(MODES)}")
    if ops is None:
        from .._ops import get_ops
        ops = get_ops()
    dev = torch.device(device)
    pose_d = torch.as_tensor(pose, dtype=torch.float64).to(dev).contiguous()
    ids_np = np.asarray(ids).reshape(-1)
    N, J = pose_d.shape[0], pose_d.shape[1]
    starts_np = window_starts(ids_np, stride, window)
    n_w = len(starts_np)
    starts = torch.from_numpy(starts_np).to(dev)
    winds = torch.empty(n_w, window, dtype=torch.int64, device=dev)
    ops.window_indices(starts, n_w, window, winds)
    parts = _flat_tree(speed_parts).to(dev)
    speed = torch.empty(n_w, dtype=torch.float64, device=dev)
    avg3 = torch.empty(n_w, 3, dtype=torch.float32, device=dev)
    heading = torch.empty(n_w, 2, dtype=torch.float32, device=dev)
    yaw = torch.empty(n_w, dtype=torch.float64, device=dev)
    ops.window_features(pose_d, starts, n_w, window, J, parts, speed, avg3, heading, yaw)
    if speed_threshold is not None:
        keep = torch.nonzero(~(speed > speed_threshold)).reshape(-1)  # reference: np.delete(where(speed > thr))
    else:
        keep = torch.arange(n_w, device=dev)
    nk = keep.numel()
    f32 = dict(dtype=torch.float32, device=dev)
    out = {
        "x6d": torch.empty(nk, window, J, 6, **f32), "root": torch.empty(nk, window, 3, **f32),
        "offsets": torch.empty(nk, window, J, 3, **f32), "target_pose": torch.empty(nk, window, J, 3, **f32),
    }
    tree = _flat_tree(kinematic_tree).to(dev)
    off = torch.as_tensor(np.asarray(offset), dtype=torch.int32).to(dev).contiguous()
    ops.preprocess_windows(pose_d, starts, keep, nk, window, J, tree, off, yaw, MODES[direction_process],
                           out["x6d"], out["root"], out["offsets"], out["target_pose"])
    out["heading"] = heading[keep]
    out["avg_speed_3d"] = avg3[keep]
    mid = torch.from_numpy(ids_np[starts_np + window // 2].astype(np.int16)).to(dev)
    out["ids"] = mid[keep]
    out["window_inds"] = winds
    out["kept"] = keep
    return out


class DevicePoseWindows:
    """The preprocessed windows resident in HBM, iterated as shuffled batches of dicts with the keys the step
    consumes (the role of MouseDataset + DataLoader, reference data/dataset.py:456-505, get/data.py:138-144,
    without the host round trip).  `ids` is served as (B, 1) like the reference's collated batch."""

    def __init__(self, data: Dict[str, torch.Tensor], batch_size: int, shuffle: bool = True, drop_last: bool = False,
                 seed: int = 0):
        self.data = {k: v for k, v in data.items() if k not in ("window_inds", "kept")}
        if "ids" in self.data and self.data["ids"].dim() == 1:
            self.data["ids"] = self.data["ids"][:, None]
        self.n = next(iter(self.data.values())).shape[0]
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(seed)

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        dev = next(iter(self.data.values())).device
        order = torch.randperm(self.n, generator=self.gen) if self.shuffle else torch.arange(self.n)
        order = order.to(dev)
        for i in range(len(self)):
            idx = order[i * self.batch_size:(i + 1) * self.batch_size]
            yield {k: v.index_select(0, idx) for k, v in self.data.items()}
