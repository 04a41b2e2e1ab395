"""Pose-window preprocessing: host windowing logic on CPU; the CUDA kernels against the
reference-generated golden fixture (tests/golden/preprocess.npz, written by make_golden.py from the unmodified
reference) and against the CPU oracle on a larger seeded input.

Tolerances: window indices, kept-window selection, ids and the integer-truncated offsets are BIT-EXACT; float32
products (x6d, root, target_pose, heading, avg_speed_3d) within 2e-5 absolute of O(1)-O(100) values: the
reference mixes float64 numpy and float32 torch ops whose libm (atan2, sin, cos) and summation order differ from
the device's in the last ulp."""
import os

import numpy as np
import pytest
import torch

import scrubvae_b200 as sv
from scrubvae_b200.data.dataset import window_starts
from oracle import scvae_oracle as orc


def _g(golden_dir):
    z = np.load(os.path.join(golden_dir, "preprocess.npz"))
    return {k: z[k] for k in z.files}


def test_window_starts_match_reference_golden(golden_dir):
    g = _g(golden_dir)
    W = g["window_inds"].shape[1]
    assert np.array_equal(window_starts(g["in.ids"], 2, W), g["window_inds"][:, 0])
    assert np.array_equal(window_starts(g["in.ids2"], 1, W), g["window_inds2_s1"][:, 0])
    assert np.array_equal(window_starts(g["in.ids2"], 3, W), g["window_inds2_s3"][:, 0])
    for ids, stride in ((g["in.ids"], 2), (g["in.ids2"], 1), (g["in.ids2"], 3)):  # and against the oracle
        assert np.array_equal(window_starts(ids, stride, W), orc.window_indices(ids, stride, W)[:, 0])


def test_window_starts_edge_cases():
    with pytest.raises(RuntimeError):  # no run long enough: the reference's torch.cat([]) error
        window_starts(np.array([0] * 50 + [1] * 50), 2, 51)
    with pytest.raises(RuntimeError):
        window_starts(np.array([], dtype=np.int64), 2, 51)
    assert np.array_equal(window_starts(np.zeros(51, dtype=np.int64), 2, 51), [0])  # exactly one window
    assert np.array_equal(window_starts(np.array([3] * 55 + [4] * 51), 2, 51), [0, 2, 4, 55])  # ragged runs


def _synth_frames(n, seed):
    """A random-walk skeleton: plausible bone directions, several animals, a few fast (outlier) stretches."""
    rng = np.random.default_rng(seed)
    off = np.asarray(orc.OFFSET, dtype=np.float64)
    parents = [0] * 18
    for chain in orc.KINEMATIC_TREE:
        for j in range(1, len(chain)):
            parents[chain[j]] = chain[j - 1]
    pose = np.zeros((n, 18, 3))
    root = np.cumsum(rng.normal(0, 0.4, (n, 3)), 0)
    jump = rng.random(n) < 0.01
    root += np.cumsum(jump[:, None] * rng.normal(0, 30, (n, 3)), 0)
    pose[:, 0] = root
    dirs = off[None] + np.cumsum(rng.normal(0, 0.02, (n, 18, 3)), 0) + rng.normal(0, 0.3, (1, 18, 3))
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True) + 1e-9
    lens = rng.uniform(5.3, 29.7, 18)
    for chain in orc.KINEMATIC_TREE:
        for j in chain[1:]:
            pose[:, j] = pose[:, parents[j]] + dirs[:, j] * lens[j]
    ids = np.repeat(np.arange(6), [n // 6] * 5 + [n - 5 * (n // 6)])
    return pose, ids


def _check(out, ref, kept_ref=None):
    assert np.array_equal(out["window_inds"].cpu().numpy(), np.asarray(ref["window_inds"]))
    for k in ("x6d", "root", "offsets", "target_pose", "heading", "avg_speed_3d"):
        a, b = out[k].cpu().numpy(), np.asarray(ref[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if k == "offsets":
            assert np.array_equal(a, b), k  # integer-truncated: bit-exact
        else:
            assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(b).max()), (k, np.abs(a - b).max())
    assert np.array_equal(out["ids"].cpu().numpy(), np.asarray(ref["ids"]))


@pytest.mark.gpu
def test_preprocess_matches_reference_golden(golden_dir):
    g = _g(golden_dir)
    out = sv.data.preprocess_windows(g["in.pose"], g["in.ids"], orc.KINEMATIC_TREE, orc.OFFSET, window=51, stride=2,
                                     speed_threshold=2.25, direction_process="midfwd")
    ref = {k[4:]: v for k, v in g.items() if k.startswith("out.")}
    ref["window_inds"] = g["window_inds"]
    _check(out, ref)
    # the reference's own known answers: mid-frame root at xy = 0, bone lengths = |offsets|
    mid = out["root"][:, 25, :2].abs().max().item()
    assert mid < 1e-4
    tp, offs = out["target_pose"], out["offsets"]
    for chain in orc.KINEMATIC_TREE:
        for a, b in zip(chain[:-1], chain[1:]):
            bl = (tp[:, :, b] - tp[:, :, a]).norm(dim=-1)
            assert (bl - offs[:, :, b].norm(dim=-1)).abs().max().item() < 1e-3


def _modes(golden_dir):
    z = np.load(os.path.join(golden_dir, "preprocess_modes.npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("mode", ["x360", None])
def test_oracle_matches_reference_golden_other_modes(golden_dir, mode):
    """direction_process "x360" (root centred on the mid frame, no rotation) and None (raw), reference dataset.py:383-413:
    the oracle against fixtures generated from the unmodified reference (make_golden.py)."""
    g, gm = _g(golden_dir), _modes(golden_dir)
    ref = orc.preprocess(g["in.pose"], g["in.ids"], window=51, stride=2, speed_threshold=2.25, direction_process=mode)
    for k in ("x6d", "root", "target_pose"):
        want = torch.from_numpy(gm[f"{mode}.{k}"])
        assert (torch.as_tensor(ref[k]).float() - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item()), (mode, k)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["x360", None])
def test_preprocess_matches_reference_golden_other_modes(golden_dir, mode):
    g, gm = _g(golden_dir), _modes(golden_dir)
    out = sv.data.preprocess_windows(g["in.pose"], g["in.ids"], orc.KINEMATIC_TREE, orc.OFFSET, window=51, stride=2,
                                     speed_threshold=2.25, direction_process=mode)
    for k in ("x6d", "root", "target_pose"):
        want = torch.from_numpy(gm[f"{mode}.{k}"])
        got = out[k].float().cpu()
        assert got.shape == want.shape, (mode, k)
        assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item()), (mode, k)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,thr,stride", [("midfwd", 2.25, 2), ("x360", None, 5), (None, 2.25, 1)])
def test_preprocess_matches_oracle_large(mode, thr, stride):
    pose, ids = _synth_frames(6000, seed=7)
    ref = orc.preprocess(pose, ids, window=51, stride=stride, speed_threshold=thr, direction_process=mode)
    out = sv.data.preprocess_windows(pose, ids, orc.KINEMATIC_TREE, orc.OFFSET, window=51, stride=stride,
                                     speed_threshold=thr, direction_process=mode)
    assert 0 < out["x6d"].shape[0] <= out["window_inds"].shape[0]
    _check(out, {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in ref.items()})


@pytest.mark.gpu
def test_device_windows_loader_feeds_the_step():
    pose, ids = _synth_frames(1500, seed=3)
    data = sv.data.preprocess_windows(pose, ids % 4, orc.KINEMATIC_TREE, orc.OFFSET, speed_threshold=None)
    loader = sv.data.DevicePoseWindows(data, batch_size=32, shuffle=True, drop_last=True, seed=1)
    assert len(loader) >= 2
    seen = 0
    for batch in loader:
        assert batch["x6d"].shape == (32, 51, 18, 6) and batch["ids"].shape == (32, 1)
        assert batch["x6d"].is_cuda
        seen += 1
    assert seen == len(loader)


def test_skeleton_constants_match_the_reference_config():
    """data/skeleton.py carries the mouse skeleton of the reference's configs/mouse_skeleton.yaml; the staged copy
    (oracle/_ref/configs, build container: /root/reference/configs) must parse to the same tree and offsets."""
    import os
    from scrubvae_b200.data import skeleton
    from oracle import scvae_oracle as orc
    assert skeleton.KINEMATIC_TREE == orc.KINEMATIC_TREE and skeleton.OFFSET == orc.OFFSET
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in (os.path.join(root, "oracle", "_ref", "configs", "mouse_skeleton.yaml"),
                 "/root/reference/configs/mouse_skeleton.yaml"):
        if os.path.exists(path):
            cfg = skeleton.read_skeleton(path)
            assert [list(c) for c in cfg["KINEMATIC_TREE"]] == skeleton.KINEMATIC_TREE
            assert [list(o) for o in cfg["OFFSET"]] == skeleton.OFFSET
            break
