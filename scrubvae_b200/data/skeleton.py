"""Skeleton definition: what the reference reads from `configs/mouse_skeleton.yaml` through
`neuroposelib.read.config` (reference params/read.py:9, data/dataset.py:348, get/data.py:23) —
`KINEMATIC_TREE` (chains of joint indices, :86-92) and the integer unit `OFFSET` per joint (:95-112).

`read_skeleton(path)` loads any file of that format; the module constants are the 18-keypoint mouse of the
reference's own config, so that synthetic runs (bench.py, smoke) need no file.  The tree and offsets are DATA handed
to the kernels (scv_recon_loss / scv_preprocess_windows take them as device arrays), nothing is baked in."""
from __future__ import annotations

from typing import Dict

KINEMATIC_TREE = [[0, 1, 2, 3, 4], [0, 5], [1, 6, 7, 8], [1, 9, 10, 11], [5, 12, 13, 14], [5, 15, 16, 17]]
OFFSET = [[0, 0, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0], [-1, 0, 0],
          [0, 1, 0], [0, 1, 0], [0, 1, 0], [0, -1, 0], [0, -1, 0], [0, -1, 0],
          [0, 1, 0], [0, 1, 0], [0, 1, 0], [0, -1, 0], [0, -1, 0], [0, -1, 0]]
N_KEYPTS = 18
# arena bounds used by the synthetic benchmark inputs: [[min xyz], [max xyz]]
ARENA = [[-100.0, -100.0, 0.0], [100.0, 100.0, 50.0]]


def read_skeleton(path: str) -> Dict:
    """YAML skeleton config -> dict with at least KINEMATIC_TREE and OFFSET (the keys the hot path consumes)."""
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f)
    for key in ("KINEMATIC_TREE", "OFFSET"):
        if key not in cfg:
            raise KeyError(f"{path}: skeleton config lacks {key}")
    n = len(cfg["OFFSET"])
    for chain in cfg["KINEMATIC_TREE"]:
        if not all(0 <= int(j) < n for j in chain):
            raise ValueError(f"{path}: KINEMATIC_TREE joint index out of range for {n} keypoints")
    return cfg
