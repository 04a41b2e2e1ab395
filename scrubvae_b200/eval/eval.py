"""Generative restrictiveness (reference eval/eval.py:22-120): decode the latent with a RESAMPLED conditional
variable, run forward kinematics on the decoded window and re-extract that variable from the pose — the R^2 between the
resampled value and the re-extracted one says how strictly the decoder obeys its conditioning.

Same signature, same random draws (torch's generator on z's device, same call order) and the same in-place update of
`data[key]` as the reference; decode runs on the engine (eval-mode BatchNorm, z given), FK + feature extraction in one
kernel (scv_gen_features) instead of ~700 small ATen launches."""
from __future__ import annotations

import numpy as np
import torch

# part definition and normalisation constants of reference eval/eval.py:77-81 and :105-115
SPEED_PARTS = [[0, 1, 2, 3, 4, 5], [1, 6, 7, 8, 9, 10, 11], [5, 12, 13, 14, 15, 16, 17]]
SPEED_MEAN = [0.4993, 0.7112, 0.6663]
SPEED_STD = [0.4038, 0.3586, 0.4169]
SPEED_MIN = [-1.2323, -1.9734, -1.5858]
SPEED_MAX = [4.6167, 4.6437, 4.2551]


def _flat(tree):
    out = [len(tree)]
    for chain in tree:
        out += [len(chain)] + [int(j) for j in chain]
    return out


def generative_restrictiveness(model, z, data, key, kinematic_tree):
    """Returns (pred, target): the variable re-extracted from the decoded pose and the resampled conditional."""
    n_keypts = data["x6d"].shape[-2]
    window = data["x6d"].shape[1]
    batch_size = data["x6d"].shape[0]
    var_true = data[key]
    dev = z.device
    if key == "heading":
        rand_yaw = (torch.rand(batch_size, dtype=torch.float32, device=dev) * 2 - 1)[:, None] * np.pi
        rand_angle2D = torch.cat([torch.sin(rand_yaw), torch.cos(rand_yaw)], axis=-1)
        data["heading"] = rand_angle2D.reshape(rand_yaw.shape[:-1] + (-1,))
    elif key == "avg_speed_3d":
        spd_std = torch.tensor(SPEED_STD, dtype=torch.float32, device=dev)
        rand_jitter = torch.randn((batch_size, 1), dtype=torch.float32, device=dev) * spd_std * 1.5 + 0.5
        mins = torch.tensor(SPEED_MIN, dtype=torch.float32, device=dev)
        maxes = torch.tensor(SPEED_MAX, dtype=torch.float32, device=dev)
        data["avg_speed_3d"] = torch.clamp(var_true + rand_jitter, min=mins, max=maxes)
    else:
        raise KeyError(key)  # the reference defines the measure for these two variables only (pred would be unbound)

    data_o = model.decode(z, data)
    plan = data_o["_plan"]
    eng = plan.eng
    cache = eng.__dict__.setdefault("_gen_cache", {})
    ck = (id(kinematic_tree), str(dev))
    if ck not in cache:
        cache[ck] = (torch.tensor(_flat(kinematic_tree), dtype=torch.int32, device=dev),
                     torch.tensor(_flat(SPEED_PARTS), dtype=torch.int32, device=dev),
                     torch.tensor(SPEED_MEAN + SPEED_STD, dtype=torch.float32, device=dev))
    tree, parts, norm = cache[ck]
    offsets = data["offsets"].to(torch.float32).contiguous()
    pred = torch.empty(batch_size, 2 if key == "heading" else 3, dtype=torch.float32, device=dev)
    eng.ops.gen_features(plan.xh, eng.C0, plan.root_hat, offsets, tree, tree.numel(), parts, batch_size, window, n_keypts,
                         norm=norm if key == "avg_speed_3d" else None,
                         heading=pred if key == "heading" else None, avg3=pred if key == "avg_speed_3d" else None)
    return pred, data[key]


def r2_score(y_true, y_pred):
    """sklearn.metrics.r2_score(y_true, y_pred) with its defaults (per-output R^2, uniform average), as the reference's
    test_epoch calls it (train/trainer.py:296-300)."""
    yt = torch.as_tensor(y_true, dtype=torch.float64).reshape(len(y_true), -1)
    yp = torch.as_tensor(y_pred, dtype=torch.float64).reshape(len(y_pred), -1)
    num = ((yt - yp) ** 2).sum(0)
    den = ((yt - yt.mean(0, keepdim=True)) ** 2).sum(0)
    r2 = torch.where(den > 0, 1 - num / den.clamp_min(1e-300), torch.where(num > 0, torch.zeros_like(num), torch.ones_like(num)))
    return float(r2.mean())
