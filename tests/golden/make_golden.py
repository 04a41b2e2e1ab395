"""Generate golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The reference has no tests/golden vectors of its own
(SURVEY.md §4), so these fixtures — outputs of the reference itself on seeded inputs —
are what pins the oracle (oracle/scvae_oracle.py) and, through it, the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import _refimport  # noqa: E402
from oracle import scvae_oracle as orc  # noqa: E402  (only for the synthetic input generator)


def build_ref_model(sv, ch, z_dim, feats_cond, feats_gr, discrete_classes=None, window=51, seed=1, diag=False):
    mc = dict(type="rcnn", channel=list(ch), kernel=5, z_dim=z_dim, window=window,
              activation="prelu", diag=diag, init_dilation=None, prior="gaussian",
              load_model=None, start_epoch=None)
    dc = dict(method={"conditional": list(feats_cond), "grad_reversal": list(feats_gr)},
              features=sorted(set(feats_cond) | set(feats_gr)), alpha=1.0)
    torch.manual_seed(seed)
    m = sv.get.model(mc, None, None, dc, 18, "midfwd", arena_size=torch.tensor(orc.ARENA),
                     kinematic_tree=orc.KINEMATIC_TREE, discrete_classes=discrete_classes or {},
                     device="cpu", verbose=0)
    # perturb BN affine / PReLU so that parity sees them
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if ".residual.1." in n or ".add.0." in n:
                p.add_(0.1 * (torch.rand(p.shape, generator=g) * 2 - 1))
    return m, dc


def run_step(sv, m, dc, data, eps, loss_scale, lr=1e-4):
    from scrubvae.train import trainer
    from scrubvae.train.losses import get_batch_loss
    m.train()
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: eps.to(t)  # inject the reparameterisation noise
    try:
        data_o = trainer.predict_batch(m, data, m.disentangle_keys)
    finally:
        torch.randn_like = orig
    losses = get_batch_loss(m, data, data_o, loss_scale, dc)
    for p in m.parameters():
        p.grad = None
    losses["total"].backward()
    grads = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p))
             for n, p in m.named_parameters()}
    opt = torch.optim.AdamW(m.parameters(), lr=lr)
    torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=1e6)
    opt.step()
    return data_o, losses, grads


def golden_step(sv, tag, ch, z_dim, cond, gr, B, discrete_classes=None, full=True, diag=False):
    m, dc = build_ref_model(sv, ch, z_dim, cond, gr, discrete_classes, diag=diag)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    data = orc.synth_batch(B, seed=0)
    eps = orc.synth_eps(B, z_dim, seed=2)
    loss_scale = {"prior": 1e-4, "jpe": 1.0, "root": 1.0}
    loss_scale.update({k + "_gr": 1.0 for k in gr})
    data_o, losses, grads = run_step(sv, m, dc, data, eps, loss_scale)
    sd1 = m.state_dict()
    out = {"meta_ch": np.array(ch), "meta_z": np.array(z_dim), "meta_B": np.array(B)}
    for k, v in losses.items():
        out["loss." + k] = v.detach().numpy()
    if full:
        for k, v in sd0.items():
            out["sd0." + k] = v.numpy()
        for k, v in grads.items():
            out["grad." + k] = v.numpy()
        for k, v in sd1.items():  # post-step state: full for BN buffers, digest for the rest
            if "running_" in k or "num_batches" in k:
                out["sd1." + k] = v.numpy()
            else:
                out["sd1sum." + k] = np.array([v.double().sum().item(), v.double().norm().item()])
        for k in ("mu", "L", "z", "root", "x6d"):
            out["out." + k] = data_o[k].detach().numpy()
        for k in gr:
            for i, e in enumerate(data_o["disentangle"]["grad_reversal"][k]):
                out[f"out.gr.{k}.{i}"] = e.detach().numpy()
    else:  # digest only: per-tensor (sum, l2) of grads
        for k, v in grads.items():
            out["gradsum." + k] = np.array([v.double().sum().item(), v.double().norm().item()])
    path = os.path.join(HERE, f"step_{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", {k: float(v) for k, v in losses.items()})


def golden_preprocess(sv):
    """Random-walk mouse-like poses -> reference preprocess chain (dataset.py:313-454)."""
    from scrubvae.data import dataset as ds
    import yaml
    skel = yaml.safe_load(open(os.path.join(_refimport.REF_ROOT, "configs/mouse_skeleton.yaml")))
    rng = np.random.default_rng(7)
    N = 400
    ids = np.concatenate([np.zeros(150, int), np.ones(30, int), np.full(120, 2), np.full(100, 3)])
    # smooth skeleton: FK of slowly varying random rotations with fixed bone lengths
    q = rng.normal(size=(N, 18, 4)); q = np.cumsum(0.004 * q, axis=0) + rng.normal(size=(1, 18, 4))
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x6 = orc.quat_to_cont6d(torch.tensor(q, dtype=torch.float32))
    lens = torch.tensor(rng.uniform(5, 30, size=(1, 18, 1)), dtype=torch.float32)
    offs = torch.tensor(orc.OFFSET, dtype=torch.float32)[None] * lens
    root = torch.tensor(np.cumsum(rng.normal(scale=0.15, size=(N, 3)), axis=0), dtype=torch.float32)
    pose = orc.fwd_kin(x6, offs.expand(N, 18, 3), root).double().numpy()
    pose[230:] += np.array([150.0, -40.0, 0.0])  # a tracking jump: windows over it are speed outliers

    # re-implement preprocess_save_data's body call with our in-memory pose (read.pose_h5 is I/O)
    import neuroposelib
    neuroposelib.read.pose_h5 = lambda path: (pose.copy(), ids.copy())
    keys = ["x6d", "root", "offsets", "target_pose", "heading", "avg_speed_3d", "ids"]
    out = ds.preprocess_save_data("", skel, "4_mice", 51, 2, keys, 2.25, "midfwd")
    winds = ds.get_window_indices(ids, 2, 51).numpy()
    save = {"in.pose": pose, "in.ids": ids, "window_inds": winds}
    for k, v in out.items():
        save["out." + k] = v.numpy()
    # a second window-index case with ragged runs (some shorter than the window)
    ids2 = np.concatenate([np.zeros(51, int), np.ones(50, int), np.full(53, 2), np.full(7, 3), np.full(60, 1)])
    save["in.ids2"] = ids2
    save["window_inds2_s1"] = ds.get_window_indices(ids2, 1, 51).numpy()
    save["window_inds2_s3"] = ds.get_window_indices(ids2, 3, 51).numpy()
    path = os.path.join(HERE, "preprocess.npz")
    np.savez_compressed(path, **save)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", out["x6d"].shape)
    # the other two direction_process modes (dataset.py:383-413) on the same frames: "x360" (centred, not rotated) and None
    modes = {}
    for mode in ("x360", None):
        o = ds.preprocess_save_data("", skel, "4_mice", 51, 2, keys, 2.25, mode)
        for k in ("x6d", "root", "target_pose"):
            modes[f"{mode}.{k}"] = o[k].numpy().astype(np.float32)
    path = os.path.join(HERE, "preprocess_modes.npz")
    np.savez_compressed(path, **modes)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    sv = _refimport.import_reference()
    golden_step(sv, "small_heading", [8, 16, 32, 64, 128], 8, ["heading"], ["heading"], B=6)
    golden_step(sv, "small_3head", [8, 16, 32, 64, 128], 8, ["heading", "avg_speed_3d", "ids"],
                ["heading", "avg_speed_3d", "ids"], B=5, discrete_classes={"ids": [0, 1, 2, 3]})
    golden_step(sv, "small_diag", [8, 16, 32, 64, 128], 8, ["heading"], ["heading"], B=6, diag=True)
    golden_step(sv, "default_heading_digest", [64, 128, 256, 512, 1024], 64, ["heading"], ["heading"],
                B=4, full=False)
    golden_preprocess(sv)
