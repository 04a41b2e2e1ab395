"""Runs the UNMODIFIED reference step (oracle/refimport.py) — baseline / checker infrastructure only.

  * `time_train_epoch` times the reference's own `train_test_epoch(mode="train")` (reference train/trainer.py:101-212,
    with `predict_batch` :92-99 and `get_batch_loss` train/losses.py:182-324) on the host cores or on a CUDA device
    (PyTorch-eager: cuDNN / cuBLAS / ATen), with the precision switches `train()` sets (trainer.py:323-325:
    float32 matmul precision "medium", cudnn.benchmark) and anomaly detection OFF.
  * `step` runs one forward / loss / backward with injected reparameterisation noise and returns outputs, losses and
    gradients — what the GPU parity tests compare the CUDA path with at the benchmarked batch size.
Nothing here is imported by the scrubvae_b200 package.
"""
import contextlib
import io
import os
import time

import torch

from . import refimport
from . import scvae_oracle as orc

DEFAULT_CH = (64, 128, 256, 512, 1024)


def build_model(device="cpu", ch=DEFAULT_CH, z_dim=64, cond=("heading",), gr=("heading",), discrete_classes=None,
                window=51, seed=1, diag=False, state_dict=None):
    """reference get.model(...) (get/model.py:4-151) with the synthetic-benchmark defaults of SURVEY.md §8(d)."""
    sv = refimport.import_reference()
    mc = dict(type="rcnn", channel=list(ch), kernel=5, z_dim=z_dim, window=window, activation="prelu", diag=diag,
              init_dilation=None, prior="gaussian", load_model=None, start_epoch=None)
    dc = dict(method={"conditional": list(cond), "grad_reversal": list(gr)},
              features=sorted(set(cond) | set(gr)), alpha=1.0)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = sv.get.model(mc, None, None, dc, 18, "midfwd", arena_size=torch.tensor(orc.ARENA),
                         kinematic_tree=orc.KINEMATIC_TREE, discrete_classes=discrete_classes or {}, device="cpu",
                         verbose=0)
    if state_dict is not None:
        m.load_state_dict(state_dict)
    return m.to(device), dc


@contextlib.contextmanager
def precision(mode):
    """"fp32": exact fp32 everywhere; "tf32": what reference train() sets (trainer.py:323-325) + torch's conv default."""
    old = (torch.get_float32_matmul_precision(), torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
           torch.backends.cudnn.benchmark)
    try:
        if mode == "fp32":
            torch.set_float32_matmul_precision("highest")
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
        else:
            torch.set_float32_matmul_precision("medium")
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cudnn.benchmark = True
        yield
    finally:
        torch.set_float32_matmul_precision(old[0])
        torch.backends.cudnn.allow_tf32 = old[1]
        torch.backends.cuda.matmul.allow_tf32 = old[2]
        torch.backends.cudnn.benchmark = old[3]


def step(m, dc, data, eps, loss_scale):
    """One reference forward + losses + backward with the reparameterisation noise `eps` injected.
    Returns (data_o, losses, grads) — tensors on the model's device."""
    from scrubvae.train import trainer
    from scrubvae.train.losses import get_batch_loss
    m.train()
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: eps.to(t)
    try:
        data_o = trainer.predict_batch(m, data, m.disentangle_keys)
    finally:
        torch.randn_like = orig
    losses = get_batch_loss(m, data, data_o, loss_scale, dc)
    for p in m.parameters():
        p.grad = None
    losses["total"].backward()
    grads = {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
             for n, p in m.named_parameters()}
    return data_o, losses, grads


def synth_batch(B, seed=0, keys=("x6d", "root", "offsets", "target_pose", "heading")):
    chunks = [orc.synth_batch(min(256, B - i), seed=seed + i) for i in range(0, B, 256)]
    return {k: torch.cat([c[k] for c in chunks], 0).contiguous() for k in keys}


def time_train_epoch(device, B, steps, warmup, loss_scale, threads=None, host_batches=False, cond=("heading",),
                     gr=("heading",), discrete_classes=None, ch=DEFAULT_CH, z_dim=64, window=51, data=None):
    """Times `steps` iterations of the reference's train_test_epoch(mode="train") after `warmup` untimed ones.
    CPU: wall clock.  CUDA: CUDA events around the epoch call (its closing .item() synchronises).
    host_batches: the loader yields pinned host tensors (the reference's `.to(device)` then copies every step)."""
    sv = refimport.import_reference()
    from scrubvae.train import trainer
    dev = torch.device(device)
    if dev.type == "cpu":
        torch.set_num_threads(threads or os.cpu_count() or 1)
    m, dc = build_model(dev, ch=ch, z_dim=z_dim, cond=cond, gr=gr, discrete_classes=discrete_classes, window=window)
    keys = ("x6d", "root", "offsets", "target_pose") + tuple(sorted(set(cond) | set(gr)))
    if data is None:
        data = synth_batch(B, seed=0, keys=keys)
    if dev.type == "cuda":
        data = {k: (v.pin_memory() if host_batches else v.to(dev)) for k, v in data.items()}
    config = {"loss": dict(loss_scale), "disentangle": dc}
    with contextlib.redirect_stdout(io.StringIO()):
        opt, sched = trainer.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})

    def epoch(n):
        with contextlib.redirect_stdout(io.StringIO()):
            return trainer.train_test_epoch(config, m, [data] * n, dev, 1, optimizer=opt, scheduler=sched, mode="train")

    with (precision("tf32") if dev.type == "cuda" else contextlib.nullcontext()):
        if warmup:
            epoch(warmup)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            metrics = epoch(steps)
            e1.record()
            torch.cuda.synchronize(dev)
            dt = e0.elapsed_time(e1) * 1e-3
        else:
            t0 = time.perf_counter()
            metrics = epoch(steps)
            dt = time.perf_counter() - t0
    return {"windows_per_s": B * steps / dt, "ms_per_step": dt / steps * 1e3, "seconds": dt, "batch": B,
            "steps": steps, "threads": torch.get_num_threads() if dev.type == "cpu" else None,
            "loss_total": metrics.get("total"), "torch": torch.__version__}
