"""Data parallelism for the SC-VAE step: one process per GPU, full model replica, the batch sharded over
ranks, ONE exchange step — the gradient all-reduce (the reference is single-device, this is new).

What is reduced.  The engine keeps weight gradients in `gpacked` (the K-major matrices the wgrad kernels
accumulate into, laid out [encoder | decoder + scrubber heads]) and the BatchNorm / PReLU gradients in
`gflat[:n_direct]`.  Backward finishes the decoder and the heads first, so their segment of `gpacked` is
all-reduced on a side stream while the encoder backward still runs; the encoder segment and the small
direct region follow, the main stream joins, and the final gather builds `gflat` from reduced values.
The SUM is left in place: the fused optimizer folds 1/world into its gradient scale (`grad_scale`), and the
global-norm clip is computed after the reduction, as `clip_grad_norm_` would on one device.

Semantics.  Every loss is a sum over windows divided by the LOCAL batch, so the mean of the replica gradients is
the gradient of the mean of the per-shard losses.  BatchNorm statistics are per replica (torch DDP semantics);
parity is claimed against "the oracle run on each shard, gradients averaged" (tests/test_parallel_cpu.py).
The nested GR normalisation constant uses the local B (SURVEY.md §8e caveat 2), as one reference process would.

The collective is NCCL over NVLink 5 / NVSwitch through torch.distributed (`gloo` on CPU for the tests);
it is capturable in the step's CUDA graph."""
from __future__ import annotations

import torch
import torch.distributed as dist


def broadcast_parameters(model, src: int = 0):
    """Identical replicas: parameters and buffers of rank `src` to every rank (call after construction and
    after the per-epoch GRScrubber.reset_parameters, reference train/trainer.py:368-370)."""
    eng = getattr(model, "_engine", None)
    with torch.no_grad():
        if eng is not None:
            dist.broadcast(eng.flat, src)
            eng.packed_dirty = True
        else:
            for p in model.parameters():
                dist.broadcast(p.data, src)
        for b in model.buffers():
            dist.broadcast(b, src)


class GradAllReduce:
    """The `comm` hook of Plan.backward / TrainStep.  phases:
       "decoder_done"  -> all-reduce gpacked[gp_split:] on the comm stream (overlaps the encoder backward)
       "encoder_done"  -> all-reduce gpacked[:gp_split] and gflat[:n_direct]; the launching stream waits for both
       "post_backward" -> un-overlapped fallback for callers that ran the whole backward first: all-reduce gflat."""

    def __init__(self, engine, world: int, group=None):
        self.world, self.group = world, group
        self.cuda = engine.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=engine.device) if self.cuda else None
        self.bytes_per_step = 4 * (engine.gpacked.numel() + engine.n_direct)

    def _reduce(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def __call__(self, eng, phase: str, wait=()):
        """`wait`: extra streams whose work the reduced buffers depend on (the weight-gradient stream)."""
        if self.world == 1:
            return
        if phase == "post_backward":
            self._reduce(eng.gflat)
            return
        if not self.cuda:
            if phase == "decoder_done":
                self._reduce(eng.gpacked[eng.gp_split:])
            elif phase == "encoder_done":
                self._reduce(eng.gpacked[:eng.gp_split])
                self._reduce(eng.gflat[:eng.n_direct])
            return
        main = torch.cuda.current_stream()
        self.stream.wait_stream(main)
        for st in wait:
            self.stream.wait_stream(st)
        with torch.cuda.stream(self.stream):
            if phase == "decoder_done":
                self._reduce(eng.gpacked[eng.gp_split:])
            elif phase == "encoder_done":
                self._reduce(eng.gpacked[:eng.gp_split])
                self._reduce(eng.gflat[:eng.n_direct])
        if phase == "encoder_done":
            main.wait_stream(self.stream)
