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

import weakref

import torch
import torch.distributed as dist


def broadcast_parameters(model, src: int = 0):
    """Identical replicas: parameters and buffers of rank `src` to every rank (call after construction and
    after the per-epoch GRScrubber.reset_parameters, reference train/trainer.py:368-370)."""
    eng = getattr(model, "_engine", None)
    with torch.no_grad():
        if eng is not None:
            eng.ensure_flat()
            dist.broadcast(eng.flat, src)
            eng.resident_valid = False
        else:
            for p in model.parameters():
                dist.broadcast(p.data, src)
        for b in model.buffers():
            dist.broadcast(b, src)


class GradAllReduce:
    """The `comm` hook of Plan.backward / TrainStep.  Buckets follow the order in which backward finishes weight
    gradients, each launched on the comm stream the moment its last wgrad kernel is enqueued (the comm stream waits for
    the weight-gradient stream), so only the last small bucket is exposed:
       "decoder_done"  -> gpacked[gp_split:]  (decoder + scrubber heads; overlaps the whole encoder backward)
       "range" lo, hi  -> gpacked[lo:hi]      (enc.fc right after its wgrad, then encoder blocks 3, 2 as they finish)
       "encoder_done"  -> what is left of gpacked (conv_in, blocks 1 and 0) + gflat[:n_direct] (BatchNorm / PReLU
                          gradients); the launching stream then waits for every bucket
       "post_backward" -> un-overlapped fallback for callers that ran the whole backward first: all-reduce gflat."""

    def __init__(self, engine, world: int, group=None):
        self.world, self.group = world, group
        self.cuda = engine.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=engine.device) if self.cuda else None
        self.bytes_per_step = 4 * (engine.gpacked.numel() + engine.n_direct)
        self._lo = engine.gp_split  # gpacked[_lo:gp_split] of the encoder part is already reduced this step
        self.steps = weakref.WeakSet()  # every TrainStep that captured this hook's all-reduces (shutdown releases them)

    def _reduce(self, t):
        if t.numel():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _run(self, eng, phase, lo, hi):
        if phase == "decoder_done":
            self._lo = eng.gp_split
            self._reduce(eng.gpacked[eng.gp_split:])
        elif phase == "range":
            assert hi == self._lo and lo < hi, "encoder buckets must be contiguous, last layer first"
            self._reduce(eng.gpacked[lo:hi])
            self._lo = lo
        elif phase == "encoder_done":
            self._reduce(eng.gpacked[:self._lo])
            if eng.bn_sync_world() == 1:  # dp_bn "sync": the BatchNorm / PReLU gradients are already global sums
                self._reduce(eng.gflat[:eng.n_direct])
            self._lo = eng.gp_split

    def __call__(self, eng, phase: str, wait=(), lo=None, hi=None):
        """`wait`: extra streams whose work the reduced buffers depend on (the weight-gradient stream)."""
        if self.world == 1:
            return
        if phase == "post_backward":
            if eng.bn_sync_world() == 1:
                self._reduce(eng.gflat)
            else:
                self._reduce(eng.gflat[eng.n_direct:])
            return
        if not self.cuda:
            self._run(eng, phase, lo, hi)
            return
        main = torch.cuda.current_stream()
        self.stream.wait_stream(main)
        for st in wait:
            self.stream.wait_stream(st)
        with torch.cuda.stream(self.stream):
            self._run(eng, phase, lo, hi)
        if phase == "encoder_done":
            main.wait_stream(self.stream)


def setup(model, optimizer=None, group=None, bn_sync=False):
    """Public entry of data parallelism (call once per process after torch.distributed is initialised and the model is
    on its device): identical replicas (rank 0's parameters and buffers), the bucketed gradient all-reduce hooked into
    the engine's backward (both the fused TrainStep path and `total.backward()`), 1/world folded into the optimizer's
    gradient scale.  Returns the GradAllReduce hook (None when the world is one process).

    bn_sync=True selects the dp_bn "sync" mode of SURVEY.md §8e (caveat 1): every BatchNorm's batch statistics (and the
    matching backward sums) are all-reduced, so the replicas together compute exactly what ONE device would on the
    global batch — parity with the single-device reference at world x B windows — at the price of 33 small latency-bound
    all-reduces per step.  Default is per-replica statistics (standard DDP semantics)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return None
    broadcast_parameters(model)
    eng = model.engine
    comm = GradAllReduce(eng, world, group)
    eng.comm = comm
    eng._bn_sync = (world, group) if bn_sync else None
    if optimizer is not None:
        optimizer.grad_scale = 1.0 / world
    return comm


def resync(model):
    """After anything that re-draws parameters on every rank from its own RNG (the per-epoch GRScrubber.reset_parameters
    of train(), reference train/trainer.py:368-370): rank 0's values everywhere again."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        broadcast_parameters(model)


def shutdown(model=None):
    """Release the CUDA graphs that hold captured NCCL kernels BEFORE the process group goes away: destroying the
    communicator under a live graph is what stalled interpreter teardown in round 1."""
    import gc
    if model is not None:
        eng = getattr(model, "_engine", None)
        if eng is not None:
            for _, st in list(eng.__dict__.get("_train_steps", {}).values()):
                st.graph = None
            eng.__dict__.pop("_train_steps", None)
            comm = getattr(eng, "comm", None)
            for st in list(getattr(comm, "steps", ())):  # steps built outside the trainer's cache
                st.graph = None
            eng.comm = None
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
