#!/usr/bin/env python
"""Benchmark of the SC-VAE training step (BASELINE.json metric: pose windows/sec per training step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|5|5w201] [--batch B]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
  value        whole-job windows/s, inputs resident in HBM, one CUDA-graph replay per step
  e2e          the same steps through the public API scrubvae_b200.train.train_test_epoch(mode="train") over a
               loader of HOST (pinned) batches: every step's H2D copy (prefetched on a side stream) and the D2H
               read of its loss are inside the timed region
  sustained    the same step replayed back to back for >= 3 s (clocks sampled): the number a long run sees
  roofline     tensor-pipe roofline of the dominant kernel class (the tcgen05 overlapping-row GEMMs): their
               launches of one step replayed back to back from one CUDA graph between CUDA events; peak = the
               driver-measured bf16 burst figure / 2 (TF32), the sustained one is reported beside it
  hbm_kernels  every HBM-bound kernel of the step: CUDA-event time per launch (single stream), algorithmic bytes,
               fraction of the measured copy bandwidth
  gpu_eager_baseline  the UNMODIFIED reference's train_test_epoch on the same GPU (PyTorch eager: cuDNN / cuBLAS /
               ATen; staged in oracle/_ref by oracle/make_ref.sh), same batch, TF32 as reference train() sets it
  cpu_baseline the reference's train_test_epoch on the host cores (a bounded sample)
`--impl reference` times the reference's own CPU path alone (all host threads)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_CH = [64, 128, 256, 512, 1024]
FLOP_PER_WINDOW = 759.0e6  # layer GEMMs only, fwd + bwd, default architecture

# BASELINE.json configs: 2 = the headline (default arch, one scrubber head, TF32); 3 = three scrubber heads, bf16
# operands; 5 = scaled architecture (wider channels, longer window, larger latent)
CONFIGS = {
    "2": dict(ch=DEFAULT_CH, z=64, window=51, feats=["heading"], precision="tf32", batch=2048,
              workload="SC-VAE mouse_skeleton default arch (ch 64-1024, k5, window 51, z64), conditional+grad_reversal on heading"),
    "3": dict(ch=DEFAULT_CH, z=64, window=51, feats=["heading", "avg_speed_3d", "ids"], precision="bf16", batch=2048,
              workload="SC-VAE default arch, conditional+grad_reversal on heading, avg_speed_3d, ids (3 scrubber heads), bf16 operands"),
    "5": dict(ch=[128, 256, 512, 1024, 2048], z=128, window=101, feats=["heading"], precision="tf32", batch=1024,
              workload="scaled SC-VAE (ch 128-2048, k5, window 101, z128), conditional+grad_reversal on heading"),
    "5w201": dict(ch=[128, 256, 512, 1024, 2048], z=128, window=201, feats=["heading"], precision="tf32", batch=512,
                  workload="scaled SC-VAE (ch 128-2048, k5, window 201, z128), conditional+grad_reversal on heading"),
}


def loss_scale_for(feats):
    d = {"prior": 1e-4, "jpe": 1.0, "root": 1.0}
    d.update({f + "_gr": 1.0 for f in feats})
    return d


LOSS_SCALE = loss_scale_for(["heading"])
WORKLOAD = CONFIGS["2"]["workload"]


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every 5 ms from a thread while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, idx):
        self.idx, self.sm, self.mask, self.mx, self.err = idx, [], 0, None, None
        self._stop = threading.Event()
        self._thr = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as ex:
            self.err = repr(ex)[:120]
            return self

        def loop():
            while not self._stop.is_set():
                try:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception as ex:
                    self.err = repr(ex)[:120]
                    return
                time.sleep(0.005)
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["nvml unavailable: %s" % self.err], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.mx, "sm_min_mhz": sm[0],
                "reasons": sorted(n for b, n in self.REASONS.items() if self.mask & b), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- reference arms
def reference_keys(feats):
    return ("x6d", "root", "offsets", "target_pose") + tuple(feats)


def cpu_reference(cfg, B, steps, warmup, threads=None):
    """The reference's own train_test_epoch on the host cores (oracle/_ref); falls back to the oracle port when the
    reference is not staged.  Returns (windows/s, cores, ms/step, kind)."""
    from oracle import refimport
    cores = threads or os.cpu_count() or 1
    feats = cfg["feats"]
    if refimport.available():
        from oracle import ref_runner as rr
        dc = {"ids": [0, 1, 2, 3]} if "ids" in feats else None
        r = rr.time_train_epoch("cpu", B, steps, warmup, loss_scale_for(feats), threads=cores, cond=feats, gr=feats,
                                discrete_classes=dc, ch=cfg["ch"], z_dim=cfg["z"], window=cfg["window"])
        return r["windows_per_s"], cores, r["ms_per_step"], "reference"
    import torch
    from oracle import scvae_oracle as orc
    torch.set_num_threads(cores)
    ocfg = orc.Cfg(ch=tuple(cfg["ch"]), z_dim=cfg["z"], window=cfg["window"], conditional=tuple(feats),
                   grad_reversal=tuple(feats), discrete_classes={"ids": [0, 1, 2, 3]} if "ids" in feats else None)
    sd = orc.synth_state_dict(ocfg, seed=1)
    data = orc.synth_batch(B, window=cfg["window"], seed=0)
    eps = orc.synth_eps(B, ocfg.z_dim, seed=2)
    opt, ts = None, []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, sd, opt, _ = orc.train_step(sd, data, ocfg, loss_scale_for(feats), eps, opt_state=opt, lr=1e-4,
                                          optimizer="adamw", step=i + 1)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    return B / dt, cores, dt * 1e3, "port"


def run_reference(args, cfg, rank):
    """`--impl reference`: the reference's CPU path, all host threads, same workload; each step a bounded sample of the
    batch so that the whole run ends within a few minutes."""
    if rank != 0:
        return
    B = args.batch or cfg["batch"]
    # probe the host speed with one small step, then size the per-step sample: the full batch if the run then fits
    # ~150 s, else the largest power-of-two fraction that does (CPU throughput is flat in the batch size)
    v0, cores, _, kind = cpu_reference(cfg, 256, 1, 1)
    budget = 150.0
    sample = B
    while sample > 64 and (args.steps + args.warmup) * sample / max(v0, 1.0) > budget:
        sample //= 2
    v, cores, ms, kind = cpu_reference(cfg, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "pose windows/sec per training step", "value": v, "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "batch_per_gpu": B, "global_batch": B, "optimizer": "adamw",
                   "sample": f"{sample} of the {B} windows per step (host probe: {v0:.0f} windows/s)"},
        "cpu_baseline": {"value": v, "unit": "windows/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps x {sample} windows after {args.warmup} warm-up: the reference's "
                                   "train_test_epoch(mode='train') on the host cores"},
        "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- this repo's arm
def build_model(device, precision, cfg=None):
    import torch
    import scrubvae_b200 as sv
    from scrubvae_b200.data.skeleton import ARENA, KINEMATIC_TREE
    cfg = cfg or CONFIGS["2"]
    feats = cfg["feats"]
    mc = dict(type="rcnn", channel=list(cfg["ch"]), kernel=5, z_dim=cfg["z"], window=cfg["window"], activation="prelu",
              diag=False, init_dilation=None, prior="gaussian", load_model=None, start_epoch=None, precision=precision)
    dcfg = dict(method={"conditional": list(feats), "grad_reversal": list(feats)}, features=list(feats), alpha=1.0)
    torch.manual_seed(1)
    m = sv.get.model(mc, None, None, dcfg, 18, "midfwd", arena_size=torch.tensor(ARENA), kinematic_tree=KINEMATIC_TREE,
                     discrete_classes={"ids": [0, 1, 2, 3]} if "ids" in feats else {}, device=device, verbose=0)
    return m, dcfg


def synth_host_batch(B, seed, cfg=None):
    """Synthetic pose windows of the reference's shapes, pinned host memory."""
    import torch
    from oracle import scvae_oracle as orc  # synthetic-input generator only; outside every timed region
    cfg = cfg or CONFIGS["2"]
    chunks = []
    for i in range(0, B, 256):  # FK of the target pose is generated in chunks to bound host memory
        chunks.append(orc.synth_batch(min(256, B - i), window=cfg["window"], seed=seed + i))
    d = {k: torch.cat([c[k] for c in chunks], 0) for k in chunks[0]}
    keep = reference_keys(cfg["feats"])
    return {k: d[k].contiguous().pin_memory() for k in keep}


def gemm_profile(step, B, reps=5):
    """The dominant kernel class in isolation: every tensor-core GEMM launch of one step (scv_gemm / scv_wgrad with
    precision != fp32, same buffers, same order) is recorded from an eager step, captured into ONE CUDA graph and
    replayed; CUDA events around the replays give the class's device time per step without launch gaps."""
    import torch
    eng = step.eng
    ops = eng.ops
    calls = []
    orig_gemm, orig_wgrad = ops.gemm, ops.wgrad
    nnz_by_w = {}
    pk = eng.packed16 if getattr(eng, "packed16", None) is not None else eng.packed
    for g in eng.W.values():
        nnz_by_w[pk.data_ptr() + pk.element_size() * g.w] = g.nnz
        if g.wd is not None:
            nnz_by_w[pk.data_ptr() + pk.element_size() * (eng._n_fwd + g.wd)] = g.nnz_d
        nnz_by_w[("g", eng.gpacked.data_ptr() + 4 * g.w)] = g.nnz
    from scrubvae_b200._ops import _ptr

    def gemm(**kw):
        if kw.get("precision", 0):
            calls.append((orig_gemm, kw, 2.0 * kw["B"] * kw["Lo"] * nnz_by_w.get(_ptr(kw["W"]), kw["N"] * kw["K"])))
        orig_gemm(**kw)

    def wgrad(**kw):
        if kw.get("precision", 0):
            calls.append((orig_wgrad, kw, 2.0 * kw["B"] * kw["Lo"] * nnz_by_w.get(("g", _ptr(kw["dW"])), kw["N"] * kw["K"])))
        orig_wgrad(**kw)

    ops.gemm, ops.wgrad = gemm, wgrad
    saved_comm, step.comm = step.comm, None  # rank 0 alone runs this instrumented step: no collective inside
    try:
        step._sequence()
        torch.cuda.synchronize()
    finally:
        del ops.gemm, ops.wgrad
        step.comm = saved_comm
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for fn, kw, _ in calls:
                fn(**kw)
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            gr.replay()
        e1.record(s)
    torch.cuda.synchronize()
    return {"launches": len(calls), "seconds": e0.elapsed_time(e1) * 1e-3 / reps, "flops": sum(c[2] for c in calls)}


def hbm_kernel_table(step, hbm_gbs, reps=3):
    """CUDA-event time of every HBM-bound launch of one eager step on ONE stream (no concurrent kernels), with the
    algorithmic bytes each launch must move (what it reads once + what it writes once; index arrays and re-reads are
    implementation overhead and not counted).  Aggregated per kernel: us (min over reps of the per-step sum),
    achieved GB/s, fraction of the measured copy bandwidth."""
    import torch
    eng, ops, plan = step.eng, step.eng.ops, step.plan
    esz = 4

    def rows_c(kw):
        return kw["B"] * kw["L"] * kw["Cc"] * esz

    def b_bnact_fwd(a, kw):
        return rows_c(kw) * (1 + (1 if kw.get("H") is not None else 0) + (2 if kw.get("U") is not None else 0))

    def b_bwd_reduce(a, kw):
        return rows_c(kw) * (1 + (1 if kw.get("dO") is not None else 0) + (2 if kw.get("dU") is not None else 0))

    def b_bwd_apply(a, kw):
        return b_bwd_reduce(a, kw) + (rows_c(kw) if kw.get("dX") is not None else 0)

    m = step.model
    z, W, B = m.z_dim, m.window, plan.B
    nsig = z * (z + 1) // 2
    J, C0, nx = plan.J, eng.C0, plan.nx
    est = {
        "pack_input": lambda a, kw: B * W * (nx + 3 + C0) * esz,
        "bnact_fwd": b_bnact_fwd, "bnact_bwd_reduce": b_bwd_reduce, "bnact_bwd_apply": b_bwd_apply,
        "reparam_fwd": lambda a, kw: B * (z + nsig + z + eng.cond_dim + z + z * z + eng.zc_ld) * esz,
        "reparam_bwd": lambda a, kw: B * (nsig + z + 3 * z + eng.zc_ld + z * z + z + nsig) * esz,
        "kl": lambda a, kw: B * (z + z * z) * esz * (2 if (len(a) > 5 and a[5] is not None) else 1),
        "recon_loss": lambda a, kw: B * W * (C0 + J * 3 * 2 + 3 + C0) * esz,
        "out_bwd": lambda a, kw: B * W * C0 * esz * 3,
        "unpack_root": lambda a, kw: B * W * 6 * esz,
        "sumsq": lambda a, kw: a[1] * esz,
        "sumsq_packed": lambda a, kw: a[2] * esz,
        "gather": lambda a, kw: a[3] * esz * 2,  # weight repack / gradient unpack: one read + one write per element
        # p, m, v read + write; gradient read; resident mode also writes the operand copy
        "optim_step": lambda a, kw: a[4] * esz * (8 if (kw.get("pack_idx") is not None or kw.get("pack_mask") is not None) else 7),
    }
    recs = []
    saved = {}

    def wrap(name):
        fn = getattr(ops, name)
        saved[name] = fn

        def w(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*a, **kw)
            e1.record()
            recs.append((name, float(est[name](a, kw)), e0, e1))
        setattr(ops, name, w)

    old_env = {k: os.environ.get(k) for k in ("SCV_SKIP_STREAM", "SCV_WGRAD_STREAM")}
    os.environ["SCV_SKIP_STREAM"] = "0"
    os.environ["SCV_WGRAD_STREAM"] = "0"
    saved_comm, step.comm = step.comm, None
    best = {}
    try:
        for n in est:
            wrap(n)
        for _ in range(reps):
            recs.clear()
            step._sequence()
            torch.cuda.synchronize()
            agg = {}
            for name, nbytes, e0, e1 in recs:
                a = agg.setdefault(name, [0, 0.0, 0.0])
                a[0] += 1
                a[1] += e0.elapsed_time(e1) * 1e3
                a[2] += nbytes
            for name, a in agg.items():
                if name not in best or a[1] < best[name][1]:
                    best[name] = a
    finally:
        for n in saved:
            delattr(ops, n)
        step.comm = saved_comm
        for k, v in old_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    out = []
    for name, (cnt, us, nbytes) in sorted(best.items(), key=lambda kv: -kv[1][1]):
        gbs = nbytes / us / 1e3 if us > 0 else 0.0
        out.append({"kernel": name, "launches": cnt, "us": round(us, 1), "alg_bytes": int(nbytes),
                    "gbs": round(gbs, 1), "hbm_frac": round(gbs / hbm_gbs, 3)})
    return out


def gpu_eager_reference(cfg, B, steps=5, warmup=3):
    """The unmodified reference step on cuda:0 (PyTorch eager), device-resident batch and host (pinned) batch."""
    from oracle import refimport
    if not refimport.available():
        return {"value": None, "unavailable": "reference not staged (oracle/make_ref.sh)"}
    import torch
    from oracle import ref_runner as rr
    feats = cfg["feats"]
    dc = {"ids": [0, 1, 2, 3]} if "ids" in feats else None
    kw = dict(cond=feats, gr=feats, discrete_classes=dc, ch=cfg["ch"], z_dim=cfg["z"], window=cfg["window"])
    r = rr.time_train_epoch("cuda", B, steps, warmup, loss_scale_for(feats), **kw)
    torch.cuda.empty_cache()
    rh = rr.time_train_epoch("cuda", B, steps, warmup, loss_scale_for(feats), host_batches=True, **kw)
    torch.cuda.empty_cache()
    return {"value": r["windows_per_s"], "unit": "windows/s", "ms_per_step": r["ms_per_step"],
            "e2e_value": rh["windows_per_s"], "e2e_ms_per_step": rh["ms_per_step"], "batch": B, "steps": steps,
            "warmup": warmup, "torch": r["torch"],
            "what": "unmodified reference train_test_epoch(mode='train') on cuda:0, PyTorch eager (cuDNN/cuBLAS/ATen), "
                    "float32 matmul precision 'medium' + cudnn TF32 + cudnn.benchmark as reference train() sets them, "
                    "anomaly detection off; value = batch resident in HBM, e2e_value = pinned host batch copied per step"}


def run_ours(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import scrubvae_b200 as sv
    from scrubvae_b200.engine import TrainStep
    from scrubvae_b200 import parallel
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or cfg["batch"]
    precision = args.precision or cfg["precision"]
    loss_scale = loss_scale_for(cfg["feats"])
    m, dcfg = build_model(dev, precision, cfg)
    m.train()
    opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
    comm = parallel.setup(m, opt) if world > 1 else None
    if world > 1:
        torch.cuda.manual_seed(1234 + rank)  # rank-distinct reparameterisation noise
    host = synth_host_batch(B, seed=1000 * rank, cfg=cfg)
    data = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    step = TrainStep(m, opt, loss_scale, B, use_graph=not args.no_graph, comm=comm, resident=not args.no_resident)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        e0.record()
        for i in range(n):
            step.run()
            marks[i].record()  # per-step marks (diagnostic: median / max step time), same stream, no host sync
        e1.record()
        return e0, e1, marks

    # ---- device-resident timing (value)
    step.run(data)  # eager warm-up (+ graph capture on the next call)
    for _ in range(max(args.warmup, 3)):
        step.run()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the barrier: NVML start-up on one rank must not delay its peers inside the timed region
    barrier()
    e0, e1, marks = timed(args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    per_step = sorted(a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks))
    dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = dt.item()
    loss_total = float(step.plan.loss_out[-1])
    launches = step.n_launch * args.steps

    # ---- sustained: the same step back to back for >= 3 s
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(3.2 / max(dt / args.steps, 1e-4)))
        s2 = ClockSampler(local_rank)
        if rank == 0:
            s2.start()
        barrier()
        f0, f1, _ = timed(n_sus)
        barrier()
        c2 = s2.stop() if rank == 0 else None
        ds = torch.tensor([f0.elapsed_time(f1) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(ds, op=dist.ReduceOp.MAX)
        ds = ds.item()
        sustained = {"value": world * B * n_sus / ds, "unit": "windows/s", "seconds": ds, "steps": n_sus,
                     "ms_per_step": ds / n_sus * 1e3, "clocks": c2}

    # ---- end to end through the public API, host inputs (e2e): sv.train.train_test_epoch over a loader of pinned
    # host batches — every step copies its batch host->device (prefetched on a side stream while the previous
    # step computes) and reads the step's total loss back to the host (4 bytes, asynchronous)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    config = {"loss": dict(loss_scale), "disentangle": dcfg, "train": {}}
    loss_host = torch.zeros(1, pin_memory=True)

    def read_loss(i, vec):
        loss_host.copy_(vec[-1:], non_blocking=True)

    import contextlib
    import io

    def api_epoch(n):
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints per-loss epoch averages
            sv.train.train_test_epoch(config, m, [host] * n, dev, 1, optimizer=opt, scheduler=None, mode="train",
                                      step_callback=read_loss)

    e2e = None
    try:
        api_epoch(3)
        barrier()
        e0.record()
        api_epoch(args.steps)
        e1.record()
        barrier()
        dte = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(dte, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * args.steps / dte.item(), "unit": "windows/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": dte.item() / args.steps * 1e3,
               "api": "scrubvae_b200.train.train_test_epoch(mode='train') over pinned host batches"}
    except Exception as ex:  # keep the main line
        e2e = {"value": None, "unit": "windows/s", "error": repr(ex)[:200]}

    if rank != 0:
        if world > 1:
            _shutdown(dist, parallel, m)
        return
    pk, src = peaks()
    is_bf16 = precision == "bf16"
    burst = pk["bf16_tflops"] / (1.0 if is_bf16 else 2.0)
    sust_pk = pk["bf16_tflops_sustained"] / (1.0 if is_bf16 else 2.0)
    pk_src = f"{src}: bf16_tflops" + ("" if is_bf16 else " / 2 (TF32)") + " — burst figure: the class is timed in isolation for tens of ms"
    eng = m.engine
    # algorithmic GEMM FLOPs per window: SURVEY.md §8(d)'s 759.0 MFLOP for the default architecture; for the other
    # configs the launches' own count, 2 x rows x nnz(W) over forward, data-gradient and weight-gradient GEMMs
    flop_pw = FLOP_PER_WINDOW if (cfg["ch"] == DEFAULT_CH and cfg["window"] == 51 and cfg["z"] == 64) else None
    line = {
        "metric": "pose windows/sec per training step", "value": world * B * args.steps / dt, "unit": "windows/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3,
        "ms_per_step_median": per_step[len(per_step) // 2], "ms_per_step_max": per_step[-1],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32": "tf32 (fp32 storage, fp32 accumulate)", "fp32": "fp32",
                  "bf16": "bf16 operands (fp32 accumulate, fp32 master weights)"}[precision],
        "data": "synthetic",
        "config": {"workload": cfg["workload"], "batch_per_gpu": B, "global_batch": world * B, "optimizer": "adamw",
                   "parallelism": f"dp{world}", "l2": "per-step working set (>2 GB of activations) exceeds the 126 MB L2",
                   "cuda_graph": not args.no_graph, "loss_total_last": loss_total, "bench_config": args.config},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "sustained": sustained,
    }
    try:
        gp = gemm_profile(step, B)
        alg = (flop_pw * B) if flop_pw else gp["flops"]
        ach = alg / gp["seconds"] / 1e12
        traffic = None
        try:  # DRAM bytes of the same launches from the committed ncu pass (profiles/, cold cache): informational
            with open(os.path.join(ROOT, "profiles", "gemm_family_dram.json")) as f:
                traffic = json.load(f).get(args.config, {}).get("dram_bytes_per_step")
        except Exception:
            pass
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": burst, "unit": "TFLOP/s",
                            "frac": ach / burst, "frac_of_sustained_peak": ach / sust_pk, "peak_sustained": sust_pk,
                            "traffic": traffic,
                            "traffic_note": "STATIC: dram__bytes_read+write summed over the family's launches of one step "
                                            "from the committed ncu pass (profiles/gemm_family_dram.json), not re-measured "
                                            "in this run; ncu counts a launch's DRAM writes only while it runs, so dirty "
                                            "L2 lines written back later are under-counted",
                            "kernel": "tcgen05 overlapping-row GEMM family (gemm_tc[_bnr]_kernel + wgrad_tc_kernel): all "
                                      "its launches of one step replayed back to back from one CUDA graph, CUDA events",
                            "launches_per_step": gp["launches"], "gemm_seconds_per_step": gp["seconds"],
                            "gemm_share_of_step": gp["seconds"] / (dt / args.steps),
                            "alg_flops_per_step": alg, "launch_counted_flops_per_step": gp["flops"], "peak_source": pk_src}
        step_flops = alg / B
        line["step_tensor_frac"] = {"flop_per_window": step_flops,
                                    "achieved_tflops": step_flops * B * args.steps / dt / 1e12, "peak_tflops": burst,
                                    "frac": step_flops * B * args.steps / dt / 1e12 / burst, "peak_source": pk_src}
    except Exception as ex:
        line["roofline"] = {"bound": "tensor", "achieved": None, "peak": burst, "unit": "TFLOP/s", "frac": None,
                            "traffic": None, "error": repr(ex)[:200]}
    try:
        line["hbm_kernels"] = hbm_kernel_table(step, pk["hbm_gbs"])
    except Exception as ex:
        line["hbm_kernels"] = {"error": repr(ex)[:200]}
    if world == 1 and not args.no_gpu_eager:
        try:
            torch.cuda.empty_cache()
            ge = gpu_eager_reference(cfg, B)
            line["gpu_eager_baseline"] = ge
            if ge.get("value"):
                line["vs_gpu_eager"] = line["value"] / ge["value"]
                if e2e and e2e.get("value") and ge.get("e2e_value"):
                    line["vs_gpu_eager_e2e"] = e2e["value"] / ge["e2e_value"]
        except Exception as ex:
            line["gpu_eager_baseline"] = {"value": None, "error": repr(ex)[:300]}
    if world == 1 and not args.no_cpu:
        try:
            v, cores, ms, kind = cpu_reference(cfg, 512, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "windows/s", "cores": cores, "kind": kind,
                                    "sample": "3 steps x 512 windows after 1 warm-up: the reference's train_test_epoch on the host cores"}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "windows/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "failed: " + repr(ex)[:160]}
    print(json.dumps(line), flush=True)
    if world > 1:
        _shutdown(dist, parallel, m)


def dist_rank_env():
    return int(os.environ.get("RANK", "0"))


def _shutdown(dist, parallel, model):
    """Leaves the process group cleanly: the captured graphs that hold NCCL kernels are released first (what stalled the
    teardown in round 1), every rank meets at a barrier (rank 0 alone ran the GEMM profile), then the group is destroyed.
    A watchdog ends the process if the teardown still does not return."""
    import torch
    sys.stdout.flush()
    sys.stderr.flush()
    t0 = time.time()

    def _watchdog():
        print("bench: teardown watchdog fired after 30 s", file=sys.stderr, flush=True)
        os._exit(0)
    threading.Timer(30.0, _watchdog).start()
    try:
        parallel.shutdown(model)
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        print(f"bench: rank {dist_rank_env()} left the process group cleanly in {time.time() - t0:.2f} s", file=sys.stderr)
    except Exception as ex:
        print("bench: teardown:", repr(ex)[:200], file=sys.stderr)
    sys.stdout.flush()
    os._exit(0)  # the watchdog thread must not keep the interpreter alive


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS), help="BASELINE.json config (2 = headline)")
    ap.add_argument("--batch", type=int, default=0, help="windows per GPU (default: the config's)")
    ap.add_argument("--precision", default=None, choices=["tf32", "fp32", "bf16"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-on-GPU comparator")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 3 s sustained leg")
    ap.add_argument("--no-resident", action="store_true",
                    help="rebuild the packed GEMM matrices from the parameter buffer every step (round-1 tail) instead of "
                         "keeping the master weights resident in the packed layout")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
