#!/usr/bin/env python
"""Benchmark of the SC-VAE training step (BASELINE.json metric: pose windows/sec per training step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
  value        whole-job windows/s, inputs resident in HBM, one CUDA-graph replay per step
  e2e          the same steps through the public API scrubvae_b200.train.train_test_epoch(mode="train") over a
               loader of HOST (pinned) batches: every step's H2D copy (prefetched on a side stream) and the D2H
               read of its loss are inside the timed region
  roofline     tensor-pipe roofline of the dominant kernel class (the tcgen05 overlapping-row GEMMs): their
               launches of one step replayed back to back from one CUDA graph between CUDA events
  cpu_baseline the CPU oracle port (oracle/scvae_oracle.py) of the reference step on the host cores
`--impl reference` times that CPU port alone (the reference is pure Python/PyTorch and its tree does
not travel to the GPU box; the port is pinned to it by tests/golden)."""
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
LOSS_SCALE = {"prior": 1e-4, "jpe": 1.0, "root": 1.0, "heading_gr": 1.0}
FLOP_PER_WINDOW = 759.0e6  # layer GEMMs only, fwd + bwd
WORKLOAD = "SC-VAE mouse_skeleton default arch (ch 64-1024, k5, window 51, z64), conditional+grad_reversal on heading"


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
            return

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

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["nvml unavailable: %s" % self.err], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.mx, "sm_min_mhz": sm[0],
                "reasons": sorted(n for b, n in self.REASONS.items() if self.mask & b), "samples": len(sm)}


def cpu_step_throughput(B, steps, warmup, threads=None):
    """The oracle port of the reference step on the host cores; returns (windows/s, cores, ms/step)."""
    import torch
    from oracle import scvae_oracle as orc
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = orc.Cfg()
    sd = orc.synth_state_dict(cfg, seed=1)
    data = orc.synth_batch(B, seed=0)
    eps = orc.synth_eps(B, cfg.z_dim, seed=2)
    opt = None
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, sd, opt, _ = orc.train_step(sd, data, cfg, LOSS_SCALE, eps, opt_state=opt, lr=1e-4, optimizer="adamw",
                                          step=i + 1)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    return B / dt, cores, dt * 1e3


def run_reference(args, rank):
    if rank != 0:
        return
    B = 128
    v, cores, ms = cpu_step_throughput(B, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "pose windows/sec per training step", "value": v, "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{B} windows per step (CPU throughput is flat in batch)"},
        "cpu_baseline": {"value": v, "unit": "windows/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {B} windows, oracle port of the reference step"},
        "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def build_model(device, precision):
    import torch
    import scrubvae_b200 as sv
    from oracle import scvae_oracle as orc  # constants only (ARENA, KINEMATIC_TREE); not on the timed path
    mc = dict(type="rcnn", channel=DEFAULT_CH, kernel=5, z_dim=64, window=51, activation="prelu", diag=False,
              init_dilation=None, prior="gaussian", load_model=None, start_epoch=None, precision=precision)
    dcfg = dict(method={"conditional": ["heading"], "grad_reversal": ["heading"]}, features=["heading"], alpha=1.0)
    torch.manual_seed(1)
    m = sv.get.model(mc, None, None, dcfg, 18, "midfwd", arena_size=torch.tensor(orc.ARENA),
                     kinematic_tree=orc.KINEMATIC_TREE, discrete_classes={}, device=device, verbose=0)
    return m, dcfg


def synth_host_batch(B, seed):
    """Synthetic pose windows of the reference's shapes, pinned host memory."""
    import torch
    from oracle import scvae_oracle as orc
    chunks = []
    for i in range(0, B, 256):  # FK of the target pose is generated in chunks to bound host memory
        chunks.append(orc.synth_batch(min(256, B - i), seed=seed + i))
    d = {k: torch.cat([c[k] for c in chunks], 0) for k in chunks[0]}
    keep = ("x6d", "root", "offsets", "target_pose", "heading")
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
    for g in eng.W.values():
        nnz_by_w[eng.packed.data_ptr() + 4 * g.w] = g.nnz
        if g.wd is not None:
            nnz_by_w[eng.packed.data_ptr() + 4 * (eng._n_fwd + g.wd)] = g.nnz_d
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


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import scrubvae_b200 as sv
    from scrubvae_b200.engine import TrainStep
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    m, dcfg = build_model(dev, args.precision)
    m.train()
    opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
    comm = None
    if world > 1:
        from scrubvae_b200.parallel import GradAllReduce, broadcast_parameters
        broadcast_parameters(m)
        comm = GradAllReduce(m.engine, world)
        m.engine.comm = comm  # public-API path: total.backward() runs the same bucketed all-reduce
        opt.grad_scale = 1.0 / world
        torch.cuda.manual_seed(1234 + rank)  # rank-distinct reparameterisation noise
    host = synth_host_batch(B, seed=1000 * rank)
    data = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    step = TrainStep(m, opt, LOSS_SCALE, B, use_graph=not args.no_graph, comm=comm)
    ops = m.engine.ops

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    step.run(data)  # eager warm-up (+ graph capture on the next call)
    for _ in range(max(args.warmup, 3)):
        step.run()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the barrier: NVML start-up on one rank must not delay its peers inside the timed region
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        step.run()
        marks[i].record()  # per-step marks (diagnostic: median / max step time), same stream, no host sync
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    per_step = sorted(a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks))
    dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = dt.item()
    loss_total = float(step.plan.loss_out[-1])
    launches = step.n_launch * args.steps

    # ---- end to end through the public API, host inputs (e2e): sv.train.train_test_epoch over a loader of pinned
    # host batches — every step copies its 91.5 MB batch host->device (prefetched on a side stream while the previous
    # step computes) and reads the step's total loss back to the host (4 bytes, asynchronous)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    config = {"loss": dict(LOSS_SCALE), "disentangle": dcfg, "train": {}}
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
            _shutdown(dist)
        return
    pk, src = peaks()
    tf32_peak = pk["bf16_tflops_sustained"] / 2.0
    line = {
        "metric": "pose windows/sec per training step", "value": world * B * args.steps / dt, "unit": "windows/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3,
        "ms_per_step_median": per_step[len(per_step) // 2], "ms_per_step_max": per_step[-1],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32": "tf32 (fp32 storage, fp32 accumulate)", "fp32": "fp32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "optimizer": "adamw",
                   "parallelism": f"dp{world}", "l2": "per-step working set (>2 GB of activations) exceeds the 126 MB L2",
                   "cuda_graph": not args.no_graph, "loss_total_last": loss_total},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "step_tensor_frac": {"flop_per_window": FLOP_PER_WINDOW, "achieved_tflops": FLOP_PER_WINDOW * B * args.steps / dt / 1e12 * 1.0,
                             "peak_tflops": tf32_peak, "frac": FLOP_PER_WINDOW * B * args.steps / dt / 1e12 / tf32_peak,
                             "peak_source": f"{src}: bf16_tflops_sustained/2 (TF32)"},
    }
    try:
        gp = gemm_profile(step, B)
        # numerator: SURVEY.md §8(d)'s algorithmic figure (759.0 MFLOP per window, layer GEMMs only).  The launches'
        # own count (2 x rows x nnz(W), boundary taps of the transposed convolutions included) is 1.1 % higher and
        # is reported beside it, not used.
        alg = FLOP_PER_WINDOW * B
        ach = alg / gp["seconds"] / 1e12
        traffic = None
        try:  # DRAM bytes of the same launches from the committed ncu pass (profiles/, cold cache): informational
            with open(os.path.join(ROOT, "profiles", "r01_gemm_family_dram.json")) as f:
                traffic = json.load(f)["dram_bytes_per_step"]
        except Exception:
            pass
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s",
                            "frac": ach / tf32_peak, "traffic": traffic,
                            "traffic_note": "dram__bytes_read+write summed over the family's launches of one step "
                                            "(profiles/r01_gemm_family_dram.json), bytes per step",
                            "kernel": "tcgen05 overlapping-row GEMM family (gemm_tc_kernel + wgrad_tc_kernel): all its "
                                      "launches of one step replayed back to back from one CUDA graph, CUDA events",
                            "launches_per_step": gp["launches"], "gemm_seconds_per_step": gp["seconds"],
                            "gemm_share_of_step": gp["seconds"] / (dt / args.steps),
                            "alg_flops_per_step": alg, "launch_counted_flops_per_step": gp["flops"], "peak_source": f"{src}: bf16_tflops_sustained/2 (TF32)"}
    except Exception as ex:
        line["roofline"] = {"bound": "tensor", "achieved": None, "peak": tf32_peak, "unit": "TFLOP/s", "frac": None,
                            "traffic": None, "error": repr(ex)[:200]}
    if world == 1 and not args.no_cpu:
        try:
            v, cores, ms = cpu_step_throughput(256, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "windows/s", "cores": cores, "kind": "port",
                                    "sample": "3 steps x 256 windows after 1 warm-up, oracle port of the reference step"}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "windows/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: " + repr(ex)[:160]}
    print(json.dumps(line), flush=True)
    if world > 1:
        _shutdown(dist)


def _shutdown(dist):
    """Leave the process group without waiting on peers that may already be gone."""
    import torch
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    # no destroy_process_group / interpreter teardown: ranks finish at different times (rank 0 alone runs the GEMM
    # profile) and CUDA graphs holding captured collectives stalled the teardown on 2 GPUs
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=2048, help="windows per GPU")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
