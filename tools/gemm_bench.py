"""Isolated timing of every GEMM-class launch of the SC-VAE step (real plan buffers, B200):
each recorded scv_gemm / scv_wgrad call is replayed `reps` times from a CUDA graph (warm L2).
   python tools/gemm_bench.py [--batch 2048] [--filter enc.1] [--reps 10] [--json out.json]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import scrubvae_b200 as sv
from scrubvae_b200.engine import TrainStep
from scrubvae_b200._ops import _ptr

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--filter", default="")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--json", default="")
ap.add_argument("--once", action="store_true", help="launch each selected call exactly once, eagerly (for ncu)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m, dcfg = bench.build_model(dev, os.environ.get("SCV_BENCH_PREC", "tf32"))
m.train()
opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
host = bench.synth_host_batch(a.batch, seed=0)
data = {k: v.to(dev) for k, v in host.items()}
step = TrainStep(m, opt, bench.LOSS_SCALE, a.batch, use_graph=False)
step.run(data)
eng, ops = step.eng, step.eng.ops
names = {}
for g in eng.W.values():
    pk = eng.packed16 if eng.packed16 is not None else eng.packed
    names[pk.data_ptr() + pk.element_size() * g.w] = (g.name + ":fwd", g.nnz)
    if g.wd is not None:
        names[pk.data_ptr() + pk.element_size() * (eng._n_fwd + g.wd)] = (g.name + ":dgrad", g.nnz_d)
    names[("g", eng.gpacked.data_ptr() + 4 * g.w)] = (g.name + ":wgrad", g.nnz)
calls = []
og, ow = ops.gemm, ops.wgrad
def rg(**kw):
    lab, nnz = names.get(_ptr(kw["W"]), ("gr.cat0:dgrad", kw["N"] * kw["K"]))
    calls.append((lab, "gemm", kw, 2.0 * kw["B"] * kw["Lo"] * nnz)); og(**kw)
def rw(**kw):
    lab, nnz = names[("g", _ptr(kw["dW"]))]
    calls.append((lab, "wgrad", kw, 2.0 * kw["B"] * kw["Lo"] * nnz)); ow(**kw)
ops.gemm, ops.wgrad = rg, rw
step._sequence()
del ops.gemm, ops.wgrad
torch.cuda.synchronize()
rows = []
sel = [c for c in calls if a.filter in c[0] and not c[0].startswith("gr.")]
if a.once:
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("once")  # ncu --nvtx --nvtx-include "once/" profiles exactly these launches
    for lab, kind, kw, fl in sel:
        (og if kind == "gemm" else ow)(**kw)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("launched", len(sel))
    sys.exit(0)
s = torch.cuda.Stream()
for lab, kind, kw, fl in sel:
    fn = og if kind == "gemm" else ow
    with torch.cuda.stream(s):
        fn(**kw)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(a.reps):
                fn(**kw)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); e1.record(s)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / a.reps
    rows.append(dict(label=lab, M=kw["B"] * kw["Lo"], N=kw["N"], K=kw["K"], us=us, tflops=fl / us / 1e6, flops=fl))
    print(f"{lab:22s} M={kw['B']*kw['Lo']:6d} N={kw['N']:5d} K={kw['K']:5d} {us:8.1f} us {fl/us/1e6:7.1f} TF/s", flush=True)
tot = sum(r["us"] for r in rows); fl = sum(r["flops"] for r in rows)
for cls in ("fwd", "dgrad", "wgrad"):
    t = sum(r["us"] for r in rows if r["label"].endswith(cls)); f = sum(r["flops"] for r in rows if r["label"].endswith(cls))
    if t: print(f"{cls:6s} {t:9.1f} us {f/t/1e6:7.1f} TF/s")
print(f"total {tot:9.1f} us {fl/tot/1e6:7.1f} TF/s")
if a.json:
    json.dump(rows, open(a.json, "w"), indent=1)
