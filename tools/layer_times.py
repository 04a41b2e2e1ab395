"""Per-launch CUDA-event timing of every kernel class in one eager SC-VAE step (B200).
   python tools/layer_times.py [--batch 2048] [--precision tf32]  ->  table on stdout"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import scrubvae_b200 as sv
from scrubvae_b200.engine import TrainStep
from scrubvae_b200._ops import _ptr

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--precision", default="tf32")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m, dcfg = bench.build_model(dev, a.precision)
m.train()
opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
host = bench.synth_host_batch(a.batch, seed=0)
data = {k: v.to(dev) for k, v in host.items()}
step = TrainStep(m, opt, bench.LOSS_SCALE, a.batch, use_graph=False)
step.run(data)
step.run()
eng, ops = step.eng, step.eng.ops
names = {}
for g in eng.W.values():
    names[eng.packed.data_ptr() + 4 * g.w] = (g.name + ":fwd", g.nnz)
    if g.wd is not None:
        names[eng.packed.data_ptr() + 4 * (eng._n_fwd + g.wd)] = (g.name + ":dgrad", g.nnz_d)
    names[("g", eng.gpacked.data_ptr() + 4 * g.w)] = (g.name + ":wgrad", g.nnz)
recs = []
import types
orig = {}
def wrap(name):
    fn = getattr(ops, name)
    orig[name] = fn
    def w(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*args, **kw); e1.record()
        label, flops, shape = name, 0.0, ""
        if name == "gemm":
            label, nnz = names.get(_ptr(kw["W"]), ("?", kw["N"] * kw["K"]))
            flops = 2.0 * kw["B"] * kw["Lo"] * nnz
            shape = f"M={kw['B']*kw['Lo']} N={kw['N']} K={kw['K']} Lo={kw['Lo']}"
        elif name == "wgrad":
            label, nnz = names.get(("g", _ptr(kw["dW"])), ("?", kw["N"] * kw["K"]))
            flops = 2.0 * kw["B"] * kw["Lo"] * nnz
            shape = f"M={kw['B']*kw['Lo']} N={kw['N']} K={kw['K']} Lo={kw['Lo']}"
        elif name.startswith("bnact"):
            shape = f"B*L={kw['B']*kw['L']} C={kw['Cc']}"
        recs.append((label, shape, flops, e0, e1))
    setattr(ops, name, w)
for n in ("gemm", "wgrad", "pack_input", "bnact_fwd", "bnact_bwd_reduce", "bnact_bwd_apply", "reparam_fwd", "reparam_bwd",
          "kl", "recon_loss", "out_bwd", "gr_loss", "gather", "sumsq", "optim_step", "loss_finalize", "unpack_root"):
    wrap(n)
tot = {}
for r in range(a.reps):
    recs.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record(); step._sequence(); t1.record(); torch.cuda.synchronize()
    for i, (label, shape, flops, e0, e1) in enumerate(recs):
        k = (i, label, shape, flops)
        tot.setdefault(k, []).append(e0.elapsed_time(e1))
    step_ms = t0.elapsed_time(t1)
print(f"eager step {step_ms:.3f} ms, {len(recs)} launches")
rows = [(k[1], k[2], k[3], min(v)) for k, v in tot.items()]
cls = {}
print(f"{'kernel':28s} {'shape':44s} {'ms':>8s} {'TFLOP/s':>8s}")
for label, shape, flops, ms in rows:
    if ms >= 0.02:
        print(f"{label:28s} {shape:44s} {ms:8.3f} {flops/ms/1e9 if flops else 0:8.1f}")
    c = label.split(":")[-1] if ":" in label else label
    if label.startswith("gr."):
        c = "gr_" + c
    cls.setdefault(c, [0.0, 0.0, 0]); cls[c][0] += ms; cls[c][1] += flops; cls[c][2] += 1
print("---- per class")
for c, (ms, fl, n) in sorted(cls.items(), key=lambda kv: -kv[1][0]):
    print(f"{c:20s} n={n:4d} {ms:8.3f} ms  {fl/ms/1e9 if fl else 0:8.1f} TFLOP/s")
print("sum of launches %.3f ms" % sum(v[0] for v in cls.values()))
