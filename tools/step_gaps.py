"""Where the captured step's wall time goes: per-stream busy time, the union over streams (GPU busy) and the idle gaps
between kernels, from a CUPTI trace (torch.profiler) of graph replays.
   python tools/step_gaps.py [--config 2] [--batch 0]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--config", default="2"); ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--reps", type=int, default=4); ap.add_argument("--top", type=int, default=14)
a = ap.parse_args()
torch.cuda.set_device(0)
import bench, scrubvae_b200 as sv
from scrubvae_b200.engine import TrainStep
cfg = bench.CONFIGS[a.config]
B = a.batch or cfg["batch"]
dev = torch.device("cuda", 0)
m, dcfg = bench.build_model(dev, cfg.get("precision", "tf32"), cfg)
m.train()
opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
data = {k: v.cuda() for k, v in bench.synth_host_batch(B, seed=0, cfg=cfg).items()}
step = TrainStep(m, opt, bench.loss_scale_for(cfg["feats"]), B, use_graph=True, resident=True)
for _ in range(4):
    step.run(data)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.reps):
        step.run()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda x: x[0])
t0, t1 = ks[0][0], max(k[1] for k in ks)
# union busy
busy, cur_s, cur_e = 0.0, None, None
gaps = []
for s, e, n in ks:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, n))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
wall = t1 - t0
tot = sum(e - s for s, e, _ in ks)
print(f"replays {a.reps}: wall {wall / a.reps:.1f} us/step, GPU busy (union over streams) {busy / a.reps:.1f} us/step, "
      f"idle {(wall - busy) / a.reps:.1f} us/step, sum of kernel durations {tot / a.reps:.1f} us/step, "
      f"{len(ks) / a.reps:.0f} launches/step")
import collections, re
def kind(n):
    m_ = re.search(r"::(\w+)\(", n) or re.search(r"(\w+)\(", n)
    return m_.group(1) if m_ else n[:48]
g = collections.defaultdict(lambda: [0, 0.0])
for d, n in gaps:
    key = kind(n)
    g[key][0] += 1; g[key][1] += d
print("idle time before kernels of each kind (us/step, count/step):")
for k, (c, d) in sorted(g.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"  {d / a.reps:8.1f} {c / a.reps:6.1f}  {k}")
d = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    key = kind(n)
    d[key][0] += 1; d[key][1] += e - s
print("kernel time by kind (us/step, count/step):")
for k, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1])[:a.top + 6]:
    print(f"  {t / a.reps:8.1f} {c / a.reps:6.1f}  {k}")
