"""Timeline of CTA 0 of one tensor-core GEMM launch (debug hook SCV_TC_TRACE in scv_gemm_tc.cu).
   python tools/tc_trace.py --filter enc.0.skip:fwd"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--filter", default="enc.0.skip:fwd"); ap.add_argument("--batch", type=int, default=2048)
a = ap.parse_args()
torch.cuda.set_device(0)
buf = torch.zeros(1 + 2 * 4000, dtype=torch.int64, device="cuda")
import bench, scrubvae_b200 as sv
from scrubvae_b200.engine import TrainStep
from scrubvae_b200._ops import _ptr
m, dcfg = bench.build_model(torch.device("cuda", 0), "tf32"); m.train()
opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
data = {k: v.cuda() for k, v in bench.synth_host_batch(a.batch, seed=0).items()}
step = TrainStep(m, opt, bench.LOSS_SCALE, a.batch, use_graph=False); step.run(data)
eng, ops = step.eng, step.eng.ops
names = {eng.packed.data_ptr() + 4 * g.w: g.name + ":fwd" for g in eng.W.values()}
names.update({eng.packed.data_ptr() + 4 * (eng._n_fwd + g.wd): g.name + ":dgrad" for g in eng.W.values() if g.wd is not None})
wnames = {eng.gpacked.data_ptr() + 4 * g.w: g.name + ":wgrad" for g in eng.W.values()}
calls = []
og, ow = ops.gemm, ops.wgrad
def rg(**kw):
    calls.append((names.get(_ptr(kw["W"]), "?"), kw)); og(**kw)
def rw(**kw):
    calls.append((wnames.get(_ptr(kw["dW"]), "?"), kw)); ow(**kw)
ops.gemm, ops.wgrad = rg, rw; step._sequence(); del ops.gemm, ops.wgrad
torch.cuda.synchronize()
kw = [c for c in calls if a.filter in c[0]][0][1]
if a.filter.endswith(":wgrad"):
    og = ow
og(**kw); torch.cuda.synchronize()
os.environ["SCV_TC_TRACE"] = str(buf.data_ptr())
og(**kw); torch.cuda.synchronize()
n = int(buf[0]); rec = buf[1:1 + 2 * min(n, 4000)].view(-1, 2).cpu().tolist()
rec.sort(key=lambda r: r[1]); t0 = rec[0][1]
role = {0: "prod", 1: "mma ", 2: "epi "}
evn = {(0, 0): "tile start", (0, 1): "first slot free", (0, 2): "last slot free", (1, 0): "tile start", (1, 1): "tmem free",
       (1, 2): "first stage full", (1, 3): "last stage full", (2, 0): "wait tfull", (2, 1): "tfull", (2, 2): "done"}
for tag, clk in rec[:80]:
    r, e, t = tag >> 40, (tag >> 32) & 0xff, tag & 0xffffffff
    print(f"{clk - t0:8d}  {role[r]} tile {t:5d}  {evn[(r, e)]}")
