"""Names + GEMM shapes of every libscv launch of one training step, in launch order (CPU, nothing executes).
   python tools/launch_order.py [B] -> JSON list on stdout"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import scrubvae_b200 as sv
from scrubvae_b200.engine import Engine, TrainStep
from scrubvae_b200._ops import _ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048


class Rec:
    name = "rec"

    def __init__(self):
        self.calls = []

    def launch_count(self):
        return len(self.calls)

    def __getattr__(self, k):
        def f(*a, **kw):
            self.calls.append((k, a, kw))
        return f


torch.set_default_device("meta") if False else None
m, dcfg = bench.build_model("cpu", "tf32")
rec = Rec()
m._engine = Engine(m, ops=rec)
m.train()
opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
step = TrainStep(m, opt, bench.LOSS_SCALE, 2, use_graph=False)
rec.calls.clear()
step._sequence()
eng = step.eng
names = {}
for g in eng.W.values():
    names[eng.packed.data_ptr() + 4 * g.w] = (g.name + ":fwd", g.nnz)
    if g.wd is not None:
        names[eng.packed.data_ptr() + 4 * (eng._n_fwd + g.wd)] = (g.name + ":dgrad", g.nnz_d)
    names[("g", eng.gpacked.data_ptr() + 4 * g.w)] = (g.name + ":wgrad", g.nnz)
out = []
for k, a, kw in rec.calls:
    d = {"op": k}
    if k == "gemm":
        lab, nnz = names.get(_ptr(kw["W"]), ("gr.cat0:dgrad", kw["N"] * kw["K"]))
        d.update(label=lab, Lo=kw["Lo"], M=B * kw["Lo"], N=kw["N"], K=kw["K"], flops=2.0 * B * kw["Lo"] * nnz,
                 R=kw.get("R") is not None, stats=kw.get("stats") is not None, prec=kw.get("precision", 0))
    elif k == "wgrad":
        lab, nnz = names[("g", _ptr(kw["dW"]))]
        d.update(label=lab, Lo=kw["Lo"], M=B * kw["Lo"], N=kw["N"], K=kw["K"], flops=2.0 * B * kw["Lo"] * nnz,
                 bias=kw.get("dbias") is not None, prec=kw.get("precision", 0))
    elif k.startswith("bnact"):
        d.update(L=kw["L"], C=kw["Cc"], M=B * kw["L"])
    out.append(d)
json.dump(out, sys.stdout)
