"""GPU tf32 step vs CPU ideal-TF32 emulation: forward outputs, losses, per-parameter gradients"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import scrubvae_b200 as sv
from scrubvae_b200.engine import Engine
from oracle import scvae_oracle as orc
from emu_ops import EmuOps
from test_engine_cpu import build_model, _rel

ch, zd, B = [8, 16, 32, 64, 128], 8, 6
cond = gr = ["heading"]
torch.manual_seed(1)
m0, dcfg = build_model(ch, zd, cond, gr, None)
sd = {k: v.clone() for k, v in m0.state_dict().items()}
data = orc.synth_batch(B, seed=0); eps = orc.synth_eps(B, zd, seed=2)
scale = {"prior": 1e-4, "jpe": 1.0, "root": 1.0, "heading_gr": 1.0}
res = {}
for dev in ("cpu", "cuda"):
    m, dcfg = build_model(ch, zd, cond, gr, None)
    m.precision = "tf32"
    m.load_state_dict(sd)
    if dev == "cpu":
        m._engine = Engine(m, ops=EmuOps())
    else:
        m = m.to("cuda")
    m.train()
    m._noise = eps.to(dev)
    d = {k: v.to(dev) for k, v in data.items()}
    data_o = sv.train.predict_batch(m, d, m.disentangle_keys)
    outs = {k: data_o[k].detach().cpu().clone() for k in ("mu", "L", "z", "root", "x6d")}
    plan = data_o["_plan"]
    acts = {}
    for name in ("ms", "xh", "zc"):
        acts[name] = getattr(plan, name).detach().cpu().clone()
    losses = sv.train.get_batch_loss(m, d, data_o, scale, dcfg)
    for p in m.parameters():
        p.grad = None
    losses["total"].backward()
    acts["dxh"] = plan.dxh.detach().cpu().clone(); acts["dms"] = plan.dms.detach().cpu().clone(); acts["dzc"] = plan.dzc.detach().cpu().clone()
    res[dev] = (outs, {k: v.item() for k, v in losses.items()}, {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters()}, acts)
for k in res["cpu"][0]:
    print("out", k, _rel(res["cuda"][0][k], res["cpu"][0][k]))
for k in res["cpu"][3]:
    print("act", k, _rel(res["cuda"][3][k], res["cpu"][3][k]))
for k in res["cpu"][1]:
    print("loss", k, res["cuda"][1][k], res["cpu"][1][k])
for n in res["cpu"][2]:
    print(f"{n:60s} {_rel(res['cuda'][2][n], res['cpu'][2][n]):.2e}")
