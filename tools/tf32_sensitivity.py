"""How far does IDEAL TF32 arithmetic (operands rounded to nearest TF32, exact fp32 accumulation — emulated on
CPU by tests/emu_ops.py) move the step's losses and gradients away from the fp32 reference golden?
This is the error floor of any TF32 tensor-core implementation of the step, ours or cuDNN/cuBLAS's."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import scrubvae_b200 as sv
from scrubvae_b200.engine import Engine
from oracle import scvae_oracle as orc
from emu_ops import EmuOps
from test_engine_cpu import build_model, _rel

name, cond, gr = "step_small_heading.npz", ["heading"], ["heading"]
z = np.load(os.path.join(ROOT, "tests", "golden", name)); g = {k: z[k] for k in z.files}
ch, zd, B = [int(c) for c in g["meta_ch"]], int(g["meta_z"]), int(g["meta_B"])
res = {}
for precision in ("fp32", "tf32"):
    m, dcfg = build_model(ch, zd, cond, gr, None)
    m.precision = precision
    m.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd0.")})
    m._engine = Engine(m, ops=EmuOps())
    m.train()
    data = orc.synth_batch(B, seed=0)
    m._noise = orc.synth_eps(B, zd, seed=2)
    scale = {"prior": 1e-4, "jpe": 1.0, "root": 1.0, **{k + "_gr": 1.0 for k in gr}}
    data_o = sv.train.predict_batch(m, data, m.disentangle_keys)
    for k in ("mu", "x6d"):
        print(precision, "out", k, _rel(data_o[k], g["out." + k]))
    losses = sv.train.get_batch_loss(m, data, data_o, scale, dcfg)
    for p in m.parameters():
        p.grad = None
    losses["total"].backward()
    res[precision] = {n: p.grad.clone() for n, p in m.named_parameters()}
tot_e = tot_n = 0.0
for n in res["fp32"]:
    ref = torch.from_numpy(g["grad." + n])
    e = (res["tf32"][n].double() - ref.double()).norm().item(); tot_e += e * e; tot_n += ref.double().norm().item() ** 2
    if "weight" in n and ("conv" in n or "residual.0" in n or "fc" in n):
        print(f"{n:55s} fp32 {_rel(res['fp32'][n], ref):.2e}  emulated-tf32 {_rel(res['tf32'][n], ref):.2e}")
print("global gradient rel err (emulated tf32):", (tot_e / tot_n) ** 0.5)
