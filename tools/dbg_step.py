"""per-parameter gradient error of the tf32 tensor-core step vs the reference golden (run on the GPU box)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import scrubvae_b200 as sv
from oracle import scvae_oracle as orc
from test_engine_cpu import build_model, _rel

name, cond, gr = "step_small_heading.npz", ["heading"], ["heading"]
z = np.load(os.path.join(ROOT, "tests", "golden", name)); g = {k: z[k] for k in z.files}
ch, zd, B = [int(c) for c in g["meta_ch"]], int(g["meta_z"]), int(g["meta_B"])
print("ch", ch, "z", zd, "B", B)
res = {}
for precision in ("fp32", "tf32"):
    m, dcfg = build_model(ch, zd, cond, gr, None, device="cpu")
    m.precision = precision
    m.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd0.")})
    m = m.to("cuda").train()
    data = {k: v.cuda() for k, v in orc.synth_batch(B, seed=0).items()}
    m._noise = orc.synth_eps(B, zd, seed=2).cuda()
    scale = {"prior": 1e-4, "jpe": 1.0, "root": 1.0, **{k + "_gr": 1.0 for k in gr}}
    data_o = sv.train.predict_batch(m, data, m.disentangle_keys)
    for k in ("mu", "L", "z", "root", "x6d"):
        print(precision, "out", k, _rel(data_o[k].cpu(), g["out." + k]))
    losses = sv.train.get_batch_loss(m, data, data_o, scale, dcfg)
    for k in list(scale) + ["total"]:
        print(precision, "loss", k, losses[k].item(), float(g["loss." + k]))
    for p in m.parameters():
        p.grad = None
    losses["total"].backward()
    res[precision] = {n: p.grad.cpu().clone() for n, p in m.named_parameters()}
for n in res["fp32"]:
    ref = torch.from_numpy(g["grad." + n])
    print(f"{n:60s} fp32 {_rel(res['fp32'][n], ref):.2e}  tf32 {_rel(res['tf32'][n], ref):.2e}  norm {ref.norm():.3e}")
