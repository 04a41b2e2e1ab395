"""Which GEMM operands of the TF32 step are not TF32-representable when the kernel reads them? (GPU box)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import scrubvae_b200 as sv
from scrubvae_b200 import _ops
from oracle import scvae_oracle as orc
from test_engine_cpu import build_model

def view(ref, shape, strides):
    t, off = (ref, 0) if isinstance(ref, torch.Tensor) else (ref.t, ref.off)
    return torch.as_strided(t, shape, strides, off)

def lowbits(x):
    return (x.contiguous().view(torch.int32) & 0x1FFF).ne(0).float().mean().item()

ops = _ops.get_ops()
og, ow = ops.gemm, ops.wgrad
def gemm(A, a_bs, a_ls, B, Lo, K, N, W, Y, y_bs, y_ls, **kw):
    if kw.get("precision", 0):
        fa = lowbits(view(A, (B, Lo, K), (a_bs, a_ls, 1))); fw = lowbits(view(W, (N, K), (K, 1)))
        if fa > 0 or fw > 0:
            print(f"gemm  B={B} Lo={Lo} K={K} N={N}: unrounded A {fa:.3f} W {fw:.3f}")
    return og(A, a_bs, a_ls, B, Lo, K, N, W, Y, y_bs, y_ls, **kw)
def wgrad(A, a_bs, a_ls, B, Lo, K, N, dY, y_bs, y_ls, dW, **kw):
    if kw.get("precision", 0):
        fa = lowbits(view(A, (B, Lo, K), (a_bs, a_ls, 1))); fy = lowbits(view(dY, (B, Lo, N), (y_bs, y_ls, 1)))
        if fa > 0 or fy > 0:
            print(f"wgrad B={B} Lo={Lo} K={K} N={N}: unrounded A {fa:.3f} dY {fy:.3f}")
    return ow(A, a_bs, a_ls, B, Lo, K, N, dY, y_bs, y_ls, dW, **kw)
ops.gemm, ops.wgrad = gemm, wgrad

torch.manual_seed(1)
B = 16
m, dcfg = build_model([64, 128, 256, 512, 1024], 64, ["heading"], ["heading"], device="cpu")
m.precision = "tf32"
m = m.to("cuda").train()
data = {k: v.cuda() for k, v in orc.synth_batch(B, seed=0).items()}
m._noise = orc.synth_eps(B, 64, seed=2).cuda()
scale = {"prior": 1e-4, "jpe": 1.0, "root": 1.0, "heading_gr": 1.0}
data_o = sv.train.predict_batch(m, data, m.disentangle_keys)
losses = sv.train.get_batch_loss(m, data, data_o, scale, dcfg)
losses["total"].backward()
torch.cuda.synchronize()
print("done")
