"""debug probe for the tcgen05 weight-gradient kernel (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scrubvae_b200._ops import get_ops
ops = get_ops()

def run(B, Lo, rows, C, s, taps, N, mode):
    K = taps * C
    a_bs, a_ls = rows * C, s * C
    if mode == "ones":
        A = torch.ones(B * rows * C + K); dY = torch.ones(B * Lo * N)
    elif mode == "ncol":   # dY[m][n] = n, A = 1  -> dW[n][k] = M*n
        A = torch.ones(B * rows * C + K); dY = torch.arange(N).float().repeat(B * Lo)
    elif mode == "kcol":   # A[..k] = position, dY = 1
        A = torch.arange(B * rows * C + K).float() % 7; dY = torch.ones(B * Lo * N)
    else:
        g = torch.Generator().manual_seed(1)
        A = torch.randn(B * rows * C + K, generator=g); dY = torch.randn(B * Lo * N, generator=g)
    outs = []
    for prec in (0, 1):
        dW = torch.zeros(N * K, device="cuda"); db = torch.zeros(N, device="cuda")
        ops.wgrad(A.cuda(), a_bs, a_ls, B, Lo, K, N, dY.cuda(), Lo * N, N, dW, dbias=db, bias_mod=N, bias_n=N, precision=prec)
        torch.cuda.synchronize()
        outs.append((dW.cpu().view(N, K), db.cpu()))
    (w0, b0), (w1, b1) = outs
    err = (w0 - w1).norm() / w0.norm()
    print(f"B={B} Lo={Lo} C={C} taps={taps} N={N} mode={mode}: rel err {err:.3e}  db err {(b0-b1).norm()/b0.norm():.3e}")
    if err > 1e-2:
        print(" ref[:3,:8]", w0[:3, :8].tolist())
        print(" got[:3,:8]", w1[:3, :8].tolist())
        print(" nonzero frac got", (w1 != 0).float().mean().item(), " got norm", w1.norm().item(), "ref norm", w0.norm().item())
        bad = ((w0 - w1).abs() > 1e-2 * w0.abs().max()).float()
        print(" bad rows frac per 32-row block:", bad.view(-1, 32 if N % 32 == 0 else 1, K).mean((1, 2))[:8].tolist())
        print(" bad cols frac per 32-col block:", bad.view(N, -1, 32).mean((0, 2))[:24].tolist())

for mode in ("ones", "ncol", "kcol", "rand"):
    run(130, 4, 8, 128, 1, 5, 256, mode)
run(256, 1, 1, 64, 1, 1, 128, "rand")
run(64, 13, 17, 64, 1, 5, 128, "rand")
run(64, 7, 16, 32, 2, 5, 64, "rand")
