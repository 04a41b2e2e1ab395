"""Every CUDA entry point of libscv.so (called through the C ABI via ctypes) against the torch-on-CPU
statement of its semantics (tests/emu_ops.py) on the same seeded inputs."""
import math

import pytest
import torch

from emu_ops import EmuOps, rtf32
from scrubvae_b200._ops import Ref, BN, PRELU, TRAIN, ACT_ACCUM, ACT_NONE, ACT_RELU, ACT_TANH, ACT_RELUMASK

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from scrubvae_b200._ops import get_ops
    return get_ops()


def run_both(ops, tensors, call, tol=1e-5, check=None):
    """tensors: name -> CPU tensor.  call(ops, T) launches with T[name] on the right device."""
    cpu = {k: v.clone() for k, v in tensors.items()}
    gpu = {k: v.clone().cuda() for k, v in tensors.items()}
    call(EmuOps(), cpu)
    call(ops, gpu)
    torch.cuda.synchronize()
    for k in (check or tensors.keys()):
        a, b = gpu[k].cpu().double(), cpu[k].double()
        err = (a - b).norm().item() / (b.norm().item() + 1e-20)
        assert err < tol or (a - b).abs().max().item() < tol * 1e-2, (k, err, (a - b).abs().max().item())


def g(seed):
    return torch.Generator().manual_seed(seed)


GEMM_CASES = [
    # B, Lo, rows(per batch), C(=a_ls/stride), stride, taps, N, n_last, act, resid, stats
    (3, 26, 56, 8, 2, 5, 16, None, ACT_NONE, False, True),
    (5, 7, 9, 64, 1, 3, 128, 64, ACT_NONE, True, True),
    (4, 51, 57, 112, 1, 7, 64, None, ACT_NONE, False, False),
    (6, 1, 1, 512, 1, 1, 44, None, ACT_RELU, False, False),
    (2, 51, 59, 8, 1, 9, 112, None, ACT_TANH, False, False),
    (7, 1, 1, 64, 1, 1, 4, None, ACT_RELUMASK, True, False),
    (130, 4, 8, 128, 1, 5, 256, None, ACT_NONE, True, True),
    # tensor-core tiling edges: several N tiles whose width is not a multiple of 32; K tiles of 208; ragged rows
    (40, 4, 8, 32, 1, 5, 720, None, ACT_RELU, True, True),
    (8, 51, 57, 112, 1, 7, 64, None, ACT_NONE, False, True),
    (300, 1, 1, 96, 1, 1, 272, None, ACT_NONE, False, False),
    (37, 13, 18, 64, 1, 6, 48, None, ACT_NONE, True, True),
    (33, 7, 16, 64, 2, 5, 160, 80, ACT_NONE, True, True),
]


@pytest.mark.parametrize("prec", [0, 1, 2])
@pytest.mark.parametrize("case", GEMM_CASES)
def test_gemm(ops, case, prec):
    B, Lo, rows, C, s, taps, N, n_last, act, resid, stats = case
    if prec == 2 and N < 16:
        pytest.skip("bf16 operands exist only on the tensor-core path (N >= 16)")
    K = taps * C
    a_bs, a_ls = rows * C, s * C
    T = {
        "A": torch.randn(B * rows * C + K, generator=g(1)),
        "W": torch.randn(N * K, generator=g(2)) / math.sqrt(K),
        "bias": torch.randn(N, generator=g(3)),
        "Y": torch.zeros(B * Lo * N),
        "R": torch.randn(B * Lo * N, generator=g(4)),
        "stats": torch.zeros(2 * N, dtype=torch.double),
    }
    if n_last is not None:
        bias_mod = N // 2
    else:
        bias_mod = N
    if prec == 1:  # operands already TF32-representable: the tensor-core result must then be fp32-exact arithmetic
        T["A"], T["W"] = rtf32(T["A"]), rtf32(T["W"])
    elif prec == 2:  # bf16 operand buffers (SCV_PREC_BF16): products of bf16 values are exact in fp32
        T["A"], T["W"] = T["A"].bfloat16(), T["W"].bfloat16()

    def call(o, t):
        o.gemm(t["A"], a_bs, a_ls, B, Lo, K, N, t["W"], t["Y"], Lo * N, N, bias=t["bias"], bias_mod=bias_mod,
               bias_n=N, n_last=n_last, R=t["R"] if resid else None, r_bs=Lo * N, r_ls=N, act=act, out_scale=0.5,
               stats=t["stats"] if stats else None, precision=prec)
    run_both(ops, T, call, tol=1e-5 if prec == 0 else 2e-5, check=["Y", "stats"])


# more work items than CTA pairs, long K: the pair kernel walks several items per pair
MANY_CASES_FWD = [
    (2600, 4, 8, 256, 1, 5, 272, None, ACT_NONE, True, True),   # 82 row tiles x 2 n tiles, K = 1280, residual + statistics
]

# many row tiles: the cluster-multicast kernels (two CTAs share the W tile / the A slabs), odd tile counts (one CTA of
# the last pair has nothing to store), two row tiles per item, several n tiles, ragged edges
BIG_CASES = [
    (701, 4, 8, 128, 1, 5, 528, None, ACT_NONE, True, True),     # 22 row tiles, 3 n tiles of 176
    (330, 7, 16, 64, 2, 5, 160, 80, ACT_NONE, True, True),       # stride 2, ragged last row
    (1100, 1, 1, 256, 1, 1, 272, None, ACT_RELU, False, False),  # linear layer, 9 row tiles
    (257, 13, 18, 64, 1, 6, 48, None, ACT_NONE, False, True),    # narrow N
    (513, 2, 6, 512, 1, 5, 1040, None, ACT_NONE, False, False),  # K = 2560, 5 n tiles
]


@pytest.mark.parametrize("mc", ["0", "1"])
@pytest.mark.parametrize("sub", ["0", "1", "2"])
@pytest.mark.parametrize("case", BIG_CASES)
def test_gemm_multitile(ops, case, sub, mc, monkeypatch):
    monkeypatch.setenv("SCV_TC_MC", mc)
    monkeypatch.setenv("SCV_TC_SUB", sub)
    monkeypatch.setenv("SCV_TC_PAIR", "0")  # the CTA-pair kernel (chosen by default for N >= 128, K >= 1024) has its own test
    test_gemm(ops, case, 1)
    if mc == "0":
        test_gemm(ops, case, 2)


@pytest.mark.parametrize("pair", ["1", "auto"])
@pytest.mark.parametrize("case", BIG_CASES + MANY_CASES_FWD)
def test_gemm_cta_pairs(ops, case, pair, monkeypatch):
    """gemm_tc2_kernel (cta_group::2): forced on every eligible shape, and as the default rule picks it."""
    if pair == "auto":
        monkeypatch.delenv("SCV_TC_PAIR", raising=False)
    else:
        monkeypatch.setenv("SCV_TC_PAIR", pair)
    test_gemm(ops, case, 1)
    test_gemm(ops, case, 1)


@pytest.mark.parametrize("mc", ["0", "1"])
@pytest.mark.parametrize("case", BIG_CASES)
def test_wgrad_multitile(ops, case, mc, monkeypatch):
    monkeypatch.setenv("SCV_TC_WMC", mc)
    test_wgrad(ops, case, 1)
    if mc == "0":
        test_wgrad(ops, case, 2)


# more work items than SMs (148): several items per CTA, under the dynamic work distribution (items drawn from a global
# counter, self-resetting: the second launch of each pair must find it clean) and the static round robin
MANY_CASES = [
    (5000, 4, 8, 64, 1, 5, 272, None, ACT_NONE, True, True),    # 157 row tiles x 2 n tiles, residual + statistics
    (2600, 13, 18, 32, 1, 3, 48, None, ACT_RELU, False, False),  # 265 row tiles, narrow N
]


@pytest.mark.parametrize("dyn", ["0", "1"])
@pytest.mark.parametrize("sub", ["1", "2"])
@pytest.mark.parametrize("case", MANY_CASES)
def test_gemm_many_items(ops, case, sub, dyn, monkeypatch):
    monkeypatch.setenv("SCV_TC_DYN", dyn)
    monkeypatch.setenv("SCV_TC_SUB", sub)
    for prec in (1, 1, 2):
        test_gemm(ops, case, prec)


@pytest.mark.parametrize("dyn", ["0", "1"])
@pytest.mark.parametrize("case", MANY_CASES)
def test_wgrad_many_items(ops, case, dyn, monkeypatch):
    monkeypatch.setenv("SCV_TC_DYN", dyn)
    for prec in (1, 1, 2):
        test_wgrad(ops, case, prec)


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("mc", ["0", "1"])
@pytest.mark.parametrize("B,K,N", [(2048, 4096, 68), (300, 1024, 44), (64, 8192, 272)])
def test_gemm_accumulate_split_k(ops, B, K, N, mc, prec, monkeypatch):
    """SCV_ACT_ACCUM: Y += A.W^T + bias; on the tensor-core path the reduction is split over CTAs (red.global.add)."""
    monkeypatch.setenv("SCV_TC_MC", mc)
    T = {"A": torch.randn(B * K + K, generator=g(1)), "W": torch.randn(N * K, generator=g(2)) / math.sqrt(K),
         "bias": torch.randn(N, generator=g(3)), "Y": torch.randn(B * N, generator=g(5))}
    if prec:
        T["A"], T["W"] = rtf32(T["A"]), rtf32(T["W"])

    def call(o, t):
        o.gemm(t["A"], K, 0, B, 1, K, N, t["W"], t["Y"], N, 0, bias=t["bias"], bias_mod=N, bias_n=N, act=ACT_ACCUM,
               out_scale=0.5, precision=prec)
    run_both(ops, T, call, tol=2e-5, check=["Y"])


@pytest.mark.parametrize("prec", [0, 1, 2])
@pytest.mark.parametrize("case", GEMM_CASES[:5] + GEMM_CASES[6:])
def test_wgrad(ops, case, prec):
    B, Lo, rows, C, s, taps, N, n_last, act, resid, stats = case
    if prec == 2 and N % 8:
        pytest.skip("bf16 dY rows must be 16-byte aligned (N % 8 == 0)")
    K = taps * C
    a_bs, a_ls = rows * C, s * C
    T = {
        "A": torch.randn(B * rows * C + K, generator=g(1)),
        "dY": torch.randn(B * Lo * N + 64, generator=g(2)),  # + read slack of one 128-byte slab (include/scv.h)
        "dW": torch.zeros(N * K),
        "db": torch.zeros(N),
    }
    bias_mod = N // 2 if n_last is not None else N
    if prec == 1:
        T["A"], T["dY"] = rtf32(T["A"]), rtf32(T["dY"])
    elif prec == 2:
        T["A"], T["dY"] = T["A"].bfloat16(), T["dY"].bfloat16()

    def call(o, t):
        o.wgrad(t["A"], a_bs, a_ls, B, Lo, K, N, t["dY"], Lo * N, N, t["dW"], dbias=t["db"], bias_mod=bias_mod,
                bias_n=N, precision=prec)
    run_both(ops, T, call, tol=2e-5, check=["dW", "db"])


def test_pack_input(ops):
    B, W, J = 5, 51, 18
    nx, C, halo = J * 6, 112, 3
    T = {"x": torch.randn(B, W, nx, generator=g(1)), "r": torch.rand(B, W, 3, generator=g(2)) * 50,
         "arena": torch.tensor([[-100.0, -100, 0], [100, 100, 50]]), "out": torch.zeros(B * (W + 2 * halo) * C)}
    run_both(ops, T, lambda o, t: o.pack_input(t["x"], t["r"], t["arena"], t["out"], B, W, nx, C, halo))


@pytest.mark.parametrize("B,L,C,fold,up", [(4, 13, 64, 1, True), (3, 7, 16, 2, False), (9, 51, 8, 1, True),
                                           (2, 4, 1024, 1, True)])
@pytest.mark.parametrize("mode", [BN | PRELU | TRAIN, BN | PRELU, PRELU, 0])
def test_bnact_fwd_bwd(ops, B, L, C, fold, up, mode):
    x = torch.randn(B * L * C, generator=g(1)) * 2 + 0.3
    xv = x.view(B * L, C).double()
    if fold == 2:  # statistics arrive as two column groups (polyphase GEMM): split rows in two halves
        h = (B * L) // 2
        st = torch.cat([xv[:h].sum(0), xv[h:].sum(0), (xv[:h] ** 2).sum(0), (xv[h:] ** 2).sum(0)])
    else:
        st = torch.cat([xv.sum(0), (xv ** 2).sum(0)])
    T = {
        "X": x, "stats": st, "gamma": 1 + 0.1 * torch.randn(C, generator=g(2)), "beta": 0.1 * torch.randn(C, generator=g(3)),
        "rm": torch.randn(C, generator=g(4)), "rv": torch.rand(C, generator=g(5)) + 0.5,
        "slope": torch.tensor([0.25]), "H": torch.zeros(B * L * C), "U": torch.zeros(B * 2 * L * C),
        "dO": torch.randn(B * L * C, generator=g(6)), "dU": torch.randn(B * 2 * L * C, generator=g(7)),
        "sums": torch.zeros(2 * C + 1, dtype=torch.double), "dX": torch.zeros(B * L * C),
        "dgamma": torch.zeros(C), "dbeta": torch.zeros(C), "dslope": torch.zeros(1),
    }
    cnt = float(B * L)

    def fwd(o, t):
        o.bnact_fwd(t["X"], L * C, C, B, L, C, mode, stats=t["stats"], fold=fold, count=cnt, eps=1e-4, momentum=0.1,
                    gamma=t["gamma"], beta=t["beta"], running_mean=t["rm"], running_var=t["rv"], slope=t["slope"],
                    H=t["H"], h_bs=L * C, h_ls=C, U=t["U"] if up else None, u_bs=2 * L * C, u_ls=C)
    run_both(ops, T, fwd, check=["H", "U", "rm", "rv"])
    if not (mode & BN) or (mode & TRAIN):
        def bwd(o, t):
            kw = dict(stats=t["stats"], fold=fold, count=cnt, eps=1e-4, gamma=t["gamma"], beta=t["beta"],
                      slope=t["slope"], dO=t["dO"], o_bs=L * C, o_ls=C, dU=t["dU"] if up else None, u_bs=2 * L * C,
                      u_ls=C)
            if mode & 3:
                o.bnact_bwd_reduce(t["X"], L * C, C, B, L, C, mode, t["sums"], **kw)
            o.bnact_bwd_apply(t["X"], L * C, C, B, L, C, mode, sums=t["sums"] if mode & 3 else None, dX=t["dX"],
                              d_bs=L * C, d_ls=C, dgamma=t["dgamma"], dbeta=t["dbeta"], dslope=t["dslope"], **kw)
        run_both(ops, T, bwd, tol=2e-5, check=["sums", "dX", "dgamma", "dbeta", "dslope"])


@pytest.mark.parametrize("z,nvar,with_eps", [(64, 2, True), (8, 9, True), (16, 0, False)])
def test_reparam_kl(ops, z, nvar, with_eps):
    B = 7
    nsig = z * (z + 1) // 2
    ms_ld, zc_ld = (z + nsig + 15) // 16 * 16, (z + nvar + 3) // 4 * 4
    T = {"ms": torch.randn(B, ms_ld, generator=g(1)), "eps": torch.randn(B, z, generator=g(2)),
         "var": torch.randn(B, max(nvar, 1), generator=g(3)), "mu": torch.zeros(B, z), "L": torch.zeros(B, z, z),
         "zc": torch.zeros(B, zc_ld), "loss": torch.zeros(1, dtype=torch.double), "gs": torch.tensor([0.37]),
         "dmu": torch.zeros(B, z), "dL": torch.zeros(B, z, z), "dmu2": torch.randn(B, z, generator=g(4)),
         "dz": torch.randn(B, zc_ld, generator=g(5)), "dms": torch.zeros(B, ms_ld)}
    T["ms"][:, z + 5] = 25.0  # softplus linear branch

    def call(o, t):
        e = t["eps"] if with_eps else None
        o.reparam_fwd(t["ms"], ms_ld, e, t["var"] if nvar else None, nvar, t["mu"], t["L"], t["zc"], zc_ld, B, z)
        o.kl(t["mu"], t["L"], t["loss"], None, None, None, B, z)
        o.kl(t["mu"], t["L"], None, t["gs"], t["dmu"], t["dL"], B, z)
        o.reparam_bwd(t["ms"], ms_ld, e, t["dmu"], t["dmu2"], -1.5, t["dz"], zc_ld, t["dL"], t["dms"], ms_ld, B, z)
    run_both(ops, T, call, tol=2e-5, check=["mu", "L", "zc", "loss", "dmu", "dL", "dms"])


def test_recon_loss_and_out_bwd(ops):
    from oracle import scvae_oracle as orc
    B, W, J, C = 3, 51, 18, 112
    F = B * W
    d = orc.synth_batch(B, seed=3)
    tr = [len(orc.KINEMATIC_TREE)]
    for c in orc.KINEMATIC_TREE:
        tr += [len(c)] + list(c)
    T = {"xh": torch.tanh(torch.randn(F, C, generator=g(1))), "off": d["offsets"].reshape(F, J, 3).contiguous(),
         "tgt": d["target_pose"].reshape(F, J, 3).contiguous(), "root": d["root"].reshape(F, 3).contiguous(),
         "arena": torch.tensor(orc.ARENA), "tree": torch.tensor(tr, dtype=torch.int32),
         "loss": torch.zeros(2, dtype=torch.double), "rh": torch.zeros(F, 3), "rh2": torch.zeros(F, 3),
         "dxh": torch.zeros(F, C), "gj": torch.tensor([0.7]), "gr": torch.tensor([1.3]),
         "draw": torch.zeros(B * (W + 6) * C)}

    def call(o, t):
        o.recon_loss(t["xh"], C, t["off"], t["tgt"], t["root"], t["arena"], t["tree"], len(tr), t["loss"], t["rh"],
                     t["dxh"], F, B, J)
        o.unpack_root(t["xh"], C, J * 6, t["arena"], t["rh2"], F)
        o.out_bwd(t["xh"], t["dxh"], C, t["gj"], t["gr"], J * 6, Ref(t["draw"], 3 * C), (W + 6) * C, C, B, W)
    run_both(ops, T, call, tol=3e-5, check=["loss", "rh", "rh2", "dxh", "draw"])


@pytest.mark.parametrize("tree,J", [
    ([[0, 1, 2, 3, 4, 5, 6], [0, 7], [7, 8, 9]], 10),                      # a chain longer than the fast path takes
    ([[0, 1], [0, 2], [0, 3], [1, 4], [2, 5], [3, 6], [4, 7], [5, 8], [6, 9]], 10),   # more chains than the fast path
    ([[0, 1, 2], [1, 3, 4], [3, 5]], 6),                                    # small skeleton on the fast path, 3 levels
])
def test_recon_loss_other_skeletons(ops, tree, J):
    """Both FK kernels on skeletons other than the mouse: the generic lane-per-joint kernel (long / many chains) and
    the lane-per-chain fast path with nested attachment levels, against the oracle FK through the ABI emulation."""
    from oracle import scvae_oracle as orc
    F, B = 37, 5
    C = (J * 6 + 3 + 3) // 4 * 4
    tr = [len(tree)]
    for c in tree:
        tr += [len(c)] + list(c)
    T = {"xh": torch.tanh(torch.randn(F, C, generator=g(1))), "off": torch.randn(F, J, 3, generator=g(2)) * 10,
         "tgt": torch.randn(F, J, 3, generator=g(3)) * 10, "root": torch.randn(F, 3, generator=g(4)) * 50,
         "arena": torch.tensor(orc.ARENA), "tree": torch.tensor(tr, dtype=torch.int32),
         "loss": torch.zeros(2, dtype=torch.double), "rh": torch.zeros(F, 3), "dxh": torch.zeros(F, C)}

    def call(o, t):
        o.recon_loss(t["xh"], C, t["off"], t["tgt"], t["root"], t["arena"], t["tree"], len(tr), t["loss"], t["rh"],
                     t["dxh"], F, B, J)
    run_both(ops, T, call, tol=3e-5, check=["loss", "rh", "dxh"])


@pytest.mark.parametrize("d,ce", [(2, False), (3, False), (4, True)])
def test_gr_loss(ops, d, ce):
    B, ld, n = 37, 4, 4
    T = {f"p{i}": torch.randn(B, ld, generator=g(i)) for i in range(n)}
    T.update({f"d{i}": torch.zeros(B, ld) for i in range(n)})
    T.update({"tgt": torch.randn(B, d, generator=g(9)), "lab": torch.randint(0, d, (B,), generator=g(10)),
              "loss": torch.zeros(1, dtype=torch.double), "gs": torch.tensor([0.9])})

    def call(o, t):
        preds = [Ref(t[f"p{i}"]) for i in range(n)]
        dps = [Ref(t[f"d{i}"]) for i in range(n)]
        o.gr_loss(preds, None, ld, None if ce else t["tgt"], t["lab"] if ce else None, B, d, 3, t["loss"], None)
        o.gr_loss(preds, dps, ld, None if ce else t["tgt"], t["lab"] if ce else None, B, d, 3, None, t["gs"])
    run_both(ops, T, call, tol=2e-5, check=["loss"] + [f"d{i}" for i in range(n)])


def test_gather_sumsq_optim_finalize(ops):
    n = 100003
    idx = torch.randint(-1, n, (n + 5,), generator=g(1)).to(torch.int32)
    T = {"src": torch.randn(n, generator=g(2)), "idx": idx, "dst": torch.ones(n + 5), "dst2": torch.ones(n + 5),
         "ss": torch.zeros(1, dtype=torch.double), "p": torch.randn(n, generator=g(3)),
         "m": torch.randn(n, generator=g(4)) * 0.01, "v": torch.rand(n, generator=g(5)) * 1e-4,
         "hyper": torch.tensor([3e-4, 7.0], dtype=torch.double),
         "acc": torch.rand(5, dtype=torch.double, generator=g(6)), "sc": torch.tensor([1.0, 0.0, 1e-4, 2.0, 1.0]),
         "lout": torch.zeros(6)}
    for kind in (0, 1, 2):
        def call(o, t, kind=kind):
            o.gather(t["src"], t["idx"], t["dst"], n + 5, False)
            o.gather(t["src"], t["idx"], t["dst2"], n + 5, True)
            t["ss"].zero_()
            o.sumsq(t["src"], n, t["ss"])
            o.optim_step(t["p"], t["src"], t["m"], t["v"], n, t["ss"], 10.0, 0.5, 1e-3, 0.9, 0.999, 1e-8, 0.01, 3,
                         kind, hyper=t["hyper"] if kind == 1 else None)
            o.loss_finalize(t["acc"], t["sc"], t["lout"], 5)
        run_both(ops, T, call, tol=1e-5, check=["dst", "dst2", "ss", "p", "m", "v", "lout"])


@pytest.mark.parametrize("use_mask", [False, True])
@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("rnd,bf16", [(False, False), (True, False), (False, True)])
def test_resident_packed_optimizer(ops, kind, rnd, bf16, use_mask):
    """The optimizer in the packed GEMM layout (resident master): padding positions (pack_idx < 0) keep their values in
    every array, live ones get the same update as the flat kernel, and the operand copies are written alongside."""
    n = 4 * 25013
    idx = torch.randint(-3, n, (n,), generator=g(1)).to(torch.int32)  # ~3/n negative would be too few: force a pattern
    idx[::7] = -1
    idx[5:40] = -1
    T = {"g": torch.randn(n, generator=g(2)), "idx": idx, "p": torch.randn(n, generator=g(3)),
         "m": torch.randn(n, generator=g(4)) * 0.01, "v": torch.rand(n, generator=g(5)) * 1e-4,
         "po": torch.full((n,), 7.0), "ss": torch.zeros(1, dtype=torch.double),
         "hyper": torch.tensor([3e-4, 7.0], dtype=torch.double)}
    if bf16:
        T["p16"] = torch.full((n,), 7.0).to(torch.bfloat16)
    if use_mask:  # the liveness bitmask instead of the index array (bit j of word w: position 32 w + j)
        live = torch.zeros((n + 31) // 32 * 32, dtype=torch.int64)
        live[:n] = (idx >= 0).to(torch.int64)
        words = (live.view(-1, 32) << torch.arange(32, dtype=torch.int64)[None, :]).sum(1)
        T["mask"] = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)

    def call(o, t):
        t["ss"].zero_()
        o.sumsq_packed(t["g"], None if use_mask else t["idx"], n, t["ss"], pack_mask=t.get("mask"))
        o.optim_step(t["p"], t["g"], t["m"], t["v"], n, t["ss"], 10.0, 0.5, 1e-3, 0.9, 0.999, 1e-8, 0.01, 3, kind,
                     hyper=t["hyper"] if kind == 1 else None, pack_idx=None if use_mask else t["idx"], packed_out=t["po"],
                     packed16_out=t.get("p16"), round_tf32=rnd, pack_mask=t.get("mask"))
    run_both(ops, T, call, tol=1e-5, check=["ss", "p", "m", "v", "po"])
    if bf16:
        run_both(ops, T, call, tol=4e-3, check=["p16"])
    # padding untouched on the device (bit-exact)
    gpu = {k: v.clone().cuda() for k, v in T.items()}
    call(ops, gpu)
    torch.cuda.synchronize()
    dead = (idx < 0)
    for k in ("p", "m", "v"):
        if not (kind == 2 and k == "v"):
            assert torch.equal(gpu[k].cpu()[dead], T[k][dead]), k


@pytest.mark.parametrize("z,ny,bias,l2", [(64, 2, True, 0.0), (128, 3, True, 0.05), (8, 2, False, 0.01), (64, 16, False, 0.0)])
def test_moving_avg_lsq_kernels(ops, z, ny, bias, l2):
    """scv_mals_solve / loss / finalize / update (csrc/scv_mals.cu) against the torch emulation (torch.linalg.solve):
    well-conditioned running covariances as they look after a few hundred updates."""
    B, nx = 300, z + int(bias)
    gg = g(7)
    X = torch.randn(900, nx, generator=gg)
    if bias:
        X[:, -1] = 1.0
    Wt = torch.randn(nx, ny, generator=gg)
    Yt = X @ Wt + 0.1 * torch.randn(900, ny, generator=gg)
    T = {"Sxx0": 0.5 * torch.eye(nx) + X[:600].T @ X[:600], "Sxy0": X[:600].T @ Yt[:600],
         "Sxx1": 0.7 * torch.eye(nx) + X.T @ X, "Sxy1": X.T @ Yt,
         "W0": torch.zeros(nx, ny), "W1": torch.zeros(nx, ny), "mu": torch.randn(B, z, generator=gg),
         "y": torch.randn(B, ny, generator=gg), "l01": torch.zeros(2, dtype=torch.double), "yh0": torch.zeros(B, ny),
         "yh1": torch.zeros(B, ny), "gs": torch.tensor([0.37]), "dmu": torch.randn(B, z, generator=gg),
         "lam0": torch.tensor([0.9]), "lam1": torch.tensor([1.0]), "loss": torch.zeros(1, dtype=torch.double)}
    T["y"] = (torch.cat([T["mu"], torch.ones(B, 1)], 1) if bias else T["mu"]) @ Wt + 0.3 * T["y"]

    def call(o, t):
        o.mals_solve(t["Sxx0"], t["Sxy0"], t["Sxx1"], t["Sxy1"], l2, bias, nx, ny, t["W0"], t["W1"])
        o.mals_loss(t["mu"], z, t["y"], ny, t["W0"], t["W1"], bias, B, z, ny, l01=t["l01"], yhat0=t["yh0"], yhat1=t["yh1"])
        o.mals_loss(t["mu"], z, t["y"], ny, t["W0"], t["W1"], bias, B, z, ny, gscale=t["gs"], dmu=t["dmu"], d_ld=z)
        o.mals_finalize(t["l01"], t["lam0"], t["lam1"], 1e-4, 0.1, B, loss=t["loss"])
        o.mals_update(t["mu"], z, t["y"], ny, bias, B, z, ny, t["lam0"], t["lam1"], t["Sxx0"], t["Sxy0"], t["Sxx1"], t["Sxy1"])
    run_both(ops, T, call, tol=2e-4, check=["W0", "W1", "yh0", "yh1", "l01", "dmu", "loss"])
    run_both(ops, T, call, tol=1e-5, check=["lam0", "lam1", "Sxx0", "Sxy0", "Sxx1", "Sxy1"])


@pytest.mark.parametrize("z,nc,B", [(64, 4, 300), (128, 2, 130), (8, 3, 37)])
def test_qda_kernels(ops, z, nc, B):
    """scv_qda_factor / loss / finalize / update (csrc/scv_qda.cu) against the torch emulation (torch.linalg.inv, logdet,
    torch.cov): running covariances as they look after some updates (identity blended with class covariances)."""
    gg = g(9)
    def spd():
        A = torch.randn(nc, z, 3 * z, generator=gg)
        return 0.6 * torch.eye(z)[None] + 0.4 * (A @ A.transpose(1, 2)) / (3 * z)
    T = {"x": torch.randn(B, z, generator=gg), "y": (torch.arange(B) % nc).to(torch.long), "cls": torch.arange(nc, dtype=torch.long),
         "SinvT": torch.zeros(4, nc, z, z), "logdet": torch.zeros(4, nc), "acc": torch.zeros(4 * nc, dtype=torch.double),
         "gs": torch.tensor([0.41]), "dx": torch.randn(B, z, generator=gg), "lama": torch.full((nc,), 0.2),
         "lamb": torch.full((nc,), 0.21), "loss": torch.zeros(1, dtype=torch.double), "stat": torch.zeros(2 * nc * (z + 1))}
    for q in range(4):
        T[f"m{q}"] = 0.3 * torch.randn(nc, z, generator=gg)
        T[f"S{q}"] = spd()

    def call(o, t):
        m4, S4 = [t[f"m{q}"] for q in range(4)], [t[f"S{q}"] for q in range(4)]
        o.qda_factor(S4, nc, z, t["SinvT"], t["logdet"])
        o.qda_loss(t["x"], z, t["y"], t["cls"], m4, t["SinvT"], t["logdet"], nc, z, B, acc=t["acc"])
        o.qda_loss(t["x"], z, t["y"], t["cls"], m4, t["SinvT"], t["logdet"], nc, z, B, gscale=t["gs"], dx=t["dx"], d_ld=z)
        o.qda_finalize(t["acc"], t["lama"], t["lamb"], 1e-3, 1e-2, nc, B, loss=t["loss"])
        o.qda_update(t["x"], z, t["y"], t["cls"], nc, z, B, t["lama"], t["lamb"], m4, S4, t["stat"])
    run_both(ops, T, call, tol=2e-4, check=["SinvT", "logdet", "acc", "dx", "loss"])
    run_both(ops, T, call, tol=2e-5, check=["lama", "lamb"] + [f"m{q}" for q in range(4)] + [f"S{q}" for q in range(4)])


@pytest.mark.parametrize("z,nc,B", [(64, 4, 300), (128, 3, 131), (8, 2, 9)])
def test_moving_avg_kernels(ops, z, nc, B):
    """scv_ma_loss / backward / update (csrc/scv_qda.cu) against the torch emulation of MovingAverageFilter."""
    gg = g(11)
    T = {"x": torch.randn(B, z, generator=gg), "y": (torch.arange(B) % nc).to(torch.long), "cls": torch.arange(nc, dtype=torch.long),
         "m1": 0.3 * torch.randn(nc, z, generator=gg), "m2": 0.3 * torch.randn(nc, z, generator=gg),
         "lam1": torch.full((nc,), 0.5), "lam2": torch.full((nc,), 0.51), "stat": torch.zeros(nc * (z + 1)),
         "coef": torch.zeros(nc, z), "loss": torch.zeros(1, dtype=torch.double), "gs": torch.tensor([1.7]),
         "dx": torch.randn(B, z, generator=gg)}
    T["x"] += 0.5 * T["y"][:, None].float()

    def call(o, t):
        o.ma_loss(t["x"], z, t["y"], t["cls"], nc, z, B, t["m1"], t["m2"], t["lam1"], t["lam2"], 1e-3, 1e-2, t["stat"], t["coef"],
                  loss=t["loss"])
        o.ma_backward(t["y"], t["cls"], t["coef"], t["gs"], nc, z, B, t["dx"], z)
        o.ma_update(t["x"], z, t["y"], t["cls"], nc, z, B, t["lam1"], t["lam2"], t["m1"], t["m2"], t["stat"])
    run_both(ops, T, call, tol=2e-5, check=["loss", "coef", "dx", "lam1", "lam2", "m1", "m2"])


def test_library_fails_loudly_when_missing(tmp_path):
    from scrubvae_b200 import _ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ops.load_library(str(tmp_path / "nope.so"))


@pytest.mark.parametrize("W", [51, 101])
def test_gen_features(ops, W):
    """FK of decoded windows + re-extracted heading / avg_speed_3d (scv_gen_features) against the torch statement."""
    from oracle import scvae_oracle as orc
    B, J, ld = 7, 18, 112
    d = orc.synth_batch(B, window=W, seed=3)
    xh = torch.zeros(B * W, ld)
    xh[:, :J * 6] = torch.tanh(torch.randn(B * W, J * 6, generator=g(5)))
    tree = [len(orc.KINEMATIC_TREE)]
    for c in orc.KINEMATIC_TREE:
        tree += [len(c)] + list(c)
    parts = [3, 6, 0, 1, 2, 3, 4, 5, 7, 1, 6, 7, 8, 9, 10, 11, 7, 5, 12, 13, 14, 15, 16, 17]
    T = {"xh": xh, "root": torch.randn(B * W, 3, generator=g(6)) * 20, "off": d["offsets"].reshape(B * W, J, 3).contiguous(),
         "tree": torch.tensor(tree, dtype=torch.int32), "parts": torch.tensor(parts, dtype=torch.int32),
         "norm": torch.tensor([0.4993, 0.7112, 0.6663, 0.4038, 0.3586, 0.4169]),
         "pose": torch.zeros(B, W, J, 3), "head": torch.zeros(B, 2), "avg3": torch.zeros(B, 3)}
    run_both(ops, T, lambda o, t: o.gen_features(t["xh"], ld, t["root"], t["off"], t["tree"], len(tree), t["parts"], B, W, J,
                                                 norm=t["norm"], pose_out=t["pose"], heading=t["head"], avg3=t["avg3"]),
             tol=2e-5, check=["pose", "head", "avg3"])


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("B,Lo,C,fold,taps,resid,bn,act_", [(300, 13, 64, 2, 3, True, True, True), (700, 1, 256, 4, 1, False, True, True),
                                                            (260, 26, 32, 1, 5, False, True, True), (129, 7, 64, 2, 3, True, False, True),
                                                            (64, 4, 128, 1, 5, False, True, False)])
def test_gemm_fused_bn_backward_reduce(ops, B, Lo, C, fold, taps, resid, bn, act_, prec):
    """scv_gemm with bnr_*: the data-gradient GEMM accumulates the BatchNorm / PReLU backward sums of the layer whose
    output gradient it produces (tensor-core epilogue; FFMA path = GEMM + stand-alone reduction)."""
    N, Cin = fold * C, 32
    K = taps * Cin
    rows = Lo + taps - 1
    n_last = N - C if (fold == 2 and Lo > 1) else None
    T = {"A": torch.randn(B * rows * Cin + K, generator=g(1)), "W": torch.randn(N * K, generator=g(2)) / math.sqrt(K),
         "Y": torch.zeros(B * Lo * N), "R": torch.randn(B * Lo * N, generator=g(4)),
         "X": torch.randn(B * Lo * N, generator=g(5)) * 1.5 + 0.2,
         "chan": torch.cat([torch.rand(C, generator=g(6)) + 0.5, torch.randn(C, generator=g(7)) * 0.3,
                            torch.randn(C, generator=g(8)) * 0.2, torch.rand(C, generator=g(9)) + 0.5]),
         "slope": torch.tensor([0.25]), "sums": torch.zeros(2 * C + 1, dtype=torch.double)}
    if prec:
        T["A"], T["W"] = rtf32(T["A"]), rtf32(T["W"])

    def call(o, t):
        o.gemm(t["A"], rows * Cin, Cin, B, Lo, K, N, t["W"], t["Y"], Lo * N, N, n_last=n_last,
               R=t["R"] if resid else None, r_bs=Lo * N, r_ls=N, precision=prec,
               bnr_x=t["X"], bnr_bs=Lo * N, bnr_ls=N, bnr_chan=t["chan"] if bn else None,
               bnr_slope=t["slope"] if act_ else None, bnr_c=C, bnr_sums=t["sums"])
    run_both(ops, T, call, tol=2e-5, check=["Y", "sums"])


@pytest.mark.parametrize("diag", [False, True])
@pytest.mark.parametrize("B,S,z,dy", [(37, 50, 16, 2), (128, 128, 64, 9), (5, 300, 96, 3)])
def test_mi_loss_and_update(ops, B, S, z, dy, diag):
    """scv_mi_update + scv_mi_loss (MutInfoEstimator, "mcmi" loss): value, gradient into x, and the not-yet-valid case."""
    T = {"mu": torch.randn(S, z, generator=g(1)), "L": torch.tril(torch.randn(S, z, z, generator=g(2)) * 0.3) + torch.eye(z) * 0.8,
         "var": torch.randn(S, dy + 2, generator=g(3)), "xs": torch.zeros(S, z), "ys": torch.zeros(S, dy), "var_s": torch.zeros(S, z),
         "logAx": torch.zeros(S), "valid": torch.zeros(1), "x": torch.randn(B, z, generator=g(4)),
         "y": torch.randn(B, dy + 2, generator=g(5)), "loss": torch.zeros(2, dtype=torch.double), "gs": torch.tensor([0.7]),
         "dx": torch.randn(B, z, generator=g(6))}

    def call(o, t):
        vs = t["var_s"] if diag else None
        o.mi_loss(t["x"], t["y"], dy + 2, t["xs"], t["ys"], vs, t["logAx"], 0.6, S, B, z, dy, valid=t["valid"], loss=Ref(t["loss"], 1),
                  gscale=t["gs"], dx=t["dx"])  # no estimator yet: nothing may change
        o.mi_update(t["mu"], t["L"], t["var"], dy + 2, t["xs"], t["ys"], vs, t["logAx"], 0.6, S, z, dy, valid=t["valid"])
        o.mi_loss(t["x"], t["y"], dy + 2, t["xs"], t["ys"], vs, t["logAx"], 0.6, S, B, z, dy, valid=t["valid"], loss=t["loss"])
        o.mi_loss(t["x"], t["y"], dy + 2, t["xs"], t["ys"], vs, t["logAx"], 0.6, S, B, z, dy, valid=t["valid"], gscale=t["gs"],
                  dx=t["dx"])
    run_both(ops, T, call, tol=2e-5, check=["xs", "ys", "logAx", "valid", "loss", "dx"] + (["var_s"] if diag else []))
