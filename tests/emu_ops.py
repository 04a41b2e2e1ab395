"""TEST INFRASTRUCTURE — a torch-on-CPU statement of the *semantics* of every entry point of
include/scv.h (same method signatures as scrubvae_b200._ops.CudaOps).

Used only by tests: (1) CPU tests inject it into the engine to check the host logic (buffer
geometry, weight index maps, forward/backward sequencing) against the oracle without a GPU;
(2) GPU tests compare each CUDA kernel against the matching method on the same inputs.
The product never imports this file.
"""
import math

import torch

ACT_NONE, ACT_RELU, ACT_TANH, ACT_RELUMASK = 0, 1, 2, 3
ACT_ROUND_TF32 = 16
ACT_ACCUM = 32
BN, PRELU, TRAIN, ROUND_TF32 = 1, 2, 4, 8


def rtf32(x):
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) to a 10-bit mantissa"""
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def _v(ref, shape, strides):
    if ref is None:
        return None
    t, off = (ref, 0) if isinstance(ref, torch.Tensor) else (ref.t, ref.off)
    return torch.as_strided(t, shape, strides, t.storage_offset() + off)  # offsets count from the tensor (it may be a view)


def _tail(ref):
    """1-D view of a buffer from the referenced element to the end of its storage"""
    t, off = (ref, 0) if isinstance(ref, torch.Tensor) else (ref.t, ref.off)
    return t.reshape(-1)[off:]


class EmuOps:
    name = "emu"

    def __init__(self):
        self.n = 0

    def launch_count(self):
        return self.n

    def gemm(self, A, a_bs, a_ls, B, Lo, K, N, W, Y, y_bs, y_ls, bias=None, bias_mod=1, bias_n=0, n_last=None,
             R=None, r_bs=0, r_ls=0, act=ACT_NONE, out_scale=1.0, stats=None, precision=0,
             bnr_x=None, bnr_bs=0, bnr_ls=0, bnr_chan=None, bnr_slope=None, bnr_c=0, bnr_sums=None):
        self.n += 1
        rnd, accum, act = bool(act & ACT_ROUND_TF32), bool(act & ACT_ACCUM), act & 15
        n_last = N if n_last is None else n_last
        a = _v(A, (B, Lo, K), (a_bs, a_ls, 1))
        w = _v(W, (N, K), (K, 1))
        if precision == 2:  # bf16 operands (the buffers hold bf16), fp32 accumulation
            a, w = a.float(), w.float()
        elif precision != 0:  # TF32 tensor-core arithmetic: operands carry 10 mantissa bits, fp32 accumulation
            a, w = rtf32(a), rtf32(w)
        y = out_scale * (a.reshape(B * Lo, K) @ w.t()).reshape(B, Lo, N)
        if bias is not None:
            bb = _v(bias, (bias_mod,), (1,))
            n = torch.arange(N)
            y = y + torch.where(n < bias_n, bb[n % bias_mod], torch.zeros(()))
        r = _v(R, (B, Lo, N), (r_bs, r_ls, 1)) if R is not None else None
        valid = torch.ones(B, Lo, N, dtype=torch.bool)
        valid[:, Lo - 1, n_last:] = False
        if r is not None and act != ACT_RELUMASK:
            y = y + torch.where(valid, r, torch.zeros(()))
        if stats is not None:
            st = _v(stats, (2, N), (N, 1))
            yd = torch.where(valid, y, torch.zeros(())).double()
            st[0] += yd.sum((0, 1))
            st[1] += (yd * yd).sum((0, 1))
        if act == ACT_RELU:
            y = torch.relu(y)
        elif act == ACT_TANH:
            y = torch.tanh(y)
        elif act == ACT_RELUMASK:
            y = torch.where(torch.where(valid, r, torch.zeros(())) > 0, y, torch.zeros(()))
        if rnd:
            y = rtf32(y)
        if bnr_sums is not None:  # fused BatchNorm(+PReLU) backward reduction (scv_gemm_t bnr_*)
            C = bnr_c
            x = _v(bnr_x, (B, Lo, N), (bnr_bs, bnr_ls, 1))
            ch = torch.arange(N) % C
            if bnr_chan is not None:
                tb = _v(bnr_chan, (4, C), (C, 1))
                sc, sh, mu, rs = tb[0][ch], tb[1][ch], tb[2][ch], tb[3][ch]
            else:
                sc, sh, mu, rs = torch.ones(N), torch.zeros(N), torch.zeros(N), torch.ones(N)
            v = x * sc + sh
            g = torch.where(valid, y, torch.zeros(()))
            sm = _v(bnr_sums, (2 * C + 1,), (1,))
            if bnr_slope is not None:
                sl = _v(bnr_slope, (1,), (1,))
                sm[2 * C] += torch.where(valid & (v < 0), g * v, torch.zeros(())).double().sum()
                g = torch.where(v < 0, g * sl, g)
            if bnr_chan is not None:
                xh = torch.where(valid, (x - mu) * rs, torch.zeros(()))
                sm[:C] += g.double().sum((0, 1)).reshape(N // C, C).sum(0)
                sm[C:2 * C] += (g * xh).double().sum((0, 1)).reshape(N // C, C).sum(0)
        out = _v(Y, (B, Lo, N), (y_bs, y_ls, 1))
        if accum:
            y = y + out
        if n_last == N:
            out.copy_(y)
        else:
            out[:, :Lo - 1].copy_(y[:, :Lo - 1])
            out[:, Lo - 1, :n_last].copy_(y[:, Lo - 1, :n_last])

    def gemm_group(self, problems):
        for kw in problems:
            self.gemm(**{**kw, "precision": 0})

    def wgrad_group(self, problems):
        for kw in problems:
            self.wgrad(**{**kw, "precision": 0})

    def wgrad(self, A, a_bs, a_ls, B, Lo, K, N, dY, y_bs, y_ls, dW, dbias=None, bias_mod=1, bias_n=0, precision=0):
        self.n += 1
        a = _v(A, (B, Lo, K), (a_bs, a_ls, 1)).reshape(B * Lo, K)
        g = _v(dY, (B, Lo, N), (y_bs, y_ls, 1)).reshape(B * Lo, N)
        if precision == 2:
            a, g = a.float(), g.float()
            gm = g
        elif precision != 0:
            a, gm = rtf32(a), rtf32(g)
        else:
            gm = g
        _v(dW, (N, K), (K, 1)).add_(gm.t() @ a)
        if dbias is not None:
            db = _v(dbias, (bias_mod,), (1,))
            s = g.sum(0)
            for n in range(min(N, bias_n)):
                db[n % bias_mod] += s[n]

    def pack_input(self, x6d, root, arena, out, B, W, nx, Cc, halo, round_tf32=False):
        self.n += 1
        x = _v(x6d, (B, W, nx), (W * nx, nx, 1))
        r = _v(root, (B, W, 3), (W * 3, 3, 1))
        a = _v(arena, (2, 3), (3, 1))
        o = _v(out, (B, W, Cc), ((W + 2 * halo) * Cc, Cc, 1))
        if not isinstance(out, torch.Tensor):
            o = torch.as_strided(out.t, (B, W, Cc), ((W + 2 * halo) * Cc, Cc, 1), out.off + halo * Cc)
        else:
            o = torch.as_strided(out, (B, W, Cc), ((W + 2 * halo) * Cc, Cc, 1), halo * Cc)
        o.zero_()
        o[..., :nx] = x
        o[..., nx:nx + 3] = 2 * (r - a[0]) / (a[1] - a[0]) - 1
        if round_tf32 == 1:  # (2: the output buffer is bf16 — the assignments above already rounded)
            o.copy_(rtf32(o.clone()))

    # ---- BN + PReLU
    def _chan(self, Cc, mode, stats, fold, count, eps, gamma, beta, rm, rv):
        if not (mode & BN):
            one = torch.ones(Cc)
            return one, torch.zeros(Cc), torch.zeros(Cc), one
        if mode & TRAIN:
            st = _v(stats, (2, fold, Cc), (fold * Cc, Cc, 1))
            mean = st[0].sum(0) / count
            var = (st[1].sum(0) / count - mean * mean).clamp_min(0)
        else:
            mean, var = _v(rm, (Cc,), (1,)).double(), _v(rv, (Cc,), (1,)).double()
        rstd = 1.0 / torch.sqrt(var + eps)
        g, b = _v(gamma, (Cc,), (1,)), _v(beta, (Cc,), (1,))
        meanf, rstdf = mean.float(), rstd.float()
        scale = g * rstdf
        shift = b - meanf * g * rstdf
        return scale, shift, meanf, rstdf

    def bnact_fwd(self, X, x_bs, x_ls, B, L, Cc, mode, stats=None, fold=1, count=1.0, eps=1e-4, momentum=0.1,
                  gamma=None, beta=None, running_mean=None, running_var=None, slope=None,
                  H=None, h_bs=0, h_ls=0, U=None, u_bs=0, u_ls=0, chan_out=None):
        self.n += 1
        scale, shift, mean_c, rstd_c = self._chan(Cc, mode, stats, fold, count, eps, gamma, beta, running_mean, running_var)
        if chan_out is not None:
            _v(chan_out, (4, Cc), (Cc, 1)).copy_(torch.stack([scale, shift, mean_c, rstd_c]))
        if (mode & 5) == 5 and running_mean is not None:
            st = _v(stats, (2, fold, Cc), (fold * Cc, Cc, 1))
            mean = st[0].sum(0) / count
            var = (st[1].sum(0) / count - mean * mean).clamp_min(0)
            unb = var * count / (count - 1.0) if count > 1 else var
            rm, rv = _v(running_mean, (Cc,), (1,)), _v(running_var, (Cc,), (1,))
            rm.copy_(((1 - momentum) * rm.double() + momentum * mean).float())
            rv.copy_(((1 - momentum) * rv.double() + momentum * unb).float())
        x = _v(X, (B, L, Cc), (x_bs, x_ls, 1))
        a = x * scale + shift
        if mode & PRELU:
            s = _v(slope, (1,), (1,))
            a = torch.where(a < 0, s * a, a)
        rnd = (lambda t: rtf32(t)) if (mode & ROUND_TF32) else (lambda t: t)
        if H is not None:
            _v(H, (B, L, Cc), (h_bs, h_ls, 1)).copy_(rnd(a))
        if U is not None:
            am = torch.cat([a[:, :1], a[:, :-1]], 1)
            ap = torch.cat([a[:, 1:], a[:, -1:]], 1)
            u = _v(U, (B, L, 2, Cc), (u_bs, 2 * u_ls, u_ls, 1))
            u[:, :, 0] = rnd(0.25 * am + 0.75 * a)
            u[:, :, 1] = rnd(0.75 * a + 0.25 * ap)

    def _dout(self, B, L, Cc, dO, o_bs, o_ls, dU, u_bs, u_ls):
        g = torch.zeros(B, L, Cc)
        if dO is not None:
            g = g + _v(dO, (B, L, Cc), (o_bs, o_ls, 1))
        if dU is not None:
            u = _v(dU, (B, L, 2, Cc), (u_bs, 2 * u_ls, u_ls, 1))
            e, o = u[:, :, 0], u[:, :, 1]
            m = torch.cat([e[:, :1], o[:, :-1]], 1)      # dU[2l-1] (l>0) else dU[0]
            n = torch.cat([e[:, 1:], o[:, -1:]], 1)      # dU[2l+2] (l<L-1) else dU[2L-1]
            g = g + 0.75 * (e + o) + 0.25 * (m + n)
        return g

    def _bwd_common(self, X, x_bs, x_ls, B, L, Cc, mode, stats, fold, count, eps, gamma, beta, slope, dO, o_bs, o_ls,
                    dU, u_bs, u_ls):
        scale, shift, mean, rstd = self._chan(Cc, mode, stats, fold, count, eps, gamma, beta, None, None)
        x = _v(X, (B, L, Cc), (x_bs, x_ls, 1))
        v = x * scale + shift
        g = self._dout(B, L, Cc, dO, o_bs, o_ls, dU, u_bs, u_ls)
        ds = torch.zeros((), dtype=torch.double)
        if mode & PRELU:
            s = _v(slope, (1,), (1,))
            ds = torch.where(v < 0, g * v, torch.zeros(())).double().sum()
            g = torch.where(v < 0, g * s, g)
        xh = (x - mean) * rstd
        return g, xh, scale, ds

    def bnact_bwd_reduce(self, X, x_bs, x_ls, B, L, Cc, mode, sums, stats=None, fold=1, count=1.0, eps=1e-4,
                         gamma=None, beta=None, slope=None, dO=None, o_bs=0, o_ls=0, dU=None, u_bs=0, u_ls=0):
        self.n += 1
        g, xh, _, ds = self._bwd_common(X, x_bs, x_ls, B, L, Cc, mode, stats, fold, count, eps, gamma, beta, slope,
                                        dO, o_bs, o_ls, dU, u_bs, u_ls)
        sm = _v(sums, (2 * Cc + 1,), (1,))
        if mode & BN:
            sm[:Cc] += g.double().sum((0, 1))
            sm[Cc:2 * Cc] += (g * xh).double().sum((0, 1))
        if mode & PRELU:
            sm[2 * Cc] += ds

    def bnact_bwd_apply(self, X, x_bs, x_ls, B, L, Cc, mode, sums=None, stats=None, fold=1, count=1.0, eps=1e-4,
                        gamma=None, beta=None, slope=None, dO=None, o_bs=0, o_ls=0, dU=None, u_bs=0, u_ls=0,
                        dX=None, d_bs=0, d_ls=0, dgamma=None, dbeta=None, dslope=None):
        self.n += 1
        g, xh, scale, _ = self._bwd_common(X, x_bs, x_ls, B, L, Cc, mode, stats, fold, count, eps, gamma, beta, slope,
                                           dO, o_bs, o_ls, dU, u_bs, u_ls)
        sm = _v(sums, (2 * Cc + 1,), (1,)) if sums is not None else None
        if (mode & PRELU) and dslope is not None:
            _v(dslope, (1,), (1,)).add_(sm[2 * Cc].float())
        d = g
        if mode & BN:
            if dgamma is not None:
                _v(dgamma, (Cc,), (1,)).add_(sm[Cc:2 * Cc].float())
            if dbeta is not None:
                _v(dbeta, (Cc,), (1,)).add_(sm[:Cc].float())
            if (mode & 5) == 5:
                mg, mgx = (sm[:Cc] / count).float(), (sm[Cc:2 * Cc] / count).float()
                d = scale * (g - mg - xh * mgx)
            else:
                d = scale * g
        if dX is not None:
            _v(dX, (B, L, Cc), (d_bs, d_ls, 1)).copy_(rtf32(d) if (mode & ROUND_TF32) else d)

    # ---- latent
    def reparam_fwd(self, ms, ms_ld, eps, var, nvar, mu, L, zc, zc_ld, B, z, round_tf32=False):
        self.n += 1
        nsig = z * (z + 1) // 2
        row = _v(ms, (B, z + nsig), (ms_ld, 1))
        m, sig = row[:, :z], row[:, z:]
        idx = torch.tril_indices(z, z)
        Ld = torch.zeros(B, z, z)
        Ld[:, idx[0], idx[1]] = sig
        dg = torch.nn.functional.softplus(torch.diagonal(Ld, dim1=-2, dim2=-1))
        Ld = Ld - torch.diag_embed(torch.diagonal(Ld, dim1=-2, dim2=-1)) + torch.diag_embed(dg)
        if mu is not None:
            _v(mu, (B, z), (z, 1)).copy_(m)
        if L is not None:
            _v(L, (B, z, z), (z * z, z, 1)).copy_(Ld)
        if zc is not None:
            o = _v(zc, (B, zc_ld), (zc_ld, 1))
            o.zero_()
            if eps is not None:
                e = _v(eps, (B, z), (z, 1))
                o[:, :z] = (Ld @ e[..., None]).squeeze(-1) + m
            else:
                o[:, :z] = m
            if nvar > 0:
                o[:, z:z + nvar] = _v(var, (B, nvar), (nvar, 1))
            if round_tf32 == 1:
                o.copy_(rtf32(o.clone()))

    def reparam_bwd(self, ms, ms_ld, eps, dmu, dmu2, dmu2_scale, dz, dz_ld, dL, dms, dms_ld, B, z, round_tf32=False):
        self.n += 1
        nsig = z * (z + 1) // 2
        sig = _v(ms, (B, z + nsig), (ms_ld, 1))[:, z:]
        out = _v(dms, (B, dms_ld), (dms_ld, 1))
        out.zero_()
        g = torch.zeros(B, z)
        if dmu is not None:
            g = g + _v(dmu, (B, z), (z, 1))
        if dmu2 is not None:
            g = g + float(dmu2_scale) * _v(dmu2, (B, z), (z, 1))
        gz = None
        if dz is not None:
            gz = _v(dz, (B, z), (dz_ld, 1))
            g = g + gz
        out[:, :z] = g
        gL = torch.zeros(B, z, z)
        if dL is not None:
            gL = gL + _v(dL, (B, z, z), (z * z, z, 1))
        if gz is not None and eps is not None:
            e = _v(eps, (B, z), (z, 1))
            gL = gL + gz[:, :, None] * e[:, None, :]
        idx = torch.tril_indices(z, z)
        gs = gL[:, idx[0], idx[1]]
        isd = idx[0] == idx[1]
        sp = torch.where(sig > 20, torch.ones(()), torch.sigmoid(sig))
        gs = torch.where(isd[None, :], gs * sp, gs)
        out[:, z:z + nsig] = gs
        if round_tf32 == 1:
            out.copy_(rtf32(out.clone()))

    def kl(self, mu, L, loss, gscale, dmu, dL, B, z):
        self.n += 1
        m = _v(mu, (B, z), (z, 1))
        Lm = _v(L, (B, z, z), (z * z, z, 1))
        tril = torch.tril(Lm)
        dg = torch.diagonal(Lm, dim1=-2, dim2=-1)
        if loss is not None:
            val = (-0.5 * (1 + 2 * torch.log(dg) - m * m).double().sum() + 0.5 * (tril * tril).double().sum()) / B
            _v(loss, (1,), (1,)).add_(val)
        gs = float(_v(gscale, (1,), (1,))) if gscale is not None else 1.0
        if dmu is not None:
            _v(dmu, (B, z), (z, 1)).copy_(m * (gs / B))
        if dL is not None:
            g = tril - torch.diag_embed(1.0 / dg)
            _v(dL, (B, z, z), (z * z, z, 1)).copy_(g * (gs / B))

    # ---- reconstruction
    @torch.enable_grad()
    def recon_loss(self, xh, ld, offsets, target, root, arena, tree, n_tree, loss, root_hat, dxh, F, B, J, tree_kind=0):
        self.n += 1
        from oracle import scvae_oracle as orc  # test infrastructure may use the oracle
        tr = _v(tree, (n_tree,), (1,)).tolist()
        chains, pos = [], 1
        for _ in range(tr[0]):
            ln = tr[pos]
            chains.append(tr[pos + 1:pos + 1 + ln])
            pos += 1 + ln
        nx = J * 6
        rows = _v(xh, (F, ld), (ld, 1))
        x = rows[:, :nx].reshape(F, J, 6).detach().clone().requires_grad_(True)
        nr = rows[:, nx:nx + 3].detach().clone().requires_grad_(True)
        off = _v(offsets, (F, J, 3), (J * 3, 3, 1))
        tgt = _v(target, (F, J, 3), (J * 3, 3, 1))
        a = _v(arena, (2, 3), (3, 1))
        pose = orc.fwd_kin(x, off, torch.zeros(F, 3), tree=chains, eps=1e-8)
        ljpe = ((tgt - pose) ** 2).sum() / (B * 3 * J)
        rh = 0.5 * (nr + 1) * (a[1] - a[0]) + a[0]
        lroot = ((rh - _v(root, (F, 3), (3, 1))) ** 2).sum() / B
        gx, = torch.autograd.grad(ljpe, x)
        gr, = torch.autograd.grad(lroot, nr)
        lo = _v(loss, (2,), (1,))
        lo[0] += ljpe.detach().double()
        lo[1] += lroot.detach().double()
        if root_hat is not None:
            _v(root_hat, (F, 3), (3, 1)).copy_(rh.detach())
        d = _v(dxh, (F, ld), (ld, 1))
        d.zero_()
        d[:, :nx] = gx.reshape(F, nx)
        d[:, nx:nx + 3] = gr

    def out_bwd(self, xh, dxh, ld, g_jpe, g_root, nx, draw, d_bs, d_ls, B, W, round_tf32=False):
        self.n += 1
        y = _v(xh, (B, W, ld), (W * ld, ld, 1))
        d = _v(dxh, (B, W, ld), (W * ld, ld, 1))
        gj = float(_v(g_jpe, (1,), (1,))) if g_jpe is not None else 0.0
        gr = float(_v(g_root, (1,), (1,))) if g_root is not None else 0.0
        sc = torch.zeros(ld)
        sc[:nx] = gj
        sc[nx:nx + 3] = gr
        dv = d * sc * (1 - y * y)
        _v(draw, (B, W, ld), (d_bs, d_ls, 1)).copy_(rtf32(dv) if round_tf32 == 1 else dv)

    @torch.enable_grad()
    def gr_loss(self, preds, dpreds, ld, target, labels, B, d, num_keys, loss, gscale):
        self.n += 1
        n = len(preds)
        c = float(n * num_keys * B)
        gs = float(_v(gscale, (1,), (1,))) if gscale is not None else 1.0
        tot = torch.zeros((), dtype=torch.double)
        for e in range(n):
            w = c ** (-(n - e))
            p = _v(preds[e], (B, d), (ld, 1)).detach().clone().requires_grad_(True)
            if labels is not None:
                le = torch.nn.functional.cross_entropy(p, _v(labels, (B,), (1,)), reduction="sum")
            else:
                le = ((p - _v(target, (B, d), (d, 1))) ** 2).sum()
            tot += le.detach().double() * w
            if dpreds is not None:
                g, = torch.autograd.grad(le, p)
                o = _v(dpreds[e], (B, ld), (ld, 1))
                o.zero_()
                o[:, :d] = g * (w * gs)
        if loss is not None:
            _v(loss, (1,), (1,)).add_(tot)

    def gather(self, src, idx, dst, n, skip_neg=False, round_tf32=False):
        self.n += 1
        i = _v(idx, (n,), (1,)).long()
        s = src if isinstance(src, torch.Tensor) else src.t[src.off:]
        o = _v(dst, (n,), (1,))
        ok = i >= 0
        vals = s[i.clamp_min(0)]
        if round_tf32 == 1:
            vals = rtf32(vals)
        vals = vals.to(o.dtype)
        if skip_neg:
            o[ok] = vals[ok]
        else:
            o.copy_(torch.where(ok, vals, torch.zeros((), dtype=vals.dtype)))

    def sumsq(self, g, n, out):
        self.n += 1
        _v(out, (1,), (1,)).add_((_v(g, (n,), (1,)).double() ** 2).sum())

    @staticmethod
    def _live(pack_idx, pack_mask, n):
        """liveness of the packed positions: the bitmask (bit j of word w = position 32 w + j) when given, else idx >= 0"""
        if pack_mask is not None:
            words = _v(pack_mask, ((n + 31) // 32,), (1,)).to(torch.int64) & 0xFFFFFFFF
            bits = (words[:, None] >> torch.arange(32)[None, :]) & 1
            return bits.reshape(-1)[:n].bool()
        return _v(pack_idx, (n,), (1,)) >= 0

    def sumsq_packed(self, gpacked, pack_idx, n, out, pack_mask=None):
        self.n += 1
        live = self._live(pack_idx, pack_mask, n)
        _v(out, (1,), (1,)).add_((_v(gpacked, (n,), (1,))[live].double() ** 2).sum())

    def zero(self, t):
        t.zero_()

    def optim_step(self, p, g, m, v, n, sumsq, max_norm, gscale, lr, beta1, beta2, eps, weight_decay, step, kind,
                   hyper=None, pack_idx=None, packed_out=None, packed16_out=None, round_tf32=False, pack_mask=None):
        self.n += 1
        if pack_idx is not None or pack_mask is not None:  # resident-packed mode: only the live positions + operand copies
            live = self._live(pack_idx, pack_mask, n)
            P, G, M = _v(p, (n,), (1,)), _v(g, (n,), (1,)), _v(m, (n,), (1,))
            V = _v(v, (n,), (1,)) if v is not None else None
            sub = [P[live].clone(), G[live].clone(), M[live].clone(), V[live].clone() if V is not None else None]
            k = sub[0].numel()
            self.optim_step(sub[0], sub[1], sub[2], sub[3], k, sumsq, max_norm, gscale, lr, beta1, beta2, eps,
                            weight_decay, step, kind, hyper=hyper)
            self.n -= 1
            P[live], M[live] = sub[0], sub[2]
            if V is not None:
                V[live] = sub[3]
            # the operand copies mirror the master at EVERY position (padding included: the master holds zeros there)
            if packed_out is not None:
                _v(packed_out, (n,), (1,)).copy_(rtf32(P) if round_tf32 else P)
            if packed16_out is not None:
                _v(packed16_out, (n,), (1,)).copy_(P.to(torch.bfloat16))
            return
        if hyper is not None:
            hy = _v(hyper, (2,), (1,))
            lr, step = float(hy[0]), int(round(float(hy[1])))
        P, G = _v(p, (n,), (1,)), _v(g, (n,), (1,))
        coef = gscale
        if sumsq is not None:
            norm = math.sqrt(float(_v(sumsq, (1,), (1,)))) * gscale
            coef = min(1.0, max_norm / (norm + 1e-6)) * gscale
        gg = G * coef
        M = _v(m, (n,), (1,))
        if kind == 2:
            buf = gg.clone() if step == 1 else beta1 * M + gg
            M.copy_(buf)
            P.sub_(lr * (gg + beta1 * buf))
            return
        V = _v(v, (n,), (1,))
        if kind == 1:
            P.mul_(1 - lr * weight_decay)
        elif weight_decay != 0:
            gg = gg + weight_decay * P
        M.lerp_(gg, 1 - beta1)
        V.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        denom = V.sqrt() / math.sqrt(bc2) + eps
        P.addcdiv_(M, denom, value=-(lr / bc1))

    def loss_finalize(self, acc, scale, out, n):
        self.n += 1
        a, sc, o = _v(acc, (n,), (1,)), _v(scale, (n,), (1,)), _v(out, (n + 1,), (1,))
        o[:n] = a.float()
        o[n] = torch.where(sc != 0, sc.double() * a, torch.zeros((), dtype=torch.double)).sum().float()

    def unpack_root(self, xh, ld, nx, arena, root_hat, F):
        self.n += 1
        a = _v(arena, (2, 3), (3, 1))
        t, off = (xh, 0) if isinstance(xh, torch.Tensor) else (xh.t, xh.off)
        nr = torch.as_strided(t, (F, 3), (ld, 1), off + nx)
        _v(root_hat, (F, 3), (3, 1)).copy_(0.5 * (nr + 1) * (a[1] - a[0]) + a[0])

    def mi_loss(self, x, y, y_ld, xs, ys, var_s, logAx, bandwidth, S, B, z, dy, valid=None, loss=None, gscale=None, dx=None):
        """MutInfoEstimator.forward (model/disentangle.py:277-317) and its gradient w.r.t. x"""
        self.n += 1
        if valid is not None and float(_v(valid, (1,), (1,))[0]) == 0.0:
            return
        X = _v(x, (B, z), (z, 1)).clone().requires_grad_(True)
        Y = _v(y, (B, dy), (y_ld, 1))
        XS, YS = _v(xs, (S, z), (z, 1)), _v(ys, (S, dy), (dy, 1))
        VS = _v(var_s, (S, z), (z, 1)) if var_s is not None else torch.tensor([bandwidth])
        LA = _v(logAx, (S,), (1,))[None, :] if var_s is not None else _v(logAx, (1,), (1,))
        logAy = dy * (math.log(2 * math.pi) + math.log(bandwidth))
        with torch.enable_grad():
            dxm = X[:, None, :] - XS[None, :, :]
            dym = Y[:, None, :] - YS[None, :, :]
            sdx = ((dxm / VS) * dxm).sum(-1)
            sdy = ((dym / bandwidth) * dym).sum(-1)
            val = (torch.logsumexp(-0.5 * (LA + logAy + sdx + sdy), -1) - torch.logsumexp(-0.5 * (LA + sdx), -1)
                   - torch.logsumexp(-0.5 * (logAy + sdy), -1)).mean()
        if loss is not None:
            _v(loss, (1,), (1,)).add_(val.detach().double())
        if dx is not None:
            g, = torch.autograd.grad(val, X)
            sc = float(_v(gscale, (1,), (1,))[0]) if gscale is not None else 1.0
            _v(dx, (B, z), (z, 1)).add_(sc * g)

    def mi_update(self, mu, L, var, var_ld, xs, ys, var_s, logAx, bandwidth, S, z, dy, valid=None):
        self.n += 1
        _v(xs, (S, z), (z, 1)).copy_(_v(mu, (S, z), (z, 1)))
        _v(ys, (S, dy), (dy, 1)).copy_(_v(var, (S, dy), (var_ld, 1)))
        log2pi = math.log(2 * math.pi)
        if var_s is not None:
            v = torch.diagonal(_v(L, (S, z, z), (z * z, z, 1)), dim1=-2, dim2=-1) ** 2 + bandwidth
            _v(var_s, (S, z), (z, 1)).copy_(v)
            _v(logAx, (S,), (1,)).copy_(z * log2pi + torch.log(v).sum(-1))
        else:
            _v(logAx, (1,), (1,)).fill_(z * (log2pi + math.log(bandwidth)))
        if valid is not None:
            _v(valid, (1,), (1,)).fill_(1.0)

    # ---- moving_avg scrubber (reference model/disentangle.py:9-88)
    @staticmethod
    def _ma_means(x, x_ld, y, classes, nc, z, B):
        X, Y, cls = _v(x, (B, z), (x_ld, 1)), _v(y, (B,), (1,)), _v(classes, (nc,), (1,))
        sel = [(Y == cls[c]) for c in range(nc)]
        return torch.stack([X[s].mean(0) for s in sel]), torch.stack([s.sum() for s in sel]).float(), sel

    def ma_loss(self, x, x_ld, y, classes, nc, z, B, m1, m2, lam1, lam2, delta, lamdiff, stat, coef, loss=None):
        self.n += 1
        xbar, cnt, _ = self._ma_means(x, x_ld, y, classes, nc, z, B)
        M1, M2 = _v(m1, (nc, z), (z, 1)), _v(m2, (nc, z), (z, 1))
        l1, l2 = _v(lam1, (nc,), (1,)), _v(lam2, (nc,), (1,))
        for c in range(nc):
            if torch.linalg.norm(xbar[c] - M1[c]) < torch.linalg.norm(xbar[c] - M2[c]):
                l1[c] = torch.clamp(l1[c] - delta, 0.0, 1.0)
                l2[c] = l1[c] + lamdiff
            else:
                l2[c] = torch.clamp(l2[c] + delta, 0.0, 1.0)
                l1[c] = l2[c] - lamdiff
        est = 0.5 * (((1 - l1)[:, None] * xbar + l1[:, None] * M1) + ((1 - l2)[:, None] * xbar + l2[:, None] * M2))
        d = torch.triu(est.T[..., None] - est.T[..., None, :], diagonal=1)
        nrm = torch.linalg.norm(d)
        if loss is not None:
            _v(loss, (1,), (1,)).add_(nrm.double())
        _v(coef, (nc, z), (z, 1)).copy_(0.5 * (2 - l1 - l2)[:, None] * (nc * est - est.sum(0, keepdim=True)) / nrm / cnt[:, None])
        st = _v(stat, (nc, z + 1), (z + 1, 1))
        st[:, :z] = xbar
        st[:, z] = cnt

    def ma_backward(self, y, classes, coef, gscale, nc, z, B, dx, d_ld):
        self.n += 1
        Y, cls = _v(y, (B,), (1,)), _v(classes, (nc,), (1,))
        g = float(_v(gscale, (1,), (1,))) if gscale is not None else 1.0
        D, Cf = _v(dx, (B, z), (d_ld, 1)), _v(coef, (nc, z), (z, 1))
        for c in range(nc):
            D[Y == cls[c]] += g * Cf[c]

    def ma_update(self, x, x_ld, y, classes, nc, z, B, lam1, lam2, m1, m2, stat):
        self.n += 1
        xbar, cnt, _ = self._ma_means(x, x_ld, y, classes, nc, z, B)
        M1, M2 = _v(m1, (nc, z), (z, 1)), _v(m2, (nc, z), (z, 1))
        l1, l2 = _v(lam1, (nc,), (1,)), _v(lam2, (nc,), (1,))
        M1.copy_((1 - l1)[:, None] * xbar + l1[:, None] * M1)
        M2.copy_((1 - l2)[:, None] * xbar + l2[:, None] * M2)

    # ---- qda scrubber (reference model/disentangle.py:90-232)
    def qda_factor(self, S4, nc, z, SinvT, logdet):
        self.n += 1
        out, ld = _v(SinvT, (4, nc, z, z), (nc * z * z, z * z, z, 1)), _v(logdet, (4, nc), (nc, 1))
        for q in range(4):
            S = _v(S4[q], (nc, z, z), (z * z, z, 1))
            out[q] = torch.linalg.inv(S).transpose(-1, -2)
            ld[q] = torch.logdet(S)

    def qda_loss(self, x, x_ld, y, classes, m4, SinvT, logdet, nc, z, B, acc=None, gscale=None, dx=None, d_ld=0):
        self.n += 1
        X = _v(x, (B, z), (x_ld, 1))
        Y = _v(y, (B,), (1,))
        cls = _v(classes, (nc,), (1,))
        Si, ld = _v(SinvT, (4, nc, z, z), (nc * z * z, z * z, z, 1)), _v(logdet, (4, nc), (nc, 1))
        g = torch.zeros(B, z)
        for c in range(nc):
            i1 = (Y == cls[c])
            s = (i1.float() * 2 - 1)
            ll, t = [], []
            for q in range(4):
                r = X - _v(m4[q], (nc, z), (z, 1))[c:c + 1]
                tq = r @ Si[q, c]  # rows: (S^-1 r_b)^T = r_b^T S^-T ; Si holds S^-T
                t.append(tq)
                ll.append(-0.5 * (ld[q, c] + (r * tq).sum(1)))
            if acc is not None:
                a = _v(acc, (nc, 4), (4, 1))
                a[c, 0] += torch.where(i1, ll[1], ll[0]).double().sum()
                a[c, 1] += torch.where(i1, ll[3], ll[2]).double().sum()
                a[c, 2] += (s * (ll[1] - ll[0])).double().sum()
                a[c, 3] += (s * (ll[3] - ll[2])).double().sum()
            g += s[:, None] * ((t[0] - t[1]) + (t[2] - t[3]))
        if dx is not None:
            _v(dx, (B, z), (d_ld, 1)).add_(float(_v(gscale, (1,), (1,))) * 0.5 / (nc * B) * g)

    def qda_finalize(self, acc, lama, lamb, delta, lamdiff, nc, B, loss=None):
        self.n += 1
        a = _v(acc, (nc, 4), (4, 1))
        la, lb = _v(lama, (nc,), (1,)), _v(lamb, (nc,), (1,))
        for c in range(nc):
            if float(a[c, 0].float()) > float(a[c, 1].float()):
                la[c] = torch.clamp(la[c] - delta, 0.0, 1.0)
                lb[c] = la[c] + lamdiff
            else:
                lb[c] = torch.clamp(lb[c] + delta, 0.0, 1.0)
                la[c] = lb[c] - lamdiff
        if loss is not None:
            _v(loss, (1,), (1,)).add_(((a[:, 2] + a[:, 3]) * 0.5).sum() / nc / B)

    def qda_update(self, x, x_ld, y, classes, nc, z, B, lama, lamb, m4, S4, stat):
        self.n += 1
        X = _v(x, (B, z), (x_ld, 1))
        Y = _v(y, (B,), (1,))
        cls = _v(classes, (nc,), (1,))
        la, lb = _v(lama, (nc,), (1,)), _v(lamb, (nc,), (1,))
        M = [_v(t, (nc, z), (z, 1)) for t in m4]
        S = [_v(t, (nc, z, z), (z * z, z, 1)) for t in S4]
        for c in range(nc):
            for side in (0, 1):
                sel = X[(Y == cls[c]) if side else (Y != cls[c])]
                mean = sel.mean(0)
                cov = torch.cov(sel.T, correction=0) if sel.shape[0] else torch.full((z, z), float("nan"))
                for lam, off in ((la[c], 0), (lb[c], 2)):
                    M[off + side][c] = (1 - lam) * M[off + side][c] + lam * mean
                    S[off + side][c] = (1 - lam) * S[off + side][c] + lam * cov

    # ---- moving_avg_lsq scrubber (reference model/disentangle.py:393-538, polynomial order 1)
    @staticmethod
    def _mals_x(mu, mu_ld, bias, B, z):
        x = _v(mu, (B, z), (mu_ld, 1))
        return torch.cat([x, torch.ones(B, 1, dtype=x.dtype)], 1) if bias else x

    def mals_solve(self, Sxx0, Sxy0, Sxx1, Sxy1, l2_reg, bias, nx, ny, W0, W1):
        self.n += 1
        l2 = torch.ones(nx) * l2_reg
        if bias:
            l2[-1] = 0
        for Sxx, Sxy, W in ((Sxx0, Sxy0, W0), (Sxx1, Sxy1, W1)):
            A = _v(Sxx, (nx, nx), (nx, 1))
            A = A.diagonal_scatter(A.diagonal() + l2)
            _v(W, (nx, ny), (ny, 1)).copy_(torch.linalg.solve(A, _v(Sxy, (nx, ny), (ny, 1))))

    def mals_loss(self, mu, mu_ld, y, y_ld, W0, W1, bias, B, z, ny, l01=None, yhat0=None, yhat1=None, gscale=None, dmu=None,
                  d_ld=0):
        self.n += 1
        nx = z + (1 if bias else 0)
        x = self._mals_x(mu, mu_ld, bias, B, z)
        Y = _v(y, (B, ny), (y_ld, 1))
        w0, w1 = _v(W0, (nx, ny), (ny, 1)), _v(W1, (nx, ny), (ny, 1))
        p0, p1 = x @ w0, x @ w1
        if l01 is not None:
            acc = _v(l01, (2,), (1,))
            acc[0] += ((Y - p0).double() ** 2).sum()
            acc[1] += ((Y - p1).double() ** 2).sum()
        if yhat0 is not None:
            _v(yhat0, (B, ny), (ny, 1)).copy_(p0)
        if yhat1 is not None:
            _v(yhat1, (B, ny), (ny, 1)).copy_(p1)
        if dmu is not None:
            g = float(_v(gscale, (1,), (1,))) / B
            _v(dmu, (B, z), (d_ld, 1)).add_(g * ((p0 - Y) @ w0[:z].T + (p1 - Y) @ w1[:z].T))

    def mals_finalize(self, l01, lam0, lam1, delta, lamdiff, B, loss=None):
        self.n += 1
        acc = _v(l01, (2,), (1,))
        a0, a1 = _v(lam0, (1,), (1,)), _v(lam1, (1,), (1,))
        if float(acc[0].float()) < float(acc[1].float()):
            a0.copy_(torch.clamp(a0 - delta, 0.0, 1.0))
            a1.copy_(a0 + lamdiff)
        else:
            a1.copy_(torch.clamp(a1 + delta, 0.0, 1.0))
            a0.copy_(a1 - lamdiff)
        if loss is not None:
            _v(loss, (1,), (1,)).add_((acc[0] + acc[1]) * 0.5 / B)

    def mals_update(self, mu, mu_ld, y, y_ld, bias, B, z, ny, lam0, lam1, Sxx0, Sxy0, Sxx1, Sxy1):
        self.n += 1
        nx = z + (1 if bias else 0)
        x = self._mals_x(mu, mu_ld, bias, B, z)
        Y = _v(y, (B, ny), (y_ld, 1))
        xx, xy = x.T @ x, x.T @ Y
        for lam, Sxx, Sxy in ((lam0, Sxx0, Sxy0), (lam1, Sxx1, Sxy1)):
            a = float(_v(lam, (1,), (1,)))
            A, Bm = _v(Sxx, (nx, nx), (nx, 1)), _v(Sxy, (nx, ny), (ny, 1))
            A.copy_(a * A + xx)
            Bm.copy_(a * Bm + xy)

    def gen_features(self, xh, ld, root_hat, offsets, tree, n_tree, parts, B, W, J, norm=None, pose_out=None, heading=None,
                     avg3=None):
        """eval/eval.py:58-118 (FK through the oracle's fwd_kin, then the feature formulas)"""
        self.n += 1
        from oracle import scvae_oracle as orc
        tr = _v(tree, (n_tree,), (1,)).tolist()
        chains, pos = [], 1
        for _ in range(tr[0]):
            chains.append(tr[pos + 1:pos + 1 + tr[pos]])
            pos += 1 + tr[pos]
        x6 = _v(xh, (B * W, J, 6), (ld, 6, 1))
        off = _v(offsets, (B * W, J, 3), (J * 3, 3, 1))
        root = _v(root_hat, (B * W, 3), (3, 1)) if root_hat is not None else torch.zeros(B * W, 3)
        pose = orc.fwd_kin(x6, off, root, tree=chains, eps=1e-8).reshape(B, W, J, 3)
        if pose_out is not None:
            _v(pose_out, (B, W, J, 3), (W * J * 3, J * 3, 3, 1)).copy_(pose)
        if heading is not None:
            fwd = pose[:, W // 2, 1] - pose[:, W // 2, 0]
            fwd = fwd / fwd.norm(dim=-1, keepdim=True)
            yaw = -torch.atan2(fwd[:, 1], fwd[:, 0])
            _v(heading, (B, 2), (2, 1)).copy_(torch.stack([torch.sin(yaw), torch.cos(yaw)], -1))
        if avg3 is not None:
            pr = (parts if isinstance(parts, torch.Tensor) else parts.t).reshape(-1).tolist()
            plist, pos = [], 1
            for _ in range(pr[0]):
                plist.append(pr[pos + 1:pos + 1 + pr[pos]])
                pos += 1 + pr[pos]
            root_spd = (pose[:, 1:, 0] - pose[:, :-1, 0]).norm(dim=-1).mean(-1)
            d = [(pose[:, 1:, p[1:]] - pose[:, :-1, p[1:]]).norm(dim=-1).mean((-1, -2)) for p in plist]
            pred = torch.stack([root_spd, d[0], 0.5 * (d[1] + d[2])], -1)
            if norm is not None:
                nm = _v(norm, (6,), (1,))
                pred = (pred - nm[:3]) / nm[3:]
            _v(avg3, (B, 3), (3, 1)).copy_(pred)

    def d2f(self, src, dst, n):
        self.n += 1
        _v(dst, (n,), (1,)).copy_(_v(src, (n,), (1,)).float())
