"""Kernel sequencing of the SC-VAE training step on libscv.so (include/scv.h).

The engine owns
  * ONE flat fp32 parameter buffer (the nn.Parameters of the drop-in modules are views into it, so
    state_dict keys/shapes stay the reference's) and ONE flat gradient buffer (p.grad are views);
  * the packed K-major weight matrices of every overlapping-row GEMM (forward + dgrad layouts),
    refreshed from the flat buffer by a single gather kernel (layout.py);
  * per-batch-size "plans": statically allocated halo-padded channels-last activation buffers and
    the pre-bound kernel launch lists for forward / loss / backward.  Everything a launch needs is
    fixed at plan time, so a whole step can be replayed from a CUDA graph.

Reference semantics followed (paths relative to /root/reference/src/scrubvae):
  ResVAE.forward/encode/decode model/residual.py:318-362,438-491; ResidualEncoder/Decoder
  :183-292; ResidualBlock(:71-119)/ResidualBlockTranspose(:122-180); CholeskyL :39-68; sampling
  :305-316; GRScrubber model/disentangle.py:541-660; get_batch_loss train/losses.py:182-324.

There is no PyTorch compute fallback: torch is used for allocation, tiny host-side plumbing
(one-hot / dtype casts of the conditioning features, RNG for the reparameterisation noise) and
streams.  `ops` is the C-ABI call table (scrubvae_b200._ops.CudaOps); tests inject a CPU emulation
of the ABI to check this host logic without a GPU.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import layout as lay
from ._ops import (ACT_ACCUM, ACT_NONE, ACT_RELU, ACT_RELUMASK, ACT_ROUND_TF32, ACT_TANH, BN, OUT_BF16, PREC, PRELU,
                   ROUND_TF32, TRAIN, Ref)

FEAT_FLOAT = ("avg_speed", "part_speed", "frame_speed", "avg_speed_3d", "heading", "heading_change", "fluorescence")


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def pad16(n: int) -> int:
    return (n + 15) // 16 * 16


class Act:
    """Halo-padded channels-last activation  t[b][hl + l][c]  (halo rows stay zero forever)."""

    def __init__(self, dev, B, L, C, hl=0, hr=0, even=False, slack=0, dtype=torch.float32):
        rows = hl + L + hr
        if even and rows % 2:
            rows += 1
        self.B, self.L, self.C, self.hl, self.rows = B, L, C, hl, rows
        self.bs, self.ls = rows * C, C
        self.t = torch.zeros(B * rows * C + slack, device=dev, dtype=dtype)

    def at(self, l: int = 0) -> Ref:
        return Ref(self.t, (self.hl + l) * self.C)

    def view(self) -> torch.Tensor:
        return self.t[:self.B * self.rows * self.C].view(self.B, self.rows, self.C)[:, self.hl:self.hl + self.L]


class SideLaunch:
    """A backward launch that only feeds the weight gradients (scv_wgrad): nothing downstream on the main stream
    reads its result before the final gather, so on CUDA it runs on a second stream and fills the SMs the
    dependent chain (dgrad -> BN backward -> dgrad ...) leaves idle: wave tails, HBM-bound passes, launch gaps."""
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn

    def __call__(self):
        self.fn()


class SideData:
    """A backward data-gradient launch off the dependent chain (the skip path's dgrad): third stream, joined by the
    next AfterSide launch (its consumer)."""
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn

    def __call__(self):
        self.fn()


class AfterSide:
    """A forward launch that consumes the result of an earlier SideLaunch (the block's skip convolution)."""
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn

    def __call__(self):
        self.fn()


class GemmW:
    """One GEMM layer's packed weights: forward matrix [N][K] (+ bias) and optional dgrad matrix."""
    __slots__ = ("name", "N", "K", "w", "b", "bias_mod", "bias_n", "gw", "gb", "dN", "dK", "wd", "nnz", "nnz_d")


class Engine:
    def __init__(self, model: nn.Module, ops=None):
        if ops is None:
            from ._ops import get_ops
            ops = get_ops()
        self.ops = ops
        self.m = model
        self.precision = PREC[getattr(model, "precision", "tf32")]
        # tensor-core path: every kernel that PRODUCES a GEMM operand rounds it to TF32 (round-to-nearest; the MMA itself
        # would truncate, which biases sums that cancel: BatchNorm backward subtracts batch means) or, in bf16 mode,
        # stores it as bf16.  `rnd` is the producers' output mode: 0 fp32, 1 TF32-rounded fp32, 2 bf16 operand buffers.
        self.rnd = self.precision
        self.op_dtype = torch.bfloat16 if self.precision == 2 else torch.float32
        self.device = next(model.parameters()).device
        if ops.name == "cuda" and self.device.type != "cuda":
            raise RuntimeError("scrubvae_b200: the model must live on a CUDA device (there is no CPU path); "
                               "call .to('cuda') first")
        self.plans: Dict[int, "Plan"] = {}
        self._flatten_params()
        self._build_weights()
        self.step_count = 0
        self.comm = None  # data-parallel gradient all-reduce hook (parallel.GradAllReduce) for the public-API path

    # ------------------------------------------------------------------ parameters
    def invalidate(self):
        self.m._engine = None

    def _flatten_params(self):
        dev = self.device
        names, params = zip(*list(self.m.named_parameters()))
        # flat layout: the parameters whose gradients the elementwise kernels write straight into gflat
        # (BatchNorm gamma/beta, PReLU slopes) come first, [0, n_direct); the GEMM weights/biases, whose
        # gradients arrive through gpacked + the final gather, follow.  Data parallelism all-reduces
        # gflat[:n_direct] and gpacked, never both copies of a weight gradient (parallel.py).
        direct_mods = {id(q) for mod in self.m.modules() if isinstance(mod, (nn.BatchNorm1d, nn.PReLU))
                       for q in mod.parameters(recurse=False)}
        offs, tot = [0] * len(params), 0
        for want_direct in (True, False):
            for i, p in enumerate(params):
                if p.dtype != torch.float32:
                    raise RuntimeError("scrubvae_b200: parameters must be float32")
                if (id(p) in direct_mods) == want_direct:
                    offs[i] = tot
                    tot += pad4(p.numel())
            if want_direct:
                self.n_direct = tot
        self.n_flat = tot
        self.flat = torch.zeros(tot, device=dev)
        self.gflat = torch.zeros(tot, device=dev)
        self.poff: Dict[str, int] = {}
        self.gviews: List[torch.Tensor] = []
        self.params = list(params)
        with torch.no_grad():
            for n, p, o in zip(names, params, offs):
                v = self.flat[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                p.grad = None
                self.poff[n] = o
                self.gviews.append(self.gflat[o:o + p.numel()].view(p.shape))
        # one counter tensor behind all BatchNorm num_batches_tracked buffers (one add per step)
        bns = [(n, mod) for n, mod in self.m.named_modules() if isinstance(mod, nn.BatchNorm1d)]
        self.nbt = torch.zeros(max(1, len(bns)), dtype=torch.long, device=dev)
        for i, (_, mod) in enumerate(bns):
            self.nbt[i] = mod.num_batches_tracked.to(dev)
            mod.num_batches_tracked = self.nbt[i]
        self.bn_index = {n: i for i, (n, _) in enumerate(bns)}
        # encode() / decode() alone advance only their own half's num_batches_tracked (torch counts per module call)
        enc = torch.tensor([1 if n.startswith("encoder.") else 0 for n, _ in bns] or [0], dtype=torch.long)
        self.nbt_enc, self.nbt_dec = enc.to(dev), (1 - enc).to(dev)

    def pref(self, name: str) -> Ref:
        return Ref(self.flat, self.poff[name])

    def gref(self, name: str) -> Ref:
        return Ref(self.gflat, self.poff[name])

    def _idx(self, name: str) -> torch.Tensor:
        p = dict(self.m.named_parameters())[name]
        return (self.poff[name] + torch.arange(p.numel())).view(p.shape)

    # ------------------------------------------------------------------ packed weights
    def _build_weights(self):
        m = self.m
        ch, k, z, W = m.ch, m.kernel, m.z_dim, m.window
        self.C0 = pad4(m.in_channels)
        self.nblk = len(ch) - 1
        self.W: Dict[str, GemmW] = {}
        self._pack_parts: List[torch.Tensor] = []   # forward matrices + biases (have gradients)
        self._pack_parts_d: List[torch.Tensor] = []  # dgrad matrices
        self._n_fwd = 0
        self._n_d = 0
        I = self._idx

        def add(name, w_idx, b_idx=None, bias_mod=None, d_idx=None):
            g = GemmW()
            g.name = name
            g.N, g.K = w_idx.shape
            g.nnz = int((w_idx >= 0).sum())  # algorithmic MACs per GEMM row (structural zeros excluded)
            g.nnz_d = int((d_idx >= 0).sum()) if d_idx is not None else 0
            assert g.K % 4 == 0, (name, g.K)
            g.w = self._n_fwd
            self._pack_parts.append(w_idx.reshape(-1))
            self._n_fwd += pad4(w_idx.numel())
            if w_idx.numel() % 4:
                self._pack_parts.append(torch.full((pad4(w_idx.numel()) - w_idx.numel(),), -1, dtype=torch.long))
            if b_idx is not None:
                nb = pad4(b_idx.numel())
                g.b = self._n_fwd
                g.bias_mod = bias_mod or b_idx.numel()
                self._pack_parts.append(lay.pad_cols(b_idx.reshape(1, -1), nb).reshape(-1))
                self._n_fwd += nb
            else:
                g.b, g.bias_mod = None, 1
            if d_idx is not None:
                g.dN, g.dK = d_idx.shape
                assert g.dK % 4 == 0, (name, g.dK)
                g.wd = self._n_d
                self._pack_parts_d.append(d_idx.reshape(-1))
                self._n_d += pad4(d_idx.numel())
                if d_idx.numel() % 4:
                    self._pack_parts_d.append(torch.full((pad4(d_idx.numel()) - d_idx.numel(),), -1, dtype=torch.long))
            else:
                g.wd = None
            self.W[name] = g
            return g

        e = "encoder."
        add("enc.conv_in", lay.conv_fprop(I(e + "conv_in.weight"), self.C0), I(e + "conv_in.bias"))
        for i in range(self.nblk):
            b = f"{e}res_layers.{i}."
            add(f"enc.{i}.skip", lay.conv_fprop(I(b + "skip.weight")), I(b + "skip.bias"),
                d_idx=lay.conv_dgrad_s2(I(b + "skip.weight")))
            add(f"enc.{i}.r0", lay.conv_fprop(I(b + "residual.0.weight")), I(b + "residual.0.bias"),
                d_idx=lay.conv_dgrad_s2(I(b + "residual.0.weight")))
            add(f"enc.{i}.r3", lay.conv_fprop(I(b + "residual.3.weight")), I(b + "residual.3.bias"),
                d_idx=lay.conv_dgrad_s1(I(b + "residual.3.weight")))
        from .model.residual import find_latent_dim, find_out_dim
        self.Ll = find_latent_dim(W, k, self.nblk)
        self.Cl = ch[-1]
        self.nsig = z * (z + 1) // 2
        self.ms_ld = pad16(z + self.nsig)
        w_sig = lay.fc_enc_fprop(I(e + "fc_sigma.0.weight"), self.Cl, self.Ll)
        b_sig = I(e + "fc_sigma.0.bias")
        if getattr(m, "is_diag", False):
            # CholeskyL(is_diag=True) (reference model/residual.py:55-56): the z outputs are the DIAGONAL of L.  The
            # kernels keep the packed lower-triangular row layout; the off-diagonal rows of the fc_sigma matrix become
            # structural zeros (weight and bias), so L's off-diagonals are exactly 0 and their gradients are dropped.
            diag_pos = torch.tensor([i * (i + 1) // 2 + i for i in range(z)])
            full_w = torch.full((self.nsig, w_sig.shape[1]), -1, dtype=torch.long)
            full_b = torch.full((self.nsig,), -1, dtype=torch.long)
            full_w[diag_pos] = w_sig
            full_b[diag_pos] = b_sig
            w_sig, b_sig = full_w, full_b
        wf = torch.cat([lay.fc_enc_fprop(I(e + "fc_mu.weight"), self.Cl, self.Ll), w_sig], 0)
        wf = lay.pad_rows(wf, self.ms_ld)
        bf = lay.pad_cols(torch.cat([I(e + "fc_mu.bias"), b_sig]).reshape(1, -1), self.ms_ld)
        add("enc.fc", wf, bf.reshape(-1), d_idx=lay.transpose_pad(wf, self.ms_ld))

        d = "decoder."
        self.cond_dim = m.conditional_dim
        self.zc_ld = pad8(z + self.cond_dim) if self.precision == 2 else pad4(z + self.cond_dim)
        wfi = lay.fc_dec_fprop(I(d + "fc_in.weight"), self.Cl, self.Ll, self.zc_ld)
        add("dec.fc_in", wfi, lay.fc_dec_bias(I(d + "fc_in.bias"), self.Cl, self.Ll),
            d_idx=lay.transpose_pad(wfi, self.Ll * self.Cl))
        for i in range(self.nblk):
            b = f"{d}res_layers.{i}."
            add(f"dec.{i}.skip", lay.conv_fprop(I(b + "skip.1.weight")), I(b + "skip.1.bias"),
                d_idx=lay.conv_dgrad_s1(I(b + "skip.1.weight")))
            add(f"dec.{i}.r0", lay.convT_fprop_s1(I(b + "residual.0.weight"), k // 2), I(b + "residual.0.bias"),
                d_idx=lay.convT_dgrad(I(b + "residual.0.weight")))
            co = I(b + "residual.3.bias").numel()
            add(f"dec.{i}.r3", lay.convT_fprop_s2(I(b + "residual.3.weight")), I(b + "residual.3.bias"), bias_mod=co,
                d_idx=lay.convT_dgrad(I(b + "residual.3.weight")))
        self.l_dec = find_out_dim(self.Ll, k, self.nblk)
        self.kf = W - self.l_dec + 7
        add("dec.conv_out", lay.convT_fprop_s1(I(d + "conv_out.weight"), 3, self.C0),
            lay.pad_cols(I(d + "conv_out.bias").reshape(1, -1), self.C0).reshape(-1),
            d_idx=lay.convT_dgrad(I(d + "conv_out.weight"), self.C0))

        # gradient-reversal heads: model/disentangle.py:583-632
        self._d_heads0 = self._n_d  # data-gradient matrices of the scrubber heads start here (they stay fp32: FFMA path)
        self.gr_keys: List[str] = []
        self.gr_layers: Dict[str, List[List[GemmW]]] = {}
        self.gr_alpha: Dict[str, float] = {}
        # moving-average least-squares scrubbers (model/disentangle.py:393-538): running covariances live in the modules'
        # buffers (state_dict keys as in the reference); the kernels read and update them in place
        self.mals_keys: List[str] = []
        self.mals_mod: Dict[str, nn.Module] = {}
        if "moving_avg_lsq" in m.disentangle:
            for key, mod in m.disentangle["moving_avg_lsq"].items():
                self.mals_keys.append(key)
                self.mals_mod[key] = mod
                mod._ops = self.ops
        # direct_lsq (train/losses.py:173-179, :253-256): no module, no state — feature -> dimension, set by get.model
        self.dlsq: Dict[str, int] = dict(getattr(m, "direct_lsq", None) or {})
        self.ma_keys: List[str] = []  # moving-average class-mean filters (model/disentangle.py:9-88)
        self.ma_mod: Dict[str, nn.Module] = {}
        if "moving_avg" in m.disentangle:
            for key, mod in m.disentangle["moving_avg"].items():
                self.ma_keys.append(key)
                self.ma_mod[key] = mod
                mod._ops = self.ops
        self.qda_keys: List[str] = []  # quadratic discriminant filters (model/disentangle.py:90-232)
        self.qda_mod: Dict[str, nn.Module] = {}
        if "qda" in m.disentangle:
            for key, mod in m.disentangle["qda"].items():
                self.qda_keys.append(key)
                self.qda_mod[key] = mod
                mod._ops = self.ops
        if "grad_reversal" in m.disentangle:
            for key, scr in m.disentangle["grad_reversal"].items():
                self.gr_keys.append(key)
                self.gr_alpha[key] = float(scr.reversal[0].alpha)
                pre = f"disentangle.grad_reversal.{key}.reversal.1."
                mlps = []
                for mi, mlp in enumerate(scr.reversal[1].members()):
                    layers = []
                    for li, mod in enumerate(mlp):
                        if not isinstance(mod, nn.Linear):
                            continue
                        wi = I(f"{pre}mlp{mi + 1}.{li}.weight")
                        assert wi.shape[1] % 4 == 0, "scrubvae_b200: z_dim must be a multiple of 8"
                        npad = pad4(wi.shape[0])
                        # (first layers have no data-gradient matrix of their own: the gradient into mu uses gr_cat below)
                        g = add(f"gr.{key}.{mi}.{li}", lay.pad_rows(wi, npad),
                                lay.pad_cols(I(f"{pre}mlp{mi + 1}.{li}.bias").reshape(1, -1), npad).reshape(-1),
                                d_idx=lay.transpose_pad(lay.pad_rows(wi, npad), npad) if li > 0 else None)
                        layers.append(g)
                    mlps.append(layers)
                self.gr_layers[key] = mlps
                # gradient reversal into mu: ONE GEMM over the concatenated first-layer output gradients of all
                # ensemble members, W_cat[z][sum N_m] = [W_0^T | W_1^T | ...]
                cat = torch.cat([lay.pad_rows(I(f"{pre}mlp{mi + 1}.0.weight"), mlps[mi][0].N).t() for mi in range(len(mlps))],
                                dim=1).contiguous()
                gc = GemmW()
                gc.name, gc.dN, gc.dK = f"gr.{key}.cat0", cat.shape[0], cat.shape[1]
                gc.nnz_d = int((cat >= 0).sum())
                gc.wd = self._n_d
                self._pack_parts_d.append(cat.reshape(-1))
                self._n_d += pad4(cat.numel())
                if cat.numel() % 4:
                    self._pack_parts_d.append(torch.full((pad4(cat.numel()) - cat.numel(),), -1, dtype=torch.long))
                self.gr_cat = getattr(self, "gr_cat", {})
                self.gr_cat[key] = gc

        dev = self.device
        fwd_idx = torch.cat(self._pack_parts)
        d_idx = torch.cat(self._pack_parts_d) if self._pack_parts_d else torch.zeros(0, dtype=torch.long)
        assert fwd_idx.numel() == self._n_fwd and d_idx.numel() == self._n_d
        self.pack_idx = torch.cat([fwd_idx, d_idx]).to(torch.int32).to(dev)
        # liveness bitmask of the forward matrices' positions (bit j of word w: position 32 w + j holds a parameter): what the
        # resident optimizer and the packed gradient norm read instead of the 4-byte index entries
        live = torch.zeros((int(fwd_idx.numel()) + 31) // 32 * 32, dtype=torch.int64)
        live[:fwd_idx.numel()] = (fwd_idx >= 0).to(torch.int64)
        words = (live.view(-1, 32) << torch.arange(32, dtype=torch.int64)[None, :]).sum(1)
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)  # two's complement into int32 storage
        self.pack_mask = words.to(torch.int32).to(dev)
        self.packed = torch.zeros(self._n_fwd + self._n_d, device=dev)
        # bf16 mode: the GEMMs read bf16 copies of the packed matrices (same element offsets); `packed` (fp32) then only
        # serves the biases, which the epilogues add in fp32
        self.packed16 = torch.zeros(self._n_fwd + self._n_d, device=dev, dtype=torch.bfloat16) if self.precision == 2 else None
        self.gpacked = torch.zeros(self._n_fwd, device=dev)
        self.inv_idx = lay.inverse_map(fwd_idx, self.n_flat).to(torch.int32).to(dev)
        self.grads_dirty = False   # gpacked / gflat hold gradients of a piecewise backward (TrainStep zeroes them)
        # RESIDENT-PACKED training state (TrainStep(resident=True)): during a run of fused steps the fp32 master of every
        # GEMM weight / bias lives in the packed layout (`pmaster`), the optimizer updates it there with coalesced accesses
        # and writes the operand copies itself — no weight repack and no gradient unpack in the step.  `flat` (the
        # reference-layout buffer the nn.Parameters view) is refreshed on demand (ensure_flat).
        self.pmaster = None
        self.idx_d_from_p = None   # data-gradient layout -> position in pmaster
        self.flat_valid = True     # `flat` holds the current weights
        self.resident_valid = False  # `pmaster` (+ packed moments) hold the current weights
        self._resident_opt = None
        del self._pack_parts, self._pack_parts_d
        self.max_k = max(max(g.K, g.dK if g.wd is not None else 0) for g in self.W.values())
        for gc in getattr(self, "gr_cat", {}).values():
            self.max_k = max(self.max_k, gc.dK)
        # gpacked = [encoder layers | decoder layers + scrubber heads]; backward finishes the second part first
        self.gp_split = self.W["dec.fc_in"].w

    def wref(self, g: GemmW) -> Ref:
        return Ref(self.packed16 if self.packed16 is not None else self.packed, g.w)

    def bref(self, g: GemmW) -> Optional[Ref]:
        return None if g.b is None else Ref(self.packed, g.b)

    def wdref(self, g: GemmW) -> Ref:
        return Ref(self.packed16 if self.packed16 is not None else self.packed, self._n_fwd + g.wd)

    # ---- resident-packed master (see __init__)
    def resident_import(self, opt):
        """flat -> pmaster / packed operand copies / packed moments (coalesced-store gathers)."""
        ops, nf = self.ops, self._n_fwd
        self.ensure_flat()
        if self.pmaster is None:
            self.pmaster = torch.zeros(nf, device=self.device)
            d_idx = self.pack_idx[nf:].long()
            comp = torch.where(d_idx >= 0, self.inv_idx.long()[d_idx.clamp_min(0)], torch.full_like(d_idx, -1))
            self.idx_d_from_p = comp.to(torch.int32)
        if getattr(opt, "mp", None) is None or opt.mp.numel() != nf:
            opt.mp, opt.vp = torch.zeros(nf, device=self.device), torch.zeros(nf, device=self.device)
        ops.gather(self.flat, self.pack_idx, self.pmaster, nf, False, round_tf32=0)
        ops.gather(opt.m, self.pack_idx, opt.mp, nf, False, round_tf32=0)
        ops.gather(opt.v, self.pack_idx, opt.vp, nf, False, round_tf32=0)
        self.repack("fwd")
        self.resident_valid = True
        self._resident_opt = opt

    def resident_dgrad_matrices(self):
        """data-gradient matrices from the resident master (side stream, beside the forward pass)"""
        nf = self._n_fwd
        n = self.packed.numel() - nf
        if n <= 0:
            return
        if self.packed16 is not None:
            self.ops.gather(self.pmaster, self.idx_d_from_p, Ref(self.packed16, nf), n, False, round_tf32=2)
            h0 = self._d_heads0
            if n > h0:
                self.ops.gather(self.pmaster, Ref(self.idx_d_from_p, h0), Ref(self.packed, nf + h0), n - h0, False, round_tf32=0)
            return
        self.ops.gather(self.pmaster, self.idx_d_from_p, Ref(self.packed, nf), n, False, round_tf32=self.rnd)

    def ensure_flat(self):
        """`flat` (and the optimizer moments in the reference layout) up to date: called by everything that reads
        parameters outside the resident fused steps (state_dict, the piecewise API path, checkpoints, epoch ends)."""
        if self.flat_valid:
            return
        ops, opt = self.ops, self._resident_opt
        ops.gather(self.pmaster, self.inv_idx, self.flat, self.n_flat, True)
        if opt is not None and getattr(opt, "mp", None) is not None:
            ops.gather(opt.mp, self.inv_idx, opt.m, self.n_flat, True)
            ops.gather(opt.vp, self.inv_idx, opt.v, self.n_flat, True)
        self.flat_valid = True

    def wref32(self, g: GemmW) -> Ref:  # scrubber heads: fp32 FFMA kernels in every precision mode
        return Ref(self.packed, g.w)

    def wdref32(self, g: GemmW) -> Ref:
        return Ref(self.packed, self._n_fwd + g.wd)

    def gwref(self, g: GemmW) -> Ref:
        return Ref(self.gpacked, g.w)

    def gbref(self, g: GemmW) -> Optional[Ref]:
        return None if g.b is None else Ref(self.gpacked, g.b)

    def repack(self, part: str = "all"):
        """Refreshes the packed GEMM matrices from the flat parameters: "fwd" = forward matrices + biases,
        "dgrad" = data-gradient matrices (only the backward pass reads them), "all" = both."""
        n0 = 0 if part in ("all", "fwd") else self._n_fwd
        n1 = self._n_fwd if part == "fwd" else self.packed.numel()
        if n1 <= n0:
            return
        if self.packed16 is not None:
            self.ops.gather(self.flat, Ref(self.pack_idx, n0), Ref(self.packed16, n0), n1 - n0, False, round_tf32=2)
            if n0 == 0:  # biases (and the scrubber heads' fp32 weights) live in the fp32 copy of the forward part
                self.ops.gather(self.flat, self.pack_idx, self.packed, self._n_fwd, False, round_tf32=0)
            h0 = self._n_fwd + self._d_heads0
            if n1 > h0:  # fp32 data-gradient matrices of the scrubber heads
                self.ops.gather(self.flat, Ref(self.pack_idx, h0), Ref(self.packed, h0), n1 - h0, False, round_tf32=0)
            return
        self.ops.gather(self.flat, Ref(self.pack_idx, n0), Ref(self.packed, n0), n1 - n0, False, round_tf32=self.rnd)

    # ------------------------------------------------------------------ API
    def plan(self, B: int) -> "Plan":
        p = self.plans.get(B)
        if p is None:
            p = Plan(self, B)
            self.plans[B] = p
        return p

    def forward(self, data, training: bool):
        B = data["x6d"].shape[0]
        return self.plan(B).forward(data, training)

    # ---- dp_bn "sync": BatchNorm statistics over the global batch (set by parallel.setup(..., bn_sync=True))
    def bn_sync_world(self) -> int:
        grp = self.__dict__.get("_bn_sync")
        return grp[0] if grp else 1

    def bn_sync_reduce(self, t):
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._bn_sync[1])

    def encode(self, data, training: bool):
        B = data["x6d"].shape[0]
        return self.plan(B).forward(data, training, upto="encode")

    def decode(self, z, data, training: bool):
        B = z.shape[0]
        return self.plan(B).forward(data, training, z_given=z)


def _chain_tree_ok(tree, J) -> bool:
    """Host twin of chain_tree_ok + the start-joint check of csrc/scv_loss.cu: can the lane-per-chain FK kernel serve this
    kinematic tree (<= 8 chains of 1..5 joints, every chain starting at joint 0 or at a joint an earlier chain placed)?"""
    if len(tree) > 8 or J > 32:
        return False
    placed = set()
    for chain in tree:
        if not 1 <= len(chain) <= 5:
            return False
        if chain[0] != 0 and chain[0] not in placed:
            return False
        placed.update(int(j) for j in chain[1:])
    return True


class Plan:
    """Static buffers + launch lists for one batch size."""

    def __init__(self, eng: Engine, B: int):
        self.eng, self.B = eng, B
        m, dev, ops = eng.m, eng.device, eng.ops
        ch, k, z, W = m.ch, m.kernel, m.z_dim, m.window
        self.J = (m.in_channels - 3) // 6
        self.nx = self.J * 6
        C0 = eng.C0
        prec = eng.precision
        rnd = eng.rnd
        rmode = {0: 0, 1: ROUND_TF32, 2: OUT_BF16}[rnd]
        bf16 = rnd == 2
        slack = eng.max_k + 64
        p2 = k // 2
        f32 = dict(device=dev, dtype=torch.float32)
        opd = dict(device=dev, dtype=eng.op_dtype)
        A = lambda L, C, hl=0, hr=0, even=False: Act(dev, B, L, C, hl, hr, even, slack)  # noqa: E731  fp32: GEMM outputs
        # GEMM OPERAND buffers (written by the elementwise producers, read by the GEMMs as A / dY): bf16 in bf16 mode
        Ao = lambda L, C, hl=0, hr=0, even=False: Act(dev, B, L, C, hl, hr, even, slack, eng.op_dtype)  # noqa: E731

        # ---- static inputs
        self.inp = {
            "x6d": torch.zeros(B, W, self.J, 6, **f32), "root": torch.zeros(B, W, 3, **f32),
            "offsets": torch.zeros(B, W, self.J, 3, **f32), "target_pose": torch.zeros(B, W, self.J, 3, **f32),
        }
        self.var = torch.zeros(B, max(1, eng.cond_dim), **f32)
        self.eps = torch.zeros(B, z, **f32)
        self.gr_target: Dict[str, torch.Tensor] = {}
        self.gr_labels: Dict[str, torch.Tensor] = {}
        self.gr_dim: Dict[str, int] = {}
        for key in eng.gr_keys:
            dimk = eng.gr_layers[key][0][-1].N  # padded
            true_d = dict(m.named_parameters())[f"disentangle.grad_reversal.{key}.reversal.1.mlp1.4.bias"].numel()
            self.gr_dim[key] = true_d
            if key == "ids":
                self.gr_labels[key] = torch.zeros(B, dtype=torch.long, device=dev)
            else:
                self.gr_target[key] = torch.zeros(B, true_d, **f32)
        if isinstance(m.kinematic_tree, (list, tuple)) and len(m.kinematic_tree) > 0:
            tr = [len(m.kinematic_tree)]
            for chain in m.kinematic_tree:
                tr += [len(chain)] + [int(j) for j in chain]
            self.tree = torch.tensor(tr, dtype=torch.int32, device=dev)
            self.tree_kind = 1 if _chain_tree_ok(m.kinematic_tree, self.J) else 2  # which scv_recon_loss kernel serves it
        else:
            self.tree = None
            self.tree_kind = 0
        arena = m.arena_size

        # ---- loss bookkeeping: acc (double) / out (float): [jpe, root, prior, gr keys...]
        self.loss_names = ["jpe", "root", "prior"] + [kk + "_gr" for kk in eng.gr_keys]
        if eng.cond_dim > 0:
            self.loss_names.append("mcmi")  # kernel mutual-information scrubbing loss (needs conditioning variables)
        self.loss_names += ([kk + "_mals" for kk in eng.mals_keys] + [kk + "_qda" for kk in eng.qda_keys] +
                            [kk + "_ma" for kk in eng.ma_keys] + [kk + "_lsq" for kk in eng.dlsq])
        self.mi = None  # estimator buffers, allocated by enable_mcmi()
        nl = len(self.loss_names)
        nbn = sum(mod.num_features for mod in m.modules() if isinstance(mod, nn.BatchNorm1d))
        n_stats, n_sums = 4 * nbn + 8, 2 * nbn + 2 * ch[0] + 64
        # every per-step accumulator in ONE buffer: [BN statistics | BN backward sums | loss terms | grad norm^2];
        # the fused step clears it with one memset, the piecewise API path clears the parts as it reaches them
        n_qda = sum(4 * len(eng.qda_mod[kk].classes) for kk in eng.qda_keys)  # (lla, llb, llra, llrb) per class
        n_mals = 2 * len(eng.mals_keys) + n_qda + 2 * len(eng.dlsq)  # + squared-error sums of the least-squares decoders
        self.zbuf = torch.zeros(n_stats + n_sums + nl + 1 + n_mals, dtype=torch.double, device=dev)
        self.loss_acc = self.zbuf[n_stats + n_sums:n_stats + n_sums + nl]
        self.sumsq = self.zbuf[n_stats + n_sums + nl:n_stats + n_sums + nl + 1]
        self.mals_l01 = self.zbuf[n_stats + n_sums + nl + 1:]
        self.qda = {}
        off = 2 * len(eng.mals_keys)
        for key in eng.qda_keys:
            mod = eng.qda_mod[key]
            ncq = len(mod.classes)
            self.qda[key] = dict(mod=mod, nc=ncq, off=off, SinvT=torch.zeros(4, ncq, z, z, **f32),
                                 logdet=torch.zeros(4, ncq, **f32), y=torch.zeros(B, dtype=torch.long, device=dev),
                                 classes=torch.tensor([int(c) for c in mod.classes], dtype=torch.long, device=dev),
                                 stat=torch.zeros(2 * ncq * (z + 1), **f32))
            off += 4 * ncq
        self.dlsq = {}
        off_l = 2 * len(eng.mals_keys) + n_qda
        for key, nyd in eng.dlsq.items():
            nxm = z + 1  # room for the bias column (chosen per call by the sign of the loss scale, as the reference)
            self.dlsq[key] = dict(ny=nyd, off=off_l, bias=False, S=[torch.zeros(nxm * nxm, **f32) for _ in range(2)],
                                  Sy=[torch.zeros(nxm * nyd, **f32) for _ in range(2)],
                                  W=[torch.zeros(nxm * nyd, **f32) for _ in range(2)], y=torch.zeros(B, nyd, **f32),
                                  zero=torch.zeros(2, **f32), lam=torch.zeros(2, **f32), gs=torch.zeros(1, **f32))
            off_l += 2
        self.ma = {}
        for key in eng.ma_keys:
            mod = eng.ma_mod[key]
            ncm = len(mod.classes)
            self.ma[key] = dict(mod=mod, nc=ncm, y=torch.zeros(B, dtype=torch.long, device=dev),
                                classes=torch.tensor([int(c) for c in mod.classes], dtype=torch.long, device=dev),
                                stat=torch.zeros(ncm * (z + 1), **f32), coef=torch.zeros(ncm, z, **f32))
        self.mals = {}
        for key in eng.mals_keys:
            mod = eng.mals_mod[key]
            nxm, nym = mod.Sxy0.shape
            self.mals[key] = dict(mod=mod, nx=nxm, ny=nym, bias=bool(mod.bias), W0=torch.zeros(nxm, nym, **f32),
                                  W1=torch.zeros(nxm, nym, **f32), yhat0=torch.zeros(B, nym, **f32),
                                  yhat1=torch.zeros(B, nym, **f32), y=torch.zeros(B, nym, **f32))
        self._fused_tail = False  # True while TrainStep drives the plan (see TrainStep._sequence)
        self.loss_out = torch.zeros(nl + 1, **f32)
        self.anchor = torch.zeros((), device=dev, requires_grad=True)  # autograd attachment point
        self.loss_scale = torch.zeros(nl, **f32)    # total = sum scale*loss
        self.gscale = torch.zeros(nl, **f32)        # d total / d loss_k used by backward
        self._scale_host = None

        # ---- BN statistics / backward sums (double), zeroed once per step
        self.stats = self.zbuf[:n_stats]
        self.sums = self.zbuf[n_stats:n_stats + n_sums]
        self._stats_n = 0
        self._sums_n = 0

        def stats(n):
            o = self._stats_n
            self._stats_n += 2 * n
            assert self._stats_n <= self.stats.numel()
            return Ref(self.stats, o)

        def sums(c):
            o = self._sums_n
            self._sums_n += 2 * c + 1
            assert self._sums_n <= self.sums.numel()
            return Ref(self.sums, o)

        F: List = []   # forward launches (training flag read at run time)
        E: List = []   # encoder part of F ends at this index
        Bk: List = []  # backward launches, appended in FORWARD order as groups; run reversed
        self._training = True

        def bn_mode(has_bn=True, has_act=True):
            return (BN if has_bn else 0) | (PRELU if has_act else 0) | rmode

        def gemm(role=None, **kw):
            """role "side": independent of the launches that follow until the next "join" launch (the skip convolution
            of a residual block runs beside residual.0 -> BN -> PReLU); role "join": consumes the side result."""
            kw.setdefault("precision", prec)
            fn = lambda: ops.gemm(**kw)  # noqa: E731
            F.append(SideLaunch(fn) if role == "side" else AfterSide(fn) if role == "join" else fn)

        # per-channel (scale, shift, mean, rstd) of every BatchNorm, written by its forward pass and read by the data-gradient
        # GEMM whose epilogue carries that layer's backward reduction (scv_gemm_t bnr_*)
        self.chan = torch.zeros(4 * nbn + 16, **f32)
        self._chan_n = 0
        chan_of: Dict[str, Ref] = {}
        fuse_bnr = os.environ.get("SCV_FUSE_BNR", "1") != "0"

        def bnact_fwd(bn_name, slope_name, X, L, Cc, st_off, fold, H=None, U=None):
            kw = dict(X=X.at(0), x_bs=X.bs, x_ls=X.ls, B=B, L=L, Cc=Cc, fold=fold, count=float(B * L), eps=1e-4,
                      momentum=0.1, slope=eng.pref(slope_name) if slope_name else None)
            if bn_name and bn_name not in chan_of:
                chan_of[bn_name] = Ref(self.chan, self._chan_n)
                self._chan_n += 4 * Cc
            if bn_name:
                kw.update(chan_out=chan_of[bn_name])
            if bn_name:
                mod = m.get_submodule(bn_name)
                kw.update(gamma=eng.pref(bn_name + ".weight"), beta=eng.pref(bn_name + ".bias"),
                          running_mean=Ref(mod.running_mean), running_var=Ref(mod.running_var), eps=mod.eps,
                          momentum=mod.momentum)
            if H is not None:
                kw.update(H=H.at(0), h_bs=H.bs, h_ls=H.ls)
            if U is not None:
                kw.update(U=U.at(0), u_bs=U.bs, u_ls=U.ls)
            base = bn_mode(bool(bn_name), bool(slope_name))

            n_st = 2 * Cc * fold

            def run():
                mode = base | (TRAIN if (self._training and bn_name) else 0)
                world = eng.bn_sync_world()
                if world > 1 and (mode & TRAIN):
                    # dp_bn "sync" (SURVEY.md §8e caveat 1): batch statistics over the GLOBAL batch — the layer's column
                    # sums are all-reduced before they are finalised, the count grows by the world size
                    eng.bn_sync_reduce(st_off.t[st_off.off:st_off.off + n_st])
                    ops.bnact_fwd(mode=mode, stats=st_off, **dict(kw, count=kw["count"] * world))
                    return
                ops.bnact_fwd(mode=mode, stats=st_off if bn_name else None, **kw)
            F.append(run)

        def bnr_ctx(bn_name, slope_name, X, Cc, sm_off, ls):
            """scv_gemm_t bnr_* arguments: the GEMM that produces the gradient w.r.t. this layer's output accumulates the
            layer's backward reduction in its epilogue (X and the GEMM output share their row geometry, row stride ls)"""
            if not fuse_bnr:
                return {}
            return dict(bnr_x=X.at(0), bnr_bs=X.bs, bnr_ls=ls, bnr_chan=chan_of[bn_name] if bn_name else None,
                        bnr_slope=eng.pref(slope_name) if slope_name else None, bnr_c=Cc, bnr_sums=sm_off)

        def bnact_bwd(bn_name, slope_name, X, L, Cc, st_off, fold, dO, dU, dX, sm_off, fused=False):
            """returns the backward launches (reduce — unless the producing GEMM's epilogue did it —, apply) for the same
            geometry"""
            fused = fused and fuse_bnr
            mode = bn_mode(bool(bn_name), bool(slope_name)) | (TRAIN if bn_name else 0)
            kw = dict(X=X.at(0), x_bs=X.bs, x_ls=X.ls, B=B, L=L, Cc=Cc, mode=mode, fold=fold, count=float(B * L),
                      eps=1e-4, slope=eng.pref(slope_name) if slope_name else None)
            if bn_name:
                kw.update(gamma=eng.pref(bn_name + ".weight"), beta=eng.pref(bn_name + ".bias"),
                          eps=m.get_submodule(bn_name).eps)
            if dO is not None:
                kw.update(dO=dO.at(0), o_bs=dO.bs, o_ls=dO.ls)
            if dU is not None:
                kw.update(dU=dU.at(0), u_bs=dU.bs, u_ls=dU.ls)
            out = []
            if (mode & 3) and not fused:
                def reduce():
                    world = eng.bn_sync_world()  # dp_bn "sync": `stats` holds the global sums
                    ops.bnact_bwd_reduce(sums=sm_off, stats=st_off if bn_name else None,
                                         **(dict(kw, count=kw["count"] * world) if world > 1 else kw))
                out.append(reduce)
            akw = dict(kw)
            akw.update(dX=dX.at(0) if dX is not None else None, d_bs=dX.bs if dX is not None else 0,
                       d_ls=dX.ls if dX is not None else 0,
                       dgamma=eng.gref(bn_name + ".weight") if bn_name else None,
                       dbeta=eng.gref(bn_name + ".bias") if bn_name else None,
                       dslope=eng.gref(slope_name) if slope_name else None)
            n_sm = 2 * Cc + 1

            def apply():
                world = eng.bn_sync_world()
                if world > 1 and (mode & 3):
                    # the backward reduction over the global batch too; the BatchNorm / PReLU parameter gradients this
                    # launch writes are then global sums on every rank (parallel.GradAllReduce leaves them alone)
                    eng.bn_sync_reduce(sm_off.t[sm_off.off:sm_off.off + n_sm])
                    ops.bnact_bwd_apply(sums=sm_off, stats=st_off if bn_name else None, **dict(akw, count=akw["count"] * world))
                    return
                ops.bnact_bwd_apply(sums=sm_off if (mode & 3) else None, stats=st_off if bn_name else None, **akw)
            out.append(apply)
            return out

        def wgrad(g: GemmW, Aref, a_bs, a_ls, Lo, dY, y_bs, y_ls, bias_n=None):
            kw = dict(A=Aref, a_bs=a_bs, a_ls=a_ls, B=B, Lo=Lo, K=g.K, N=g.N, dY=dY, y_bs=y_bs, y_ls=y_ls,
                      dW=eng.gwref(g), dbias=eng.gbref(g), bias_mod=g.bias_mod,
                      bias_n=(g.N if bias_n is None else bias_n) if g.b is not None else 0, precision=prec)
            return SideLaunch(lambda: ops.wgrad(**kw))

        def dgemm(g: GemmW, Aref, a_bs, a_ls, Lo, Y, y_bs, y_ls, n_last=None, R=None, r_bs=0, r_ls=0, act=ACT_NONE,
                  out_scale=1.0, bnr=None):
            kw = dict(A=Aref, a_bs=a_bs, a_ls=a_ls, B=B, Lo=Lo, K=g.dK, N=g.dN, W=eng.wdref(g), Y=Y, y_bs=y_bs,
                      y_ls=y_ls, n_last=n_last, R=R, r_bs=r_bs, r_ls=r_ls, act=act, out_scale=out_scale,
                      precision=prec, **(bnr or {}))
            return lambda: ops.gemm(**kw)

        WG = eng.W
        # =============================================================== encoder
        X0 = Ao(W, C0, 3, 3)
        self.X0 = X0
        F.append(lambda: ops.pack_input(self.inp["x6d"], self.inp["root"], arena, X0.t, B, W, self.nx, C0, 3,
                                        round_tf32=rnd))
        Y0 = A(W, ch[0])
        g = WG["enc.conv_in"]
        gemm(A=X0.at(-3), a_bs=X0.bs, a_ls=C0, B=B, Lo=W, K=g.K, N=g.N, W=eng.wref(g), bias=eng.bref(g),
             bias_mod=g.bias_mod, bias_n=g.N, Y=Y0.at(0), y_bs=Y0.bs, y_ls=Y0.ls)
        H = Ao(W, ch[0], p2, p2, even=True)
        bnact_fwd(None, "encoder.activation.weight", Y0, W, ch[0], None, 1, H=H)
        enc_in = dict(Y0=Y0, H0=H, g=g)
        L = W
        enc_blocks = []
        for i in range(eng.nblk):
            Ci, Co = ch[i], ch[i + 1]
            Lo = (L + 2 * p2 - k) // 2 + 1
            pre = f"encoder.res_layers.{i}."
            gs, g0, g3 = WG[f"enc.{i}.skip"], WG[f"enc.{i}.r0"], WG[f"enc.{i}.r3"]
            S, R0 = A(Lo, Co), A(Lo, Co // 2)
            gemm(role="side", A=H.at(-p2), a_bs=H.bs, a_ls=2 * Ci, B=B, Lo=Lo, K=gs.K, N=gs.N, W=eng.wref(gs),
                 bias=eng.bref(gs), bias_mod=gs.bias_mod, bias_n=gs.N, Y=S.at(0), y_bs=S.bs, y_ls=S.ls)
            st1 = stats(Co // 2)
            gemm(A=H.at(-p2), a_bs=H.bs, a_ls=2 * Ci, B=B, Lo=Lo, K=g0.K, N=g0.N, W=eng.wref(g0), bias=eng.bref(g0),
                 bias_mod=g0.bias_mod, bias_n=g0.N, Y=R0.at(0), y_bs=R0.bs, y_ls=R0.ls, stats=st1)
            R0a = Ao(Lo, Co // 2, p2, p2)
            bnact_fwd(pre + "residual.1", pre + "residual.2.weight", R0, Lo, Co // 2, st1, 1, H=R0a)
            T = A(Lo, Co)
            st2 = stats(Co)
            gemm(role="join", A=R0a.at(-p2), a_bs=R0a.bs, a_ls=Co // 2, B=B, Lo=Lo, K=g3.K, N=g3.N, W=eng.wref(g3),
                 bias=eng.bref(g3), bias_mod=g3.bias_mod, bias_n=g3.N, Y=T.at(0), y_bs=T.bs, y_ls=T.ls,
                 R=S.at(0), r_bs=S.bs, r_ls=S.ls, stats=st2)
            last = i == eng.nblk - 1
            Hn = Ao(Lo, Co) if last else Ao(Lo, Co, p2, p2, even=True)
            bnact_fwd(pre + "add.0", pre + "add.1.weight", T, Lo, Co, st2, 1, H=Hn)
            enc_blocks.append(dict(pre=pre, Hin=H, Lin=L, Ci=Ci, Co=Co, Lo=Lo, S=S, R0=R0, R0a=R0a, T=T, Hn=Hn,
                                   st1=st1, st2=st2, gs=gs, g0=g0, g3=g3))
            H, L = Hn, Lo
        assert L == eng.Ll
        Hflat = H
        self.ms = torch.zeros(B, eng.ms_ld, **f32)
        gfc = WG["enc.fc"]
        gemm(A=Hflat.at(0), a_bs=Hflat.bs, a_ls=0, B=B, Lo=1, K=gfc.K, N=gfc.N, W=eng.wref(gfc), bias=eng.bref(gfc),
             bias_mod=gfc.bias_mod, bias_n=gfc.N, Y=self.ms, y_bs=eng.ms_ld, y_ls=0)
        self.mu = torch.zeros(B, z, **f32)
        self.Lmat = torch.zeros(B, z, z, **f32)
        self.zc = torch.zeros(B * eng.zc_ld + 64, **opd)[:B * eng.zc_ld].view(B, eng.zc_ld)  # + slab read slack (scv.h)

        def reparam():
            ops.reparam_fwd(self.ms, eng.ms_ld, self.eps if self._training else None,
                            self.var if eng.cond_dim > 0 else None, eng.cond_dim, self.mu, self.Lmat, self.zc,
                            eng.zc_ld, B, z, round_tf32=rnd)
        F.append(reparam)
        for key in eng.mals_keys:  # moving_avg_lsq: decoder weights from the running covariances, predictions from mu
            F.append(lambda key=key: self._mals_forward(key))
        self._n_enc = len(F)

        # =============================================================== decoder
        Ll, Cl = eng.Ll, eng.Cl
        gin = WG["dec.fc_in"]
        H = Ao(Ll, Cl, p2, p2)
        U = Ao(2 * Ll, Cl, p2, p2)
        if bf16:
            # GEMM outputs are fp32: fc_in writes Hraw, the (BN-less, activation-less) pass that builds the upsampled
            # copy also writes the bf16 operand copy H
            Hraw = A(Ll, Cl)
            gemm(A=self.zc, a_bs=eng.zc_ld, a_ls=0, B=B, Lo=1, K=gin.K, N=gin.N, W=eng.wref(gin), bias=eng.bref(gin),
                 bias_mod=gin.bias_mod, bias_n=gin.N, Y=Hraw.at(0), y_bs=Hraw.bs, y_ls=0)
            bnact_fwd(None, None, Hraw, Ll, Cl, None, 1, H=H, U=U)
        else:
            Hraw = H
            gemm(A=self.zc, a_bs=eng.zc_ld, a_ls=0, B=B, Lo=1, K=gin.K, N=gin.N, W=eng.wref(gin), bias=eng.bref(gin),
                 bias_mod=gin.bias_mod, bias_n=gin.N, Y=H.at(0), y_bs=H.bs, y_ls=0,
                 act=ACT_ROUND_TF32 if rnd else ACT_NONE)  # fc_in output feeds the next GEMM directly
            bnact_fwd(None, None, H, Ll, Cl, None, 1, U=U)
        dec_in = dict(H=H, Hraw=Hraw, U=U, g=gin)
        wl, wr = lay.poly_window(k)
        hw = max(wl, wr)
        L = Ll
        dec_blocks = []
        for i in range(eng.nblk):
            Ci, Co = ch[-1 - i], ch[-2 - i]
            pre = f"decoder.res_layers.{i}."
            gs, g0, g3 = WG[f"dec.{i}.skip"], WG[f"dec.{i}.r0"], WG[f"dec.{i}.r3"]
            Lo2 = 2 * L - 1
            S = A(Lo2, Co)
            gemm(role="side", A=U.at(-p2), a_bs=U.bs, a_ls=Ci, B=B, Lo=Lo2, K=gs.K, N=gs.N, W=eng.wref(gs),
                 bias=eng.bref(gs), bias_mod=gs.bias_mod, bias_n=gs.N, Y=S.at(0), y_bs=S.bs, y_ls=S.ls)
            R0 = A(L, Ci // 2)
            st1 = stats(Ci // 2)
            gemm(A=H.at(-p2), a_bs=H.bs, a_ls=Ci, B=B, Lo=L, K=g0.K, N=g0.N, W=eng.wref(g0), bias=eng.bref(g0),
                 bias_mod=g0.bias_mod, bias_n=g0.N, Y=R0.at(0), y_bs=R0.bs, y_ls=R0.ls, stats=st1)
            R0a = Ao(L, Ci // 2, hw, hw)
            bnact_fwd(pre + "residual.1", pre + "residual.2.weight", R0, L, Ci // 2, st1, 1, H=R0a)
            T = A(Lo2, Co)
            st2 = stats(2 * Co)
            gemm(role="join", A=R0a.at(-wl), a_bs=R0a.bs, a_ls=Ci // 2, B=B, Lo=L, K=g3.K, N=g3.N, W=eng.wref(g3),
                 bias=eng.bref(g3), bias_mod=Co, bias_n=2 * Co, Y=T.at(0), y_bs=T.bs, y_ls=2 * Co, n_last=Co,
                 R=S.at(0), r_bs=S.bs, r_ls=2 * Co, stats=st2)
            last = i == eng.nblk - 1
            if last:
                ho = eng.kf - 1 - 3
                Hn, Un = Ao(Lo2, Co, ho, ho), None
            else:
                Hn, Un = Ao(Lo2, Co, p2, p2), Ao(2 * Lo2, Co, p2, p2)
            bnact_fwd(pre + "add.0", pre + "add.1.weight", T, Lo2, Co, st2, 2, H=Hn, U=Un)
            dec_blocks.append(dict(pre=pre, H=H, U=U, L=L, Ci=Ci, Co=Co, Lo2=Lo2, S=S, R0=R0, R0a=R0a, T=T, Hn=Hn,
                                   Un=Un, st1=st1, st2=st2, gs=gs, g0=g0, g3=g3))
            H, U, L = Hn, Un, Lo2
        assert L == eng.l_dec
        gout = WG["dec.conv_out"]
        self.xh = torch.zeros(B * W, C0, **f32)
        ho = eng.kf - 1 - 3
        gemm(A=H.at(-ho), a_bs=H.bs, a_ls=ch[0], B=B, Lo=W, K=gout.K, N=gout.N, W=eng.wref(gout),
             bias=eng.bref(gout), bias_mod=gout.bias_mod, bias_n=gout.N, Y=self.xh, y_bs=W * C0, y_ls=C0,
             act=ACT_TANH)
        self.root_hat = torch.zeros(B, W, 3, **f32)
        F.append(lambda: ops.unpack_root(self.xh, C0, self.nx, arena, self.root_hat, B * W))
        Hlast = H

        # =============================================================== scrubber heads (forward)
        # The MLPs of an ensemble are independent: their same-depth layers run as ONE grouped fp32 launch.  The
        # first-layer outputs (and their gradients) of all members share one [B x sum N] buffer so that the
        # gradient into mu is a single GEMM over the concatenation.
        from ._ops import MAX_GROUP
        self.dmu_gr = torch.zeros(B, z, **f32) if eng.gr_keys else None
        self.gr_act: Dict[str, List[List]] = {}
        self.gr_dact: Dict[str, List[List]] = {}
        self.gr_cat_ld: Dict[str, int] = {}
        self.gr_dact0: Dict[str, torch.Tensor] = {}

        def chunks(lst):
            return [lst[i:i + MAX_GROUP] for i in range(0, len(lst), MAX_GROUP)]

        fwd_levels: Dict[int, List[dict]] = {}
        for key in eng.gr_keys:
            mlps = eng.gr_layers[key]
            ld0 = sum(layers[0].N for layers in mlps)
            self.gr_cat_ld[key] = ld0
            act0 = torch.zeros(B, ld0, **f32)
            dact0 = torch.zeros(B, ld0, **f32)
            self.gr_dact0[key] = dact0
            acts_k, dacts_k = [], []
            col = 0
            for layers in mlps:
                acts, dacts = [], []   # entries: (Ref, row stride)
                hin, h_ld = Ref(self.mu), z
                for li, gl in enumerate(layers):
                    if li == 0:
                        out, dout, o_ld = Ref(act0, col), Ref(dact0, col), ld0
                        col += gl.N
                    else:
                        out, dout, o_ld = Ref(torch.zeros(B, gl.N, **f32)), Ref(torch.zeros(B, gl.N, **f32)), gl.N
                    lastl = li == len(layers) - 1
                    fwd_levels.setdefault(li, []).append(dict(
                        A=hin, a_bs=h_ld, a_ls=0, B=B, Lo=1, K=gl.K, N=gl.N, W=eng.wref32(gl), bias=eng.bref(gl),
                        bias_mod=gl.bias_mod, bias_n=gl.N, Y=out, y_bs=o_ld, y_ls=0,
                        act=ACT_NONE if lastl else ACT_RELU, precision=0))
                    acts.append((out, o_ld))
                    dacts.append((dout, o_ld))
                    hin, h_ld = out, o_ld
                acts_k.append(acts)
                dacts_k.append(dacts)
            self.gr_act[key], self.gr_dact[key] = acts_k, dacts_k
        self._f_heads0 = len(F)
        for li in sorted(fwd_levels):
            for grp in chunks(fwd_levels[li]):
                F.append(lambda grp=grp: ops.gemm_group(grp))
        self._n_head_launches = sum(len(chunks(v)) for v in fwd_levels.values())
        # the heads only read mu: on CUDA they run on a side stream, concurrently with the decoder (forward and backward)
        self.side = torch.cuda.Stream(device=dev) if (dev.type == "cuda" and eng.gr_keys) else None
        self.wside = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None  # weight-gradient stream
        self.dside = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None  # skip-path data-gradient stream

        self.F = F

        # =============================================================== loss launches
        self.dxh = torch.zeros(B * W, C0, **f32)
        Lk: List = []
        Lk.append(lambda: None if self._fused_tail else self.loss_acc.zero_())
        self._loss_has = {"jpe": False, "root": False, "prior": False}

        def recon():
            if self.tree is None:
                raise RuntimeError("scrubvae_b200: the jpe loss needs model.kinematic_tree")
            ops.recon_loss(self.xh, C0, self.inp["offsets"], self.inp["target_pose"], self.inp["root"], arena,
                           self.tree, self.tree.numel(), self.loss_acc, None, self.dxh, B * W, B, self.J,
                           tree_kind=self.tree_kind)
        Lk.append(recon)
        # fused step: the loss scales are known up front, so ONE pass over L gives the KL term and its gradients
        # (the backward launch below is then skipped); the piecewise path learns d total / d prior only in backward
        Lk.append(lambda: ops.kl(self.mu, self.Lmat, Ref(self.loss_acc, 2), Ref(self.loss_scale, 2), self.dmu_kl, self.dL_kl,
                                 B, z) if self._fused_tail else
                  ops.kl(self.mu, self.Lmat, Ref(self.loss_acc, 2), None, None, None, B, z))
        for ki, key in enumerate(eng.gr_keys):
            preds = [a[-1][0] for a in self.gr_act[key]]
            ld = self.gr_act[key][0][-1][1]
            Lk.append(lambda preds=preds, ld=ld, key=key, ki=ki: ops.gr_loss(
                preds, None, ld, self.gr_target.get(key), self.gr_labels.get(key), B, self.gr_dim[key],
                len(eng.gr_keys), Ref(self.loss_acc, 3 + ki), None))
        Lk.append(lambda: self._mi_launch(loss=True))
        for ki, key in enumerate(eng.mals_keys):
            Lk.append(lambda key=key, ki=ki: self._mals_loss(key, ki))
        for key in eng.qda_keys:
            Lk.append(lambda key=key: self._qda_loss(key))
        for key in eng.ma_keys:
            Lk.append(lambda key=key: self._ma_loss(key))
        for key in eng.dlsq:
            Lk.append(lambda key=key: self._dlsq_loss(key))
        Lk.append(lambda: ops.loss_finalize(self.loss_acc, self.loss_scale, self.loss_out, nl))
        self.Lk = Lk

        # =============================================================== backward launches
        Bw: List = []

        def zero_grads():
            if self._fused_tail:
                return  # TrainStep: accumulators cleared by its one memset, gradients zeroed by the optimizer as it reads them
            self.sums.zero_()
            eng.gpacked.zero_()
            eng.gflat.zero_()
            eng.grads_dirty = True  # this path leaves the gradients in place; TrainStep expects zeroed buffers
        Bw.append(zero_grads)
        # conv_out + tanh
        dOut = Ao(W, C0, 3, 3)
        Bw.append(lambda: ops.out_bwd(self.xh, self.dxh, C0, Ref(self.gscale, 0), Ref(self.gscale, 1), self.nx,
                                      dOut.at(0), dOut.bs, dOut.ls, B, W, round_tf32=rnd))
        Bw.append(wgrad(gout, Hlast.at(-ho), Hlast.bs, ch[0], W, dOut.at(0), dOut.bs, dOut.ls))
        dH = A(eng.l_dec, ch[0])
        lastb = dec_blocks[-1]
        sm_next = sums(lastb["Co"])  # the gradient of conv_out goes straight into the last block's add.0 (no skip term)
        Bw.append(dgemm(gout, dOut.at(-3), dOut.bs, C0, eng.l_dec, dH.at(0), dH.bs, dH.ls,
                        bnr=bnr_ctx(lastb["pre"] + "add.0", lastb["pre"] + "add.1.weight", lastb["T"], lastb["Co"], sm_next,
                                    dH.ls)))
        dU = None
        for blk in reversed(dec_blocks):
            Ci, Co, L, Lo2, pre = blk["Ci"], blk["Co"], blk["L"], blk["Lo2"], blk["pre"]
            gs, g0, g3 = blk["gs"], blk["g0"], blk["g3"]
            dT = Ao(Lo2, Co, k - p2, p2 + 1, even=True)
            lst = bnact_bwd(pre + "add.0", pre + "add.1.weight", blk["T"], Lo2, Co, blk["st2"], 2, dH, dU, dT,
                            sm_next if dU is None else sums(Co), fused=dU is None)
            if dU is not None:
                lst[0] = AfterSide(lst[0])  # dU comes from the previous block's skip dgrad on the side stream
            Bw += lst
            # skip conv (k+1, stride 1, pad k//2) on the upsampled input
            Bw.append(wgrad(gs, blk["U"].at(-p2), blk["U"].bs, Ci, Lo2, dT.at(0), dT.bs, dT.ls))
            dUin = A(2 * L, Ci)
            hl_s = (k + 1) - 1 - p2
            Bw.append(SideData(dgemm(gs, dT.at(-hl_s), dT.bs, Co, 2 * L, dUin.at(0), dUin.bs, dUin.ls)))
            # stride-2 transposed conv (polyphase forward; plain strided-window dgrad)
            Bw.append(wgrad(g3, blk["R0a"].at(-wl), blk["R0a"].bs, Ci // 2, L, dT.at(0), dT.bs, 2 * Co,
                            bias_n=2 * Co))
            dR0a = A(L, Ci // 2)
            sm1 = sums(Ci // 2)
            Bw.append(dgemm(g3, dT.at(-p2), dT.bs, 2 * Co, L, dR0a.at(0), dR0a.bs, dR0a.ls,
                            bnr=bnr_ctx(pre + "residual.1", pre + "residual.2.weight", blk["R0"], Ci // 2, sm1, dR0a.ls)))
            dR0 = Ao(L, Ci // 2, p2, p2)
            Bw += bnact_bwd(pre + "residual.1", pre + "residual.2.weight", blk["R0"], L, Ci // 2, blk["st1"], 1,
                            dR0a, None, dR0, sm1, fused=True)
            Bw.append(wgrad(g0, blk["H"].at(-p2), blk["H"].bs, Ci, L, dR0.at(0), dR0.bs, dR0.ls))
            dHin = A(L, Ci)
            Bw.append(dgemm(g0, dR0.at(-p2), dR0.bs, Ci // 2, L, dHin.at(0), dHin.bs, dHin.ls))
            dH, dU = dHin, dUin
        # fc_in (input of decoder block 0 has no BN / activation: only the upsample transpose)
        dX0 = Ao(Ll, Cl)
        lst = bnact_bwd(None, None, dec_in["Hraw"], Ll, Cl, None, 1, dH, dU, dX0, None)
        lst[0] = AfterSide(lst[0])
        Bw += lst
        Bw.append(wgrad(gin, self.zc, eng.zc_ld, 0, 1, dX0.at(0), dX0.bs, 0))
        # 16 output tiles against a 4096-long reduction: split-K, partial tiles added into the zeroed buffer
        self.dzc = torch.zeros(B, eng.zc_ld, **f32)
        Bw.append(lambda: ops.zero(self.dzc))
        Bw.append(dgemm(gin, dX0.at(0), dX0.bs, 0, 1, self.dzc, eng.zc_ld, 0, act=ACT_ACCUM))
        self._bw_heads0 = len(Bw)
        # scrubber heads: dpred from the loss, then level by level from the output side — the weight gradients and
        # the data gradients of all ensemble members at the same distance from the output are one grouped launch
        # each; the gradient-reversed gradient into mu is one GEMM per key over the concatenated first layers
        for ki, key in enumerate(eng.gr_keys):
            preds = [a[-1][0] for a in self.gr_act[key]]
            dpreds = [d[-1][0] for d in self.gr_dact[key]]
            ld = self.gr_act[key][0][-1][1]
            Bw.append(lambda preds=preds, dpreds=dpreds, ld=ld, key=key, ki=ki: ops.gr_loss(
                preds, dpreds, ld, self.gr_target.get(key), self.gr_labels.get(key), B, self.gr_dim[key],
                len(eng.gr_keys), None, Ref(self.gscale, 3 + ki)))
        depth = max((len(l) for kk in eng.gr_keys for l in eng.gr_layers[kk]), default=0)
        for t in range(depth):  # t = distance from the output layer
            wg, dg = [], []
            for key in eng.gr_keys:
                for mi, layers in enumerate(eng.gr_layers[key]):
                    li = len(layers) - 1 - t
                    if li < 0:
                        continue
                    gl = layers[li]
                    acts, dacts = self.gr_act[key][mi], self.gr_dact[key][mi]
                    hin, h_ld = acts[li - 1] if li > 0 else (Ref(self.mu), z)
                    dy, dy_ld = dacts[li]
                    wg.append(dict(A=hin, a_bs=h_ld, a_ls=0, B=B, Lo=1, K=gl.K, N=gl.N, dY=dy, y_bs=dy_ld, y_ls=0,
                                   dW=eng.gwref(gl), dbias=eng.gbref(gl), bias_mod=gl.bias_mod, bias_n=gl.N, precision=0))
                    if li > 0:
                        dprev, dp_ld = dacts[li - 1]
                        dg.append(dict(A=dy, a_bs=dy_ld, a_ls=0, B=B, Lo=1, K=gl.dK, N=gl.dN, W=eng.wdref32(gl), Y=dprev,
                                       y_bs=dp_ld, y_ls=0, R=hin, r_bs=h_ld, r_ls=0, act=ACT_RELUMASK, precision=0))
            for grp in chunks(wg):
                Bw.append(lambda grp=grp: ops.wgrad_group(grp))
            for grp in chunks(dg):
                Bw.append(lambda grp=grp: ops.gemm_group(grp))
        for ki, key in enumerate(eng.gr_keys):  # gradient reversal: dmu_gr (+)= -alpha * dact0_cat . W_cat^T
            gc = eng.gr_cat[key]
            kw2 = dict(A=self.gr_dact0[key], a_bs=self.gr_cat_ld[key], a_ls=0, B=B, Lo=1, K=gc.dK, N=gc.dN,
                       W=eng.wdref32(gc), Y=self.dmu_gr, y_bs=z, y_ls=0, R=None if ki == 0 else self.dmu_gr, r_bs=z, r_ls=0,
                       out_scale=-eng.gr_alpha[key], precision=0)
            Bw.append(lambda kw2=kw2: ops.gemm(**kw2))
        self._bw_dec_end = len(Bw)  # decoder + scrubber-head weight gradients are final from here on
        # latent
        self.dmu_kl = torch.zeros(B, z, **f32)
        self.dL_kl = torch.zeros(B, z, z, **f32)
        self.dms = torch.zeros(B * eng.ms_ld + 64, **opd)[:B * eng.ms_ld].view(B, eng.ms_ld)
        Bw.append(lambda: None if self._fused_tail else
                  ops.kl(self.mu, self.Lmat, None, Ref(self.gscale, 2), self.dmu_kl, self.dL_kl, B, z))
        Bw.append(lambda: self._mi_launch(loss=False))  # adds d mcmi / d mu into dmu_kl
        for ki, key in enumerate(eng.mals_keys):  # adds d <key>_mals / d mu into dmu_kl
            Bw.append(lambda key=key, ki=ki: self._mals_backward(key, ki))
        for key in eng.qda_keys:  # adds d <key>_qda / d mu into dmu_kl
            Bw.append(lambda key=key: self._qda_backward(key))
        for key in eng.ma_keys:  # adds d <key>_ma / d mu into dmu_kl
            Bw.append(lambda key=key: self._ma_backward(key))
        for key in eng.dlsq:  # adds d <key>_lsq / d mu into dmu_kl
            Bw.append(lambda key=key: self._dlsq_backward(key))
        Bw.append(lambda: ops.reparam_bwd(self.ms, eng.ms_ld, self.eps, self.dmu_kl, self.dmu_gr, 1.0, self.dzc,
                                          eng.zc_ld, self.dL_kl, self.dms, eng.ms_ld, B, z, round_tf32=rnd))
        Bw.append(wgrad(gfc, Hflat.at(0), Hflat.bs, 0, 1, self.dms, eng.ms_ld, 0))
        # data parallelism: all-reduce buckets of the encoder's weight gradients, launched as soon as their last wgrad is
        # enqueued — (index in Bw after which gpacked[lo:hi] is final, lo, hi), last layers first
        self._bw_buckets = [(len(Bw), gfc.w, eng.gp_split)]
        dH = A(Ll, Cl)
        laste = enc_blocks[-1]
        sm_next = sums(laste["Co"])
        Bw.append(dgemm(gfc, self.dms, eng.ms_ld, 0, 1, dH.at(0), dH.bs, 0,
                        bnr=bnr_ctx(laste["pre"] + "add.0", laste["pre"] + "add.1.weight", laste["T"], laste["Co"], sm_next, 0)))
        for bi, blk in reversed(list(enumerate(enc_blocks))):
            Ci, Co, Lo, Lin, pre = blk["Ci"], blk["Co"], blk["Lo"], blk["Lin"], blk["pre"]
            gs, g0, g3 = blk["gs"], blk["g0"], blk["g3"]
            dT = Ao(Lo, Co, p2, p2)
            Bw += bnact_bwd(pre + "add.0", pre + "add.1.weight", blk["T"], Lo, Co, blk["st2"], 1, dH, None, dT,
                            sm_next, fused=True)
            # the skip path's data gradient needs only dT: side stream, beside residual.3 dgrad -> BN backward
            dHin = A(blk["Lin"], Ci)
            Lg = (blk["Lin"] + 1) // 2
            nlast = Ci if blk["Lin"] % 2 else 2 * Ci
            Bw.append(SideData(dgemm(gs, dT.at(-wl), dT.bs, Co, Lg, dHin.at(0), dHin.bs, 2 * Ci, n_last=nlast)))
            Bw.append(wgrad(g3, blk["R0a"].at(-p2), blk["R0a"].bs, Co // 2, Lo, dT.at(0), dT.bs, dT.ls))
            dR0a = A(Lo, Co // 2)
            sm1 = sums(Co // 2)
            Bw.append(dgemm(g3, dT.at(-(k - 1 - p2)), dT.bs, Co, Lo, dR0a.at(0), dR0a.bs, dR0a.ls,
                            bnr=bnr_ctx(pre + "residual.1", pre + "residual.2.weight", blk["R0"], Co // 2, sm1, dR0a.ls)))
            dR0 = Ao(Lo, Co // 2, hw, hw)
            Bw += bnact_bwd(pre + "residual.1", pre + "residual.2.weight", blk["R0"], Lo, Co // 2, blk["st1"], 1,
                            dR0a, None, dR0, sm1, fused=True)
            Hin = blk["Hin"]
            Bw.append(wgrad(g0, Hin.at(-p2), Hin.bs, 2 * Ci, Lo, dR0.at(0), dR0.bs, dR0.ls))
            Bw.append(wgrad(gs, Hin.at(-p2), Hin.bs, 2 * Ci, Lo, dT.at(0), dT.bs, dT.ls))
            # the block's input gradient is complete in this GEMM (skip path added as R): its epilogue carries the backward
            # reduction of the layer that produced the input — the previous block's add.0, or the encoder's first PReLU
            sm_next = sums(Ci)
            if bi > 0:
                prev = enc_blocks[bi - 1]
                ctx = bnr_ctx(prev["pre"] + "add.0", prev["pre"] + "add.1.weight", prev["T"], Ci, sm_next, 2 * Ci)
            else:
                ctx = bnr_ctx(None, "encoder.activation.weight", enc_in["Y0"], Ci, sm_next, 2 * Ci)
            Bw.append(AfterSide(dgemm(g0, dR0.at(-wl), dR0.bs, Co // 2, Lg, dHin.at(0), dHin.bs, 2 * Ci, n_last=nlast,
                                      R=dHin.at(0), r_bs=dHin.bs, r_ls=2 * Ci, bnr=ctx)))
            if bi >= eng.nblk - 2:  # the two widest blocks (6.5 M and 1.6 M weights at the default widths); the small rest goes last
                self._bw_buckets.append((len(Bw), gs.w, self._bw_buckets[-1][1]))
            dH = dHin
        dY0 = Ao(W, ch[0])
        Bw += bnact_bwd(None, "encoder.activation.weight", enc_in["Y0"], W, ch[0], None, 1, dH, None, dY0, sm_next, fused=True)
        g = enc_in["g"]
        Bw.append(wgrad(g, X0.at(-3), X0.bs, C0, W, dY0.at(0), dY0.bs, dY0.ls))
        # public-API path only: weight gradients back into the reference layout (p.grad views of gflat)
        Bw.append(lambda: None if self._fused_tail else ops.gather(eng.gpacked, eng.inv_idx, eng.gflat, eng.n_flat, True))
        self.Bw = Bw

    # ------------------------------------------------------------------ moving_avg_lsq (MovingAvgLeastSquares)
    def _mals_forward(self, key):
        """reference model/disentangle.py:466-487 (forward): W_i = solve(Sxx_i + l2, Sxy_i), yhat_i = [mu | 1] W_i"""
        st, ops, z = self.mals[key], self.eng.ops, self.eng.m.z_dim
        mod = st["mod"]
        ops.mals_solve(mod.Sxx0, mod.Sxy0, mod.Sxx1, mod.Sxy1, mod.l2_reg, st["bias"], st["nx"], st["ny"], st["W0"], st["W1"])
        ops.mals_loss(self.mu, z, st["y"], st["ny"], st["W0"], st["W1"], st["bias"], self.B, z, st["ny"],
                      yhat0=st["yhat0"], yhat1=st["yhat1"])

    def _mals_loss(self, key, ki):
        """evaluate_loss (:505-538) / batch_size (train/losses.py:237-245); also moves the forgetting factors"""
        st, ops, z = self.mals[key], self.eng.ops, self.eng.m.z_dim
        mod = st["mod"]
        if not self._fused_tail:
            self.mals_l01[2 * ki:2 * ki + 2].zero_()
        l01 = Ref(self.mals_l01, 2 * ki)
        ops.mals_loss(self.mu, z, st["y"], st["ny"], st["W0"], st["W1"], st["bias"], self.B, z, st["ny"], l01=l01)
        ops.mals_finalize(l01, mod.lam0, mod.lam1, mod.delta, mod.lamdiff, self.B,
                          loss=Ref(self.loss_acc, self.loss_names.index(key + "_mals")))

    def _mals_backward(self, key, ki):
        st, ops, z = self.mals[key], self.eng.ops, self.eng.m.z_dim
        ops.mals_loss(self.mu, z, st["y"], st["ny"], st["W0"], st["W1"], st["bias"], self.B, z, st["ny"],
                      gscale=Ref(self.gscale, self.loss_names.index(key + "_mals")), dmu=self.dmu_kl, d_ld=z)

    def mals_update(self):
        """MovingAvgLeastSquares.update (:489-503) with this step's mu and targets (train/trainer.py:169-178)"""
        ops, z = self.eng.ops, self.eng.m.z_dim
        for key, st in self.mals.items():
            mod = st["mod"]
            ops.mals_update(self.mu, z, st["y"], st["ny"], st["bias"], self.B, z, st["ny"], mod.lam0, mod.lam1, mod.Sxx0,
                            mod.Sxy0, mod.Sxx1, mod.Sxy1)

    # ------------------------------------------------------------------ direct_lsq
    def _dlsq_loss(self, key):
        """direct_lsq_loss (reference train/losses.py:173-179): the batch's own least-squares decoder of y from mu,
        loss = sum (mu W - y)^2 with W = solve(mu^T mu, mu^T y), NOT divided by the batch size (:253-256).  Built from the
        moving_avg_lsq kernels: the covariance update with forgetting factor 0 is exactly mu^T [mu | y]; both decoder
        slots hold the same W, so (l0 + l1) / 2 = the loss."""
        st, ops, z = self.dlsq[key], self.eng.ops, self.eng.m.z_dim
        nx, ny, bias = z + int(st["bias"]), st["ny"], st["bias"]
        ops.mals_update(self.mu, z, st["y"], ny, bias, self.B, z, ny, st["zero"], Ref(st["zero"], 1), st["S"][0], st["Sy"][0],
                        st["S"][1], st["Sy"][1])
        ops.mals_solve(st["S"][0], st["Sy"][0], st["S"][1], st["Sy"][1], 0.0, bias, nx, ny, st["W"][0], st["W"][1])
        if not self._fused_tail:
            self.mals_l01[st["off"]:st["off"] + 2].zero_()
        l01 = Ref(self.mals_l01, st["off"])
        ops.mals_loss(self.mu, z, st["y"], ny, st["W"][0], st["W"][1], bias, self.B, z, ny, l01=l01)
        ops.mals_finalize(l01, st["lam"], Ref(st["lam"], 1), 0.0, 0.0, 1,
                          loss=Ref(self.loss_acc, self.loss_names.index(key + "_lsq")))

    def _dlsq_backward(self, key):
        """d loss / d mu = 2 (mu W - y) W^T: at the least-squares optimum the residual is orthogonal to the columns of mu, so
        the path through W contributes nothing.  scv_mals_loss adds gscale / B * (e0 W0^T + e1 W1^T): gscale is pre-scaled by B."""
        st, ops, z = self.dlsq[key], self.eng.ops, self.eng.m.z_dim
        idx = self.loss_names.index(key + "_lsq")
        torch.mul(self.gscale[idx:idx + 1], float(self.B), out=st["gs"])
        ops.mals_loss(self.mu, z, st["y"], st["ny"], st["W"][0], st["W"][1], st["bias"], self.B, z, st["ny"], gscale=st["gs"],
                      dmu=self.dmu_kl, d_ld=z)

    # ------------------------------------------------------------------ moving_avg (MovingAverageFilter)
    def _ma_loss(self, key):
        """evaluate_loss (reference model/disentangle.py:32-74; train/losses.py:286-289: not divided by the batch size)"""
        st, ops, z = self.ma[key], self.eng.ops, self.eng.m.z_dim
        mod = st["mod"]
        ops.ma_loss(self.mu, z, st["y"], st["classes"], st["nc"], z, self.B, mod.m1, mod.m2, mod.lam1, mod.lam2, mod.delta,
                    mod.lamdiff, st["stat"], st["coef"], loss=Ref(self.loss_acc, self.loss_names.index(key + "_ma")))

    def _ma_backward(self, key):
        st, ops, z = self.ma[key], self.eng.ops, self.eng.m.z_dim
        ops.ma_backward(st["y"], st["classes"], st["coef"], Ref(self.gscale, self.loss_names.index(key + "_ma")), st["nc"], z,
                        self.B, self.dmu_kl, z)

    def ma_update(self):
        """MovingAverageFilter.update (:76-88) with this step's mu and labels (train/trainer.py:169-178)"""
        ops, z = self.eng.ops, self.eng.m.z_dim
        for key, st in self.ma.items():
            mod = st["mod"]
            ops.ma_update(self.mu, z, st["y"], st["classes"], st["nc"], z, self.B, mod.lam1, mod.lam2, mod.m1, mod.m2, st["stat"])

    # ------------------------------------------------------------------ qda (QuadraticDiscriminantFilter)
    def _qda_loss(self, key):
        """evaluate_loss (reference model/disentangle.py:173-232) / batch_size (train/losses.py:247-251)"""
        st, ops, z = self.qda[key], self.eng.ops, self.eng.m.z_dim
        mod = st["mod"]
        acc = Ref(self.mals_l01, st["off"])
        if not self._fused_tail:
            self.mals_l01[st["off"]:st["off"] + 4 * st["nc"]].zero_()
        ops.qda_factor(mod.S4(), st["nc"], z, st["SinvT"], st["logdet"])
        ops.qda_loss(self.mu, z, st["y"], st["classes"], mod.m4(), st["SinvT"], st["logdet"], st["nc"], z, self.B, acc=acc)
        ops.qda_finalize(acc, mod.lama, mod.lamb, mod.delta, mod.lamdiff, st["nc"], self.B,
                         loss=Ref(self.loss_acc, self.loss_names.index(key + "_qda")))

    def _qda_backward(self, key):
        st, ops, z = self.qda[key], self.eng.ops, self.eng.m.z_dim
        mod = st["mod"]
        ops.qda_loss(self.mu, z, st["y"], st["classes"], mod.m4(), st["SinvT"], st["logdet"], st["nc"], z, self.B,
                     gscale=Ref(self.gscale, self.loss_names.index(key + "_qda")), dx=self.dmu_kl, d_ld=z)

    def qda_update(self):
        """QuadraticDiscriminantFilter.update (:133-171) with this step's mu and labels (train/trainer.py:169-178)"""
        ops, z = self.eng.ops, self.eng.m.z_dim
        for key, st in self.qda.items():
            mod = st["mod"]
            ops.qda_update(self.mu, z, st["y"], st["classes"], st["nc"], z, self.B, mod.lama, mod.lamb, mod.m4(), mod.S4(),
                           st["stat"])

    # ------------------------------------------------------------------ mcmi (MutInfoEstimator)
    def enable_mcmi(self, bandwidth: float, var_mode: str = "sphere"):
        """Allocates the estimator's stored-sample buffers (reference model/disentangle.py:234-275): S = this plan's batch."""
        eng, m = self.eng, self.eng.m
        if eng.cond_dim <= 0:
            raise RuntimeError("scrubvae_b200: the mcmi loss needs conditional features (data_o['var'])")
        key = (float(bandwidth), var_mode)
        if self.mi is not None and self.mi["key"] == key:
            return
        if var_mode not in ("sphere", "diagonal"):
            raise ValueError(var_mode)
        f32 = dict(device=eng.device, dtype=torch.float32)
        B, z = self.B, m.z_dim
        self.mi = dict(key=key, bandwidth=float(bandwidth), diag=var_mode == "diagonal", xs=torch.zeros(B, z, **f32),
                       ys=torch.zeros(B, eng.cond_dim, **f32), var_s=torch.zeros(B, z, **f32), logAx=torch.zeros(B, **f32),
                       valid=torch.zeros(1, **f32))

    def _mi_launch(self, loss: bool):
        mi = self.mi
        if mi is None:
            return
        eng, z = self.eng, self.eng.m.z_dim
        idx = self.loss_names.index("mcmi")
        self.eng.ops.mi_loss(self.mu, self.var, self.var.shape[1], mi["xs"], mi["ys"], mi["var_s"] if mi["diag"] else None,
                             mi["logAx"], mi["bandwidth"], self.B, self.B, z, eng.cond_dim, valid=mi["valid"],
                             loss=Ref(self.loss_acc, idx) if loss else None,
                             gscale=None if loss else Ref(self.gscale, idx), dx=None if loss else self.dmu_kl)

    def mi_update(self):
        """Estimator rebuild from the plan's current mu / L / var (call after an encode with the UPDATED weights)."""
        mi, eng = self.mi, self.eng
        eng.ops.mi_update(self.mu, self.Lmat, self.var, self.var.shape[1], mi["xs"], mi["ys"],
                          mi["var_s"] if mi["diag"] else None, mi["logAx"], mi["bandwidth"], self.B, eng.m.z_dim, eng.cond_dim,
                          valid=mi["valid"])

    def mi_set(self, est):
        """Piecewise API path: the estimator object the trainer built (model.mi_estimator) or None."""
        mi = self.mi
        if est is None:
            mi["valid"].zero_()
            return
        mi["xs"].copy_(est.x_s)
        mi["ys"].copy_(est.y_s)
        if mi["diag"]:
            mi["var_s"].copy_(est.var_s)
            mi["logAx"].copy_(est.logA_x.reshape(-1))
        else:
            mi["logAx"][:1].copy_(est.logA_x.reshape(-1)[:1])
        mi["valid"].fill_(1.0)

    # ------------------------------------------------------------------ running
    def load_inputs(self, data, need_loss_inputs: bool, need_cond: bool = True):
        """Copies the batch into the static input buffers (skipped for tensors that already ARE them)."""
        eng, m = self.eng, self.eng.m
        for kname in ("x6d", "root") + (("offsets", "target_pose") if need_loss_inputs else ()):
            if kname in data:
                src = data[kname]
                if src.data_ptr() != self.inp[kname].data_ptr():
                    self.inp[kname].copy_(src.reshape(self.inp[kname].shape), non_blocking=True)
        # encode(data) needs x6d and root only (reference model/residual.py:437-459): the conditioning variables are
        # loaded when the batch carries them
        if eng.cond_dim > 0 and (need_cond or all(kk in data for kk in m.conditional_keys)):
            parts = []
            for kk in m.conditional_keys:
                if kk in m.discrete_classes:
                    parts.append(torch.nn.functional.one_hot(data[kk].ravel().long(),
                                                             len(m.discrete_classes[kk])).to(torch.float32))
                else:
                    parts.append(data[kk].to(torch.float32))
            if len(parts) == 1:
                self.var.copy_(parts[0], non_blocking=True)
            else:
                torch.cat(parts, dim=-1, out=self.var)

    def load_targets(self, data):
        for key, st in self.mals.items():
            st["y"].copy_(data[key].reshape(st["y"].shape), non_blocking=True)
        for key, st in self.qda.items():
            st["y"].copy_(data[key].ravel(), non_blocking=True)
        for key, st in self.ma.items():
            st["y"].copy_(data[key].ravel(), non_blocking=True)
        for key, st in self.dlsq.items():
            st["y"].copy_(data[key].reshape(st["y"].shape), non_blocking=True)
        for key in self.eng.gr_keys:
            if key == "ids":
                self.gr_labels[key].copy_(data[key].ravel(), non_blocking=True)
            else:
                self.gr_target[key].copy_(data[key].reshape(self.gr_target[key].shape), non_blocking=True)

    def forward(self, data, training: bool, upto: Optional[str] = None, z_given=None):
        eng, m = self.eng, self.eng.m
        eng.ensure_flat()  # this path packs the GEMM matrices from `flat`
        self._training = training
        if z_given is None:
            self.load_inputs(data, need_loss_inputs=False, need_cond=upto != "encode")
            if training:
                if m._noise is not None:
                    self.eps.copy_(m._noise)
                else:
                    self.eps.normal_()
                self.stats.zero_()
                if upto == "encode":
                    eng.nbt.add_(eng.nbt_enc)
                else:
                    eng.nbt.add_(1)
            eng.repack()
            self.run_forward(upto=upto)
        else:  # decode(z, data): decoder only
            self.load_inputs(data, need_loss_inputs=False)
            if training:
                self.stats.zero_()
                eng.nbt.add_(eng.nbt_dec)
            eng.repack()
            # zc = [z | var | 0] through the same kernel as the full forward (eps = None: z = "mu" row), so the given
            # latent gets the operand rounding the tensor-core GEMM expects; mu / L outputs go to scratch, never to
            # the buffers an earlier encode() handed out
            if getattr(self, "_dec_scratch", None) is None:
                f32 = dict(device=eng.device, dtype=torch.float32)
                self._dec_scratch = (torch.zeros(self.B, eng.ms_ld, **f32), torch.zeros(self.B, m.z_dim, **f32),
                                     torch.zeros(self.B, m.z_dim, m.z_dim, **f32))
            ms_t, mu_t, L_t = self._dec_scratch
            ms_t[:, :m.z_dim].copy_(z_given)
            eng.ops.reparam_fwd(ms_t, eng.ms_ld, None, self.var if eng.cond_dim > 0 else None, eng.cond_dim, mu_t, L_t,
                                self.zc, eng.zc_ld, self.B, m.z_dim, round_tf32=eng.rnd)
            self.run_forward(decode_only=True)
        out = {}
        if z_given is None:
            out["mu"], out["L"] = self.mu, self.Lmat
            if upto == "encode":
                return out
            out["z"] = self.zc[:, :m.z_dim] if eng.precision != 2 else self.zc[:, :m.z_dim].float()
        if eng.cond_dim > 0:
            out["var"] = self.var
        B, W = self.B, m.window
        out["root"] = self.root_hat
        out["x6d"] = self.xh.view(B, W, -1)[..., :self.nx].unflatten(-1, (self.J, 6))
        if z_given is not None:
            out["_plan"] = self  # eval.generative_restrictiveness reads the decoded rows straight from the plan
        if z_given is None:
            out["disentangle"] = {}
            if eng.gr_keys:
                out["disentangle"]["grad_reversal"] = {
                    key: [a[-1][0].t[:, :self.gr_dim[key]] for a in self.gr_act[key]] for key in eng.gr_keys}
            if eng.mals_keys:
                out["disentangle"]["moving_avg_lsq"] = {key: [st["yhat0"], st["yhat1"]] for key, st in self.mals.items()}
            out["_plan"] = self
        return out

    def set_loss_scale(self, loss_scale: Dict[str, float]):
        for key, st in self.dlsq.items():  # reference train/losses.py:255: bias = loss_scale[key + "_lsq"] < 0
            st["bias"] = float(loss_scale.get(key + "_lsq", 0.0) or 0.0) < 0
        host = tuple(float(loss_scale.get(n, 0.0) or 0.0) for n in self.loss_names)
        if host != self._scale_host:
            self._scale_host = host
            t = torch.tensor(host, dtype=torch.float32)
            self.loss_scale.copy_(t)
        return host

    def loss(self, data, loss_scale):
        self.load_inputs(data, need_loss_inputs=True)
        self.load_targets(data)
        self.set_loss_scale(loss_scale)
        for f in self.Lk:
            f()
        return self.loss_out

    def _run_main(self, fns):
        """Launches in order on the current stream; SideLaunch entries go to the second stream (forked here), the
        matching AfterSide entry waits for them."""
        ws = self.wside if os.environ.get("SCV_SKIP_STREAM", "1") != "0" else None
        if ws is None:
            for f in fns:
                f()
            return
        main = torch.cuda.current_stream()
        pending = False
        for f in fns:
            if isinstance(f, SideLaunch):
                ws.wait_stream(main)
                with torch.cuda.stream(ws):
                    f()
                pending = True
                continue
            if isinstance(f, AfterSide) and pending:
                main.wait_stream(ws)
                pending = False
            f()
        if pending:
            main.wait_stream(ws)

    def run_forward(self, upto=None, decode_only=False):
        """All forward launches; the scrubber heads fork onto the side stream right after the latent is final."""
        if decode_only:
            self._run_main(self.F[self._n_enc:self._f_heads0])
            return
        self._run_main(self.F[:self._n_enc])
        if upto == "encode":
            return
        if self.side is None:
            self._run_main(self.F[self._n_enc:])
            return
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            for f in self.F[self._f_heads0:]:
                f()
        self._run_main(self.F[self._n_enc:self._f_heads0])
        main.wait_stream(self.side)

    def backward(self, comm=None):
        """Runs the backward launch list; `self.gscale` must hold d total / d loss_k.
        `comm(engine, phase)` (data parallelism, parallel.py) is called when the decoder + scrubber-head weight
        gradients are final ("decoder_done": their all-reduce overlaps the encoder backward) and again before
        the final gather of the packed gradients ("encoder_done")."""
        last = len(self.Bw) - 1
        main = torch.cuda.current_stream() if (self.side is not None or self.wside is not None) else None
        wside = self.wside if os.environ.get("SCV_WGRAD_STREAM", "1") != "0" else None
        dside = self.dside if os.environ.get("SCV_SKIP_STREAM", "1") != "0" else None
        dpending = False
        for i, f in enumerate(self.Bw):
            if self.side is not None and i == 1:  # gradients zeroed: the head backward forks off, beside the decoder's
                self.side.wait_stream(main)
                with torch.cuda.stream(self.side):
                    for g in self.Bw[self._bw_heads0:self._bw_dec_end]:
                        g()
            if i == self._bw_dec_end and self.side is not None:
                main.wait_stream(self.side)
            if wside is not None and i == last:
                main.wait_stream(wside)  # the final gather reads the weight gradients
            if comm is not None:
                for at, lo, hi in self._bw_buckets:
                    if at == i:
                        comm(self.eng, "range", wait=[wside] if wside is not None else [], lo=lo, hi=hi)
                if i == self._bw_dec_end:
                    # the all-reduce stream (not the main stream) waits for the decoder's weight gradients
                    comm(self.eng, "decoder_done", wait=[wside] if wside is not None else [])
                if i == last:
                    comm(self.eng, "encoder_done")
            if self.side is not None and self._bw_heads0 <= i < self._bw_dec_end:
                continue  # launched on the side stream above
            if wside is not None and isinstance(f, SideLaunch):
                wside.wait_stream(main)  # its dY is ready at this point of the main stream
                with torch.cuda.stream(wside):
                    f()
                continue
            if dside is not None and isinstance(f, SideData):
                dside.wait_stream(main)
                with torch.cuda.stream(dside):
                    f()
                dpending = True
                continue
            if dpending and isinstance(f, AfterSide):
                main.wait_stream(dside)
                dpending = False
            f()


class TrainStep:
    """One whole training step — forward, losses, backward, grad-norm clip, optimizer — as a single
    launch sequence over a plan's static buffers, optionally replayed from a CUDA graph.  Same kernels
    and order as the public-API path (train/trainer.py: predict_batch -> get_batch_loss -> backward ->
    clip_grad_norm_ -> optimizer.step), minus the Python between them.

    `comm` (optional): callable(engine, phase) hook used by data parallelism (parallel.py) to launch the
    bucketed gradient all-reduce from inside the backward launch list (Plan.backward)."""

    def __init__(self, model, optimizer, loss_scale, B, max_norm=1e6, use_graph=True, comm=None, keep_grads=False,
                 resident=False, mi=None):
        self.model, self.opt = model, optimizer
        # mi = dict(bandwidth=..., var_mode=...): the "mcmi" loss is active — after the optimizer the encoder runs once more
        # with the updated weights and the mutual-information estimator is rebuilt from it (reference trainer.py:184-199)
        self.mi = mi
        # resident=True: the weights' fp32 master stays in the packed GEMM layout between steps (Engine.resident_import /
        # ensure_flat); nn.Parameter values are then refreshed only on demand — call sync() (or Engine.ensure_flat())
        # before reading them, and nothing else may edit the parameters between the steps of one run
        self.resident = resident
        # keep_grads (tests / debugging): the step's (all-reduced, unscaled) gradients are copied out in the reference
        # parameter layout before the optimizer consumes them — see named_grads()
        self.keep_grads = keep_grads
        self.grad_snapshot = None
        self.eng = model.engine
        self.plan = self.eng.plan(B)
        self.max_norm = float(max_norm)
        self.comm = comm
        if comm is not None and hasattr(comm, "steps"):
            comm.steps.add(self)
        if self.mi is not None:
            self.plan.enable_mcmi(self.mi["bandwidth"], self.mi.get("var_mode") or "sphere")
        self.plan.set_loss_scale(loss_scale)
        self.opt._bind()
        self.graph = None
        self.use_graph = use_graph and self.eng.device.type == "cuda"
        self.n_launch = None

    def _sequence(self):
        plan, eng, opt = self.plan, self.eng, self.opt
        m = self.model
        ops = eng.ops
        plan._training = True
        plan._fused_tail = True
        try:
            ops.zero(plan.zbuf)  # BN statistics, BN backward sums, loss terms, grad norm: one memset
            if m._noise is not None:
                plan.eps.copy_(m._noise)
            else:
                plan.eps.normal_()
            eng.nbt.add_(1)
            # packed forward matrices: resident mode — already current (the optimizer wrote them); otherwise gathered from
            # the flat parameters (0.1 ms).  The data-gradient matrices, which only backward reads, and the clearing of the
            # weight-gradient accumulators run on a side stream beside the forward pass.
            res = self.resident
            if not res:
                eng.repack("fwd")

            def side_work():
                if res:
                    eng.resident_dgrad_matrices()
                else:
                    eng.repack("dgrad")
                ops.zero(eng.gpacked)
                ops.zero(eng.gflat[:eng.n_direct])  # BatchNorm / PReLU gradients accumulate
            if plan.dside is not None:
                main = torch.cuda.current_stream()
                plan.dside.wait_stream(main)
                with torch.cuda.stream(plan.dside):
                    side_work()
                plan.run_forward()
                for f in plan.Lk:
                    f()
                main.wait_stream(plan.dside)
            else:
                side_work()
                plan.run_forward()
                for f in plan.Lk:
                    f()
            plan.gscale.copy_(plan.loss_scale)
            plan.backward(self.comm)
            grp = opt.param_groups[0]
            from .train.optim import KIND
            b1 = grp["momentum"] if opt.kind == "sgd" else grp["betas"][0]
            hp = (self.max_norm, opt.grad_scale, float(grp["lr"]), b1, grp["betas"][1], grp["eps"], grp["weight_decay"], 1,
                  KIND[opt.kind])
            if self.keep_grads:  # tests: the step's gradients in the reference layout
                if self.grad_snapshot is None:
                    self.grad_snapshot = torch.zeros_like(eng.gflat)
                self.grad_snapshot.copy_(eng.gflat)
                ops.gather(eng.gpacked, eng.inv_idx, self.grad_snapshot, eng.n_flat, True)
            if res:
                # resident tail: global norm over the packed gradients + the direct ones, then the optimizer in the packed
                # layout (writes master, moments and the operand copies; all coalesced) and on the small direct region
                nf, nd = eng._n_fwd, eng.n_direct
                ops.sumsq_packed(eng.gpacked, None, nf, plan.sumsq, pack_mask=eng.pack_mask)
                ops.sumsq(eng.gflat, nd, plan.sumsq)
                ops.optim_step(eng.pmaster, eng.gpacked, opt.mp, opt.vp, nf, plan.sumsq, *hp, hyper=opt.hyper,
                               pack_mask=eng.pack_mask, packed_out=eng.packed,
                               packed16_out=eng.packed16, round_tf32=eng.rnd == 1)
                ops.optim_step(eng.flat, eng.gflat, opt.m, opt.v, nd, plan.sumsq, *hp, hyper=opt.hyper)
            else:
                # tail: weight gradients back to the parameter layout (coalesced-store gather), global norm, clip +
                # optimizer.  Measured alternatives (profiles/r02_optimizer_tail.md): an optimizer that WRITES the packed
                # matrices with scattered 4-byte stores took 3.8 ms; one that READS the packed gradients by gather 348 us
                # against 110 (gather) + 132 (optimizer) here.
                ops.gather(eng.gpacked, eng.inv_idx, eng.gflat, eng.n_flat, True)
                ops.sumsq(eng.gflat, eng.n_flat, plan.sumsq)
                ops.optim_step(eng.flat, eng.gflat, opt.m, opt.v, eng.n_flat, plan.sumsq, *hp, hyper=opt.hyper)
            if plan.mals:  # running covariances of the moving_avg_lsq scrubbers: this step's mu (train/trainer.py:169-178)
                plan.mals_update()
            if plan.qda:
                plan.qda_update()
            if plan.ma:
                plan.ma_update()
            if self.mi is not None:
                # updated encode (train mode: batch statistics again, running statistics advance a second time, as the
                # reference's model.encode(data) does) -> new stored samples of the estimator
                if not res:
                    eng.repack("fwd")
                ops.zero(plan.stats)
                eng.nbt.add_(eng.nbt_enc)
                plan.run_forward(upto="encode")
                plan.mi_update()
        finally:
            plan._fused_tail = False

    def run(self, data=None):
        """Returns the static loss vector [jpe, root, prior, <gr>..., total] (device, overwritten each step)."""
        plan, opt = self.plan, self.opt
        if data is not None:
            plan.load_inputs(data, need_loss_inputs=True)
            plan.load_targets(data)
        opt._steps += 1
        opt.push_hyper()  # ring of pinned slots: safe when the host runs several steps ahead of the device
        if self.resident:
            if not self.eng.resident_valid or self.eng._resident_opt is not opt:
                self.eng.resident_import(opt)
            self.eng.flat_valid = False  # from here on only the packed master is current
        else:
            self.eng.ensure_flat()
            self.eng.resident_valid = False  # this step updates `flat`
        if not self.use_graph:
            n0 = self.eng.ops.launch_count()
            self._sequence()
            self.n_launch = self.eng.ops.launch_count() - n0
        elif self.graph is None:
            # warm-up once eagerly on a side stream (allocations, lazy init), then capture
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                n0 = self.eng.ops.launch_count()
                self._sequence()
                self.n_launch = self.eng.ops.launch_count() - n0
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            # the eager warm-up WAS this step; capture for the following ones
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._sequence()
            self.graph = g
        else:
            self.graph.replay()
        return plan.loss_out

    def sync(self):
        """nn.Parameter values (and the optimizer moments in the reference layout) up to date after resident steps."""
        self.eng.ensure_flat()

    def named_grads(self):
        """{parameter name: gradient of the last step} (needs keep_grads=True); under data parallelism the SUM over ranks."""
        if self.grad_snapshot is None:
            raise RuntimeError("TrainStep(keep_grads=True) is required to read the step's gradients")
        eng = self.eng
        return {n: self.grad_snapshot[eng.poff[n]:eng.poff[n] + p.numel()].view(p.shape)
                for (n, p) in self.model.named_parameters()}

    def losses(self):
        """dict view of the last step's losses (0-d device tensors)."""
        v = self.plan.loss_out
        d = {n: v[i] for i, n in enumerate(self.plan.loss_names)}
        d["total"] = v[len(self.plan.loss_names)]
        return d
