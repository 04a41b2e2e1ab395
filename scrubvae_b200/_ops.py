"""ctypes binding of libscv.so (include/scv.h).  This is the ONLY compute backend of the package:
there is no CPU or PyTorch fallback — if the library is missing or a call fails, an exception is
raised.

Every method takes `Ref`s (tensor + element offset) for device buffers and enqueues one kernel on
torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libscv.so")

ACT_NONE, ACT_RELU, ACT_TANH, ACT_RELUMASK = 0, 1, 2, 3
ACT_ROUND_TF32 = 16  # OR-ed into act: the stored output is rounded to TF32 (it feeds a tensor-core GEMM)
ACT_ACCUM = 32  # OR-ed into act: Y += result (split-K on the tensor-core path; the caller zeroes Y)
PREC = {"fp32": 0, "tf32": 1, "bf16": 2}
MAX_GROUP = 6  # problems per scv_gemm_group / scv_wgrad_group launch
BN, PRELU, TRAIN, ROUND_TF32, OUT_BF16 = 1, 2, 4, 8, 16  # bnact mode bits


class Ref:
    """A device buffer position: tensor `t` (1-D storage view) + element offset."""
    __slots__ = ("t", "off")

    def __init__(self, t: torch.Tensor, off: int = 0):
        self.t, self.off = t, int(off)

    def __add__(self, k: int) -> "Ref":
        return Ref(self.t, self.off + int(k))


def _ptr(r: Optional[Ref]):
    if r is None:
        return None
    if isinstance(r, torch.Tensor):
        return r.data_ptr()
    return r.t.data_ptr() + r.off * r.t.element_size()


_i64, _f64, _vp = C.c_int64, C.c_double, C.c_void_p


def _struct(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


GemmT = _struct("GemmT", [
    ("A", _vp), ("a_bs", _i64), ("a_ls", _i64), ("B", _i64), ("Lo", _i64), ("K", _i64), ("N", _i64),
    ("W", _vp), ("bias", _vp), ("bias_mod", _i64), ("bias_n", _i64),
    ("Y", _vp), ("y_bs", _i64), ("y_ls", _i64), ("n_last", _i64),
    ("R", _vp), ("r_bs", _i64), ("r_ls", _i64), ("act", _i64), ("out_scale", _f64),
    ("stats", _vp), ("precision", _i64),
    ("bnr_x", _vp), ("bnr_bs", _i64), ("bnr_ls", _i64), ("bnr_chan", _vp), ("bnr_slope", _vp), ("bnr_c", _i64),
    ("bnr_sums", _vp)])
WgradT = _struct("WgradT", [
    ("A", _vp), ("a_bs", _i64), ("a_ls", _i64), ("B", _i64), ("Lo", _i64), ("K", _i64), ("N", _i64),
    ("dY", _vp), ("y_bs", _i64), ("y_ls", _i64), ("dW", _vp), ("dbias", _vp), ("bias_mod", _i64),
    ("bias_n", _i64), ("precision", _i64)])
BnactT = _struct("BnactT", [
    ("X", _vp), ("x_bs", _i64), ("x_ls", _i64), ("B", _i64), ("L", _i64), ("C", _i64),
    ("stats", _vp), ("fold", _i64), ("count", _f64), ("eps", _f64), ("momentum", _f64),
    ("gamma", _vp), ("beta", _vp), ("running_mean", _vp), ("running_var", _vp), ("slope", _vp),
    ("H", _vp), ("h_bs", _i64), ("h_ls", _i64), ("U", _vp), ("u_bs", _i64), ("u_ls", _i64), ("mode", _i64),
    ("chan_out", _vp)])
BnactBwdT = _struct("BnactBwdT", [
    ("X", _vp), ("x_bs", _i64), ("x_ls", _i64), ("B", _i64), ("L", _i64), ("C", _i64),
    ("stats", _vp), ("fold", _i64), ("count", _f64), ("eps", _f64),
    ("gamma", _vp), ("beta", _vp), ("slope", _vp),
    ("dO", _vp), ("o_bs", _i64), ("o_ls", _i64), ("dU", _vp), ("u_bs", _i64), ("u_ls", _i64),
    ("sums", _vp), ("dX", _vp), ("d_bs", _i64), ("d_ls", _i64),
    ("dgamma", _vp), ("dbeta", _vp), ("dslope", _vp), ("mode", _i64), ("chan", _vp)])
OptimT = _struct("OptimT", [
    ("p", _vp), ("g", _vp), ("m", _vp), ("v", _vp), ("n", _i64), ("sumsq", _vp), ("max_norm", _f64),
    ("gscale", _f64), ("lr", _f64), ("beta1", _f64), ("beta2", _f64), ("eps", _f64),
    ("weight_decay", _f64), ("step", _i64), ("hyper", _vp), ("kind", _i64),
    ("pack_idx", _vp), ("packed_out", _vp), ("packed16_out", _vp), ("flags", _i64), ("pack_mask", _vp)])

_SIGS = {
    "scv_version": (C.c_int, []),
    "scv_last_error": (C.c_char_p, []),
    "scv_launch_count": (_i64, []),
    "scv_gemm": (C.c_int, [C.POINTER(GemmT), _vp]),
    "scv_wgrad": (C.c_int, [C.POINTER(WgradT), _vp]),
    "scv_gemm_group": (C.c_int, [C.POINTER(GemmT), _i64, _vp]),
    "scv_wgrad_group": (C.c_int, [C.POINTER(WgradT), _i64, _vp]),
    "scv_pack_input": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp]),
    "scv_bnact_fwd": (C.c_int, [C.POINTER(BnactT), _vp]),
    "scv_bnact_bwd_reduce": (C.c_int, [C.POINTER(BnactBwdT), _vp]),
    "scv_bnact_bwd_apply": (C.c_int, [C.POINTER(BnactBwdT), _vp]),
    "scv_reparam_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "scv_reparam_bwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _f64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "scv_kl": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "scv_recon_loss": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "scv_out_bwd": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "scv_gr_loss": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), _i64, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "scv_gather": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "scv_sumsq": (C.c_int, [_vp, _i64, _vp, _vp]),
    "scv_optim_step": (C.c_int, [C.POINTER(OptimT), _vp]),
    "scv_mi_loss": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _f64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "scv_mi_update": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _f64, _i64, _i64, _i64, _vp, _vp]),
    "scv_qda_factor": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "scv_qda_loss": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp]),
    "scv_qda_finalize": (C.c_int, [_vp, _vp, _vp, _f64, _f64, _i64, _i64, _vp, _vp]),
    "scv_qda_update": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp] + [_vp] * 8 + [_vp, _vp]),
    "scv_ma_loss": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _f64, _f64, _vp, _vp, _vp, _vp]),
    "scv_ma_backward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "scv_ma_update": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scv_mals_solve": (C.c_int, [_vp, _vp, _vp, _vp, _f64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "scv_mals_loss": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "scv_mals_finalize": (C.c_int, [_vp, _vp, _vp, _f64, _f64, _i64, _vp, _vp]),
    "scv_mals_update": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scv_gen_features": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "scv_zero": (C.c_int, [_vp, _i64, _vp]),
    "scv_sumsq_packed": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "scv_d2f": (C.c_int, [_vp, _vp, _i64, _vp]),
    "scv_loss_finalize": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "scv_unpack_root": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _i64, _vp]),
    "scv_window_indices": (C.c_int, [_vp, _i64, _i64, _vp, _vp]),
    "scv_window_features": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scv_preprocess_windows": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
}

EXPORTS = tuple(_SIGS)


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise RuntimeError(
            f"scrubvae_b200: CUDA kernel library not found at {path}. Build it with "
            "`python -m scrubvae_b200.build` (there is no CPU fallback).")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
        fn.restype, fn.argtypes = res, args
    return lib


class CudaOps:
    """Thin call layer over the C ABI; one method per exported kernel entry point."""

    name = "cuda"

    def __init__(self, lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()

    # -- helpers
    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed (rc={rc}): {self.lib.scv_last_error().decode()}")

    def launch_count(self) -> int:
        return int(self.lib.scv_launch_count())

    # -- kernels
    @staticmethod
    def _gemm_struct(A, a_bs, a_ls, B, Lo, K, N, W, Y, y_bs, y_ls, bias=None, bias_mod=1, bias_n=0, n_last=None,
                     R=None, r_bs=0, r_ls=0, act=ACT_NONE, out_scale=1.0, stats=None, precision=0,
                     bnr_x=None, bnr_bs=0, bnr_ls=0, bnr_chan=None, bnr_slope=None, bnr_c=0, bnr_sums=None):
        return GemmT(_ptr(A), a_bs, a_ls, B, Lo, K, N, _ptr(W), _ptr(bias), bias_mod, bias_n, _ptr(Y), y_bs, y_ls,
                     N if n_last is None else n_last, _ptr(R), r_bs, r_ls, act, out_scale, _ptr(stats), precision,
                     _ptr(bnr_x), bnr_bs, bnr_ls, _ptr(bnr_chan), _ptr(bnr_slope), bnr_c, _ptr(bnr_sums))

    @staticmethod
    def _wgrad_struct(A, a_bs, a_ls, B, Lo, K, N, dY, y_bs, y_ls, dW, dbias=None, bias_mod=1, bias_n=0, precision=0):
        return WgradT(_ptr(A), a_bs, a_ls, B, Lo, K, N, _ptr(dY), y_bs, y_ls, _ptr(dW), _ptr(dbias), bias_mod, bias_n,
                      precision)

    def gemm(self, *a, **kw):
        p = self._gemm_struct(*a, **kw)
        self._check(self.lib.scv_gemm(C.byref(p), self._stream()), "scv_gemm")

    def wgrad(self, *a, **kw):
        p = self._wgrad_struct(*a, **kw)
        self._check(self.lib.scv_wgrad(C.byref(p), self._stream()), "scv_wgrad")

    def gemm_group(self, problems: Sequence[dict]):
        """Up to MAX_GROUP independent small problems (keyword dicts of `gemm`) in one fp32 launch."""
        arr = (GemmT * len(problems))(*[self._gemm_struct(**kw) for kw in problems])
        self._check(self.lib.scv_gemm_group(arr, len(problems), self._stream()), "scv_gemm_group")

    def wgrad_group(self, problems: Sequence[dict]):
        arr = (WgradT * len(problems))(*[self._wgrad_struct(**kw) for kw in problems])
        self._check(self.lib.scv_wgrad_group(arr, len(problems), self._stream()), "scv_wgrad_group")

    def pack_input(self, x6d, root, arena, out, B, W, nx, Cc, halo, round_tf32=False):
        self._check(self.lib.scv_pack_input(_ptr(x6d), _ptr(root), _ptr(arena), _ptr(out), B, W, nx, Cc, halo,
                                            int(round_tf32), self._stream()), "scv_pack_input")

    def bnact_fwd(self, X, x_bs, x_ls, B, L, Cc, mode, stats=None, fold=1, count=1.0, eps=1e-4, momentum=0.1,
                  gamma=None, beta=None, running_mean=None, running_var=None, slope=None,
                  H=None, h_bs=0, h_ls=0, U=None, u_bs=0, u_ls=0, chan_out=None):
        p = BnactT(_ptr(X), x_bs, x_ls, B, L, Cc, _ptr(stats), fold, count, eps, momentum, _ptr(gamma), _ptr(beta),
                   _ptr(running_mean), _ptr(running_var), _ptr(slope), _ptr(H), h_bs, h_ls, _ptr(U), u_bs, u_ls, mode,
                   _ptr(chan_out))
        self._check(self.lib.scv_bnact_fwd(C.byref(p), self._stream()), "scv_bnact_fwd")

    def _bwd_struct(self, X, x_bs, x_ls, B, L, Cc, mode, stats, fold, count, eps, gamma, beta, slope, dO, o_bs, o_ls,
                    dU, u_bs, u_ls, sums, dX, d_bs, d_ls, dgamma, dbeta, dslope):
        return BnactBwdT(_ptr(X), x_bs, x_ls, B, L, Cc, _ptr(stats), fold, count, eps, _ptr(gamma), _ptr(beta),
                         _ptr(slope), _ptr(dO), o_bs, o_ls, _ptr(dU), u_bs, u_ls, _ptr(sums), _ptr(dX), d_bs, d_ls,
                         _ptr(dgamma), _ptr(dbeta), _ptr(dslope), mode, None)

    def bnact_bwd_reduce(self, X, x_bs, x_ls, B, L, Cc, mode, sums, stats=None, fold=1, count=1.0, eps=1e-4,
                         gamma=None, beta=None, slope=None, dO=None, o_bs=0, o_ls=0, dU=None, u_bs=0, u_ls=0):
        p = self._bwd_struct(X, x_bs, x_ls, B, L, Cc, mode, stats, fold, count, eps, gamma, beta, slope, dO, o_bs,
                             o_ls, dU, u_bs, u_ls, sums, None, 0, 0, None, None, None)
        self._check(self.lib.scv_bnact_bwd_reduce(C.byref(p), self._stream()), "scv_bnact_bwd_reduce")

    def bnact_bwd_apply(self, X, x_bs, x_ls, B, L, Cc, mode, sums=None, stats=None, fold=1, count=1.0, eps=1e-4,
                        gamma=None, beta=None, slope=None, dO=None, o_bs=0, o_ls=0, dU=None, u_bs=0, u_ls=0,
                        dX=None, d_bs=0, d_ls=0, dgamma=None, dbeta=None, dslope=None):
        p = self._bwd_struct(X, x_bs, x_ls, B, L, Cc, mode, stats, fold, count, eps, gamma, beta, slope, dO, o_bs,
                             o_ls, dU, u_bs, u_ls, sums, dX, d_bs, d_ls, dgamma, dbeta, dslope)
        self._check(self.lib.scv_bnact_bwd_apply(C.byref(p), self._stream()), "scv_bnact_bwd_apply")

    def reparam_fwd(self, ms, ms_ld, eps, var, nvar, mu, L, zc, zc_ld, B, z, round_tf32=False):
        self._check(self.lib.scv_reparam_fwd(_ptr(ms), ms_ld, _ptr(eps), _ptr(var), nvar, _ptr(mu), _ptr(L), _ptr(zc),
                                             zc_ld, B, z, int(round_tf32), self._stream()), "scv_reparam_fwd")

    def reparam_bwd(self, ms, ms_ld, eps, dmu, dmu2, dmu2_scale, dz, dz_ld, dL, dms, dms_ld, B, z, round_tf32=False):
        self._check(self.lib.scv_reparam_bwd(_ptr(ms), ms_ld, _ptr(eps), _ptr(dmu), _ptr(dmu2), float(dmu2_scale),
                                             _ptr(dz), dz_ld, _ptr(dL), _ptr(dms), dms_ld, B, z, int(round_tf32),
                                             self._stream()),
                    "scv_reparam_bwd")

    def kl(self, mu, L, loss, gscale, dmu, dL, B, z):
        self._check(self.lib.scv_kl(_ptr(mu), _ptr(L), _ptr(loss), _ptr(gscale), _ptr(dmu), _ptr(dL), B, z,
                                    self._stream()), "scv_kl")

    def recon_loss(self, xh, ld, offsets, target, root, arena, tree, n_tree, loss, root_hat, dxh, F, B, J, tree_kind=0):
        self._check(self.lib.scv_recon_loss(_ptr(xh), ld, _ptr(offsets), _ptr(target), _ptr(root), _ptr(arena),
                                            _ptr(tree), n_tree, _ptr(loss), _ptr(root_hat), _ptr(dxh), F, B, J,
                                            int(tree_kind), self._stream()), "scv_recon_loss")

    def out_bwd(self, xh, dxh, ld, g_jpe, g_root, nx, draw, d_bs, d_ls, B, W, round_tf32=False):
        self._check(self.lib.scv_out_bwd(_ptr(xh), _ptr(dxh), ld, _ptr(g_jpe), _ptr(g_root), nx, _ptr(draw), d_bs,
                                         d_ls, B, W, int(round_tf32), self._stream()), "scv_out_bwd")

    def gr_loss(self, preds: Sequence[Ref], dpreds: Optional[Sequence[Ref]], ld, target, labels, B, d, num_keys,
                loss, gscale):
        n = len(preds)
        pa = (_vp * n)(*[_ptr(p) for p in preds])
        da = (_vp * n)(*[_ptr(p) for p in dpreds]) if dpreds is not None else None
        self._check(self.lib.scv_gr_loss(pa, da, ld, n, _ptr(target), _ptr(labels), B, d, num_keys, _ptr(loss),
                                         _ptr(gscale), self._stream()), "scv_gr_loss")

    def gather(self, src, idx, dst, n, skip_neg=False, round_tf32=False):
        # round_tf32: 0 plain, 1 round to TF32, 2 dst holds bf16 elements
        self._check(self.lib.scv_gather(_ptr(src), _ptr(idx), _ptr(dst), n,
                                        int(bool(skip_neg)) | (4 if round_tf32 == 2 else 2 if round_tf32 else 0),
                                        self._stream()),
                    "scv_gather")

    def sumsq(self, g, n, out):
        self._check(self.lib.scv_sumsq(_ptr(g), n, _ptr(out), self._stream()), "scv_sumsq")

    def optim_step(self, p, g, m, v, n, sumsq, max_norm, gscale, lr, beta1, beta2, eps, weight_decay, step, kind,
                   hyper=None, pack_idx=None, packed_out=None, packed16_out=None, round_tf32=False, pack_mask=None):
        s = OptimT(_ptr(p), _ptr(g), _ptr(m), _ptr(v), n, _ptr(sumsq), max_norm, gscale, lr, beta1, beta2, eps,
                   weight_decay, step, _ptr(hyper), kind, _ptr(pack_idx), _ptr(packed_out), _ptr(packed16_out),
                   int(bool(round_tf32)), _ptr(pack_mask))
        self._check(self.lib.scv_optim_step(C.byref(s), self._stream()), "scv_optim_step")

    def sumsq_packed(self, gpacked, pack_idx, n, out, pack_mask=None):
        self._check(self.lib.scv_sumsq_packed(_ptr(gpacked), _ptr(pack_idx), _ptr(pack_mask), n, _ptr(out), self._stream()),
                    "scv_sumsq_packed")

    def zero(self, t):
        """cudaMemsetAsync over a whole tensor (a memset node when captured)."""
        self._check(self.lib.scv_zero(t.data_ptr(), t.numel() * t.element_size(), self._stream()), "scv_zero")

    def loss_finalize(self, acc, scale, out, n):
        self._check(self.lib.scv_loss_finalize(_ptr(acc), _ptr(scale), _ptr(out), n, self._stream()),
                    "scv_loss_finalize")

    def unpack_root(self, xh, ld, nx, arena, root_hat, F):
        self._check(self.lib.scv_unpack_root(_ptr(xh), ld, nx, _ptr(arena), _ptr(root_hat), F, self._stream()),
                    "scv_unpack_root")

    def window_indices(self, starts, n_w, window, winds):
        self._check(self.lib.scv_window_indices(_ptr(starts), n_w, window, _ptr(winds), self._stream()),
                    "scv_window_indices")

    def window_features(self, pose, starts, n_w, window, J, parts, speed, avg3, heading, yaw):
        self._check(self.lib.scv_window_features(_ptr(pose), _ptr(starts), n_w, window, J, _ptr(parts), _ptr(speed),
                                                 _ptr(avg3), _ptr(heading), _ptr(yaw), self._stream()),
                    "scv_window_features")

    def preprocess_windows(self, pose, starts, keep, n_keep, window, J, tree, offset, yaw, mode, x6d, root, offsets,
                           target_pose):
        self._check(self.lib.scv_preprocess_windows(_ptr(pose), _ptr(starts), _ptr(keep), n_keep, window, J, _ptr(tree),
                                                    _ptr(offset), _ptr(yaw), mode, _ptr(x6d), _ptr(root), _ptr(offsets),
                                                    _ptr(target_pose), self._stream()), "scv_preprocess_windows")

    def mi_loss(self, x, y, y_ld, xs, ys, var_s, logAx, bandwidth, S, B, z, dy, valid=None, loss=None, gscale=None, dx=None):
        self._check(self.lib.scv_mi_loss(_ptr(x), _ptr(y), y_ld, _ptr(xs), _ptr(ys), _ptr(var_s), _ptr(logAx), float(bandwidth),
                                         S, B, z, dy, _ptr(valid), _ptr(loss), _ptr(gscale), _ptr(dx), self._stream()),
                    "scv_mi_loss")

    def mi_update(self, mu, L, var, var_ld, xs, ys, var_s, logAx, bandwidth, S, z, dy, valid=None):
        self._check(self.lib.scv_mi_update(_ptr(mu), _ptr(L), _ptr(var), var_ld, _ptr(xs), _ptr(ys), _ptr(var_s), _ptr(logAx),
                                           float(bandwidth), S, z, dy, _ptr(valid), self._stream()), "scv_mi_update")

    def qda_factor(self, S4, nc, z, SinvT, logdet):
        self._check(self.lib.scv_qda_factor(*[_ptr(t) for t in S4], nc, z, _ptr(SinvT), _ptr(logdet), self._stream()),
                    "scv_qda_factor")

    def qda_loss(self, x, x_ld, y, classes, m4, SinvT, logdet, nc, z, B, acc=None, gscale=None, dx=None, d_ld=0):
        self._check(self.lib.scv_qda_loss(_ptr(x), x_ld, _ptr(y), _ptr(classes), *[_ptr(t) for t in m4], _ptr(SinvT),
                                          _ptr(logdet), nc, z, B, _ptr(acc), _ptr(gscale), _ptr(dx), d_ld, self._stream()),
                    "scv_qda_loss")

    def qda_finalize(self, acc, lama, lamb, delta, lamdiff, nc, B, loss=None):
        self._check(self.lib.scv_qda_finalize(_ptr(acc), _ptr(lama), _ptr(lamb), float(delta), float(lamdiff), nc, B, _ptr(loss),
                                              self._stream()), "scv_qda_finalize")

    def qda_update(self, x, x_ld, y, classes, nc, z, B, lama, lamb, m4, S4, stat):
        self._check(self.lib.scv_qda_update(_ptr(x), x_ld, _ptr(y), _ptr(classes), nc, z, B, _ptr(lama), _ptr(lamb),
                                            *[_ptr(t) for t in m4], *[_ptr(t) for t in S4], _ptr(stat), self._stream()),
                    "scv_qda_update")

    def ma_loss(self, x, x_ld, y, classes, nc, z, B, m1, m2, lam1, lam2, delta, lamdiff, stat, coef, loss=None):
        self._check(self.lib.scv_ma_loss(_ptr(x), x_ld, _ptr(y), _ptr(classes), nc, z, B, _ptr(m1), _ptr(m2), _ptr(lam1),
                                         _ptr(lam2), float(delta), float(lamdiff), _ptr(stat), _ptr(coef), _ptr(loss),
                                         self._stream()), "scv_ma_loss")

    def ma_backward(self, y, classes, coef, gscale, nc, z, B, dx, d_ld):
        self._check(self.lib.scv_ma_backward(_ptr(y), _ptr(classes), _ptr(coef), _ptr(gscale), nc, z, B, _ptr(dx), d_ld,
                                             self._stream()), "scv_ma_backward")

    def ma_update(self, x, x_ld, y, classes, nc, z, B, lam1, lam2, m1, m2, stat):
        self._check(self.lib.scv_ma_update(_ptr(x), x_ld, _ptr(y), _ptr(classes), nc, z, B, _ptr(lam1), _ptr(lam2), _ptr(m1),
                                           _ptr(m2), _ptr(stat), self._stream()), "scv_ma_update")

    def mals_solve(self, Sxx0, Sxy0, Sxx1, Sxy1, l2_reg, bias, nx, ny, W0, W1):
        self._check(self.lib.scv_mals_solve(_ptr(Sxx0), _ptr(Sxy0), _ptr(Sxx1), _ptr(Sxy1), float(l2_reg), int(bias), nx, ny,
                                            _ptr(W0), _ptr(W1), self._stream()), "scv_mals_solve")

    def mals_loss(self, mu, mu_ld, y, y_ld, W0, W1, bias, B, z, ny, l01=None, yhat0=None, yhat1=None, gscale=None, dmu=None,
                  d_ld=0):
        self._check(self.lib.scv_mals_loss(_ptr(mu), mu_ld, _ptr(y), y_ld, _ptr(W0), _ptr(W1), int(bias), B, z, ny, _ptr(l01),
                                           _ptr(yhat0), _ptr(yhat1), _ptr(gscale), _ptr(dmu), d_ld, self._stream()),
                    "scv_mals_loss")

    def mals_finalize(self, l01, lam0, lam1, delta, lamdiff, B, loss=None):
        self._check(self.lib.scv_mals_finalize(_ptr(l01), _ptr(lam0), _ptr(lam1), float(delta), float(lamdiff), B, _ptr(loss),
                                               self._stream()), "scv_mals_finalize")

    def mals_update(self, mu, mu_ld, y, y_ld, bias, B, z, ny, lam0, lam1, Sxx0, Sxy0, Sxx1, Sxy1):
        self._check(self.lib.scv_mals_update(_ptr(mu), mu_ld, _ptr(y), y_ld, int(bias), B, z, ny, _ptr(lam0), _ptr(lam1),
                                             _ptr(Sxx0), _ptr(Sxy0), _ptr(Sxx1), _ptr(Sxy1), self._stream()),
                    "scv_mals_update")

    def gen_features(self, xh, ld, root_hat, offsets, tree, n_tree, parts, B, W, J, norm=None, pose_out=None, heading=None,
                     avg3=None):
        self._check(self.lib.scv_gen_features(_ptr(xh), ld, _ptr(root_hat), _ptr(offsets), _ptr(tree), n_tree, _ptr(parts),
                                              B, W, J, _ptr(norm), _ptr(pose_out), _ptr(heading), _ptr(avg3),
                                              self._stream()), "scv_gen_features")

    def d2f(self, src, dst, n):
        self._check(self.lib.scv_d2f(_ptr(src), _ptr(dst), n, self._stream()), "scv_d2f")


_OPS: Optional[CudaOps] = None


def get_ops() -> CudaOps:
    """The process-wide CUDA op table; raises if the library is absent."""
    global _OPS
    if _OPS is None:
        _OPS = CudaOps()
    return _OPS
