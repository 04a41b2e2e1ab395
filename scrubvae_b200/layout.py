"""Weight layouts of the overlapping-row GEMMs, expressed as gather index maps.

The model's parameters keep the reference's shapes (state_dict parity, model/residual.py) and live
in ONE flat fp32 buffer.  Every GEMM of the step wants its weight as a K-major matrix  W[n][k]
whose k index runs over (window tap u, input channel) in the memory order of the halo-padded
channels-last activations.  Instead of one repack kernel per layer type, each packed matrix is
described by an int32 map `idx` with  packed[i] = flat[idx[i]]  (or 0 where idx[i] < 0), so that a
single gather kernel (scv_gather) repacks all layers, and a single gather with the inverse map
scatters the packed weight gradients back into the reference layout.

All functions return LongTensors of flat-parameter indices (-1 = structural zero).
"""
from __future__ import annotations

import torch


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def pad_cols(idx: torch.Tensor, width: int) -> torch.Tensor:
    if idx.shape[-1] == width:
        return idx
    out = torch.full(idx.shape[:-1] + (width,), -1, dtype=torch.long)
    out[..., :idx.shape[-1]] = idx
    return out


def pad_rows(idx: torch.Tensor, rows: int) -> torch.Tensor:
    if idx.shape[0] == rows:
        return idx
    out = torch.full((rows,) + idx.shape[1:], -1, dtype=torch.long)
    out[:idx.shape[0]] = idx
    return out


def poly_window(k: int):
    """Stride-2 transposed conv (kernel k odd, padding k//2) as a stride-1 conv producing the
    (even | odd) output pair of position j from inputs j-wl .. j+wr."""
    p = k // 2
    return p // 2, (p + 1) // 2


# ---- Conv1d weight I[co, ci, t]  (nn.Conv1d: model/residual.py:78-109, :160-170, :198)
def conv_fprop(I: torch.Tensor, cin_pad: int | None = None) -> torch.Tensor:
    """out[o] = sum_t x[s*o - p + t] w[:, :, t]  ->  W[co][t*Cin + ci]"""
    co, ci, k = I.shape
    m = I.permute(0, 2, 1)  # (co, k, ci)
    if cin_pad is not None and cin_pad != ci:
        m = pad_cols(m, cin_pad)
    return m.reshape(co, -1)


def conv_dgrad_s1(I: torch.Tensor) -> torch.Tensor:
    """dx[i] = sum_u dy[i - (k-1-p) + u] w[:, :, k-1-u]  ->  W[ci][u*Cout + co]"""
    co, ci, k = I.shape
    return I.flip(2).permute(1, 2, 0).reshape(ci, k * co)


def conv_dgrad_s2(I: torch.Tensor) -> torch.Tensor:
    """stride-2 conv, padding p = k//2:  dx[2j+r] = sum_u dy[j - wl + u] w[:, :, r + p + 2wl - 2u]
    ->  W[r*Cin + ci][u*Cout + co]"""
    co, ci, k = I.shape
    p = k // 2
    wl, wr = poly_window(k)
    nu = wl + wr + 1
    out = torch.full((2, ci, nu, co), -1, dtype=torch.long)
    for r in range(2):
        for u in range(nu):
            t = r + p + 2 * wl - 2 * u
            if 0 <= t < k:
                out[r, :, u, :] = I[:, :, t].t()
    return out.reshape(2 * ci, nu * co)


# ---- ConvTranspose1d weight I[ci, co, t]  (model/residual.py:136-158, :286)
def convT_fprop_s1(I: torch.Tensor, pad: int, n_pad: int | None = None) -> torch.Tensor:
    """out[o] = sum_i x[i] w[:, :, o + pad - i]: window x[o - (k-1-pad) + u], tap k-1-u
    ->  W[co][u*Cin + ci]"""
    ci, co, k = I.shape
    m = I.flip(2).permute(1, 2, 0).reshape(co, k * ci)
    return pad_rows(m, n_pad) if n_pad is not None else m


def convT_dgrad(I: torch.Tensor, co_pad: int | None = None) -> torch.Tensor:
    """dx[i] = sum_o dy[o] w[:, :, o + pad - s*i]: window dy[s*i - pad + u], tap u  ->  W[ci][u*Cout + co]"""
    ci, co, k = I.shape
    m = I.permute(0, 2, 1)  # (ci, k, co)
    if co_pad is not None and co_pad != co:
        m = pad_cols(m, co_pad)
    return m.reshape(ci, -1)


def convT_fprop_s2(I: torch.Tensor) -> torch.Tensor:
    """stride-2 transposed conv, padding p: out[2j+r] = sum_u x[j - wl + u] w[:, :, r + p + 2wl - 2u]
    ->  W[r*Cout + co][u*Cin + ci]"""
    ci, co, k = I.shape
    p = k // 2
    wl, wr = poly_window(k)
    nu = wl + wr + 1
    out = torch.full((2, co, nu, ci), -1, dtype=torch.long)
    for r in range(2):
        for u in range(nu):
            t = r + p + 2 * wl - 2 * u
            if 0 <= t < k:
                out[r, :, u, :] = I[:, :, t].t()
    return out.reshape(2 * co, nu * ci)


# ---- Linear weight I[out, in]
def fc_enc_fprop(I: torch.Tensor, Cl: int, Ll: int) -> torch.Tensor:
    """nn.Flatten of (B, Cl, Ll) indexes c*Ll + l (model/residual.py:214,229); ours is l*Cl + c."""
    o = I.shape[0]
    return I.reshape(o, Cl, Ll).permute(0, 2, 1).reshape(o, Ll * Cl)


def fc_dec_fprop(I: torch.Tensor, Cl: int, Ll: int, k_pad: int) -> torch.Tensor:
    """nn.Unflatten(1, (Cl, Ll)) (model/residual.py:265): reference row c*Ll + l -> ours l*Cl + c."""
    i = I.shape[1]
    return pad_cols(I.reshape(Cl, Ll, i).permute(1, 0, 2).reshape(Ll * Cl, i), k_pad)


def fc_dec_bias(Ib: torch.Tensor, Cl: int, Ll: int) -> torch.Tensor:
    return Ib.reshape(Cl, Ll).t().reshape(-1)


def transpose_pad(Wf: torch.Tensor, k_pad: int) -> torch.Tensor:
    """dgrad matrix of a linear layer: W_d[k][n] = W_f[n][k], reduction dim padded."""
    return pad_cols(Wf.t().contiguous(), k_pad)


def inverse_map(packed_idx: torch.Tensor, n_flat: int) -> torch.Tensor:
    """inv[j] = position i with packed_idx[i] == j (each flat weight appears at most once), -1 otherwise."""
    inv = torch.full((n_flat,), -1, dtype=torch.long)
    pos = torch.nonzero(packed_idx >= 0).squeeze(1)
    src = packed_idx[pos]
    if src.numel() != torch.unique(src).numel():
        raise AssertionError("packed forward layout duplicates a parameter")
    inv[src] = pos
    return inv
