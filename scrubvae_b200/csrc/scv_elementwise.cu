// HBM-bound elementwise / reduction kernels of the SC-VAE step: input pack, BatchNorm+PReLU(+upsample)
// forward and backward, gather (weight repack), grad-norm + fused optimizer, small helpers.
// All kernels are float4-vectorised over the contiguous channel dimension (channels-last rows).
#include "scv_common.cuh"

namespace {

constexpr int NT = 256;

struct Tile2D {
  int c;  // float4 column
  bool cok;
  int64_t r0, rstep, rend;
};
// 256 threads = 8 warps; a warp covers cw float4 columns x (32/cw) rows, so that narrow layers
// (C = 64 -> 16 float4) still issue full 512-byte warp requests.
__device__ __forceinline__ Tile2D make_tile(int C4, int cw, int64_t rows, int64_t rows_per_block) {
  Tile2D t;
  int lane = threadIdx.x & 31, ly = threadIdx.x >> 5;
  int rpw = 32 / cw, lx = lane % cw, rsub = lane / cw;
  t.c = blockIdx.x * cw + lx;
  t.cok = t.c < C4;
  int64_t rbeg = (int64_t)blockIdx.y * rows_per_block;
  t.rend = rbeg + rows_per_block < rows ? rbeg + rows_per_block : rows;
  t.r0 = rbeg + ly * rpw + rsub;
  t.rstep = 8 * rpw;
  return t;
}

struct Launch2D {
  dim3 grid;
  int cw;
  int64_t rows_per_block;
};
Launch2D plan2d(int64_t C4, int64_t rows, int ctas_per_sm) {
  Launch2D l;
  int cw = 1;
  while (cw < 32 && cw < C4) cw <<= 1;
  l.cw = cw;
  int64_t gx = (C4 + cw - 1) / cw;
  int64_t rows_per_iter = 8 * (32 / cw);
  int64_t target = (int64_t)scv::sm_count() * ctas_per_sm;
  int64_t gy = target / gx;
  if (gy < 1) gy = 1;
  int64_t maxgy = (rows + rows_per_iter - 1) / rows_per_iter;
  if (gy > maxgy) gy = maxgy;
  if (gy > 65535) gy = 65535;
  int64_t rpb = (rows + gy - 1) / gy;
  rpb = (rpb + rows_per_iter - 1) / rows_per_iter * rows_per_iter;
  gy = (rows + rpb - 1) / rpb;
  l.grid = dim3((unsigned)gx, (unsigned)gy);
  l.rows_per_block = rpb;
  return l;
}

// (window, position) of a row index, advanced by a fixed row step without dividing: the division by L per row was a fifth
// of these kernels' instructions (ncu: they issue-bound at ~30 % of HBM bandwidth)
struct RowIt {
  uint32_t b, l, sb, sl, L;
  __device__ __forceinline__ void init(int64_t r, int64_t step, uint32_t L_) {
    L = L_;
    b = (uint32_t)(r / L_); l = (uint32_t)(r - (int64_t)b * L_);
    sb = (uint32_t)(step / L_); sl = (uint32_t)(step - (int64_t)sb * L_);
  }
  __device__ __forceinline__ void next() {
    l += sl; b += sb;
    if (l >= L) { l -= L; ++b; }
  }
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float a) { return make_float4(a, a, a, a); }

// per-channel affine of BN (scale, shift) and the statistics needed by backward
struct ChanBN {
  float scale[4], shift[4], mean[4], rstd[4];
};
__device__ __forceinline__ void chan_bn(ChanBN& cb, int ch0, int C, int mode, const double* stats, int fold,
                                        double count, double eps, const float* gamma, const float* beta,
                                        const float* rmean, const float* rvar) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int ch = ch0 + q;
    if (!(mode & 1)) {
      cb.scale[q] = 1.f; cb.shift[q] = 0.f; cb.mean[q] = 0.f; cb.rstd[q] = 1.f;
      continue;
    }
    double mean, var;
    if (mode & 4) {
      double s = 0.0, ss = 0.0;
      for (int j = 0; j < fold; ++j) { s += stats[ch + j * C]; ss += stats[(int64_t)fold * C + ch + j * C]; }
      mean = s / count;
      var = ss / count - mean * mean;
      if (var < 0.0) var = 0.0;
    } else {
      mean = (double)rmean[ch];
      var = (double)rvar[ch];
    }
    double rstd = 1.0 / sqrt(var + eps);
    cb.mean[q] = (float)mean;
    cb.rstd[q] = (float)rstd;
    float g = gamma[ch], b = beta[ch];
    cb.scale[q] = g * (float)rstd;
    cb.shift[q] = b - (float)mean * g * (float)rstd;
  }
}

// The per-channel constants need fp64 divisions and a square root (~1500 instructions for 4 channels): computed ONCE
// per block by the first `cw` threads (one float4 column each) and handed to the rest through shared memory —
// every thread recomputing them cost more instructions than the streaming work itself (ncu, round 1).
__device__ __forceinline__ void chan_bn_block(ChanBN& cb, int cw, int C4, int C, int mode, const double* stats, int fold,
                                              double count, double eps, const float* gamma, const float* beta,
                                              const float* rmean, const float* rvar, const float* chan = nullptr) {
  __shared__ ChanBN sh[32];
  if ((int)threadIdx.x < cw) {
    const int c = blockIdx.x * cw + threadIdx.x;
    if (c < C4) {
      if (chan && (mode & 1)) {  // constants already finalised by the forward pass (scv_bnact_fwd chan_out)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = c * 4 + q;
          sh[threadIdx.x].scale[q] = chan[ch]; sh[threadIdx.x].shift[q] = chan[C + ch];
          sh[threadIdx.x].mean[q] = chan[2 * C + ch]; sh[threadIdx.x].rstd[q] = chan[3 * C + ch];
        }
      } else {
        chan_bn(sh[threadIdx.x], c * 4, C, mode, stats, fold, count, eps, gamma, beta, rmean, rvar);
      }
    }
  }
  __syncthreads();
  cb = sh[(threadIdx.x & 31) % cw];
}

__device__ __forceinline__ float4 bn_prelu(float4 x, const ChanBN& cb, bool has_act, float slope) {
  float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float t = fmaf(v[q], cb.scale[q], cb.shift[q]);
    v[q] = (has_act && t < 0.f) ? slope * t : t;
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

template <int OCC, bool UP>
__global__ void __launch_bounds__(NT, OCC) bnact_fwd_kernel(const scv_bnact_t p, const int cw, const int64_t rpb) {
  const int C = (int)p.C, C4 = C >> 2;
  const int mode = (int)p.mode;
  const int64_t rows = p.B * p.L;
  Tile2D t = make_tile(C4, cw, rows, rpb);
  ChanBN cb;
  chan_bn_block(cb, cw, C4, C, mode, p.stats, (int)p.fold, p.count, p.eps, p.gamma, p.beta, p.running_mean,
                p.running_var);
  if (!t.cok) return;
  const bool has_act = mode & 2;
  const float slope = has_act ? __ldg(p.slope) : 0.f;
  // running statistics: one thread per channel group (block row 0, first row lane)
  if (p.chan_out && blockIdx.y == 0 && t.r0 == 0) {  // the per-channel constants, for fused backward reductions
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ch = t.c * 4 + q;
      p.chan_out[ch] = cb.scale[q]; p.chan_out[C + ch] = cb.shift[q];
      p.chan_out[2 * C + ch] = cb.mean[q]; p.chan_out[3 * C + ch] = cb.rstd[q];
    }
  }
  if ((mode & 5) == 5 && p.running_mean && blockIdx.y == 0 && t.r0 == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int ch = t.c * 4 + q;
      double s = 0.0, ss = 0.0;
      for (int j = 0; j < (int)p.fold; ++j) { s += p.stats[ch + j * C]; ss += p.stats[p.fold * C + ch + j * C]; }
      double mean = s / p.count, var = ss / p.count - mean * mean;
      if (var < 0.0) var = 0.0;
      double unb = p.count > 1.0 ? var * p.count / (p.count - 1.0) : var;
      p.running_mean[ch] = (float)((1.0 - p.momentum) * (double)p.running_mean[ch] + p.momentum * mean);
      p.running_var[ch] = (float)((1.0 - p.momentum) * (double)p.running_var[ch] + p.momentum * unb);
    }
  }
  const uint32_t L = (uint32_t)p.L;
  const int om = (mode & SCV_MODE_OUT_BF16) ? 2 : ((mode & SCV_MODE_ROUND_TF32) ? 1 : 0);
  // 4 rows per iteration, loads first: one thread keeps 4 (12 with the upsample neighbours) float4 loads in flight
  RowIt it;
  it.init(t.r0, t.rstep, L);
  for (int64_t r = t.r0; r < t.rend; r += 4 * (int64_t)t.rstep) {
    float4 xc[4], xm[UP ? 4 : 1], xp[UP ? 4 : 1];
    uint32_t bb[4], ll[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t rr = r + u * (int64_t)t.rstep;
      ok[u] = rr < t.rend;
      bb[u] = it.b;
      ll[u] = it.l;
      it.next();
      if (ok[u]) {
        const float* xr = p.X + (int64_t)bb[u] * p.x_bs + (int64_t)ll[u] * p.x_ls + t.c * 4;
        xc[u] = ld4(xr);
        if (UP && p.U) {
          xm[u] = ll[u] > 0 ? ld4(xr - p.x_ls) : xc[u];
          xp[u] = ll[u] < L - 1 ? ld4(xr + p.x_ls) : xc[u];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const float4 a = bn_prelu(xc[u], cb, has_act, slope);
      if (p.H) scv::store_out4(p.H, (int64_t)bb[u] * p.h_bs + (int64_t)ll[u] * p.h_ls + t.c * 4, a, om);
      if (UP && p.U) {
        const float4 am = bn_prelu(xm[u], cb, has_act, slope), ap = bn_prelu(xp[u], cb, has_act, slope);
        float4 e, o;
        e.x = 0.25f * am.x + 0.75f * a.x; e.y = 0.25f * am.y + 0.75f * a.y;
        e.z = 0.25f * am.z + 0.75f * a.z; e.w = 0.25f * am.w + 0.75f * a.w;
        o.x = 0.75f * a.x + 0.25f * ap.x; o.y = 0.75f * a.y + 0.25f * ap.y;
        o.z = 0.75f * a.z + 0.25f * ap.z; o.w = 0.75f * a.w + 0.25f * ap.w;
        const int64_t ui = (int64_t)bb[u] * p.u_bs + (int64_t)(2 * ll[u]) * p.u_ls + t.c * 4;
        scv::store_out4(p.U, ui, e, om);
        scv::store_out4(p.U, ui + p.u_ls, o, om);
      }
    }
  }
}

// gradient w.r.t. the activated output at (b,l): direct part + transpose of the x2 linear upsample
template <bool UP = true>
__device__ __forceinline__ float4 load_dout(const scv_bnact_bwd_t& p, int64_t b, int64_t l, int c4) {
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.dO) g = ld4(p.dO + b * p.o_bs + l * p.o_ls + c4 * 4);
  if (UP && p.dU) {
    const int64_t L = p.L;
    const float* u = p.dU + b * p.u_bs + c4 * 4;
    float4 e = ld4(u + (2 * l) * p.u_ls), o = ld4(u + (2 * l + 1) * p.u_ls);
    float4 m = ld4(u + (l > 0 ? 2 * l - 1 : 0) * p.u_ls);
    float4 n = ld4(u + (l < L - 1 ? 2 * l + 2 : 2 * L - 1) * p.u_ls);
    g.x += 0.75f * (e.x + o.x) + 0.25f * (m.x + n.x);
    g.y += 0.75f * (e.y + o.y) + 0.25f * (m.y + n.y);
    g.z += 0.75f * (e.z + o.z) + 0.25f * (m.z + n.z);
    g.w += 0.75f * (e.w + o.w) + 0.25f * (m.w + n.w);
  }
  return g;
}

__global__ void __launch_bounds__(NT, 3) bnact_bwd_reduce_kernel(const scv_bnact_bwd_t p, const int cw,
                                                              const int64_t rpb) {
  __shared__ float red[8][32][9];
  __shared__ double shd[32];
  const int C = (int)p.C, C4 = C >> 2;
  const int mode = (int)p.mode;
  const int64_t rows = p.B * p.L, L = p.L;
  Tile2D t = make_tile(C4, cw, rows, rpb);
  float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f}, ds = 0.f;
  ChanBN cb;
  chan_bn_block(cb, cw, C4, C, mode, p.stats, (int)p.fold, p.count, p.eps, p.gamma, p.beta, nullptr, nullptr, p.chan);
  if (t.cok) {
    const bool has_act = mode & 2;
    const float slope = has_act ? __ldg(p.slope) : 0.f;
    const uint32_t L32 = (uint32_t)L;
    RowIt it;
    it.init(t.r0, t.rstep, L32);
    for (int64_t r = t.r0; r < t.rend; r += 4 * (int64_t)t.rstep) {
      float4 xv[4], gv[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {  // loads first
        const int64_t rr = r + u * (int64_t)t.rstep;
        ok[u] = rr < t.rend;
        const uint32_t b = it.b, l = it.l;
        it.next();
        if (ok[u]) {
          xv[u] = ld4(p.X + (int64_t)b * p.x_bs + (int64_t)l * p.x_ls + t.c * 4);
          gv[u] = load_dout(p, b, l, t.c);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const float x[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, g[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v = fmaf(x[q], cb.scale[q], cb.shift[q]);
          float gg = g[q];
          if (has_act && v < 0.f) { ds += gg * v; gg *= slope; }
          sg[q] += gg;
          sgx[q] += gg * (x[q] - cb.mean[q]) * cb.rstd[q];
        }
      }
    }
  }
  // reduce over the 8 warps and the row sub-lanes of each warp
  const int lane = threadIdx.x & 31, ly = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) { red[ly][lane][q] = sg[q]; red[ly][lane][4 + q] = sgx[q]; }
  red[ly][lane][8] = ds;
  __syncthreads();
  const int rpw = 32 / cw;
  if (threadIdx.x < cw) {
    int c = blockIdx.x * cw + threadIdx.x;
    if (c < C4) {
      double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int w = 0; w < 8; ++w)
        for (int s = 0; s < rpw; ++s)
#pragma unroll
          for (int q = 0; q < 8; ++q) a[q] += (double)red[w][s * cw + threadIdx.x][q];
      if (mode & 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          atomicAdd(p.sums + c * 4 + q, a[q]);
          atomicAdd(p.sums + C + c * 4 + q, a[4 + q]);
        }
      }
    }
  }
  if (mode & 2) {
    double d = scv::block_sum_d((double)ds, shd);
    if (threadIdx.x == 0) atomicAdd(p.sums + 2 * C, d);
  }
}

template <int OCC, bool UP>
__global__ void __launch_bounds__(NT, OCC) bnact_bwd_apply_kernel(const scv_bnact_bwd_t p, const int cw,
                                                             const int64_t rpb) {
  const int C = (int)p.C, C4 = C >> 2;
  const int mode = (int)p.mode;
  const int64_t rows = p.B * p.L, L = p.L;
  Tile2D t = make_tile(C4, cw, rows, rpb);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (mode & 2) && p.dslope)
    p.dslope[0] += (float)p.sums[2 * C];
  ChanBN cb;
  chan_bn_block(cb, cw, C4, C, mode, p.stats, (int)p.fold, p.count, p.eps, p.gamma, p.beta, nullptr, nullptr, p.chan);
  if (!t.cok) return;
  const bool has_act = mode & 2;
  const bool train_bn = (mode & 5) == 5;
  const float slope = has_act ? __ldg(p.slope) : 0.f;
  float mg[4] = {0.f, 0.f, 0.f, 0.f}, mgx[4] = {0.f, 0.f, 0.f, 0.f};
  if (mode & 1) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int ch = t.c * 4 + q;
      double s = p.sums[ch], sx = p.sums[C + ch];
      if (train_bn) { mg[q] = (float)(s / p.count); mgx[q] = (float)(sx / p.count); }
      if (blockIdx.y == 0 && t.r0 == 0) {
        if (p.dgamma) p.dgamma[ch] += (float)sx;
        if (p.dbeta) p.dbeta[ch] += (float)s;
      }
    }
  }
  if (!p.dX) return;
  const uint32_t L32 = (uint32_t)L;
  RowIt it;
  it.init(t.r0, t.rstep, L32);
  for (int64_t r = t.r0; r < t.rend; r += 4 * (int64_t)t.rstep) {
    float4 xv[4], gv[4];
    uint32_t bb[4], ll[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // loads first
      const int64_t rr = r + u * (int64_t)t.rstep;
      ok[u] = rr < t.rend;
      bb[u] = it.b;
      ll[u] = it.l;
      it.next();
      if (ok[u]) {
        xv[u] = ld4(p.X + (int64_t)bb[u] * p.x_bs + (int64_t)ll[u] * p.x_ls + t.c * 4);
        gv[u] = load_dout<UP>(p, bb[u], ll[u], t.c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const float x[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, g[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
      float d[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = fmaf(x[q], cb.scale[q], cb.shift[q]);
        float gg = g[q];
        if (has_act && v < 0.f) gg *= slope;
        if (mode & 1) {
          float xh = (x[q] - cb.mean[q]) * cb.rstd[q];
          d[q] = cb.scale[q] * (gg - mg[q] - xh * mgx[q]);
        } else {
          d[q] = gg;
        }
      }
      const float4 dv = make_float4(d[0], d[1], d[2], d[3]);
      scv::store_out4(p.dX, (int64_t)bb[u] * p.d_bs + (int64_t)ll[u] * p.d_ls + t.c * 4, dv,
                      (mode & SCV_MODE_OUT_BF16) ? 2 : ((mode & SCV_MODE_ROUND_TF32) ? 1 : 0));
    }
  }
}

// one thread per 4 output columns (C % 4 == 0): float4 reads of the x6d row when nx % 4 == 0, 32-bit index math
__global__ void __launch_bounds__(NT) pack_input_kernel(const float* __restrict__ x6d, const float* __restrict__ root,
                                                        const float* __restrict__ arena, float* __restrict__ out,
                                                        int64_t rows, int W, int nx, int C, int halo, int rnd) {
  const uint32_t C4 = (uint32_t)C >> 2;
  const int64_t total = rows * C4;
  const bool vec = (nx & 3) == 0;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const uint32_t r = (uint32_t)(i / C4), c = ((uint32_t)i - r * C4) * 4;
    const uint32_t b = r / (uint32_t)W, w = r - b * (uint32_t)W;
    float v[4];
    if (vec && (int)c + 4 <= nx) {
      const float4 t = ld4(x6d + (int64_t)r * nx + c);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cc = (int)c + q;
        v[q] = 0.f;
        if (cc < nx) {
          v[q] = x6d[(int64_t)r * nx + cc];
        } else if (cc < nx + 3) {
          const int d = cc - nx;
          const float a0 = arena[d], a1 = arena[3 + d];
          v[q] = 2.f * (root[(int64_t)r * 3 + d] - a0) / (a1 - a0) - 1.f;
        }
      }
    }
    float4 o = make_float4(v[0], v[1], v[2], v[3]);
    scv::store_out4(out, ((int64_t)b * (W + 2 * halo) + halo + w) * C + c, o, rnd);
  }
}

__global__ void __launch_bounds__(NT) gather_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                    float* __restrict__ dst, int64_t n, int flags) {
  const int om = (flags & SCV_GATHER_OUT_BF16) ? 2 : ((flags & SCV_GATHER_ROUND_TF32) ? 1 : 0);
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
    int32_t j = idx[i];
    if (j >= 0) scv::store_out(dst, i, __ldg(src + j), om);
    else if (!(flags & SCV_GATHER_SKIP_NEG)) scv::store_out(dst, i, 0.f, om);
  }
}

__global__ void __launch_bounds__(NT) sumsq_kernel(const float* __restrict__ g, int64_t n, double* out) {
  __shared__ double sh[32];
  float acc = 0.f;
  double dacc = 0.0;
  int cnt = 0;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n4; i += (int64_t)gridDim.x * NT) {
    float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    if (++cnt == 64) { dacc += (double)acc; acc = 0.f; cnt = 0; }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { float v = g[n4 * 4 + threadIdx.x]; acc += v * v; }
  dacc += (double)acc;
  double s = scv::block_sum_d(dacc, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// liveness of 4 consecutive packed positions starting at float4 index i: from the bitmask (bit j of word w = position
// 32 w + j is live; 1/32 of the index array's traffic) when given, else from the sign of the index entries
__device__ __forceinline__ uint32_t live4(const int32_t* __restrict__ idx, const uint32_t* __restrict__ mask, int64_t i) {
  if (mask) return (__ldg(mask + (i >> 3)) >> (((uint32_t)i & 7u) * 4u)) & 0xFu;
  const int4 j = __ldg(reinterpret_cast<const int4*>(idx) + i);
  return (j.x >= 0 ? 1u : 0u) | (j.y >= 0 ? 2u : 0u) | (j.z >= 0 ? 4u : 0u) | (j.w >= 0 ? 8u : 0u);
}

__global__ void __launch_bounds__(NT) sumsq_packed_kernel(const float* __restrict__ gp, const int32_t* __restrict__ idx,
                                                          const uint32_t* __restrict__ mask, int64_t n, double* out) {
  __shared__ double sh[32];
  float acc = 0.f;
  double dacc = 0.0;
  int cnt = 0;
  const int64_t n4 = n >> 2;  // every packed matrix is padded to 4 floats
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n4; i += (int64_t)gridDim.x * NT) {
    const uint32_t lv = live4(idx, mask, i);
    const float4 v = __ldg(reinterpret_cast<const float4*>(gp) + i);
    acc += ((lv & 1u) ? v.x * v.x : 0.f) + ((lv & 2u) ? v.y * v.y : 0.f) + ((lv & 4u) ? v.z * v.z : 0.f) + ((lv & 8u) ? v.w * v.w : 0.f);
    if (++cnt == 64) { dacc += (double)acc; acc = 0.f; cnt = 0; }
  }
  dacc += (double)acc;
  double s = scv::block_sum_d(dacc, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// resident-packed optimizer: same arithmetic as optim_kernel, every array in the packed layout, plus the operand copies
__global__ void __launch_bounds__(NT) optim_packed_kernel(const scv_optim_t p) {
  const float gs = (float)p.gscale;
  float coef = gs;
  if (p.sumsq) {
    double norm = sqrt(p.sumsq[0]) * p.gscale;
    double c = p.max_norm / (norm + 1e-6);
    coef = (float)((c < 1.0 ? c : 1.0) * p.gscale);
  }
  const double lrd = p.hyper ? p.hyper[0] : p.lr;
  const double stepd = p.hyper ? p.hyper[1] : (double)p.step;
  const float lr = (float)lrd, b1 = (float)p.beta1, b2 = (float)p.beta2, eps = (float)p.eps, wd = (float)p.weight_decay;
  const double bc1d = 1.0 - pow(p.beta1, stepd), bc2d = 1.0 - pow(p.beta2, stepd);
  const float step_size = (float)(lrd / bc1d), bc2s = (float)sqrt(bc2d);
  const bool first = stepd < 1.5;
  const int kind = (int)p.kind;
  const bool rnd = (p.flags & SCV_F_ROUND_TF32) != 0;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(p.packed16_out);
  // four elements per thread, every load issued before the mask is known (padding positions are real memory in
  // all the packed arrays); masked lanes write their old values back, so the stores stay 16-byte vectors
  const int64_t n4 = p.n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n4; i += (int64_t)gridDim.x * NT) {
    const uint32_t lv = live4(p.pack_idx, p.pack_mask, i);
    const float4 g4 = ld4(p.g + 4 * i), w4 = ld4(p.p + 4 * i), m4 = ld4(p.m + 4 * i);
    const float4 v4 = kind == 2 ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(p.v + 4 * i);
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float ww[4] = {w4.x, w4.y, w4.z, w4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!((lv >> q) & 1u)) continue;  // structural zero / padding of the packed layout: left untouched
      float g = gg[q] * coef, w = ww[q];
      if (kind == 2) {
        float buf = first ? g : b1 * mm[q] + g;
        mm[q] = buf;
        ww[q] = w - lr * (g + b1 * buf);
      } else {
        if (kind == 1) w *= 1.f - lr * wd;
        else if (wd != 0.f) g += wd * w;
        float m = mm[q] + (1.f - b1) * (g - mm[q]);
        float v = b2 * vv[q] + (1.f - b2) * g * g;
        mm[q] = m;
        vv[q] = v;
        float denom = sqrtf(v) / bc2s + eps;
        ww[q] = w - step_size * (m / denom);
      }
    }
    st4(p.p + 4 * i, make_float4(ww[0], ww[1], ww[2], ww[3]));
    st4(p.m + 4 * i, make_float4(mm[0], mm[1], mm[2], mm[3]));
    if (kind != 2) st4(p.v + 4 * i, make_float4(vv[0], vv[1], vv[2], vv[3]));
    if (p.packed_out) {
      if (rnd) {
#pragma unroll
        for (int q = 0; q < 4; ++q) ww[q] = scv::round_tf32(ww[q]);
      }
      st4(p.packed_out + 4 * i, make_float4(ww[0], ww[1], ww[2], ww[3]));
    }
    if (o16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(ww[0], ww[1]), hi = __floats2bfloat162_rn(ww[2], ww[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(o16 + 4 * i) = pk;
    }
  }
}

__global__ void __launch_bounds__(NT) optim_kernel(const scv_optim_t p) {
  // clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6))  (train/trainer.py:164)
  const float gs = (float)p.gscale;
  float coef = gs;
  if (p.sumsq) {
    double norm = sqrt(p.sumsq[0]) * p.gscale;
    double c = p.max_norm / (norm + 1e-6);
    coef = (float)((c < 1.0 ? c : 1.0) * p.gscale);
  }
  const double lrd = p.hyper ? p.hyper[0] : p.lr;
  const double stepd = p.hyper ? p.hyper[1] : (double)p.step;
  const float lr = (float)lrd, b1 = (float)p.beta1, b2 = (float)p.beta2, eps = (float)p.eps,
              wd = (float)p.weight_decay;
  const double bc1d = 1.0 - pow(p.beta1, stepd), bc2d = 1.0 - pow(p.beta2, stepd);
  const float step_size = (float)(lrd / bc1d), bc2s = (float)sqrt(bc2d);
  const bool first = stepd < 1.5;
  const int kind = (int)p.kind;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * NT) {
    float g = p.g[i] * coef, w = p.p[i];
    if (kind == 2) {  // SGD, momentum beta1, nesterov (torch.optim.SGD)
      float buf = first ? g : b1 * p.m[i] + g;
      p.m[i] = buf;
      p.p[i] = w - lr * (g + b1 * buf);
      continue;
    }
    if (kind == 1) w *= 1.f - lr * wd;
    else if (wd != 0.f) g += wd * w;
    float m = p.m[i] + (1.f - b1) * (g - p.m[i]);          // torch: exp_avg.lerp_(grad, 1-beta1)
    float v = b2 * p.v[i] + (1.f - b2) * g * g;
    p.m[i] = m;
    p.v[i] = v;
    float denom = sqrtf(v) / bc2s + eps;
    p.p[i] = w - step_size * (m / denom);
  }
}

__global__ void d2f_kernel(const double* in, float* out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

__global__ void loss_finalize_kernel(const double* acc, const float* scale, float* out, int n) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double tot = 0.0;
  for (int i = 0; i < n; ++i) {
    out[i] = (float)acc[i];
    if (scale[i] != 0.f) tot += (double)scale[i] * acc[i];
  }
  out[n] = (float)tot;
}

__global__ void __launch_bounds__(NT) unpack_root_kernel(const float* __restrict__ xh, int64_t ld, int nx,
                                                         const float* __restrict__ arena,
                                                         float* __restrict__ root_hat, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    int64_t f = i / 3;
    int d = (int)(i - f * 3);
    float a0 = arena[d], a1 = arena[3 + d];
    root_hat[i] = 0.5f * (xh[f * ld + nx + d] + 1.f) * (a1 - a0) + a0;
  }
}

int grid1d(int64_t n, int per_sm) {
  int64_t want = (n + NT - 1) / NT;
  int64_t cap = (int64_t)scv::sm_count() * per_sm;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace

extern "C" {

int scv_pack_input(const float* x6d, const float* root, const float* arena, float* out, int64_t B, int64_t W,
                   int64_t nx, int64_t C, int64_t halo, int64_t flags, void* stream) {
  SCV_REQUIRE(nx + 3 <= C, "scv_pack_input: C too small");
  SCV_REQUIRE(C % 4 == 0 && scv::aligned16(out) && scv::aligned16(x6d) && B * W < (1LL << 31),
              "scv_pack_input: C must be a multiple of 4 and the buffers 16-byte aligned");
  pack_input_kernel<<<grid1d(B * W * (C / 4), 8), NT, 0, (cudaStream_t)stream>>>(x6d, root, arena, out, B * W, (int)W,
                                                                          (int)nx, (int)C, (int)halo,
                                                                          (flags & SCV_F_OUT_BF16) ? 2 : (int)(flags & SCV_F_ROUND_TF32));
  return scv::check_launch("pack_input_kernel");
}

static int check_rows(const char* who, int64_t C, const void* ptr, int64_t bs, int64_t ls) {
  SCV_REQUIRE(C % 4 == 0 && bs % 4 == 0 && ls % 4 == 0 && scv::aligned16(ptr), "%s: rows must be float4-aligned", who);
  return 0;
}

// resident CTAs per SM of the BN / PReLU streaming kernels: 3 (80 registers) by default; SCV_BNACT_OCC=4 (64 registers)
static int bnact_occ() {
  const char* e = getenv("SCV_BNACT_OCC");
  return (e && atoi(e) == 4) ? 4 : 3;
}

static bool bnact_nospec() {  // experiments: SCV_BNACT_SPEC=0 runs every launch through the upsample-capable variants
  const char* e = getenv("SCV_BNACT_SPEC");
  return e && atoi(e) == 0;
}

int scv_bnact_fwd(const scv_bnact_t* p, void* stream) {
  if (check_rows("scv_bnact_fwd X", p->C, p->X, p->x_bs, p->x_ls)) return -1;
  if (p->H && check_rows("scv_bnact_fwd H", p->C, p->H, p->h_bs, p->h_ls)) return -1;
  if (p->U && check_rows("scv_bnact_fwd U", p->C, p->U, p->u_bs, p->u_ls)) return -1;
  SCV_REQUIRE(!(p->mode & 1) || (p->gamma && p->beta), "scv_bnact_fwd: BN needs gamma/beta");
  SCV_REQUIRE(!((p->mode & 5) == 5) || p->stats, "scv_bnact_fwd: training BN needs stats");
  SCV_REQUIRE(!((p->mode & 5) == 1) || (p->running_mean && p->running_var), "scv_bnact_fwd: eval BN needs running stats");
  SCV_REQUIRE(p->B * p->L < (1LL << 31), "scv_bnact_fwd: more than 2^31 rows");
  const int occ = bnact_occ();
  Launch2D l = plan2d(p->C / 4, p->B * p->L, occ);
  if (occ == 4) bnact_fwd_kernel<4, true><<<l.grid, NT, 0, (cudaStream_t)stream>>>(*p, l.cw, l.rows_per_block);
  else if (p->U || bnact_nospec()) bnact_fwd_kernel<3, true><<<l.grid, NT, 0, (cudaStream_t)stream>>>(*p, l.cw, l.rows_per_block);
  else bnact_fwd_kernel<3, false><<<l.grid, NT, 0, (cudaStream_t)stream>>>(*p, l.cw, l.rows_per_block);
  return scv::check_launch("bnact_fwd_kernel");
}

int scv_bnact_bwd_reduce(const scv_bnact_bwd_t* p, void* stream) {
  if (check_rows("scv_bnact_bwd X", p->C, p->X, p->x_bs, p->x_ls)) return -1;
  if (p->dO && check_rows("scv_bnact_bwd dO", p->C, p->dO, p->o_bs, p->o_ls)) return -1;
  if (p->dU && check_rows("scv_bnact_bwd dU", p->C, p->dU, p->u_bs, p->u_ls)) return -1;
  SCV_REQUIRE(p->sums, "scv_bnact_bwd_reduce: sums required");
  SCV_REQUIRE(p->B * p->L < (1LL << 31), "scv_bnact_bwd_reduce: more than 2^31 rows");
  Launch2D l = plan2d(p->C / 4, p->B * p->L, 2);
  bnact_bwd_reduce_kernel<<<l.grid, NT, 0, (cudaStream_t)stream>>>(*p, l.cw, l.rows_per_block);
  return scv::check_launch("bnact_bwd_reduce_kernel");
}

int scv_bnact_bwd_apply(const scv_bnact_bwd_t* p, void* stream) {
  if (check_rows("scv_bnact_bwd X", p->C, p->X, p->x_bs, p->x_ls)) return -1;
  if (p->dO && check_rows("scv_bnact_bwd dO", p->C, p->dO, p->o_bs, p->o_ls)) return -1;
  if (p->dU && check_rows("scv_bnact_bwd dU", p->C, p->dU, p->u_bs, p->u_ls)) return -1;
  if (p->dX && check_rows("scv_bnact_bwd dX", p->C, p->dX, p->d_bs, p->d_ls)) return -1;
  SCV_REQUIRE(!(p->mode & 3) || p->sums, "scv_bnact_bwd_apply: sums required");
  SCV_REQUIRE(p->B * p->L < (1LL << 31), "scv_bnact_bwd_apply: more than 2^31 rows");
  const int occ = bnact_occ();
  Launch2D l = plan2d(p->C / 4, p->B * p->L, occ);
  if (occ == 4) bnact_bwd_apply_kernel<4, true><<<l.grid, NT, 0, (cudaStream_t)stream>>>(*p, l.cw, l.rows_per_block);
  else if (p->dU || bnact_nospec()) bnact_bwd_apply_kernel<3, true><<<l.grid, NT, 0, (cudaStream_t)stream>>>(*p, l.cw, l.rows_per_block);
  else bnact_bwd_apply_kernel<3, false><<<l.grid, NT, 0, (cudaStream_t)stream>>>(*p, l.cw, l.rows_per_block);
  return scv::check_launch("bnact_bwd_apply_kernel");
}

int scv_gather(const float* src, const int32_t* idx, float* dst, int64_t n, int64_t flags, void* stream) {
  if (n <= 0) return 0;
  gather_kernel<<<grid1d(n, 16), NT, 0, (cudaStream_t)stream>>>(src, idx, dst, n, (int)flags);
  return scv::check_launch("gather_kernel");
}

int scv_sumsq(const float* g, int64_t n, double* sumsq, void* stream) {
  SCV_REQUIRE(scv::aligned16(g), "scv_sumsq: g must be 16-byte aligned");
  sumsq_kernel<<<grid1d(n / 4 + 1, 4), NT, 0, (cudaStream_t)stream>>>(g, n, sumsq);
  return scv::check_launch("sumsq_kernel");
}

int scv_zero(void* p, int64_t bytes, void* stream) {
  if (bytes <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    scv::set_error("scv_zero: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int scv_optim_step(const scv_optim_t* p, void* stream) {
  SCV_REQUIRE(p->kind >= 0 && p->kind <= 2 && (p->hyper || p->step >= 1), "scv_optim_step: bad kind/step");
  if (p->n <= 0) return 0;
  if (p->pack_idx || p->pack_mask) {
    SCV_REQUIRE(p->n % 4 == 0 && (p->pack_mask || scv::aligned16(p->pack_idx)) && scv::aligned16(p->g) && scv::aligned16(p->p) &&
                    scv::aligned16(p->m) && (p->kind == 2 || scv::aligned16(p->v)) &&
                    (!p->packed_out || scv::aligned16(p->packed_out)) &&
                    (!p->packed16_out || (reinterpret_cast<uintptr_t>(p->packed16_out) & 7) == 0),
                "scv_optim_step: the packed arrays must be 16-byte aligned and n a multiple of 4");
    optim_packed_kernel<<<grid1d(p->n / 4, 2), NT, 0, (cudaStream_t)stream>>>(*p);
    return scv::check_launch("optim_packed_kernel");
  }
  optim_kernel<<<grid1d(p->n, 8), NT, 0, (cudaStream_t)stream>>>(*p);
  return scv::check_launch("optim_kernel");
}

int scv_sumsq_packed(const float* gpacked, const int32_t* pack_idx, const uint32_t* pack_mask, int64_t n, double* sumsq,
                     void* stream) {
  SCV_REQUIRE(gpacked && (pack_idx || pack_mask) && sumsq && n % 4 == 0 && scv::aligned16(gpacked) &&
                  (pack_mask || scv::aligned16(pack_idx)),
              "scv_sumsq_packed: gpacked / pack_idx must be 16-byte aligned and n a multiple of 4");
  if (n <= 0) return 0;
  sumsq_packed_kernel<<<grid1d(n / 4 + 1, 4), NT, 0, (cudaStream_t)stream>>>(gpacked, pack_idx, pack_mask, n, sumsq);
  return scv::check_launch("sumsq_packed_kernel");
}

int scv_loss_finalize(const double* acc, const float* scale, float* out, int64_t n, void* stream) {
  SCV_REQUIRE(n >= 1 && n <= 64, "scv_loss_finalize: n out of range");
  loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, scale, out, (int)n);
  return scv::check_launch("loss_finalize_kernel");
}

int scv_unpack_root(const float* xh, int64_t ld, int64_t nx, const float* arena, float* root_hat, int64_t F,
                    void* stream) {
  if (F <= 0) return 0;
  unpack_root_kernel<<<grid1d(F * 3, 8), NT, 0, (cudaStream_t)stream>>>(xh, ld, (int)nx, arena, root_hat, F * 3);
  return scv::check_launch("unpack_root_kernel");
}

int scv_d2f(const double* in, float* out, int64_t n, void* stream) {
  d2f_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(in, out, n);
  return scv::check_launch("d2f_kernel");
}

}  // extern "C"
