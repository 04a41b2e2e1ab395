// Kernel mutual-information estimator of the "mcmi" scrubbing loss (reference model/disentangle.py:234-317,
// train/losses.py:221-225, estimator rebuild train/trainer.py:184-199):
//   loss = mean_b [ LSE_s a_xy[b,s] - LSE_s a_x[b,s] - LSE_s a_y[b,s] ]
//   a_x  = -1/2 (logA_x[s] + sum_d (x_b - xs_s)_d^2 / var_s[s,d]),  a_y = -1/2 (logA_y + |y_b - ys_s|^2 / gamma),
//   a_xy = a_x + a_y  (Gaussian mixtures centred on the stored samples of the previous, already updated, batch)
// var_mode "sphere": var_s = bandwidth for every (s, d); "diagonal": var_s[s,d] = L_s[d,d]^2 + bandwidth.
// One block per row b: pass 1 the three running log-sum-exps over s, pass 2 the gradient
//   d loss / d x_b = 1/B sum_s (softmax_xy[s] - softmax_x[s]) * (-(x_b - xs_s) / var_s[s])      (y is data: no gradient).
#include "scv_common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAXZ = 256;

struct Lse {  // running (max, sum exp)
  float m, s;
  __device__ __forceinline__ void add(float a) {
    if (a > m) { s = s * __expf(m - a) + 1.f; m = a; }
    else s += __expf(a - m);
  }
  __device__ __forceinline__ void merge(const Lse& o) {
    if (o.m > m) { s = s * __expf(m - o.m) + o.s; m = o.m; }
    else if (o.s > 0.f) s += o.s * __expf(o.m - m);
  }
};

__device__ __forceinline__ Lse block_lse(Lse v, Lse* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Lse t;
    t.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    t.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    v.merge(t);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  Lse r = sh[0];
  for (int i = 1; i < NT / 32; ++i) r.merge(sh[i]);
  return r;
}

__global__ void __launch_bounds__(NT) mi_loss_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t y_ld,
                                                     const float* __restrict__ xs, const float* __restrict__ ys,
                                                     const float* __restrict__ var_s, const float* __restrict__ logAx,
                                                     float bandwidth, float logAy, int S, int z, int dy,
                                                     const float* __restrict__ valid, double* loss,
                                                     const float* __restrict__ gscale, float* __restrict__ dx, int B) {
  __shared__ float xb[MAXZ], yb[64], gacc[MAXZ];
  __shared__ Lse shl[NT / 32];
  if (valid && valid[0] == 0.f) return;  // no estimator yet (first batch of an epoch): loss 0, no gradient
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < z; d += NT) { xb[d] = x[(int64_t)b * z + d]; gacc[d] = 0.f; }
  for (int d = threadIdx.x; d < dy; d += NT) yb[d] = y[(int64_t)b * y_ld + d];
  __syncthreads();
  const float inv_g = 1.f / bandwidth;
  Lse lx = {-INFINITY, 0.f}, ly = lx, lxy = lx;
  for (int s = threadIdx.x; s < S; s += NT) {
    const float* xr = xs + (int64_t)s * z;
    float sdx = 0.f, sdy = 0.f;
    if (var_s) {
      const float* vr = var_s + (int64_t)s * z;
      for (int d = 0; d < z; ++d) { const float t = xb[d] - xr[d]; sdx += t / vr[d] * t; }
    } else {
      for (int d = 0; d < z; ++d) { const float t = xb[d] - xr[d]; sdx += t * inv_g * t; }
    }
    const float* yr = ys + (int64_t)s * dy;
    for (int d = 0; d < dy; ++d) { const float t = yb[d] - yr[d]; sdy += t * inv_g * t; }
    const float lax = logAx[var_s ? s : 0];
    const float ax = -0.5f * (lax + sdx), ay = -0.5f * (logAy + sdy), axy = -0.5f * (lax + logAy + sdx + sdy);
    lx.add(ax); ly.add(ay); lxy.add(axy);
  }
  lx = block_lse(lx, shl);
  ly = block_lse(ly, shl);
  lxy = block_lse(lxy, shl);
  const float Ex = lx.m + logf(lx.s), Ey = ly.m + logf(ly.s), Exy = lxy.m + logf(lxy.s);
  if (loss && threadIdx.x == 0) atomicAdd(loss, (double)(Exy - Ex - Ey) / (double)B);
  if (!dx) return;
  // pass 2: softmax weights recomputed per sample (cheaper than keeping S / NT rows of state); the gradient is accumulated
  // in registers, 64 dimensions at a time, and reduced over the block with shuffles
  const float gs = (gscale ? gscale[0] : 1.f) / (float)B;
  for (int d0 = 0; d0 < z; d0 += 64) {
    float g[64];
#pragma unroll
    for (int q = 0; q < 64; ++q) g[q] = 0.f;
    const int nd = min(64, z - d0);
    for (int s = threadIdx.x; s < S; s += NT) {
      const float* xr = xs + (int64_t)s * z;
      const float* vr = var_s ? var_s + (int64_t)s * z : nullptr;
      float sdx = 0.f, sdy = 0.f;
      for (int d = 0; d < z; ++d) { const float t = xb[d] - xr[d]; sdx += vr ? t / vr[d] * t : t * inv_g * t; }
      const float* yr = ys + (int64_t)s * dy;
      for (int d = 0; d < dy; ++d) { const float t = yb[d] - yr[d]; sdy += t * inv_g * t; }
      const float lax = logAx[var_s ? s : 0];
      const float ax = -0.5f * (lax + sdx), axy = -0.5f * (lax + logAy + sdx + sdy);
      const float w = gs * (__expf(axy - Exy) - __expf(ax - Ex));
#pragma unroll
      for (int q = 0; q < 64; ++q)
        if (q < nd) g[q] -= w * (xb[d0 + q] - xr[d0 + q]) * (vr ? 1.f / vr[d0 + q] : inv_g);
    }
#pragma unroll
    for (int q = 0; q < 64; ++q) {
      const float v = scv::warp_sum(g[q]);
      if ((threadIdx.x & 31) == 0 && q < nd) atomicAdd(&gacc[d0 + q], v);  // 8 warps per dimension
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < z; d += NT) dx[(int64_t)b * z + d] += gacc[d];
}

// estimator rebuild after the optimizer step (trainer.py:184-199): stored samples = the UPDATED encoder's mu, the batch's
// conditional variables; diagonal mode: var_s = diag(L)^2 + bandwidth, logA_x[s] = z log 2pi + sum_d log var_s[s,d]
__global__ void __launch_bounds__(NT) mi_update_kernel(const float* __restrict__ mu, const float* __restrict__ L,
                                                       const float* __restrict__ var, int64_t var_ld, float* __restrict__ xs,
                                                       float* __restrict__ ys, float* __restrict__ var_s,
                                                       float* __restrict__ logAx, float bandwidth, int S, int z, int dy,
                                                       float* __restrict__ valid) {
  const float log2pi = 1.8378770664093453f;
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    float acc = 0.f;
    for (int d = threadIdx.x; d < z; d += NT) {
      xs[(int64_t)s * z + d] = mu[(int64_t)s * z + d];
      if (var_s) {
        const float l = L[((int64_t)s * z + d) * z + d];
        const float v = l * l + bandwidth;
        var_s[(int64_t)s * z + d] = v;
        acc += logf(v);
      }
    }
    for (int d = threadIdx.x; d < dy; d += NT) ys[(int64_t)s * dy + d] = var[(int64_t)s * var_ld + d];
    if (var_s) {
      __shared__ float red[NT / 32];
      acc = scv::warp_sum(acc);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < NT / 32; ++i) t += red[i];
        logAx[s] = (float)z * log2pi + t;
      }
      __syncthreads();
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (!var_s) logAx[0] = (float)z * (log2pi + logf(bandwidth));
    if (valid) valid[0] = 1.f;
  }
}

}  // namespace

extern "C" {

int scv_mi_loss(const float* x, const float* y, int64_t y_ld, const float* xs, const float* ys, const float* var_s,
                const float* logAx, double bandwidth, int64_t S, int64_t B, int64_t z, int64_t dy, const float* valid,
                double* loss, const float* gscale, float* dx, void* stream) {
  SCV_REQUIRE(x && y && xs && ys && logAx, "scv_mi_loss: null pointer");
  SCV_REQUIRE(z >= 1 && z <= MAXZ && dy >= 1 && dy <= 64 && S >= 1 && bandwidth > 0, "scv_mi_loss: size out of range");
  if (B <= 0) return 0;
  const float logAy = (float)dy * (1.8378770664093453f + logf((float)bandwidth));
  mi_loss_kernel<<<(unsigned)B, NT, 0, (cudaStream_t)stream>>>(x, y, y_ld, xs, ys, var_s, logAx, (float)bandwidth, logAy, (int)S,
                                                             (int)z, (int)dy, valid, loss, gscale, dx, (int)B);
  return scv::check_launch("mi_loss_kernel");
}

int scv_mi_update(const float* mu, const float* L, const float* var, int64_t var_ld, float* xs, float* ys, float* var_s,
                  float* logAx, double bandwidth, int64_t S, int64_t z, int64_t dy, float* valid, void* stream) {
  SCV_REQUIRE(mu && var && xs && ys && logAx && (!var_s || L), "scv_mi_update: null pointer");
  if (S <= 0) return 0;
  const unsigned grid = (unsigned)(S < 1184 ? S : 1184);
  mi_update_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(mu, L, var, var_ld, xs, ys, var_s, logAx, (float)bandwidth, (int)S, (int)z,
                                                        (int)dy, valid);
  return scv::check_launch("mi_update_kernel");
}

}  // extern "C"
