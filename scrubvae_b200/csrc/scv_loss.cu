// Fused latent / loss kernels of the SC-VAE step: CholeskyL + reparameterisation (fwd/bwd), KL prior,
// forward-kinematics joint-position loss + root loss (with unit gradients), tanh backward,
// gradient-reversal head losses.  Reductions are accumulated in double.
#include "scv_common.cuh"
#include "scv_fk.h"

namespace {

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

// one block per sample
__global__ void __launch_bounds__(256) reparam_fwd_kernel(const float* __restrict__ ms, int64_t ms_ld,
                                                          const float* __restrict__ eps, const float* __restrict__ var,
                                                          int nvar, float* __restrict__ mu, float* __restrict__ L,
                                                          float* __restrict__ zc, int64_t zc_ld, int z, int rnd) {
  extern __shared__ float sh[];
  const int nsig = z * (z + 1) / 2;
  float* sig = sh;          // nsig
  float* e = sh + nsig;     // z
  const int64_t b = blockIdx.x;
  const float* row = ms + b * ms_ld;
  for (int i = threadIdx.x; i < nsig; i += blockDim.x) sig[i] = row[z + i];
  for (int i = threadIdx.x; i < z; i += blockDim.x) e[i] = eps ? eps[b * z + i] : 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < z; i += blockDim.x) {
    const int tri = i * (i + 1) / 2;
    float m = row[i], acc = m;
    for (int j = 0; j < i; ++j) acc = fmaf(sig[tri + j], e[j], acc);
    acc = fmaf(softplus_f(sig[tri + i]), e[i], acc);
    if (mu) mu[b * z + i] = m;
    if (zc) { float zv = eps ? acc : m; zc[b * zc_ld + i] = rnd ? scv::round_tf32(zv) : zv; }
  }
  if (zc) {
    for (int t = threadIdx.x; t < (int)zc_ld - z; t += blockDim.x)
      { float vv = t < nvar ? var[b * nvar + t] : 0.f; zc[b * zc_ld + z + t] = rnd ? scv::round_tf32(vv) : vv; }
  }
  if (L) {
    float* Lb = L + b * (int64_t)z * z;
    for (int idx = threadIdx.x; idx < z * z; idx += blockDim.x) {
      int i = idx / z, j = idx - i * z;
      float v = 0.f;
      if (j < i) v = sig[i * (i + 1) / 2 + j];
      else if (j == i) v = softplus_f(sig[i * (i + 1) / 2 + i]);
      Lb[idx] = v;
    }
  }
}

__global__ void __launch_bounds__(256) reparam_bwd_kernel(const float* __restrict__ ms, int64_t ms_ld,
                                                          const float* __restrict__ eps, const float* __restrict__ dmu,
                                                          const float* __restrict__ dmu2, float s2,
                                                          const float* __restrict__ dz, int64_t dz_ld,
                                                          const float* __restrict__ dL, float* __restrict__ dms,
                                                          int64_t dms_ld, int z, int rnd) {
  extern __shared__ float sh[];
  float* e = sh;       // z
  float* gz = sh + z;  // z
  const int nsig = z * (z + 1) / 2;
  const int64_t b = blockIdx.x;
  for (int i = threadIdx.x; i < z; i += blockDim.x) {
    e[i] = eps ? eps[b * z + i] : 0.f;
    gz[i] = (dz && eps) ? dz[b * dz_ld + i] : 0.f;
    float g = 0.f;
    if (dmu) g += dmu[b * z + i];
    if (dmu2) g += s2 * dmu2[b * z + i];
    if (dz) g += dz[b * dz_ld + i];
    dms[b * dms_ld + i] = rnd ? scv::round_tf32(g) : g;
  }
  __syncthreads();
  const float* sraw = ms + b * ms_ld + z;
  float* drow = dms + b * dms_ld + z;
  const float* dLb = dL ? dL + b * (int64_t)z * z : nullptr;
  for (int idx = threadIdx.x; idx < z * z; idx += blockDim.x) {
    int i = idx / z, j = idx - i * z;
    if (j > i) continue;
    float g = gz[i] * e[j];
    if (dLb) g += dLb[idx];
    const int t = i * (i + 1) / 2 + j;
    if (j == i) g *= softplus_grad(sraw[t]);
    drow[t] = rnd ? scv::round_tf32(g) : g;
  }
  for (int t = z + nsig + threadIdx.x; t < dms_ld; t += blockDim.x) dms[b * dms_ld + t] = 0.f;
}

__global__ void __launch_bounds__(256) kl_kernel(const float* __restrict__ mu, const float* __restrict__ L,
                                                 double* loss, const float* gscale, float* __restrict__ dmu,
                                                 float* __restrict__ dL, int B, int z) {
  __shared__ double shd[32];
  const int64_t b = blockIdx.x;
  const float invB = (gscale ? *gscale : 1.f) / (float)B;
  const float* Lb = L + b * (int64_t)z * z;
  float acc = 0.f;
  for (int idx = threadIdx.x; idx < z * z; idx += blockDim.x) {
    int i = idx / z, j = idx - i * z;
    float v = Lb[idx];
    float g = 0.f;
    if (j <= i) {
      acc += 0.5f * v * v;
      g = v * invB;
      if (j == i) {
        float m = mu[b * z + i];
        acc += -0.5f * (1.f + 2.f * logf(v) - m * m);
        g = (v - 1.f / v) * invB;
        if (dmu) dmu[b * z + i] = m * invB;
      }
    }
    if (dL) dL[b * (int64_t)z * z + idx] = g;
  }
  if (loss) {
    double s = scv::block_sum_d((double)acc, shd);
    if (threadIdx.x == 0) atomicAdd(loss, s / (double)B);
  }
}

constexpr int FK_THREADS = 64;
__global__ void __launch_bounds__(FK_THREADS) recon_loss_kernel(const float* __restrict__ xh, int64_t ld,
                                                                const float* __restrict__ offsets,
                                                                const float* __restrict__ target,
                                                                const float* __restrict__ root,
                                                                const float* __restrict__ arena,
                                                                const int32_t* __restrict__ tree, int n_tree,
                                                                double* loss, float* __restrict__ root_hat,
                                                                float* __restrict__ dxh, int64_t F, int B, int J) {
  __shared__ int32_t stree[SCV_MAX_J * 3];
  __shared__ double shd[32];
  for (int i = threadIdx.x; i < n_tree; i += blockDim.x) stree[i] = tree[i];
  __syncthreads();
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nx = J * 6;
  double ljpe = 0.0, lroot = 0.0;
  if (f < F) {
    float c6[SCV_MAX_J * 6], off[SCV_MAX_J * 3], tgt[SCV_MAX_J * 3], gc[SCV_MAX_J * 6];
    const float* xr = xh + f * ld;
    for (int q = 0; q < nx; ++q) c6[q] = xr[q];
    for (int q = 0; q < J * 3; ++q) { off[q] = offsets[f * J * 3 + q]; tgt[q] = target[f * J * 3 + q]; }
    const float scale = 1.f / ((float)B * 3.f * (float)J);
    float l = scvfk::fk_jpe_frame(c6, off, tgt, stree, J, 1e-8f, scale, gc);
    ljpe = (double)l * (double)scale;
    float* dr = dxh + f * ld;
    for (int q = 0; q < nx; ++q) dr[q] = gc[q];
    const float invB = 1.f / (float)B;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float a0 = arena[d], a1 = arena[3 + d];
      float rh = 0.5f * (xr[nx + d] + 1.f) * (a1 - a0) + a0;
      if (root_hat) root_hat[f * 3 + d] = rh;
      float diff = rh - root[f * 3 + d];
      lroot += (double)(diff * diff) * (double)invB;
      dr[nx + d] = 2.f * diff * 0.5f * (a1 - a0) * invB;
    }
    for (int q = nx + 3; q < ld; ++q) dr[q] = 0.f;
  }
  double s0 = scv::block_sum_d(ljpe, shd);
  if (threadIdx.x == 0) atomicAdd(loss, s0);
  double s1 = scv::block_sum_d(lroot, shd);
  if (threadIdx.x == 0) atomicAdd(loss + 1, s1);
}

__global__ void __launch_bounds__(256) out_bwd_kernel(const float* __restrict__ xh, const float* __restrict__ dxh,
                                                      int ld, const float* g_jpe, const float* g_root, int nx,
                                                      float* __restrict__ draw, int64_t d_bs, int64_t d_ls,
                                                      int64_t rows, int W, int rnd) {
  const float gj = g_jpe ? *g_jpe : 0.f, gr = g_root ? *g_root : 0.f;
  const int64_t total = rows * ld;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t r = i / ld;
    int c = (int)(i - r * ld);
    int64_t b = r / W;
    int w = (int)(r - b * W);
    float y = xh[i];
    float g = dxh[i] * (c < nx ? gj : (c < nx + 3 ? gr : 0.f));
    const float dv = g * (1.f - y * y);
    draw[b * d_bs + w * d_ls + c] = rnd ? scv::round_tf32(dv) : dv;
  }
}

struct GrPtrs {
  const float* pred[8];
  float* dpred[8];
  float w[8];
};
__global__ void __launch_bounds__(256) gr_loss_kernel(const GrPtrs P, int ld, int n_ens, const float* __restrict__ target,
                                                      const int64_t* __restrict__ labels, int B, int d, double* loss,
                                                      const float* gscale) {
  __shared__ double shd[32];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const float gs = gscale ? *gscale : 1.f;
  double acc = 0.0;
  if (b < B) {
    for (int e = 0; e < n_ens; ++e) {
      const float* pr = P.pred[e] + (int64_t)b * ld;
      float* dp = P.dpred[e] ? P.dpred[e] + (int64_t)b * ld : nullptr;
      float le = 0.f;
      if (labels) {  // CrossEntropyLoss(reduction="sum")
        float mx = pr[0];
        for (int q = 1; q < d; ++q) mx = fmaxf(mx, pr[q]);
        float se = 0.f;
        for (int q = 0; q < d; ++q) se += expf(pr[q] - mx);
        const int lab = (int)labels[b];
        le = logf(se) + mx - pr[lab];
        if (dp)
          for (int q = 0; q < d; ++q) dp[q] = gs * P.w[e] * (expf(pr[q] - mx) / se - (q == lab ? 1.f : 0.f));
      } else {  // MSELoss(reduction="sum")
        for (int q = 0; q < d; ++q) {
          float df = pr[q] - target[(int64_t)b * d + q];
          le += df * df;
          if (dp) dp[q] = gs * P.w[e] * 2.f * df;
        }
      }
      if (dp)
        for (int q = d; q < ld; ++q) dp[q] = 0.f;
      acc += (double)le * (double)P.w[e];
    }
  }
  if (loss) {
    double s = scv::block_sum_d(acc, shd);
    if (threadIdx.x == 0) atomicAdd(loss, s);
  }
}

}  // namespace

extern "C" {

int scv_reparam_fwd(const float* ms, int64_t ms_ld, const float* eps, const float* var, int64_t nvar, float* mu,
                    float* L, float* zc, int64_t zc_ld, int64_t B, int64_t z, int64_t flags, void* stream) {
  size_t sh = (size_t)(z * (z + 1) / 2 + z) * sizeof(float);
  SCV_REQUIRE(sh <= 48 * 1024, "scv_reparam_fwd: z_dim %lld too large", (long long)z);
  SCV_REQUIRE(!zc || zc_ld >= z + nvar, "scv_reparam_fwd: zc_ld too small");
  if (B <= 0) return 0;
  reparam_fwd_kernel<<<(unsigned)B, 256, sh, (cudaStream_t)stream>>>(ms, ms_ld, eps, var, (int)nvar, mu, L, zc, zc_ld,
                                                                    (int)z, (int)(flags & SCV_F_ROUND_TF32));
  return scv::check_launch("reparam_fwd_kernel");
}

int scv_reparam_bwd(const float* ms, int64_t ms_ld, const float* eps, const float* dmu, const float* dmu2,
                    double dmu2_scale, const float* dz, int64_t dz_ld, const float* dL, float* dms, int64_t dms_ld,
                    int64_t B, int64_t z, int64_t flags, void* stream) {
  if (B <= 0) return 0;
  reparam_bwd_kernel<<<(unsigned)B, 256, (size_t)(2 * z) * sizeof(float), (cudaStream_t)stream>>>(
      ms, ms_ld, eps, dmu, dmu2, (float)dmu2_scale, dz, dz_ld, dL, dms, dms_ld, (int)z, (int)(flags & SCV_F_ROUND_TF32));
  return scv::check_launch("reparam_bwd_kernel");
}

int scv_kl(const float* mu, const float* L, double* loss, const float* gscale, float* dmu, float* dL, int64_t B,
           int64_t z, void* stream) {
  if (B <= 0) return 0;
  kl_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(mu, L, loss, gscale, dmu, dL, (int)B, (int)z);
  return scv::check_launch("kl_kernel");
}

int scv_recon_loss(const float* xh, int64_t ld, const float* offsets, const float* target, const float* root,
                   const float* arena, const int32_t* tree, int64_t n_tree, double* loss, float* root_hat, float* dxh,
                   int64_t F, int64_t B, int64_t J, void* stream) {
  SCV_REQUIRE(J <= SCV_MAX_J && n_tree <= SCV_MAX_J * 3 && ld >= J * 6 + 3, "scv_recon_loss: bad J/tree/ld");
  if (F <= 0) return 0;
  recon_loss_kernel<<<(unsigned)((F + FK_THREADS - 1) / FK_THREADS), FK_THREADS, 0, (cudaStream_t)stream>>>(
      xh, ld, offsets, target, root, arena, tree, (int)n_tree, loss, root_hat, dxh, F, (int)B, (int)J);
  return scv::check_launch("recon_loss_kernel");
}

int scv_out_bwd(const float* xh, const float* dxh, int64_t ld, const float* g_jpe, const float* g_root, int64_t nx,
                float* draw, int64_t d_bs, int64_t d_ls, int64_t B, int64_t W, int64_t flags, void* stream) {
  int64_t total = B * W * ld;
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)scv::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return 0;
  out_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(xh, dxh, (int)ld, g_jpe, g_root, (int)nx, draw,
                                                                    d_bs, d_ls, B * W, (int)W,
                                                                    (int)(flags & SCV_F_ROUND_TF32));
  return scv::check_launch("out_bwd_kernel");
}

int scv_gr_loss(const float* const* pred, float* const* dpred, int64_t ld, int64_t n_ens, const float* target,
                const int64_t* labels, int64_t B, int64_t d, int64_t num_keys, double* loss, const float* gscale,
                void* stream) {
  SCV_REQUIRE(n_ens >= 1 && n_ens <= 8 && d <= ld, "scv_gr_loss: bad n_ens/d");
  GrPtrs P;
  const double c = (double)n_ens * (double)num_keys * (double)B;
  for (int e = 0; e < 8; ++e) {
    P.pred[e] = e < n_ens ? pred[e] : nullptr;
    P.dpred[e] = (e < n_ens && dpred) ? dpred[e] : nullptr;
    P.w[e] = e < n_ens ? (float)pow(c, -(double)(n_ens - e)) : 0.f;
  }
  if (B <= 0) return 0;
  gr_loss_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, (int)ld, (int)n_ens, target, labels,
                                                                               (int)B, (int)d, loss, gscale);
  return scv::check_launch("gr_loss_kernel");
}

}  // extern "C"
