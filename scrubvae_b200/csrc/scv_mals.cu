// "moving_avg_lsq" scrubber (MovingAvgLeastSquares, reference model/disentangle.py:393-538; loss train/losses.py:237-245;
// running-covariance update after the optimizer step train/trainer.py:169-178), polynomial order 1:
//   x   = mu (B, z)  [+ a column of ones when bias]                     nx = z + bias
//   W_i = solve(Sxx_i + diag(l2_reg, bias column excluded), Sxy_i)      i = 0, 1  (two forgetting factors lam0 < lam1)
//   yhat_i = x W_i;  l_i = sum (y - yhat_i)^2;  loss = (l0 + l1) / 2 / B
//   forgetting factors: l0 < l1 ? (lam0 = clamp(lam0 - delta), lam1 = lam0 + lamdiff)
//                               : (lam1 = clamp(lam1 + delta), lam0 = lam1 - lamdiff)
//   update: Sxx_i = lam_i Sxx_i + x^T x,  Sxy_i = lam_i Sxy_i + x^T y
// The weights are constants for autograd (the running sums are detached): d loss / d mu = ((yhat0 - y) W0^T + (yhat1 - y)
// W1^T) / B over the first z rows of W.
#include "scv_common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAXNX = 136;  // z <= 128 (+ bias), padded
constexpr int MAXNY = 16;

// Gaussian elimination with partial pivoting on the augmented matrix [A | B] in shared memory, fp32 like
// torch.linalg.solve on fp32 inputs (LAPACK getrf/getrs: same pivoting rule, unblocked order).  One block per system.
__global__ void __launch_bounds__(NT) mals_solve_kernel(const float* __restrict__ Sxx0, const float* __restrict__ Sxy0,
                                                        const float* __restrict__ Sxx1, const float* __restrict__ Sxy1,
                                                        float l2, int bias, int nx, int ny, float* __restrict__ W0,
                                                        float* __restrict__ W1) {
  extern __shared__ float sm[];
  const int ld = nx + ny;
  float* M = sm;  // nx x ld
  __shared__ int piv;
  __shared__ float pval;
  const float* Sxx = blockIdx.x == 0 ? Sxx0 : Sxx1;
  const float* Sxy = blockIdx.x == 0 ? Sxy0 : Sxy1;
  float* W = blockIdx.x == 0 ? W0 : W1;
  for (int i = threadIdx.x; i < nx * ld; i += NT) {
    const int r = i / ld, c = i - r * ld;
    float v = c < nx ? Sxx[r * nx + c] : Sxy[r * ny + (c - nx)];
    if (c == r && !(bias && r == nx - 1)) v += l2;
    M[i] = v;
  }
  __syncthreads();
  for (int k = 0; k < nx; ++k) {
    if (threadIdx.x < 32) {  // pivot search: first warp, max |M[r][k]| over r >= k (ties: smallest r, as LAPACK's isamax)
      float best = -1.f;
      int bi = k;
      for (int r = k + (int)threadIdx.x; r < nx; r += 32) {
        const float a = fabsf(M[r * ld + k]);
        if (a > best) { best = a; bi = r; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (threadIdx.x == 0) { piv = bi; pval = M[bi * ld + k]; }
    }
    __syncthreads();
    const int p = piv;
    if (p != k) {
      for (int c = threadIdx.x; c < ld; c += NT) {
        const float t = M[k * ld + c];
        M[k * ld + c] = M[p * ld + c];
        M[p * ld + c] = t;
      }
    }
    __syncthreads();
    const float inv = 1.f / pval;
    // eliminate below: one thread per (row, column) pair in strides; multipliers recomputed per element (M[r][k] is read
    // before any thread of this sweep overwrites it: column k itself is skipped and zeroed afterwards)
    const int rows = nx - k - 1, cols = ld - k - 1;
    for (int i = threadIdx.x; i < rows * cols; i += NT) {
      const int r = k + 1 + i / cols, c = k + 1 + i % cols;
      const float f = M[r * ld + k] * inv;
      M[r * ld + c] = fmaf(-f, M[k * ld + c], M[r * ld + c]);
    }
    __syncthreads();
  }
  // back substitution, one thread per right-hand side
  if ((int)threadIdx.x < ny) {
    const int j = threadIdx.x;
    for (int r = nx - 1; r >= 0; --r) {
      float s = M[r * ld + nx + j];
      for (int c = r + 1; c < nx; ++c) s = fmaf(-M[r * ld + c], M[c * ld + nx + j], s);
      M[r * ld + nx + j] = s / M[r * ld + r];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nx * ny; i += NT) W[i] = M[(i / ny) * ld + nx + (i % ny)];
}

// one warp per row: predictions, squared errors, gradient into dmu
__global__ void __launch_bounds__(NT) mals_loss_kernel(const float* __restrict__ mu, int64_t mu_ld, const float* __restrict__ y,
                                                       int64_t y_ld, const float* __restrict__ W0, const float* __restrict__ W1,
                                                       int bias, int B, int z, int ny, double* __restrict__ l01,
                                                       float* __restrict__ yhat0, float* __restrict__ yhat1,
                                                       const float* __restrict__ gscale, float* __restrict__ dmu, int64_t d_ld) {
  __shared__ float w0[MAXNX * MAXNY], w1[MAXNX * MAXNY];
  __shared__ double sh[32];
  const int nx = z + (bias ? 1 : 0);
  for (int i = threadIdx.x; i < nx * ny; i += NT) { w0[i] = W0[i]; w1[i] = W1[i]; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float g = gscale ? gscale[0] / (float)B : 0.f;
  double a0 = 0.0, a1 = 0.0;
  for (int b = blockIdx.x * (NT / 32) + warp; b < B; b += gridDim.x * (NT / 32)) {
    const float* xr = mu + (int64_t)b * mu_ld;
    float p0[MAXNY], p1[MAXNY];
#pragma unroll
    for (int j = 0; j < MAXNY; ++j) { p0[j] = 0.f; p1[j] = 0.f; }
    for (int k = lane; k < nx; k += 32) {
      const float xv = k < z ? xr[k] : 1.f;
#pragma unroll
      for (int j = 0; j < MAXNY; ++j)
        if (j < ny) { p0[j] = fmaf(xv, w0[k * ny + j], p0[j]); p1[j] = fmaf(xv, w1[k * ny + j], p1[j]); }
    }
    float e0[MAXNY], e1[MAXNY];
#pragma unroll
    for (int j = 0; j < MAXNY; ++j) {
      if (j < ny) {
        const float s0 = scv::warp_sum(p0[j]), s1 = scv::warp_sum(p1[j]);
        const float yv = y[(int64_t)b * y_ld + j];
        e0[j] = s0 - yv;
        e1[j] = s1 - yv;
        if (lane == 0) {
          a0 += (double)e0[j] * e0[j];
          a1 += (double)e1[j] * e1[j];
          if (yhat0) yhat0[(int64_t)b * ny + j] = s0;
          if (yhat1) yhat1[(int64_t)b * ny + j] = s1;
        }
      } else {
        e0[j] = 0.f; e1[j] = 0.f;
      }
    }
    if (dmu) {
      for (int k = lane; k < z; k += 32) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < MAXNY; ++j)
          if (j < ny) acc += e0[j] * w0[k * ny + j] + e1[j] * w1[k * ny + j];
        dmu[(int64_t)b * d_ld + k] += g * acc;
      }
    }
  }
  if (l01) {
    const double s0 = scv::block_sum_d(a0, sh);
    const double s1 = scv::block_sum_d(a1, sh);
    if (threadIdx.x == 0) { atomicAdd(l01, s0); atomicAdd(l01 + 1, s1); }
  }
}

__global__ void mals_finalize_kernel(const double* __restrict__ l01, float* lam0, float* lam1, float delta, float lamdiff,
                                     int B, double* loss) {
  if (threadIdx.x || blockIdx.x) return;
  const float l0 = (float)l01[0], l1 = (float)l01[1];  // the reference compares fp32 sums
  if (l0 < l1) {
    const float a = fminf(fmaxf(lam0[0] - delta, 0.f), 1.f);
    lam0[0] = a;
    lam1[0] = a + lamdiff;
  } else {
    const float a = fminf(fmaxf(lam1[0] + delta, 0.f), 1.f);
    lam1[0] = a;
    lam0[0] = a - lamdiff;
  }
  if (loss) loss[0] += (l01[0] + l01[1]) * 0.5 / (double)B;
}

// S_i = lam_i S_i + x^T [x | y]: block per row r of x^T (feature r), thread per column of [x | y]
__global__ void __launch_bounds__(NT) mals_update_kernel(const float* __restrict__ mu, int64_t mu_ld, const float* __restrict__ y,
                                                         int64_t y_ld, int bias, int B, int z, int ny,
                                                         const float* __restrict__ lam0, const float* __restrict__ lam1,
                                                         float* __restrict__ Sxx0, float* __restrict__ Sxy0,
                                                         float* __restrict__ Sxx1, float* __restrict__ Sxy1) {
  const int nx = z + (bias ? 1 : 0);
  const int r = blockIdx.x, c = threadIdx.x;
  if (c >= nx + ny) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float xr = r < z ? mu[(int64_t)b * mu_ld + r] : 1.f;
    const float v = c < nx ? (c < z ? mu[(int64_t)b * mu_ld + c] : 1.f) : y[(int64_t)b * y_ld + (c - nx)];
    acc = fmaf(xr, v, acc);
  }
  const float a0 = lam0[0], a1 = lam1[0];
  if (c < nx) {
    Sxx0[r * nx + c] = a0 * Sxx0[r * nx + c] + acc;
    Sxx1[r * nx + c] = a1 * Sxx1[r * nx + c] + acc;
  } else {
    Sxy0[r * ny + (c - nx)] = a0 * Sxy0[r * ny + (c - nx)] + acc;
    Sxy1[r * ny + (c - nx)] = a1 * Sxy1[r * ny + (c - nx)] + acc;
  }
}

}  // namespace

extern "C" {

int scv_mals_solve(const float* Sxx0, const float* Sxy0, const float* Sxx1, const float* Sxy1, double l2_reg, int64_t bias,
                   int64_t nx, int64_t ny, float* W0, float* W1, void* stream) {
  SCV_REQUIRE(nx >= 1 && nx <= MAXNX && ny >= 1 && ny <= MAXNY, "scv_mals_solve: nx <= 136, ny <= 16");
  const size_t smem = (size_t)nx * (nx + ny) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mals_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(MAXNX * (MAXNX + MAXNY) * sizeof(float)));
    if (e != cudaSuccess) {
      scv::set_error("scv_mals_solve: cannot opt in to shared memory: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return (int)e;
    }
    attr = true;
  }
  mals_solve_kernel<<<2, NT, smem, (cudaStream_t)stream>>>(Sxx0, Sxy0, Sxx1, Sxy1, (float)l2_reg, (int)bias, (int)nx, (int)ny,
                                                          W0, W1);
  return scv::check_launch("mals_solve_kernel");
}

int scv_mals_loss(const float* mu, int64_t mu_ld, const float* y, int64_t y_ld, const float* W0, const float* W1, int64_t bias,
                  int64_t B, int64_t z, int64_t ny, double* l01, float* yhat0, float* yhat1, const float* gscale, float* dmu,
                  int64_t d_ld, void* stream) {
  SCV_REQUIRE(z + (bias ? 1 : 0) <= MAXNX && ny >= 1 && ny <= MAXNY, "scv_mals_loss: nx <= 136, ny <= 16");
  if (B <= 0) return 0;
  int blocks = (int)((B + NT / 32 - 1) / (NT / 32));
  const int cap = scv::sm_count() * 4;
  if (blocks > cap) blocks = cap;
  mals_loss_kernel<<<blocks, NT, 0, (cudaStream_t)stream>>>(mu, mu_ld, y, y_ld, W0, W1, (int)bias, (int)B, (int)z, (int)ny, l01,
                                                           yhat0, yhat1, gscale, dmu, d_ld);
  return scv::check_launch("mals_loss_kernel");
}

int scv_mals_finalize(const double* l01, float* lam0, float* lam1, double delta, double lamdiff, int64_t B, double* loss,
                      void* stream) {
  mals_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(l01, lam0, lam1, (float)delta, (float)lamdiff, (int)B, loss);
  return scv::check_launch("mals_finalize_kernel");
}

int scv_mals_update(const float* mu, int64_t mu_ld, const float* y, int64_t y_ld, int64_t bias, int64_t B, int64_t z, int64_t ny,
                    const float* lam0, const float* lam1, float* Sxx0, float* Sxy0, float* Sxx1, float* Sxy1, void* stream) {
  const int nx = (int)z + (bias ? 1 : 0);
  SCV_REQUIRE(nx <= MAXNX && ny >= 1 && ny <= MAXNY && nx + ny <= NT, "scv_mals_update: nx <= 136, ny <= 16");
  mals_update_kernel<<<nx, NT, 0, (cudaStream_t)stream>>>(mu, mu_ld, y, y_ld, (int)bias, (int)B, (int)z, (int)ny, lam0, lam1,
                                                         Sxx0, Sxy0, Sxx1, Sxy1);
  return scv::check_launch("mals_update_kernel");
}

}  // extern "C"
