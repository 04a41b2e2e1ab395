// "qda" scrubber (QuadraticDiscriminantFilter, reference model/disentangle.py:90-232; loss train/losses.py:247-251;
// running-statistics update after the optimizer step train/trainer.py:169-178).  Per class c two one-vs-rest Gaussian
// classifiers A, B with forgetting factors lama < lamb, each holding (mean, covariance) of "not c" (0) and "c" (1):
// buffers m0a, m1a, m0b, m1b (nc, z) and S0a, S1a, S0b, S1b (nc, z, z); index q = 0..3 in that order below.
//   cgll_q(x) = -1/2 (logdet S_q + (x - m_q)^T S_q^-1 (x - m_q))
//   lla = sum_b [y_b != c] cgll_0a + [y_b == c] cgll_1a   (llb alike);   lla > llb ? lama -= delta, lamb = lama + lamdiff
//                                                                               : lamb += delta, lama = lamb - lamdiff
//   loss = 1/nc sum_c 1/2 (sum_b s_b (cgll_1a - cgll_0a) + sum_b s_b (cgll_1b - cgll_0b)),  s_b = +1 if y_b == c else -1
//   d loss / d x_b = 1/nc sum_c s_b / 2 [(-t_1a + t_0a) + (-t_1b + t_0b)],  t_q = S_q^-1 (x_b - m_q)
//   update: class-conditional batch mean / covariance (correction = 0) blended in with lama (A) and lamb (B).
#include "scv_common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAXZ = 128;
constexpr int MAXC = 16;
constexpr int RB = 4;  // rows per warp in the loss kernel (each S^-1 row is loaded once for RB rows)

struct QdaPtrs {
  const float* m[4];
  const float* S[4];
};
struct QdaMut {
  float* m[4];
  float* S[4];
};

// Gauss-Jordan inverse with partial pivoting of one z x z matrix per block ([S | I] in shared memory), plus log|det|
// (NaN for a negative determinant, as torch.logdet).  Writes the TRANSPOSE of the inverse: the loss kernel reads it by rows.
__global__ void __launch_bounds__(NT) qda_factor_kernel(const QdaPtrs P, int nc, int z, float* __restrict__ SinvT,
                                                        float* __restrict__ logdet) {
  extern __shared__ float sm[];
  __shared__ int piv;
  __shared__ float pval;
  const int q = blockIdx.x / nc, c = blockIdx.x - q * nc;
  const float* S = P.S[q] + (size_t)c * z * z;
  const int ld = 2 * z;
  for (int i = threadIdx.x; i < z * ld; i += NT) {
    const int r = i / ld, col = i - r * ld;
    sm[i] = col < z ? S[r * z + col] : (col - z == r ? 1.f : 0.f);
  }
  __syncthreads();
  float lsum = 0.f;
  int neg = 0;
  for (int k = 0; k < z; ++k) {
    if (threadIdx.x < 32) {
      float best = -1.f;
      int bi = k;
      for (int r = k + (int)threadIdx.x; r < z; r += 32) {
        const float a = fabsf(sm[r * ld + k]);
        if (a > best) { best = a; bi = r; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (threadIdx.x == 0) { piv = bi; pval = sm[bi * ld + k]; }
    }
    __syncthreads();
    const int p = piv;
    const float pv = pval;
    if (p != k) {
      for (int col = threadIdx.x; col < ld; col += NT) {
        const float t = sm[k * ld + col];
        sm[k * ld + col] = sm[p * ld + col];
        sm[p * ld + col] = t;
      }
      neg ^= 1;
    }
    if (pv < 0.f) neg ^= 1;
    lsum += logf(fabsf(pv));
    __syncthreads();
    // eliminate column k from every other row (columns > k of the augmented matrix; column k itself is left stale)
    const int cols = ld - k - 1;
    const float inv = 1.f / pv;
    for (int i = threadIdx.x; i < (z - 1) * cols; i += NT) {
      int r = i / cols;
      const int col = k + 1 + (i - r * cols);
      if (r >= k) ++r;
      const float f = sm[r * ld + k] * inv;
      sm[r * ld + col] = fmaf(-f, sm[k * ld + col], sm[r * ld + col]);
    }
    __syncthreads();
    for (int col = k + 1 + threadIdx.x; col < ld; col += NT) sm[k * ld + col] *= inv;  // normalise the pivot row
    __syncthreads();
  }
  float* out = SinvT + (size_t)blockIdx.x * z * z;
  for (int i = threadIdx.x; i < z * z; i += NT) {
    const int r = i / z, col = i - r * z;
    out[col * z + r] = sm[r * ld + z + col];
  }
  if (threadIdx.x == 0) logdet[blockIdx.x] = neg ? __int_as_float(0x7fc00000) : lsum;
}

// warp per RB rows, all classes and the four Gaussians; acc (double, 4 per class): lla, llb, llra, llrb
__global__ void __launch_bounds__(NT) qda_loss_kernel(const float* __restrict__ x, int64_t x_ld, const int64_t* __restrict__ y,
                                                      const int64_t* __restrict__ classes, const QdaPtrs P,
                                                      const float* __restrict__ SinvT, const float* __restrict__ logdet, int nc,
                                                      int z, int B, double* __restrict__ acc, const float* __restrict__ gscale,
                                                      float* __restrict__ dx, int64_t d_ld) {
  __shared__ float rs[NT / 32][RB][MAXZ];
  __shared__ double wacc[NT / 32][MAXC][4];  // per warp: lla, llb, llra, llrb of every class (lane 0 accumulates)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int U = (z + 31) / 32;
  const float gcoef = gscale ? gscale[0] * 0.5f / ((float)nc * (float)B) : 0.f;
  for (int i = threadIdx.x; i < (NT / 32) * MAXC * 4; i += NT) (&wacc[0][0][0])[i] = 0.0;
  __syncthreads();
  for (int b0 = (blockIdx.x * (NT / 32) + warp) * RB; b0 < B; b0 += gridDim.x * (NT / 32) * RB) {
    float xv[RB][4], gacc[RB][4];
    int64_t yb[RB];
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
      const int b = b0 + rr;
      yb[rr] = b < B ? y[b] : 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = lane + 32 * u;
        xv[rr][u] = (b < B && j < z) ? x[(int64_t)b * x_ld + j] : 0.f;
        gacc[rr][u] = 0.f;
      }
    }
    for (int c = 0; c < nc; ++c) {
      const int64_t label = classes[c];
      float ll[RB][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float* mq = P.m[q] + (size_t)c * z;
        const float* Si = SinvT + ((size_t)q * nc + c) * z * z;
        const float ld_q = logdet[q * nc + c];
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < RB; ++rr)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = lane + 32 * u;
            if (j < z) rs[warp][rr][j] = xv[rr][u] - mq[j];
          }
        __syncwarp();
        float t[RB][4];
#pragma unroll
        for (int rr = 0; rr < RB; ++rr)
#pragma unroll
          for (int u = 0; u < 4; ++u) t[rr][u] = 0.f;
        for (int k = 0; k < z; ++k) {
          float sv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) sv[u] = (u < U && lane + 32 * u < z) ? __ldg(Si + (size_t)k * z + lane + 32 * u) : 0.f;
#pragma unroll
          for (int rr = 0; rr < RB; ++rr) {
            const float rk = rs[warp][rr][k];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[rr][u] = fmaf(sv[u], rk, t[rr][u]);
          }
        }
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
          float qf = 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = lane + 32 * u;
            if (j < z) qf = fmaf(rs[warp][rr][j], t[rr][u], qf);
          }
          qf = scv::warp_sum(qf);
          ll[rr][q] = -0.5f * (ld_q + qf);
          if (dx) {  // d cgll_q / d x = -t; sign per (row, class) and the q pattern: q = 0,2 -> "not c" (+), q = 1,3 -> "c" (-)
            const float s = (b0 + rr < B) ? ((yb[rr] == label) ? 1.f : -1.f) : 0.f;
            const float w = (q & 1) ? -s : s;
#pragma unroll
            for (int u = 0; u < 4; ++u) gacc[rr][u] = fmaf(w, t[rr][u], gacc[rr][u]);
          }
        }
      }
      if (acc && lane == 0) {
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
          if (b0 + rr >= B) continue;
          const bool is1 = yb[rr] == label;
          const float s = is1 ? 1.f : -1.f;
          wacc[warp][c][0] += (double)(is1 ? ll[rr][1] : ll[rr][0]);
          wacc[warp][c][1] += (double)(is1 ? ll[rr][3] : ll[rr][2]);
          wacc[warp][c][2] += (double)(s * (ll[rr][1] - ll[rr][0]));
          wacc[warp][c][3] += (double)(s * (ll[rr][3] - ll[rr][2]));
        }
      }
    }
    if (dx) {
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const int b = b0 + rr;
        if (b >= B) continue;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = lane + 32 * u;
          if (j < z) dx[(int64_t)b * d_ld + j] += gcoef * gacc[rr][u];
        }
      }
    }
  }
  __syncthreads();
  if (acc && (int)threadIdx.x < 4 * nc) {
    const int c = threadIdx.x >> 2, k = threadIdx.x & 3;
    double sum = 0.0;
    for (int w = 0; w < NT / 32; ++w) sum += wacc[w][c][k];
    atomicAdd(acc + threadIdx.x, sum);
  }
}

__global__ void qda_finalize_kernel(const double* __restrict__ acc, float* lama, float* lamb, float delta, float lamdiff, int nc,
                                    int B, double* loss) {
  if (threadIdx.x || blockIdx.x) return;
  double tot = 0.0;
  for (int c = 0; c < nc; ++c) {
    const float lla = (float)acc[4 * c], llb = (float)acc[4 * c + 1];
    if (lla > llb) {
      const float a = fminf(fmaxf(lama[c] - delta, 0.f), 1.f);
      lama[c] = a;
      lamb[c] = a + lamdiff;
    } else {
      const float a = fminf(fmaxf(lamb[c] + delta, 0.f), 1.f);
      lamb[c] = a;
      lama[c] = a - lamdiff;
    }
    tot += (acc[4 * c + 2] + acc[4 * c + 3]) * 0.5;
  }
  if (loss) loss[0] += tot / (double)nc / (double)B;
}

// class-conditional batch means: block per (class, side), thread per feature; stat[(c*2+side)*(z+1) + j], count at [z]
__global__ void __launch_bounds__(NT) qda_mean_kernel(const float* __restrict__ x, int64_t x_ld, const int64_t* __restrict__ y,
                                                      const int64_t* __restrict__ classes, int z, int B, float* __restrict__ stat) {
  const int c = blockIdx.x >> 1, side = blockIdx.x & 1;
  const int64_t label = classes[c];
  float* out = stat + (size_t)blockIdx.x * (z + 1);
  for (int j = threadIdx.x; j <= z; j += NT) {
    float s = 0.f;
    int n = 0;
    for (int b = 0; b < B; ++b) {
      const bool in = (y[b] == label) == (side == 1);
      if (in) { ++n; if (j < z) s += x[(int64_t)b * x_ld + j]; }
    }
    out[j] = j < z ? s / (float)n : (float)n;  // empty subset: 0 / 0 = NaN, as torch.mean of an empty selection
  }
}

// covariance (correction = 0) of each subset around its mean, blended into the four running (mean, covariance) pairs:
// grid (z, 2 nc): block row r of the covariance of (class, side); thread per column
__global__ void __launch_bounds__(NT) qda_update_kernel(const float* __restrict__ x, int64_t x_ld, const int64_t* __restrict__ y,
                                                        const int64_t* __restrict__ classes, int z, int B,
                                                        const float* __restrict__ stat, const float* __restrict__ lama,
                                                        const float* __restrict__ lamb, const QdaMut P) {
  const int r = blockIdx.x, cs = blockIdx.y, c = cs >> 1, side = cs & 1;
  const int col = threadIdx.x;
  if (col >= z) return;
  const int64_t label = classes[c];
  const float* st = stat + (size_t)cs * (z + 1);
  const float mr = st[r], mc = st[col], n = st[z];
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const bool in = (y[b] == label) == (side == 1);
    if (in) s = fmaf(x[(int64_t)b * x_ld + r] - mr, x[(int64_t)b * x_ld + col] - mc, s);
  }
  const float cov = s / n;
  const float la = lama[c], lb = lamb[c];
  float* Sa = P.S[side] + ((size_t)c * z + r) * z + col;      // S0a / S1a
  float* Sb = P.S[2 + side] + ((size_t)c * z + r) * z + col;  // S0b / S1b
  *Sa = (1.f - la) * *Sa + la * cov;
  *Sb = (1.f - lb) * *Sb + lb * cov;
  if (r == 0) {
    float* ma = P.m[side] + (size_t)c * z + col;
    float* mb = P.m[2 + side] + (size_t)c * z + col;
    *ma = (1.f - la) * *ma + la * mc;
    *mb = (1.f - lb) * *mb + lb * mc;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// "moving_avg" scrubber (MovingAverageFilter, reference model/disentangle.py:9-88; loss train/losses.py:286-289; update
// train/trainer.py:169-178): per class two running means m1, m2 of the latent mean with forgetting factors lam1 < lam2.
//   xbar_c = mean of x over the members of class c;  ||xbar_c - m1_c|| < ||xbar_c - m2_c|| ? (lam1 -= delta, lam2 = lam1 +
//   lamdiff) : (lam2 += delta, lam1 = lam2 - lamdiff);  e_c = ((1-lam1) xbar + lam1 m1 + (1-lam2) xbar + lam2 m2) / 2
//   loss = sqrt(sum_{c<c'} ||e_c - e_c'||^2);  d loss / d x_b = (2 - lam1 - lam2)/2 (nc e_c - sum_c' e_c') / loss / n_c
// ---------------------------------------------------------------------------------------------------------------------
// class means: block per class, thread per feature; stat[c*(z+1) + j], member count at [z]
__global__ void __launch_bounds__(NT) ma_mean_kernel(const float* __restrict__ x, int64_t x_ld, const int64_t* __restrict__ y,
                                                     const int64_t* __restrict__ classes, int z, int B, float* __restrict__ stat) {
  const int64_t label = classes[blockIdx.x];
  float* out = stat + (size_t)blockIdx.x * (z + 1);
  for (int j = threadIdx.x; j <= z; j += NT) {
    float s = 0.f;
    int n = 0;
    for (int b = 0; b < B; ++b)
      if (y[b] == label) { ++n; if (j < z) s += x[(int64_t)b * x_ld + j]; }
    out[j] = j < z ? s / (float)n : (float)n;
  }
}

// one block: forgetting factors, estimates, loss, per-class gradient rows coef (nc, z)
__global__ void __launch_bounds__(NT) ma_eval_kernel(const float* __restrict__ stat, const float* __restrict__ m1,
                                                     const float* __restrict__ m2, float* lam1, float* lam2, float delta,
                                                     float lamdiff, int nc, int z, float* __restrict__ coef, double* loss) {
  __shared__ float est[MAXC][MAXZ];
  __shared__ float red[NT / 32][2];
  __shared__ float l1s[MAXC], l2s[MAXC], tot[1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < nc; ++c) {
    const float* xb = stat + (size_t)c * (z + 1);
    float d1 = 0.f, d2 = 0.f;
    for (int j = threadIdx.x; j < z; j += NT) {
      const float a = xb[j] - m1[c * z + j], b = xb[j] - m2[c * z + j];
      d1 += a * a;
      d2 += b * b;
    }
    d1 = scv::warp_sum(d1);
    d2 = scv::warp_sum(d2);
    if (lane == 0) { red[warp][0] = d1; red[warp][1] = d2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float s1 = 0.f, s2 = 0.f;
      for (int w = 0; w < NT / 32; ++w) { s1 += red[w][0]; s2 += red[w][1]; }
      float a = lam1[c], b = lam2[c];
      if (sqrtf(s1) < sqrtf(s2)) {
        a = fminf(fmaxf(a - delta, 0.f), 1.f);
        b = a + lamdiff;
      } else {
        b = fminf(fmaxf(b + delta, 0.f), 1.f);
        a = b - lamdiff;
      }
      lam1[c] = a; lam2[c] = b;
      l1s[c] = a; l2s[c] = b;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < z; j += NT)
      est[c][j] = 0.5f * (((1.f - l1s[c]) * xb[j] + l1s[c] * m1[c * z + j]) + ((1.f - l2s[c]) * xb[j] + l2s[c] * m2[c * z + j]));
    __syncthreads();
  }
  float acc = 0.f;
  for (int j = threadIdx.x; j < z; j += NT)
    for (int a = 0; a < nc; ++a)
      for (int b = a + 1; b < nc; ++b) {
        const float d = est[a][j] - est[b][j];
        acc += d * d;
      }
  acc = scv::warp_sum(acc);
  if (lane == 0) red[warp][0] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < NT / 32; ++w) t += red[w][0];
    tot[0] = sqrtf(t);
    if (loss) loss[0] += (double)tot[0];
  }
  __syncthreads();
  const float nrm = tot[0];
  for (int j = threadIdx.x; j < z; j += NT) {
    float sum = 0.f;
    for (int c = 0; c < nc; ++c) sum += est[c][j];
    for (int c = 0; c < nc; ++c) {
      const float n = stat[(size_t)c * (z + 1) + z];
      coef[c * z + j] = 0.5f * (2.f - l1s[c] - l2s[c]) * ((float)nc * est[c][j] - sum) / nrm / n;
    }
  }
}

__global__ void __launch_bounds__(NT) ma_bwd_kernel(const int64_t* __restrict__ y, const int64_t* __restrict__ classes,
                                                    const float* __restrict__ coef, const float* __restrict__ gscale, int nc, int z,
                                                    int B, float* __restrict__ dx, int64_t d_ld) {
  const float g = gscale ? gscale[0] : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < (int64_t)B * z; i += (int64_t)gridDim.x * NT) {
    const int b = (int)(i / z), j = (int)(i - (int64_t)b * z);
    const int64_t yb = y[b];
    int c = -1;
    for (int k = 0; k < nc; ++k)
      if (classes[k] == yb) c = k;
    if (c >= 0) dx[(int64_t)b * d_ld + j] += g * coef[c * z + j];
  }
}

__global__ void __launch_bounds__(NT) ma_update_kernel(const float* __restrict__ stat, const float* __restrict__ lam1,
                                                       const float* __restrict__ lam2, int nc, int z, float* m1, float* m2) {
  for (int i = threadIdx.x + blockIdx.x * NT; i < nc * z; i += gridDim.x * NT) {
    const int c = i / z, j = i - c * z;
    const float xb = stat[(size_t)c * (z + 1) + j];
    m1[i] = (1.f - lam1[c]) * xb + lam1[c] * m1[i];
    m2[i] = (1.f - lam2[c]) * xb + lam2[c] * m2[i];
  }
}

}  // namespace

extern "C" {

int scv_qda_factor(const float* S0a, const float* S1a, const float* S0b, const float* S1b, int64_t nc, int64_t z, float* SinvT,
                   float* logdet, void* stream) {
  const float* S4[4] = {S0a, S1a, S0b, S1b};
  SCV_REQUIRE(z >= 1 && z <= MAXZ && nc >= 1 && nc <= MAXC, "scv_qda_factor: z <= 128, classes <= 16");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(qda_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(MAXZ * 2 * MAXZ * sizeof(float)));
    if (e != cudaSuccess) {
      scv::set_error("scv_qda_factor: cannot opt in to shared memory: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return (int)e;
    }
    attr = true;
  }
  QdaPtrs P;
  for (int q = 0; q < 4; ++q) { P.S[q] = S4[q]; P.m[q] = nullptr; }
  qda_factor_kernel<<<(unsigned)(4 * nc), NT, (size_t)z * 2 * z * sizeof(float), (cudaStream_t)stream>>>(P, (int)nc, (int)z, SinvT,
                                                                                                       logdet);
  return scv::check_launch("qda_factor_kernel");
}

int scv_qda_loss(const float* x, int64_t x_ld, const int64_t* y, const int64_t* classes, const float* m0a, const float* m1a,
                 const float* m0b, const float* m1b, const float* SinvT, const float* logdet, int64_t nc, int64_t z, int64_t B,
                 double* acc, const float* gscale, float* dx, int64_t d_ld, void* stream) {
  const float* m4[4] = {m0a, m1a, m0b, m1b};
  SCV_REQUIRE(z >= 1 && z <= MAXZ && nc >= 1 && nc <= MAXC, "scv_qda_loss: z <= 128, classes <= 16");
  if (B <= 0) return 0;
  QdaPtrs P;
  for (int q = 0; q < 4; ++q) { P.m[q] = m4[q]; P.S[q] = nullptr; }
  const int64_t rows_per_block = (NT / 32) * RB;
  int blocks = (int)((B + rows_per_block - 1) / rows_per_block);
  const int cap = scv::sm_count() * 2;
  if (blocks > cap) blocks = cap;
  qda_loss_kernel<<<blocks, NT, 0, (cudaStream_t)stream>>>(x, x_ld, y, classes, P, SinvT, logdet, (int)nc, (int)z, (int)B, acc,
                                                          gscale, dx, d_ld);
  return scv::check_launch("qda_loss_kernel");
}

int scv_qda_finalize(const double* acc, float* lama, float* lamb, double delta, double lamdiff, int64_t nc, int64_t B, double* loss,
                     void* stream) {
  qda_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, lama, lamb, (float)delta, (float)lamdiff, (int)nc, (int)B, loss);
  return scv::check_launch("qda_finalize_kernel");
}

int scv_qda_update(const float* x, int64_t x_ld, const int64_t* y, const int64_t* classes, int64_t nc, int64_t z, int64_t B,
                   const float* lama, const float* lamb, float* m0a, float* m1a, float* m0b, float* m1b, float* S0a, float* S1a,
                   float* S0b, float* S1b, float* stat, void* stream) {
  float* m4[4] = {m0a, m1a, m0b, m1b};
  float* S4[4] = {S0a, S1a, S0b, S1b};
  SCV_REQUIRE(z >= 1 && z <= MAXZ && nc >= 1 && nc <= MAXC, "scv_qda_update: z <= 128, classes <= 16");
  QdaMut P;
  for (int q = 0; q < 4; ++q) { P.m[q] = m4[q]; P.S[q] = S4[q]; }
  qda_mean_kernel<<<(unsigned)(2 * nc), NT, 0, (cudaStream_t)stream>>>(x, x_ld, y, classes, (int)z, (int)B, stat);
  int rc = scv::check_launch("qda_mean_kernel");
  if (rc) return rc;
  qda_update_kernel<<<dim3((unsigned)z, (unsigned)(2 * nc)), NT, 0, (cudaStream_t)stream>>>(x, x_ld, y, classes, (int)z, (int)B, stat,
                                                                                          lama, lamb, P);
  return scv::check_launch("qda_update_kernel");
}

int scv_ma_loss(const float* x, int64_t x_ld, const int64_t* y, const int64_t* classes, int64_t nc, int64_t z, int64_t B,
                const float* m1, const float* m2, float* lam1, float* lam2, double delta, double lamdiff, float* stat, float* coef,
                double* loss, void* stream) {
  SCV_REQUIRE(z >= 1 && z <= MAXZ && nc >= 1 && nc <= MAXC, "scv_ma_loss: z <= 128, classes <= 16");
  ma_mean_kernel<<<(unsigned)nc, NT, 0, (cudaStream_t)stream>>>(x, x_ld, y, classes, (int)z, (int)B, stat);
  int rc = scv::check_launch("ma_mean_kernel");
  if (rc) return rc;
  ma_eval_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(stat, m1, m2, lam1, lam2, (float)delta, (float)lamdiff, (int)nc, (int)z, coef,
                                                    loss);
  return scv::check_launch("ma_eval_kernel");
}

int scv_ma_backward(const int64_t* y, const int64_t* classes, const float* coef, const float* gscale, int64_t nc, int64_t z,
                    int64_t B, float* dx, int64_t d_ld, void* stream) {
  if (B <= 0) return 0;
  int blocks = (int)((B * z + NT - 1) / NT);
  const int cap = scv::sm_count() * 4;
  if (blocks > cap) blocks = cap;
  ma_bwd_kernel<<<blocks, NT, 0, (cudaStream_t)stream>>>(y, classes, coef, gscale, (int)nc, (int)z, (int)B, dx, d_ld);
  return scv::check_launch("ma_bwd_kernel");
}

int scv_ma_update(const float* x, int64_t x_ld, const int64_t* y, const int64_t* classes, int64_t nc, int64_t z, int64_t B,
                  const float* lam1, const float* lam2, float* m1, float* m2, float* stat, void* stream) {
  SCV_REQUIRE(z >= 1 && z <= MAXZ && nc >= 1 && nc <= MAXC, "scv_ma_update: z <= 128, classes <= 16");
  ma_mean_kernel<<<(unsigned)nc, NT, 0, (cudaStream_t)stream>>>(x, x_ld, y, classes, (int)z, (int)B, stat);
  int rc = scv::check_launch("ma_mean_kernel");
  if (rc) return rc;
  ma_update_kernel<<<(unsigned)((nc * z + NT - 1) / NT), NT, 0, (cudaStream_t)stream>>>(stat, lam1, lam2, (int)nc, (int)z, m1, m2);
  return scv::check_launch("ma_update_kernel");
}

}  // extern "C"
