// fp32 (FFMA) overlapping-row GEMM + weight-gradient kernels.
// This is the exact-fp32 path (SCV_PREC_FP32) and the path for shapes the tcgen05 kernels do
// not take (tiny N/K of the scrubber heads).  128x128x16 tiles, 256 threads, 8x8 per thread,
// register prefetch + double-buffered shared memory.
#include "scv_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, LDS = BM + 4;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float apply_act(float v, int act, float r) {
  if (act == SCV_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == SCV_ACT_TANH) return tanhf(v);
  if (act == SCV_ACT_RELUMASK) return r > 0.f ? v : 0.f;
  return v;
}

constexpr int kMaxGroup = 6;  // independent problems per grouped launch (scv_gemm_group / scv_wgrad_group)
struct GemmGroupArgs { scv_gemm_t p[kMaxGroup]; int vec[kMaxGroup]; };
struct WgradGroupArgs { scv_wgrad_t p[kMaxGroup]; int64_t rps[kMaxGroup]; int vecy[kMaxGroup]; int zbeg[kMaxGroup + 1]; };

__device__ __forceinline__ void gemm_ffma_body(const scv_gemm_t& p, const int vec) {
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Bs[2][BK][LDS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t M = p.B * p.Lo;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = (int)p.K, N = (int)p.N;

  const int kq = tid & 3;
  const float* aptr[2];
  const float* wptr[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int row = (tid + i * NT) >> 2;
    int64_t m = m0 + row;
    if (m < M) {
      int64_t b = m / p.Lo, l = m - b * p.Lo;
      aptr[i] = p.A + b * p.a_bs + l * p.a_ls;
    } else {
      aptr[i] = nullptr;
    }
    int n = n0 + row;
    wptr[i] = n < N ? p.W + (int64_t)n * K : nullptr;
  }
  float4 ra[2], rb[2];
  auto load = [&](int k0) {
    int kk = k0 + kq * 4;
    bool kok = kk < K;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ra[i] = (kok && aptr[i]) ? ldg4(aptr[i] + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = (kok && wptr[i]) ? ldg4(wptr[i] + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int row = (tid + i * NT) >> 2;
      As[buf][kq * 4 + 0][row] = ra[i].x; As[buf][kq * 4 + 1][row] = ra[i].y;
      As[buf][kq * 4 + 2][row] = ra[i].z; As[buf][kq * 4 + 3][row] = ra[i].w;
      Bs[buf][kq * 4 + 0][row] = rb[i].x; Bs[buf][kq * 4 + 1][row] = rb[i].y;
      Bs[buf][kq * 4 + 2][row] = rb[i].z; Bs[buf][kq * 4 + 3][row] = rb[i].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load(0);
  store(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  const float osc = (float)p.out_scale;
  const int act = (int)(p.act & 15);
  const bool rnd = p.act & SCV_ACT_ROUND_TF32;
  const bool accum = p.act & SCV_ACT_ACCUM;  // Y += result (one thread owns each output element here: plain read-add-write)
  float csum[8], csq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) csum[j] = csq[j] = 0.f;
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
    bv[j] = (p.bias && n < p.bias_n) ? __ldg(p.bias + (n % p.bias_mod)) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
    int64_t b = m / p.Lo, l = m - b * p.Lo;
    const int ncap = (l == p.Lo - 1) ? (int)p.n_last : N;
    float* yrow = p.Y + b * p.y_bs + l * p.y_ls;
    const float* rrow = p.R ? p.R + b * p.r_bs + l * p.r_ls : nullptr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int nb = n0 + (h ? 64 + tx * 4 : tx * 4);
      if (nb >= ncap) continue;
      float v[4], r[4] = {0.f, 0.f, 0.f, 0.f};
      if (rrow) {
        if (vec) {
          float4 t = *reinterpret_cast<const float4*>(rrow + nb);
          r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) r[q] = (nb + q < ncap) ? rrow[nb + q] : 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float t = osc * acc[i][h * 4 + q] + bv[h * 4 + q];
        if (act != SCV_ACT_RELUMASK) t += r[q];
        if (nb + q < ncap) { csum[h * 4 + q] += t; csq[h * 4 + q] += t * t; }
        v[q] = apply_act(t, act, r[q]);
        if (rnd) v[q] = scv::round_tf32(v[q]);
      }
      if (accum) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (nb + q < ncap) yrow[nb + q] += v[q];
      } else if (vec) {
        *reinterpret_cast<float4*>(yrow + nb) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (nb + q < ncap) yrow[nb + q] = v[q];
      }
    }
  }
  if (p.stats) {
    float* red0 = &As[0][0][0];  // [16][128]
    float* red1 = &Bs[0][0][0];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      red0[ty * BN + c] = csum[j];
      red1[ty * BN + c] = csq[j];
    }
    __syncthreads();
    if (tid < BN && n0 + tid < N) {
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int t = 0; t < 16; ++t) { s += (double)red0[t * BN + tid]; q += (double)red1[t * BN + tid]; }
      atomicAdd(p.stats + n0 + tid, s);
      atomicAdd(p.stats + N + n0 + tid, q);
    }
  }
}

__global__ void __launch_bounds__(NT, 2) gemm_ffma_kernel(const scv_gemm_t p, const int vec) { gemm_ffma_body(p, vec); }

// several small independent problems in one launch: blockIdx.z picks the problem
__global__ void __launch_bounds__(NT, 2) gemm_ffma_group_kernel(const __grid_constant__ GemmGroupArgs g) {
  const scv_gemm_t& p = g.p[blockIdx.z];
  if ((int64_t)blockIdx.x * BM >= p.B * p.Lo || (int64_t)blockIdx.y * BN >= p.N) return;
  gemm_ffma_body(p, g.vec[blockIdx.z]);
}

// dW[n][k] += sum_m dY[m][n] * A[m][k] over this block's m-range (split-M, fp32 atomics)
__device__ __forceinline__ void wgrad_ffma_body(const scv_wgrad_t& p, const int64_t rows_per_split, const int vecy,
                                                const int split) {
  __shared__ __align__(16) float Ys[2][BK][LDS];
  __shared__ __align__(16) float As[2][BK][LDS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t M = p.B * p.Lo;
  const int k0 = blockIdx.x * BM;  // K tile
  const int n0 = blockIdx.y * BN;  // N tile
  const int64_t mbeg = (int64_t)split * rows_per_split;
  const int64_t mend = mbeg + rows_per_split < M ? mbeg + rows_per_split : M;
  const int K = (int)p.K, N = (int)p.N;
  const int c4 = (tid & 31) * 4;

  float4 ra[2], ry[2];
  auto load = [&](int64_t mb) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = (tid + i * NT) >> 5;
      int64_t m = mb + r;
      ra[i] = ry[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < mend) {
        int64_t b = m / p.Lo, l = m - b * p.Lo;
        if (k0 + c4 < K) ra[i] = ldg4(p.A + b * p.a_bs + l * p.a_ls + k0 + c4);
        const float* yp = p.dY + b * p.y_bs + l * p.y_ls + n0 + c4;
        if (vecy) {
          if (n0 + c4 < N) ry[i] = ldg4(yp);
        } else {
          if (n0 + c4 + 0 < N) ry[i].x = __ldg(yp + 0);
          if (n0 + c4 + 1 < N) ry[i].y = __ldg(yp + 1);
          if (n0 + c4 + 2 < N) ry[i].z = __ldg(yp + 2);
          if (n0 + c4 + 3 < N) ry[i].w = __ldg(yp + 3);
        }
      }
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = (tid + i * NT) >> 5;
      *reinterpret_cast<float4*>(&As[buf][r][c4]) = ra[i];
      *reinterpret_cast<float4*>(&Ys[buf][r][c4]) = ry[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float bsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
  const bool do_bias = p.dbias != nullptr && blockIdx.x == 0 && tx == 0;

  if (mbeg < mend) {
    const int nit = (int)((mend - mbeg + BK - 1) / BK);
    load(mbeg);
    store(0);
    __syncthreads();
    for (int it = 0; it < nit; ++it) {
      const int buf = it & 1;
      if (it + 1 < nit) load(mbeg + (int64_t)(it + 1) * BK);
#pragma unroll
      for (int r = 0; r < BK; ++r) {
        float4 y0 = *reinterpret_cast<const float4*>(&Ys[buf][r][ty * 4]);
        float4 y1 = *reinterpret_cast<const float4*>(&Ys[buf][r][64 + ty * 4]);
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][r][tx * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][r][64 + tx * 4]);
        float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(y[i], a[j], acc[i][j]);
          bsum[i] += y[i];
        }
      }
      if (it + 1 < nit) store(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int n = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (n >= N) continue;
    float* wrow = p.dW + (int64_t)n * K;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (k < K) atomicAdd(wrow + k, acc[i][j]);
    }
    if (do_bias && n < p.bias_n) atomicAdd(p.dbias + (n % p.bias_mod), bsum[i]);
  }
}

__global__ void __launch_bounds__(NT, 2) wgrad_ffma_kernel(const scv_wgrad_t p, const int64_t rows_per_split,
                                                        const int vecy) {
  wgrad_ffma_body(p, rows_per_split, vecy, blockIdx.z);
}

__global__ void __launch_bounds__(NT, 2) wgrad_ffma_group_kernel(const __grid_constant__ WgradGroupArgs g) {
  int i = 0;
  while (i + 1 < kMaxGroup && (int)blockIdx.z >= g.zbeg[i + 1]) ++i;
  const scv_wgrad_t& p = g.p[i];
  if ((int64_t)blockIdx.x * BM >= p.K || (int64_t)blockIdx.y * BN >= p.N) return;
  wgrad_ffma_body(p, g.rps[i], g.vecy[i], (int)blockIdx.z - g.zbeg[i]);
}

}  // namespace

namespace scv {

static int gemm_vec(const scv_gemm_t* p) {
  int vec = p->N % 4 == 0 && p->n_last % 4 == 0 && p->y_bs % 4 == 0 && p->y_ls % 4 == 0 && aligned16(p->Y);
  if (p->R) vec = vec && p->r_bs % 4 == 0 && p->r_ls % 4 == 0 && aligned16(p->R);
  return vec;
}

int gemm_ffma_group(const scv_gemm_t* p, int n, cudaStream_t st) {
  SCV_REQUIRE(n >= 1 && n <= kMaxGroup, "scv_gemm_group: 1..%d problems per launch", kMaxGroup);
  GemmGroupArgs g;
  unsigned gx = 1, gy = 1;
  for (int i = 0; i < n; ++i) {
    SCV_REQUIRE(p[i].K % 4 == 0 && p[i].a_bs % 4 == 0 && p[i].a_ls % 4 == 0 && aligned16(p[i].A) && aligned16(p[i].W),
                "scv_gemm_group: A/W rows must be 16-byte aligned (problem %d)", i);
    g.p[i] = p[i];
    g.vec[i] = gemm_vec(p + i);
    const unsigned x = (unsigned)((p[i].B * p[i].Lo + BM - 1) / BM), y = (unsigned)((p[i].N + BN - 1) / BN);
    gx = x > gx ? x : gx;
    gy = y > gy ? y : gy;
  }
  SCV_REQUIRE(gy <= 65535, "scv_gemm_group: N too large");
  gemm_ffma_group_kernel<<<dim3(gx, gy, (unsigned)n), NT, 0, st>>>(g);
  return check_launch("gemm_ffma_group_kernel");
}

int gemm_ffma(const scv_gemm_t* p, cudaStream_t st) {
  const int64_t M = p->B * p->Lo;
  SCV_REQUIRE(p->K % 4 == 0 && p->a_bs % 4 == 0 && p->a_ls % 4 == 0 && aligned16(p->A) && aligned16(p->W),
              "scv_gemm: A/W rows must be 16-byte aligned (K=%lld a_bs=%lld a_ls=%lld)", (long long)p->K,
              (long long)p->a_bs, (long long)p->a_ls);
  const int vec = gemm_vec(p);
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((p->N + BN - 1) / BN));
  SCV_REQUIRE(grid.y <= 65535, "scv_gemm: N too large");
  gemm_ffma_kernel<<<grid, NT, 0, st>>>(*p, vec);
  return check_launch("gemm_ffma_kernel");
}

// row splits of one weight-gradient problem: ~`ctas` CTAs in flight over its tiles, at least 8 k-steps per split
static void wgrad_splits(const scv_wgrad_t* p, int64_t ctas, int64_t& S, int64_t& rps) {
  const int64_t M = p->B * p->Lo;
  int64_t tiles = ((p->K + BM - 1) / BM) * ((p->N + BN - 1) / BN);
  int64_t want = (ctas + tiles - 1) / tiles;
  int64_t maxs = (M + 8 * BK - 1) / (8 * BK);
  S = want < 1 ? 1 : want;
  if (S > maxs) S = maxs;
  if (S > 65535) S = 65535;
  if (S < 1) S = 1;
  rps = (M + S - 1) / S;
  rps = (rps + BK - 1) / BK * BK;
  S = (M + rps - 1) / rps;
}

int wgrad_ffma_group(const scv_wgrad_t* p, int n, cudaStream_t st) {
  SCV_REQUIRE(n >= 1 && n <= kMaxGroup, "scv_wgrad_group: 1..%d problems per launch", kMaxGroup);
  WgradGroupArgs g;
  unsigned gx = 1, gy = 1;
  int z = 0;
  for (int i = 0; i < kMaxGroup + 1; ++i) g.zbeg[i] = 0x7fffffff;
  for (int i = 0; i < n; ++i) {
    SCV_REQUIRE(p[i].K % 4 == 0 && p[i].a_bs % 4 == 0 && p[i].a_ls % 4 == 0 && aligned16(p[i].A),
                "scv_wgrad_group: A rows must be 16-byte aligned (problem %d)", i);
    g.p[i] = p[i];
    g.vecy[i] = p[i].N % 4 == 0 && p[i].y_bs % 4 == 0 && p[i].y_ls % 4 == 0 && aligned16(p[i].dY);
    int64_t S, rps;
    wgrad_splits(p + i, 4LL * sm_count() / n, S, rps);
    g.rps[i] = rps;
    g.zbeg[i] = z;
    z += (int)S;
    const unsigned x = (unsigned)((p[i].K + BM - 1) / BM), y = (unsigned)((p[i].N + BN - 1) / BN);
    gx = x > gx ? x : gx;
    gy = y > gy ? y : gy;
  }
  SCV_REQUIRE(gy <= 65535 && z <= 65535, "scv_wgrad_group: problem too large");
  wgrad_ffma_group_kernel<<<dim3(gx, gy, (unsigned)z), NT, 0, st>>>(g);
  return check_launch("wgrad_ffma_group_kernel");
}

int wgrad_ffma(const scv_wgrad_t* p, cudaStream_t st) {
  const int64_t M = p->B * p->Lo;
  SCV_REQUIRE(p->K % 4 == 0 && p->a_bs % 4 == 0 && p->a_ls % 4 == 0 && aligned16(p->A),
              "scv_wgrad: A rows must be 16-byte aligned");
  int vecy = p->N % 4 == 0 && p->y_bs % 4 == 0 && p->y_ls % 4 == 0 && aligned16(p->dY);
  int64_t S, rps;
  wgrad_splits(p, 4LL * sm_count(), S, rps);
  (void)M;
  dim3 grid((unsigned)((p->K + BM - 1) / BM), (unsigned)((p->N + BN - 1) / BN), (unsigned)S);
  SCV_REQUIRE(grid.y <= 65535, "scv_wgrad: N too large");
  wgrad_ffma_kernel<<<grid, NT, 0, st>>>(*p, rps, vecy);
  return check_launch("wgrad_ffma_kernel");
}

}  // namespace scv
