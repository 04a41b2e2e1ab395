// Main-loop throughput of a TMA -> tcgen05.mma pipeline on STREAMING operands (every chunk is new data from L2):
//   mode 0  one CTA per SM, cta_group::1: per chunk A 128x32 (16 KB) + W 256x32 (32 KB), 4 MMAs M=128 N=256
//   mode 1  the same with two A tiles per chunk (64 KB, 8 MMAs)                                ("sub = 2")
//   mode 2  CTA pair, cta_group::2: per CTA per chunk A 128x32 (16 KB) + half of W 128x32 (16 KB), 4 MMAs M=256 N=256
//   mode 3  CTA pair with two A tiles per CTA per chunk (32 KB + 16 KB, 8 MMAs M=256)
// Operands are TF32 zeros in a 64 MB global array (L2-resident after the first pass).  Prints cycles per chunk on CTA 0.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "scv_tc.cuh"
using namespace scv::tc;
namespace scv { void set_error(const char* f, ...) { fprintf(stderr, "%s\n", f); } int64_t g_launches = 0; int sm_count() { return 148; } }

static int tmap2d(CUtensorMap* tm, const float* base, int64_t K, int64_t rows, int box_rows) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto fn = (encode_tiled_fn)p;
  cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)rows}, gs[1] = {(cuuint64_t)K * 4};
  cuuint32_t bx[2] = {32, (cuuint32_t)box_rows}, es[2] = {1, 1};
  return (int)fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

constexpr int kStagesMax = 6;
struct Ctl { uint64_t full[kStagesMax], empty[kStagesMax], done; uint32_t tmem; };

template <int kCg>
__device__ void body(const CUtensorMap& tmA, const CUtensorMap& tmW, int sub, int chunks, int stages, int rows_total, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kCg == 2 ? cluster_ctarank() : 0u;
  const uint32_t a_bytes = 16384u * sub, w_rows = 256 / kCg, w_bytes = w_rows * 128u, stage_bytes = a_bytes + w_bytes;
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + (size_t)stages * stage_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&ctl->full[s]), 1); mbar_init(smem_u32(&ctl->empty[s]), 1); }
    mbar_init(smem_u32(&ctl->done), 1);
    fence_barrier_init();
  }
  if (warp == 2) { if (kCg == 2) tmem_alloc2(smem_u32(&ctl->tmem), 512); else tmem_alloc(smem_u32(&ctl->tmem), 512); }
  tc_fence_before();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl->tmem;
  // every CTA streams its own row range of A and its own W rows, wrapping inside the array
  const int row0 = (int)((blockIdx.x * 977) % (rows_total - 1024));
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int c = 0; c < chunks; ++c) {
      mbar_wait(smem_u32(&ctl->empty[s]), ph ^ 1);
      const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
      const int kc = (c * 32) % 4096, r = row0 + ((c / 128) * 256) % 512;
      if (kCg == 2) {
        const uint32_t fb = mapa_rank(smem_u32(&ctl->full[s]), 0);
        if (rank == 0) mbar_expect_tx(smem_u32(&ctl->full[s]), 2 * stage_bytes);
        for (int j = 0; j < sub; ++j) tma2_load_2d(sa + j * 16384, &tmA, fb, kc, r + j * 128);
        tma2_load_2d(sa + a_bytes, &tmW, fb, kc, (r + 300) % (rows_total - 256) + (int)rank * 128);
      } else {
        const uint32_t fb = smem_u32(&ctl->full[s]);
        mbar_expect_tx(fb, stage_bytes);
        for (int j = 0; j < sub; ++j) tma_load_2d(sa + j * 16384, &tmA, fb, kc, r + j * 128);
        tma_load_2d(sa + a_bytes, &tmW, fb, kc, (r + 300) % (rows_total - 256));
      }
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = idesc_tf32(128 * kCg, 256, 0, 0);
    const uint64_t d0 = smem_desc(smem_u32(smem), 16, 1024);
    const uint32_t lo0 = (uint32_t)d0, hi = (uint32_t)(d0 >> 32);
    int s = 0; uint32_t ph = 0;
    long long t0 = 0;
    for (int c = 0; c < chunks; ++c) {
      mbar_wait(smem_u32(&ctl->full[s]), ph);
      tc_fence_after();
      if (c == 16) t0 = clock64();
      const uint32_t a_lo = lo0 + (uint32_t)s * (stage_bytes >> 4), b_lo = a_lo + (a_bytes >> 4);
      for (int j = 0; j < sub; ++j) {
        const uint32_t aj = a_lo + j * 1024, dj = tmem + j * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kCg == 2) umma2_tf32_lh(dj, aj + 2 * k, hi, b_lo + 2 * k, hi, idesc, 1u);
          else umma_tf32_lh(dj, aj + 2 * k, hi, b_lo + 2 * k, hi, idesc, 1u);
        }
      }
      if (kCg == 2) umma2_commit(smem_u32(&ctl->empty[s])); else umma_commit(smem_u32(&ctl->empty[s]));
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    if (kCg == 2) umma2_commit(smem_u32(&ctl->done)); else umma_commit(smem_u32(&ctl->done));
    mbar_wait(smem_u32(&ctl->done), 0);
    out[blockIdx.x / kCg] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_before();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) { tc_fence_after(); if (kCg == 2) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
}
__global__ void __launch_bounds__(128, 1) k1(const __grid_constant__ CUtensorMap a, const __grid_constant__ CUtensorMap w, int sub, int chunks, int stages, int rows, long long* out) { body<1>(a, w, sub, chunks, stages, rows, out); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(const __grid_constant__ CUtensorMap a, const __grid_constant__ CUtensorMap w, int sub, int chunks, int stages, int rows, long long* out) { body<2>(a, w, sub, chunks, stages, rows, out); }

int main() {
  const int64_t K = 4096, rows = 4096;
  float* g; cudaMalloc(&g, K * rows * 4); cudaMemset(g, 0, K * rows * 4);
  long long* d; cudaMalloc(&d, 148 * 8);
  CUtensorMap tmA, tmW1, tmW2;
  if (tmap2d(&tmA, g, K, rows, 128) || tmap2d(&tmW1, g, K, rows, 256) || tmap2d(&tmW2, g, K, rows, 128)) { printf("tmap failed\n"); return 1; }
  const int smem = 227 * 1024;
  cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int chunks = 2000;
  for (int mode = 0; mode < 4; ++mode) {
    const int cg = mode >= 2 ? 2 : 1, sub = (mode & 1) ? 2 : 1;
    const int stage_bytes = 16384 * sub + (256 / cg) * 128;
    int stages = (smem - 2048) / stage_bytes; if (stages > kStagesMax) stages = kStagesMax;
    for (int rep = 0; rep < 2; ++rep) {
      if (cg == 1) k1<<<148, 128, smem>>>(tmA, tmW1, sub, chunks, stages, (int)rows, d);
      else k2<<<148, 128, smem>>>(tmA, tmW2, sub, chunks, stages, (int)rows, d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0] / (chunks - 16);
    const double mma_cyc = 512.0 * sub;  // 4 MMAs x 128 cycles per A tile
    printf("mode %d (cta_group::%d, %d A tile%s): %d stages of %3d KB per CTA, %7.1f cycles/chunk (MMA floor %4.0f) -> %4.0f%% of the tensor pipe, "
           "ingest %5.1f B/cycle/SM  (%s)\n", mode, cg, sub, sub > 1 ? "s" : " ", stages, stage_bytes / 1024, cyc, mma_cyc, 100.0 * mma_cyc / cyc,
           stage_bytes / cyc, cudaGetErrorString(e));
  }
  return 0;
}
