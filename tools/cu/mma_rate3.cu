// tcgen05.mma issue / execution rate with PRECOMPUTED descriptors (the issuing thread only adds to the low word), for
//   cta_group::1 (M=128) and cta_group::2 (CTA pair, M=256), kind::tf32 and kind::f16 (bf16), N = 128 / 256,
// alone and while a second warp streams bulk copies (L2 -> shared memory) into the same SM — the question being whether
// operand reads (MMA) and TMA writes share one 128 B/cycle shared-memory port, and whether pairing (each CTA holds half
// of the B operand) lifts the rate.     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I scrubvae_b200/csrc ...
#include <cstdio>
#include <cuda_runtime.h>
#include "scv_tc.cuh"
using namespace scv::tc;
namespace scv { void set_error(const char*, ...) {} int64_t g_launches = 0; int sm_count() { return 148; } }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode bit0: cta_group::2, bit1: bf16, bit2: concurrent bulk copies; `sub2`: alternate two A tiles / two accumulators
template <int kCg>
__device__ void body(int N, int iters, long long* out, int bf16, int traffic, int sub2, const float* gsrc, int copy_kb) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, cbar[4];
  __shared__ uint32_t tbase;
  __shared__ volatile int stop;
  const uint32_t rank = kCg == 2 ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&cbar[i]), 1);
    stop = 0;
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { if (kCg == 2) tmem_alloc2(smem_u32(&tbase), 512); else tmem_alloc(smem_u32(&tbase), 512); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long copied = 0;
  if (warp == 1 && lane == 0 && traffic) {
    // ring of 4 buffers of copy_kb KB behind the operand region (96 KB): keep 4 copies in flight until told to stop
    const uint32_t cb = smem_u32(smem) + 96 * 1024;
    const uint32_t bytes = (uint32_t)copy_kb * 1024;
    uint32_t ph[4] = {0, 0, 0, 0};
    const char* src = reinterpret_cast<const char*>(gsrc) + (size_t)(blockIdx.x % 16) * 65536;
    for (int i = 0; i < 4; ++i) { mbar_expect_tx(smem_u32(&cbar[i]), bytes); bulk_g2s(cb + i * bytes, src + i * bytes, bytes, smem_u32(&cbar[i])); }
    int i = 0;
    while (!stop) {
      mbar_wait(smem_u32(&cbar[i]), ph[i]);
      ph[i] ^= 1;
      copied += bytes;
      mbar_expect_tx(smem_u32(&cbar[i]), bytes);
      bulk_g2s(cb + i * bytes, src + i * bytes, bytes, smem_u32(&cbar[i]));
      i = (i + 1) & 3;
    }
    for (int j = 0; j < 4; ++j) mbar_wait(smem_u32(&cbar[j]), ph[j]);
    out[148 + blockIdx.x] = copied;
  }
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = bf16 ? idesc_bf16(128 * kCg, N) : idesc_tf32(128 * kCg, N, 0, 0);
    const uint64_t d0 = smem_desc(smem_u32(smem), 16, 1024);
    const uint32_t lo0 = (uint32_t)d0, hi = (uint32_t)(d0 >> 32);
    const uint32_t a_units = (16384u * 2) >> 4;  // B operand behind two A tiles
    const uint32_t stage_units = (96u * 1024 / 1) >> 4;
    (void)stage_units;
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      // one "stage": 4 k-steps of 32 B inside the 128-byte swizzle span
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_lo = lo0 + 2 * k, b_lo = lo0 + a_units + 2 * k;
        if (kCg == 2) {
          if (bf16) asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tbase), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(hi), "r"(idesc), "r"(1u) : "memory");
          else umma2_tf32_lh(tbase, a_lo, hi, b_lo, hi, idesc, 1u);
        } else {
          if (bf16) asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tbase), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(hi), "r"(idesc), "r"(1u) : "memory");
          else umma_tf32_lh(tbase, a_lo, hi, b_lo, hi, idesc, 1u);
          if (sub2) {
            const uint32_t a2 = a_lo + (16384u >> 4);
            if (bf16) asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                                   "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tbase + 256), "r"(a2), "r"(hi), "r"(b_lo), "r"(hi), "r"(idesc), "r"(1u) : "memory");
            else umma_tf32_lh(tbase + 256, a2, hi, b_lo, hi, idesc, 1u);
          }
        }
      }
    }
    if (kCg == 2) umma2_commit(smem_u32(&bar)); else umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x / kCg] = t1 - t0;
    stop = 1;
  }
  if (kCg == 2 && threadIdx.x == 0 && rank == 1 && traffic) {
    // the peer's copy warp stops when the leader is done: poll the leader's flag through the cluster
    // (simplest: fixed spin on clock, the leader's run time is bounded)
    long long t0 = clock64();
    while (clock64() - t0 < (long long)iters * 140) {}
    stop = 1;
  }
  __syncthreads();
  tc_fence_before();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); if (kCg == 2) tmem_dealloc2(tbase, 512); else tmem_dealloc(tbase, 512); }
}

__global__ void __launch_bounds__(128, 1) k1(int N, int iters, long long* out, int bf16, int traffic, int sub2, const float* g, int kb) {
  body<1>(N, iters, out, bf16, traffic, sub2, g, kb);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(int N, int iters, long long* out, int bf16, int traffic, int sub2, const float* g, int kb) {
  body<2>(N, iters, out, bf16, traffic, sub2, g, kb);
}

int main() {
  long long* d; cudaMalloc(&d, 2 * 148 * 8);
  float* g; cudaMalloc(&g, 4 << 20); cudaMemset(g, 0, 4 << 20);
  const int smem = 96 * 1024 + 4 * 24 * 1024 + 2048;
  cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 8000;
  printf("clock-cycles per MMA instruction (per SM for cta_group::1, per PAIR for cta_group::2); 148 CTAs\n");
  for (int bf16 = 0; bf16 < 2; ++bf16)
    for (int cg = 1; cg <= 2; ++cg)
      for (int N : {128, 256})
        for (int sub2 = 0; sub2 < (cg == 1 ? 2 : 1); ++sub2)
          for (int kb : {0, 8, 16, 24}) {
            const int traffic = kb > 0;
            cudaMemset(d, 0, 2 * 148 * 8);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
              cudaEventRecord(e0);
              if (cg == 1) k1<<<148, 128, smem>>>(N, iters, d, bf16, traffic, sub2, g, kb ? kb : 8);
              else k2<<<148, 128, smem>>>(N, iters, d, bf16, traffic, sub2, g, kb ? kb : 8);
              cudaEventRecord(e1);
            }
            cudaError_t e = cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            const double n_mma = (double)iters * (sub2 ? 2 : 1);
            const double cyc = (double)h[0] / n_mma;
            const int kk = bf16 ? 16 : 8;
            const double flop_per_mma = 2.0 * 128 * cg * N * kk;
            const double tf_chip = flop_per_mma * n_mma * (148 / cg) / ((double)h[0] / 1.965e9) / 1e12;
            const double copy_bpc = traffic ? (double)h[148] / (double)h[0] : 0.0;
            printf("%s cta_group::%d N %3d %s copies %2d KB: %6.1f cyc/MMA  -> %6.0f TFLOP/s chip @1.965GHz; bulk-copy %5.1f B/cyc/SM  (%s)\n",
                   bf16 ? "bf16" : "tf32", cg, N, sub2 ? "2 A tiles" : "1 A tile ", kb, cyc, tf_chip, copy_bpc, cudaGetErrorString(e));
          }
  return 0;
}
