// tcgen05.mma execution rate, branch-free issue loop (templated), precomputed descriptor words:
// cta_group::1 (M=128) vs cta_group::2 (pair, M=256), kind::tf32 / kind::f16 (bf16), N = 64 / 128 / 256, K-major operands;
// one or two accumulators.  Cycles are per MMA INSTRUCTION (per pair for cta_group::2).
#include <cstdio>
#include <cuda_runtime.h>
#include "scv_tc.cuh"
using namespace scv::tc;
namespace scv { void set_error(const char*, ...) {} int64_t g_launches = 0; int sm_count() { return 148; } }
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int kCg, int kBf16>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  if (kCg == 2) {
    if (kBf16) asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %4, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
  } else {
    if (kBf16) asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
  }
}
template <int kCg, int kBf16, int kAcc>
__device__ void body(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t rank = kCg == 2 ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { if (kCg == 2) tmem_alloc2(smem_u32(&tbase), 512); else tmem_alloc(smem_u32(&tbase), 512); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = kBf16 ? idesc_bf16(128 * kCg, N) : idesc_tf32(128 * kCg, N, 0, 0);
    const uint64_t d0 = smem_desc(smem_u32(smem), 16, 1024);
    const uint32_t a0 = (uint32_t)d0, hi = (uint32_t)(d0 >> 32), b0 = a0 + (32768u >> 4), a1 = a0 + (16384u >> 4);
    const uint32_t t0a = tbase, t1a = tbase + 256;
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mma<kCg, kBf16>(t0a, a0 + 2 * k, b0 + 2 * k, hi, idesc);
        mma<kCg, kBf16>(kAcc == 2 ? t1a : t0a, (kAcc == 2 ? a1 : a0) + 2 * k, b0 + 2 * k, hi, idesc);
      }
    }
    if (kCg == 2) umma2_commit(smem_u32(&bar)); else umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x / kCg] = t1 - t0;
  }
  tc_fence_before();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); if (kCg == 2) tmem_dealloc2(tbase, 512); else tmem_dealloc(tbase, 512); }
}
#define KERN(name, cg, bf, acc, ...) __global__ void __VA_ARGS__ __launch_bounds__(128, 1) name(int N, int iters, long long* out) { body<cg, bf, acc>(N, iters, out); }
KERN(k_1_t_1, 1, 0, 1) KERN(k_1_t_2, 1, 0, 2) KERN(k_1_b_1, 1, 1, 1) KERN(k_1_b_2, 1, 1, 2)
KERN(k_2_t_1, 2, 0, 1, __cluster_dims__(2, 1, 1)) KERN(k_2_t_2, 2, 0, 2, __cluster_dims__(2, 1, 1))
KERN(k_2_b_1, 2, 1, 1, __cluster_dims__(2, 1, 1)) KERN(k_2_b_2, 2, 1, 2, __cluster_dims__(2, 1, 1))
typedef void (*kfn)(int, int, long long*);
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  struct { kfn f; int cg, bf, acc; } ks[] = {{k_1_t_1, 1, 0, 1}, {k_1_t_2, 1, 0, 2}, {k_2_t_1, 2, 0, 1}, {k_2_t_2, 2, 0, 2},
                                            {k_1_b_1, 1, 1, 1}, {k_1_b_2, 1, 1, 2}, {k_2_b_1, 2, 1, 1}, {k_2_b_2, 2, 1, 2}};
  const int iters = 8000, smem = 66 * 1024;
  for (auto& k : ks) {
    cudaFuncSetAttribute(k.f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int N : {64, 128, 256}) {
      for (int rep = 0; rep < 2; ++rep) k.f<<<148, 128, smem>>>(N, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double cyc = (double)h / iters;
      const double flop = 2.0 * 128 * k.cg * N * (k.bf ? 16 : 8);
      printf("%s cta_group::%d N %3d accumulators %d: %6.1f cyc/MMA -> %6.0f TFLOP/s chip at 1.965 GHz (%s)\n", k.bf ? "bf16" : "tf32", k.cg, N,
             k.acc, cyc, flop * (148 / k.cg) / cyc * 1.965e9 / 1e12, cudaGetErrorString(e));
    }
  }
  return 0;
}
