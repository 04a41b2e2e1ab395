// tcgen05.mma execution rate with MN-MAJOR operands (the weight-gradient kernel's layout: both operands transposed,
// slabs of R rows x 128 B; fp32: 128-byte swizzle with 32-byte atoms (layout type 1, SBO 512 B), bf16: plain 128-byte
// swizzle (layout type 2, SBO 1024 B)) against the K-major rate of mma_rate4.cu.  Zero operands resident in shared memory,
// branch-free issue loop; cycles per MMA instruction.
#include <cstdio>
#include <cuda_runtime.h>
#include "scv_tc.cuh"
using namespace scv::tc;
namespace scv { void set_error(const char*, ...) {} int64_t g_launches = 0; int sm_count() { return 148; } }
template <int kBf16>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  if (kBf16) asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
  else asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, 1;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
}
// mode 0: both MN-major; 1: A MN-major, B K-major; 2: A K-major, B MN-major; 3: both K-major
template <int kBf16>
__global__ void __launch_bounds__(128, 1) kern(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tbase), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const bool at = mode == 0 || mode == 1, bt = mode == 0 || mode == 2;
    const uint32_t idesc = kBf16 ? idesc_bf16(128, N, at, bt) : idesc_tf32(128, N, at, bt);
    constexpr int R = kBf16 ? 64 : 32;             // reduction rows per stage
    const uint32_t slab = R * 128;
    const uint64_t dmn = kBf16 ? smem_desc(smem_u32(smem), slab, 1024, 2) : smem_desc(smem_u32(smem), slab, 512, 1);
    const uint64_t dk = smem_desc(smem_u32(smem), 16, 1024);
    const uint32_t a_lo = (uint32_t)(at ? dmn : dk), a_hi = (uint32_t)((at ? dmn : dk) >> 32);
    const uint32_t b_lo = (uint32_t)(bt ? dmn : dk) + (16384u >> 4), b_hi = (uint32_t)((bt ? dmn : dk) >> 32);
    const uint32_t a_adv = at ? 64u : 2u, b_adv = bt ? 64u : 2u;  // per MMA: 8 (16) reduction rows of 128 B, or 32 B along K
    (void)a_hi; (void)b_hi;  // the high words differ only in LBO / SBO / layout type: pass per operand
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kBf16) asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t}" ::"r"(tbase), "r"(a_lo + a_adv * k), "r"(a_hi), "r"(b_lo + b_adv * k), "r"(b_hi), "r"(idesc) : "memory");
        else asm volatile("{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, 1;\n\t}" ::"r"(tbase), "r"(a_lo + a_adv * k), "r"(a_hi), "r"(b_lo + b_adv * k), "r"(b_hi), "r"(idesc) : "memory");
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int iters = 8000, smem = 66 * 1024;
  cudaFuncSetAttribute(kern<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(kern<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* mn[] = {"A MN-major, B MN-major", "A MN-major, B K-major ", "A K-major,  B MN-major", "A K-major,  B K-major "};
  for (int bf = 0; bf < 2; ++bf)
    for (int mode = 0; mode < 4; ++mode)
      for (int N : {64, 128, 224, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
          if (bf) kern<1><<<148, 128, smem>>>(N, mode, iters, d); else kern<0><<<148, 128, smem>>>(N, mode, iters, d);
        }
        cudaError_t e = cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double cyc = (double)h / iters;
        const double flop = 2.0 * 128 * N * (bf ? 16 : 8);
        printf("%s %s N %3d: %6.1f cyc/MMA -> %6.0f TFLOP/s chip at 1.965 GHz (%s)\n", bf ? "bf16" : "tf32", mn[mode], N, cyc,
               flop * 148 / cyc * 1.965e9 / 1e12, cudaGetErrorString(e));
      }
  return 0;
}
