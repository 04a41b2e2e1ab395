// Micro-benchmark: issue rate of tcgen05.mma (cta_group::1, M=128) for kind::tf32 and kind::f16 (bf16) on sm_100a.
// One CTA per SM, one thread issues `iters` back-to-back MMAs on fixed shared-memory operands; cycles by clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../scrubvae_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "scv_tc.cuh"
using namespace scv::tc;
namespace scv { void set_error(const char*, ...) {} int64_t g_launches = 0; int sm_count() { return 148; } }

__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// kind::f16 with bf16 operands: a/b format = 1 (BF16), c format = 1 (F32)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int kind, int N, int iters, long long* out, int mn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tbase), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const uint32_t idesc = kind == 0 ? idesc_tf32(128, N, mn, mn) : idesc_bf16(128, N);
    uint64_t da[4], db[4];
    for (int k = 0; k < 4; ++k) {
      da[k] = mn ? smem_desc(sa + k * 1024, 4096, 512, 1) : smem_desc(sa + k * 32, 16, 1024);
      db[k] = mn ? smem_desc(sb + k * 1024, 4096, 512, 1) : smem_desc(sb + k * 32, 16, 1024);
    }
    long long t0 = clock64();
    if (kind == 0) {
      for (int i = 0; i < iters; i += 4) {
        umma_tf32(tbase, da[0], db[0], idesc, 1u);
        umma_tf32(tbase, da[1], db[1], idesc, 1u);
        umma_tf32(tbase, da[2], db[2], idesc, 1u);
        umma_tf32(tbase, da[3], db[3], idesc, 1u);
      }
    } else {
      for (int i = 0; i < iters; i += 4) {
        umma_f16(tbase, da[0], db[0], idesc, 1u);
        umma_f16(tbase, da[1], db[1], idesc, 1u);
        umma_f16(tbase, da[2], db[2], idesc, 1u);
        umma_f16(tbase, da[3], db[3], idesc, 1u);
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
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4000;
  for (int mn = 0; mn < 2; ++mn) for (int grid : {148}) for (int kind = 0; kind < 2 - mn; ++kind) for (int N : {16, 32, 64, 128, 192, 256}) {
    rate_kernel<<<grid, 128, 64 * 1024>>>(kind, N, iters, d, mn);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<<<grid, 128, 64 * 1024>>>(kind, N, iters, d, mn);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    const int kk = kind == 0 ? 8 : 16;
    double cyc = (double)h[0] / iters;
    double tf = 2.0 * 128 * N * kk * iters * grid / (ms * 1e-3) / 1e12;
    printf("%s grid %3d kind %s N %3d: %.1f cycles/MMA (128x%dx%d), %.1f TFLOP/s by events (%s)\n", mn ? "MN-major" : "K-major ", grid, kind == 0 ? "tf32" : "bf16", N, cyc,
           N, kk, tf, cudaGetErrorString(e));
  }
  return 0;
}
