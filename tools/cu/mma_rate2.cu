// Issue rate of tcgen05.mma.cta_group::2 (CTA pair, M=256) for kind::tf32, K-major and MN-major operands.
#include <cstdio>
#include <cuda_runtime.h>
#include "scv_tc.cuh"
using namespace scv::tc;
namespace scv { void set_error(const char*, ...) {} int64_t g_launches = 0; int sm_count() { return 148; } }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(int N, int iters, long long* out, int mn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc2(smem_u32(&tbase), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const uint32_t idesc = idesc_tf32(256, N, mn, mn);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t koff = (i & 3) * 32;
      if (mn) umma2_tf32(tbase, smem_desc(sa + (i & 3) * 1024, 4096, 512, 1), smem_desc(sb + (i & 3) * 1024, 4096, 512, 1), idesc, 1u);
      else umma2_tf32(tbase, smem_desc(sa + koff, 16, 1024), smem_desc(sb + koff, 16, 1024), idesc, 1u);
    }
    umma2_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x / 2] = t1 - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc2(tbase, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4000;
  for (int mn = 0; mn < 2; ++mn) for (int grid : {2, 148}) for (int N : {64, 128, 256}) {
    rate2_kernel<<<grid, 128, 64 * 1024>>>(N, iters, d, mn);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate2_kernel<<<grid, 128, 64 * 1024>>>(N, iters, d, mn);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h[0] / iters;
    double tf = 2.0 * 256 * N * 8 * iters * (grid / 2) / (ms * 1e-3) / 1e12;
    printf("%s pairs %3d tf32 N %3d: %.1f cycles/MMA (256x%dx8), %.1f TFLOP/s by events (%s)\n", mn ? "MN-major" : "K-major ", grid / 2, N, cyc, N, tf,
           cudaGetErrorString(e));
  }
  return 0;
}
