// Shared helpers for libscv (sm_100a).  No torch headers: the library is a plain C-ABI .so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/scv.h"

namespace scv {

void set_error(const char* fmt, ...);
extern int64_t g_launches;

inline int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define SCV_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      scv::set_error(__VA_ARGS__);        \
      return -1;                          \
    }                                     \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int sm_count();

// round-to-nearest (ties away) to the 10-bit TF32 mantissa; the result is an fp32 value tcgen05 truncates exactly
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 round_tf32(float4 v) {
  return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
}

// Stores of kernels that PRODUCE tensor-core GEMM operands.  omode 0: fp32 as is; 1: fp32 rounded to TF32; 2: bf16 — `base`
// then addresses bf16 elements (idx counts elements either way).
__device__ __forceinline__ int out_mode(int64_t flags) { return (flags & 2) ? 2 : (int)(flags & 1); }
__device__ __forceinline__ void store_out(float* base, int64_t idx, float v, int omode) {
  if (omode == 2) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else base[idx] = omode == 1 ? round_tf32(v) : v;
}
__device__ __forceinline__ void store_out4(float* base, int64_t idx, float4 v, int omode) {  // idx % 4 == 0
  if (omode == 2) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  } else {
    *reinterpret_cast<float4*>(base + idx) = omode == 1 ? round_tf32(v) : v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of a double, result valid in thread 0; `sh` needs 32 doubles
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = l < nw ? sh[l] : 0.0;
    v = warp_sum_d(v);
  }
  return v;
}

}  // namespace scv
