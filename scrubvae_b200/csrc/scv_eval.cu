// Evaluation-side kernel: forward kinematics of decoded windows + re-extraction of the conditioned kinematic variables
// (generative restrictiveness, reference eval/eval.py:22-120 -> fwd_kin_cont6d_torch data/dataset.py:83-116):
//   heading       (sin, cos) of the yaw of joint0 -> joint1 at the mid frame (:67-72)
//   avg_speed_3d  [root speed, spine/head part speed, mean of the two limb part speeds], normalised with the reference's
//                 constants (:73-118)
// One block per window: every thread runs the FK of one frame into shared memory, then the block reduces over frames.
#include "scv_common.cuh"
#include "scv_fk.h"

namespace {

constexpr int NT = 64;

__global__ void __launch_bounds__(NT) gen_features_kernel(const float* __restrict__ xh, int64_t ld,
                                                          const float* __restrict__ root_hat,
                                                          const float* __restrict__ offsets,
                                                          const int32_t* __restrict__ tree, int n_tree,
                                                          const int32_t* __restrict__ parts, int W, int J,
                                                          const float* __restrict__ norm, float* __restrict__ pose_out,
                                                          float* __restrict__ heading, float* __restrict__ avg3) {
  extern __shared__ float pose[];  // [W][J][3]
  __shared__ int32_t s_tree[SCV_MAX_J * 3];
  __shared__ float red[NT / 32][4];
  const int64_t b = blockIdx.x;
  for (int i = threadIdx.x; i < n_tree; i += NT) s_tree[i] = tree[i];
  __syncthreads();
  for (int f = threadIdx.x; f < W; f += NT) {
    const int64_t fr = b * W + f;
    const float* c6 = xh + fr * ld;
    const float* off = offsets + fr * J * 3;
    float* P = pose + (size_t)f * J * 3;
    float M0[9];
    scvfk::c6d_to_mat(c6, 1e-8f, M0);
    for (int q = 0; q < 3; ++q) P[q] = root_hat ? root_hat[fr * 3 + q] : 0.f;
    int pos = 1;
    for (int ch = 0; ch < s_tree[0]; ++ch) {
      const int len = s_tree[pos];
      const int32_t* cj = s_tree + pos + 1;
      float R[9], Mj[9], T[9];
      for (int q = 0; q < 9; ++q) R[q] = M0[q];
      for (int i = 1; i < len; ++i) {
        const int j = cj[i], pj = cj[i - 1];
        scvfk::c6d_to_mat(c6 + j * 6, 1e-8f, Mj);
        scvfk::mat_mul(R, Mj, T);
        for (int q = 0; q < 9; ++q) R[q] = T[q];
        for (int r = 0; r < 3; ++r)
          P[j * 3 + r] = R[r * 3] * off[j * 3] + R[r * 3 + 1] * off[j * 3 + 1] + R[r * 3 + 2] * off[j * 3 + 2] + P[pj * 3 + r];
      }
      pos += 1 + len;
    }
  }
  __syncthreads();
  if (pose_out)
    for (int i = threadIdx.x; i < W * J * 3; i += NT) pose_out[b * (int64_t)W * J * 3 + i] = pose[i];
  if (heading && threadIdx.x == 0) {
    const float* P = pose + (size_t)(W / 2) * J * 3;
    float fx = P[3] - P[0], fy = P[4] - P[1], fz = P[5] - P[2];
    const float n = sqrtf(fx * fx + fy * fy + fz * fz);
    fx /= n; fy /= n;
    const float yaw = -atan2f(fy, fx);
    heading[b * 2] = sinf(yaw);
    heading[b * 2 + 1] = cosf(yaw);
  }
  if (avg3) {
    // accumulators: 0 root speed, 1..3 part speeds (sums of |pose[t+1][j] - pose[t][j]| over frames and part joints)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int cnt[4] = {W - 1, 0, 0, 0};
    auto dist = [&](int t, int j) {
      const float* a = pose + ((size_t)t * J + j) * 3;
      const float* c = a + J * 3;
      const float dx = c[0] - a[0], dy = c[1] - a[1], dz = c[2] - a[2];
      return sqrtf(dx * dx + dy * dy + dz * dz);
    };
    for (int t = threadIdx.x; t < W - 1; t += NT) acc[0] += dist(t, 0);
    int pos = 1;
    for (int pi = 0; pi < parts[0] && pi < 3; ++pi) {
      const int len = parts[pos];
      cnt[1 + pi] = (W - 1) * (len - 1);
      for (int e = threadIdx.x; e < (W - 1) * (len - 1); e += NT) {
        const int t = e / (len - 1), jj = e - t * (len - 1);
        acc[1 + pi] += dist(t, parts[pos + 2 + jj]);  // part[1:]: the first joint of a part is only its reference
      }
      pos += 1 + len;
    }
    for (int q = 0; q < 4; ++q) {
      const float v = scv::warp_sum(acc[q]);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float s[4];
      for (int q = 0; q < 4; ++q) {
        s[q] = 0.f;
        for (int w = 0; w < NT / 32; ++w) s[q] += red[w][q];
        s[q] /= (float)(cnt[q] > 0 ? cnt[q] : 1);
      }
      const float pred[3] = {s[0], s[1], 0.5f * (s[2] + s[3])};
      for (int q = 0; q < 3; ++q) avg3[b * 3 + q] = norm ? (pred[q] - norm[q]) / norm[3 + q] : pred[q];
    }
  }
}

}  // namespace

extern "C" int scv_gen_features(const float* xh, int64_t ld, const float* root_hat, const float* offsets,
                                const int32_t* tree, int64_t n_tree, const int32_t* parts, int64_t B, int64_t W, int64_t J,
                                const float* norm, float* pose_out, float* heading, float* avg3, void* stream) {
  SCV_REQUIRE(xh && offsets && tree, "scv_gen_features: null pointer");
  SCV_REQUIRE(J >= 2 && J <= SCV_MAX_J && n_tree <= SCV_MAX_J * 3 && W >= 2, "scv_gen_features: skeleton / window out of range");
  SCV_REQUIRE(!avg3 || parts, "scv_gen_features: avg_speed_3d needs the part definition");
  if (B <= 0) return 0;
  const size_t sh = (size_t)W * J * 3 * sizeof(float);
  SCV_REQUIRE(sh <= 200 * 1024, "scv_gen_features: window too long for shared memory");
  if (sh > 48 * 1024) cudaFuncSetAttribute(gen_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
  gen_features_kernel<<<(unsigned)B, NT, sh, (cudaStream_t)stream>>>(xh, ld, root_hat, offsets, tree, (int)n_tree, parts, (int)W,
                                                                   (int)J, norm, pose_out, heading, avg3);
  return scv::check_launch("gen_features_kernel");
}
