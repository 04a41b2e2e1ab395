// Pose-window preprocessing (sm_100a): the reference's preprocess_save_data chain, data/dataset.py:313-454,
// as three kernels over the raw float64 keypoint frames resident in HBM.
//   window_indices_kernel   get_window_indices :198-233 (the index matrix; run boundaries are host metadata)
//   window_features_kernel  get_speed_outliers :299-309, get_speed_parts :134-163 + :362-374,
//                           get_frame_yaw :236-243, get_angle2D :260-267      (one block per window, fp64)
//   preprocess_kernel       root centring :383-392, inv_kin :11-46 (qbetween/qmul/qinv data/quaternion.py:409,34,17),
//                           mid-forward rotation :405-413 (qrot :55-74), quaternion_to_cont6d :325-334,
//                           get_segment_len :279-296 (int64 truncation quirk), target_pose FK :438-449
// One block per window: the window's W x J x 3 doubles are staged in shared memory with coalesced loads, one
// thread per frame does the per-frame kinematics, and the outputs leave through shared memory as contiguous
// rows.  Precision follows the reference step by step: float64 where numpy computes, float32 after the
// `.float()` casts of the *_np helpers.  This file is compiled with -fmad=false so that fp32 products and sums
// round separately, as the reference's elementwise torch ops do.
#include "scv_common.cuh"
#include "scv_fk.h"

namespace {

struct Q { float w, x, y, z; };

__device__ __forceinline__ Q qmul(Q q, Q r) {  // Hamilton product q*r as data/quaternion.py:34-52 orders the terms
  Q o;
  o.w = r.w * q.w - r.x * q.x - r.y * q.y - r.z * q.z;
  o.x = r.w * q.x + r.x * q.w - r.y * q.z + r.z * q.y;
  o.y = r.w * q.y + r.x * q.z + r.y * q.w - r.z * q.x;
  o.z = r.w * q.z - r.x * q.y + r.y * q.x + r.z * q.w;
  return o;
}
__device__ __forceinline__ Q qinv(Q q) { return Q{q.w, -q.x, -q.y, -q.z}; }

// quaternion rotating v0 onto v1: data/quaternion.py:409-420
__device__ __forceinline__ Q qbetween(float ax, float ay, float az, float bx, float by, float bz) {
  float vx = ay * bz - az * by, vy = az * bx - ax * bz, vz = ax * by - ay * bx;
  float w = sqrtf((ax * ax + ay * ay + az * az) * (bx * bx + by * by + bz * bz)) + (ax * bx + ay * by + az * bz);
  float n = sqrtf(w * w + vx * vx + vy * vy + vz * vz);
  return Q{w / n, vx / n, vy / n, vz / n};
}

__global__ void window_indices_kernel(const int64_t* __restrict__ starts, int64_t n_w, int64_t window,
                                      int64_t* __restrict__ winds) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_w * window) return;
  const int64_t w = i / window;
  winds[i] = starts[w] + (i - w * window);
}

// parts: [n_parts, len0, j.., len1, j..]; the first joint of a part is its reference joint
__global__ void __launch_bounds__(128) window_features_kernel(const double* __restrict__ pose, const int64_t* __restrict__ starts,
                                                              int64_t window, int J, const int32_t* __restrict__ parts,
                                                              double* __restrict__ speed, float* __restrict__ avg3,
                                                              float* __restrict__ heading, double* __restrict__ yaw_out) {
  extern __shared__ double sp[];  // W x J x 3
  __shared__ double red[32];
  const int64_t w = blockIdx.x;
  const int W = (int)window;
  const double* src = pose + starts[w] * J * 3;
  for (int i = threadIdx.x; i < W * J * 3; i += blockDim.x) sp[i] = src[i];
  __syncthreads();
  const int npairs = (W - 1) * J;
  // mean keypoint speed (:299-309)
  double acc = 0.0;
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    const int t = i / J, j = i - t * J;
    const double* a = sp + (t * J + j) * 3;
    const double* b = a + J * 3;
    const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
    acc += sqrt(dx * dx + dy * dy + dz * dz);
  }
  acc = scv::block_sum_d(acc, red);
  if (threadIdx.x == 0) speed[w] = acc / (double)npairs;
  // root speed (:140-143)
  acc = 0.0;
  for (int t = threadIdx.x; t < W - 1; t += blockDim.x) {
    const double* a = sp + (t * J) * 3;
    const double* b = a + J * 3;
    const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
    acc += sqrt(dx * dx + dy * dy + dz * dz);
  }
  acc = scv::block_sum_d(acc, red);
  double root_spd = acc / (double)(W - 1);
  // part speeds relative to the root (:144-163): cen = pose - root joint; for parts whose first joint is not 0 the
  // reference subtracts cen[:, part[0]] along the FRAME axis (a frame-constant), reproduced literally
  const int n_parts = parts[0];
  double part_spd[8];
  int pos = 1;
  for (int pi = 0; pi < n_parts && pi < 8; ++pi) {
    const int len = parts[pos];
    const int32_t* pj = parts + pos + 1;
    const int f0 = pj[0];
    const int nj = len - 1;
    acc = 0.0;
    for (int i = threadIdx.x; i < (W - 1) * nj; i += blockDim.x) {
      const int t = i / nj, j = pj[1 + (i - t * nj)];
      double d2 = 0.0;
      for (int c = 0; c < 3; ++c) {
        double a = sp[(t * J + j) * 3 + c] - sp[(t * J) * 3 + c];
        double b = sp[((t + 1) * J + j) * 3 + c] - sp[((t + 1) * J) * 3 + c];
        if (f0 != 0 && f0 < W) {
          const double ref = sp[(f0 * J + j) * 3 + c] - sp[(f0 * J) * 3 + c];
          a -= ref;
          b -= ref;
        }
        const double d = b - a;
        d2 += d * d;
      }
      acc += sqrt(d2);
    }
    acc = scv::block_sum_d(acc, red);
    part_spd[pi] = acc / (double)((W - 1) * nj);
    pos += 1 + len;
  }
  if (threadIdx.x == 0) {
    // avg_speed_3d = [root, first part, mean of the remaining parts] (:362-374)
    double rest = 0.0;
    for (int pi = 1; pi < n_parts; ++pi) rest += part_spd[pi];
    avg3[w * 3 + 0] = (float)root_spd;
    avg3[w * 3 + 1] = (float)part_spd[0];
    avg3[w * 3 + 2] = (float)(rest / (double)(n_parts - 1));
    // yaw of the mid frame (:236-243 with root_i = 0, front_i = 1), heading = (sin, cos) (:260-267)
    const double* m0 = sp + ((W / 2) * J) * 3;
    const double fx = m0[3] - m0[0], fy = m0[4] - m0[1], fz = m0[5] - m0[2];
    const double n = sqrt(fx * fx + fy * fy + fz * fz);
    const double yaw = -atan2(fy / n, fx / n);
    yaw_out[w] = yaw;
    heading[w * 2 + 0] = (float)sin(yaw);
    heading[w * 2 + 1] = (float)cos(yaw);
  }
}

struct PreParams {
  const double* pose;
  const int64_t* starts;
  const int64_t* keep;
  const double* yaw;
  const int32_t* tree;
  const int32_t* offset;  // J x 3 integer unit offsets (mouse_skeleton.yaml OFFSET)
  float *x6d, *root, *offsets, *target;
  int W, J, mode;  // mode 0 none, 1 midfwd, 2 x360 (centre only)
};

__global__ void __launch_bounds__(64) preprocess_kernel(const PreParams p) {
  extern __shared__ double smem_d[];
  const int W = p.W, J = p.J;
  double* sp = smem_d;                                       // W x J x 3 doubles
  float* sx = reinterpret_cast<float*>(sp + (size_t)W * J * 3);  // W x J x 6
  float* so = sx + (size_t)W * J * 6;                        // W x J x 3 offsets
  float* st = so + (size_t)W * J * 3;                        // W x J x 3 target pose
  float* sr = st + (size_t)W * J * 3;                        // W x 3 root
  __shared__ int parent[SCV_MAX_J];
  const int64_t i = blockIdx.x;
  const int64_t w = p.keep ? p.keep[i] : i;
  const double* src = p.pose + p.starts[w] * J * 3;
  for (int k = threadIdx.x; k < W * J * 3; k += blockDim.x) sp[k] = src[k];
  if (threadIdx.x == 0) {
    for (int j = 0; j < J; ++j) parent[j] = j == 0 ? -1 : 0;
    int pos = 1;
    for (int ch = 0; ch < p.tree[0]; ++ch) {
      const int len = p.tree[pos];
      for (int k = 1; k < len; ++k) parent[p.tree[pos + 1 + k]] = p.tree[pos + k];
      pos += 1 + len;
    }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < W; f += blockDim.x) {
    const double* P = sp + (size_t)f * J * 3;
    const double yaw = p.yaw[w];
    const Q fq{(float)cos(yaw / 2), 0.f, 0.f, (float)sin(yaw / 2)};
    // root trajectory: centre on the mid frame's xy (:383-392), rotate to mid-forward (:411-413)
    double rx = P[0], ry = P[1], rz = P[2];
    if (p.mode != 0) {
      const double* Pm = sp + (size_t)(W / 2) * J * 3;
      rx -= Pm[0];
      ry -= Pm[1];
    }
    float ox = (float)rx, oy = (float)ry, oz = (float)rz;
    if (p.mode == 1) {  // qrot(fq, v) = v + 2 (w (qv x v) + qv x (qv x v))
      const float ux = fq.y * oz - fq.z * oy, uy = fq.z * ox - fq.x * oz, uz = fq.x * oy - fq.y * ox;
      const float uux = fq.y * uz - fq.z * uy, uuy = fq.z * ux - fq.x * uz, uuz = fq.x * uy - fq.y * ux;
      ox = ox + 2.f * (fq.w * ux + uux);
      oy = oy + 2.f * (fq.w * uy + uuy);
      oz = oz + 2.f * (fq.w * uz + uuz);
    }
    sr[f * 3 + 0] = ox; sr[f * 3 + 1] = oy; sr[f * 3 + 2] = oz;
    // inverse kinematics (:11-46, forward_indices = [1, 0])
    double dx = P[0] - P[3], dy = P[1] - P[4], dz = P[2] - P[5];
    double dn = sqrt(dx * dx + dy * dy + dz * dz);
    Q root_q = qbetween((float)(dx / dn), (float)(dy / dn), (float)(dz / dn), 1.f, 0.f, 0.f);
    if (i == 0 && f == 0) root_q = Q{1.f, 0.f, 0.f, 0.f};  // quirk (iv): only the very first frame of the data set
    float* X = sx + (size_t)f * J * 6;
    auto store6 = [&](int j, Q q) {  // quaternion_to_matrix :291-317, first two columns (:325-334)
      const float s = 2.0f / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
      X[j * 6 + 0] = 1.f - s * (q.y * q.y + q.z * q.z);
      X[j * 6 + 1] = s * (q.x * q.y + q.z * q.w);
      X[j * 6 + 2] = s * (q.x * q.z - q.y * q.w);
      X[j * 6 + 3] = s * (q.x * q.y - q.z * q.w);
      X[j * 6 + 4] = 1.f - s * (q.x * q.x + q.z * q.z);
      X[j * 6 + 5] = s * (q.y * q.z + q.x * q.w);
    };
    store6(0, p.mode == 1 ? qmul(fq, root_q) : root_q);
    int pos = 1;
    for (int ch = 0; ch < p.tree[0]; ++ch) {
      const int len = p.tree[pos];
      const int32_t* cj = p.tree + pos + 1;
      Q R = root_q;
      for (int k = 0; k + 1 < len; ++k) {
        const int ja = cj[k], jb = cj[k + 1];
        const double ex = P[jb * 3] - P[ja * 3], ey = P[jb * 3 + 1] - P[ja * 3 + 1], ez = P[jb * 3 + 2] - P[ja * 3 + 2];
        const double en = sqrt(ex * ex + ey * ey + ez * ez);
        const Q ruv = qbetween((float)p.offset[jb * 3], (float)p.offset[jb * 3 + 1], (float)p.offset[jb * 3 + 2],
                               (float)(ex / en), (float)(ey / en), (float)(ez / en));
        const Q Rloc = qmul(qinv(R), ruv);
        store6(jb, Rloc);
        R = qmul(R, Rloc);
      }
      pos += 1 + len;
    }
    // segment lengths (:279-296): unit offset x bone length, truncated toward zero (int64 store, quirk iii)
    float* O = so + (size_t)f * J * 3;
    for (int j = 0; j < J; ++j) {
      double ln = 0.0;
      if (j > 0) {
        const int pj = parent[j];
        const double ex = P[j * 3] - P[pj * 3], ey = P[j * 3 + 1] - P[pj * 3 + 1], ez = P[j * 3 + 2] - P[pj * 3 + 2];
        ln = sqrt(ex * ex + ey * ey + ez * ez);
      }
      for (int c = 0; c < 3; ++c) {
        const int64_t oi = p.offset[j * 3 + c];
        O[j * 3 + c] = j == 0 ? (float)oi : (float)(int64_t)(ln * (double)oi);
      }
    }
    // target pose: FK of (x6d, offsets) with zero root, eps 1e-8 (:438-449 -> fwd_kin_cont6d_torch :83-116)
    float* T = st + (size_t)f * J * 3;
    for (int c = 0; c < 3; ++c) T[c] = 0.f;
    float M0[9];
    scvfk::c6d_to_mat(X, 1e-8f, M0);
    pos = 1;
    for (int ch = 0; ch < p.tree[0]; ++ch) {
      const int len = p.tree[pos];
      const int32_t* cj = p.tree + pos + 1;
      float R[9], Mj[9], Tm[9];
      for (int q = 0; q < 9; ++q) R[q] = M0[q];
      for (int k = 1; k < len; ++k) {
        const int j = cj[k], pj = cj[k - 1];
        scvfk::c6d_to_mat(X + j * 6, 1e-8f, Mj);
        scvfk::mat_mul(R, Mj, Tm);
        for (int q = 0; q < 9; ++q) R[q] = Tm[q];
        for (int r = 0; r < 3; ++r)
          T[j * 3 + r] = R[r * 3] * O[j * 3] + R[r * 3 + 1] * O[j * 3 + 1] + R[r * 3 + 2] * O[j * 3 + 2] + T[pj * 3 + r];
      }
      pos += 1 + len;
    }
  }
  __syncthreads();
  float* gx = p.x6d + (size_t)i * W * J * 6;
  for (int k = threadIdx.x; k < W * J * 6; k += blockDim.x) gx[k] = sx[k];
  float* go = p.offsets + (size_t)i * W * J * 3;
  float* gt = p.target + (size_t)i * W * J * 3;
  for (int k = threadIdx.x; k < W * J * 3; k += blockDim.x) {
    go[k] = so[k];
    gt[k] = st[k];
  }
  float* gr = p.root + (size_t)i * W * 3;
  for (int k = threadIdx.x; k < W * 3; k += blockDim.x) gr[k] = sr[k];
}

int g_pre_attr = 0;

}  // namespace

extern "C" {

int scv_window_indices(const int64_t* starts, int64_t n_w, int64_t window, int64_t* winds, void* stream) {
  SCV_REQUIRE(n_w >= 0 && window > 0, "scv_window_indices: bad sizes");
  if (n_w == 0) return 0;
  SCV_REQUIRE(starts && winds, "scv_window_indices: null pointer");
  const int64_t n = n_w * window;
  window_indices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(starts, n_w, window, winds);
  return scv::check_launch("window_indices_kernel");
}

int scv_window_features(const double* pose, const int64_t* starts, int64_t n_w, int64_t window, int64_t J,
                        const int32_t* parts, double* speed, float* avg_speed_3d, float* heading, double* yaw,
                        void* stream) {
  SCV_REQUIRE(n_w >= 0 && window > 1 && J > 1 && J <= SCV_MAX_J, "scv_window_features: bad sizes");
  if (n_w == 0) return 0;
  SCV_REQUIRE(pose && starts && parts && speed && avg_speed_3d && heading && yaw, "scv_window_features: null pointer");
  const size_t smem = (size_t)window * J * 3 * sizeof(double);
  SCV_REQUIRE(smem <= 200 * 1024, "scv_window_features: window x joints too large for shared memory");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(window_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { scv::set_error("scv_window_features: %s", cudaGetErrorString(e)); return (int)e; }
  }
  window_features_kernel<<<(unsigned)n_w, 128, smem, (cudaStream_t)stream>>>(pose, starts, window, (int)J, parts, speed,
                                                                            avg_speed_3d, heading, yaw);
  return scv::check_launch("window_features_kernel");
}

int scv_preprocess_windows(const double* pose, const int64_t* starts, const int64_t* keep, int64_t n_keep, int64_t window,
                           int64_t J, const int32_t* tree, const int32_t* offset, const double* yaw, int64_t mode,
                           float* x6d, float* root, float* offsets, float* target_pose, void* stream) {
  SCV_REQUIRE(n_keep >= 0 && window > 0 && J > 1 && J <= SCV_MAX_J, "scv_preprocess_windows: bad sizes");
  SCV_REQUIRE(mode >= 0 && mode <= 2, "scv_preprocess_windows: mode must be 0 (none), 1 (midfwd) or 2 (x360 centring)");
  if (n_keep == 0) return 0;
  SCV_REQUIRE(pose && starts && tree && offset && yaw && x6d && root && offsets && target_pose,
              "scv_preprocess_windows: null pointer");
  const size_t smem = (size_t)window * J * 3 * sizeof(double) + (size_t)window * (J * 12 + 3) * sizeof(float);
  SCV_REQUIRE(smem <= 220 * 1024, "scv_preprocess_windows: window x joints too large for shared memory");
  if (smem > 48 * 1024 && g_pre_attr < (int)smem) {
    cudaError_t e = cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { scv::set_error("scv_preprocess_windows: %s", cudaGetErrorString(e)); return (int)e; }
    g_pre_attr = (int)smem;
  }
  PreParams q{pose, starts, keep, yaw, tree, offset, x6d, root, offsets, target_pose, (int)window, (int)J, (int)mode};
  preprocess_kernel<<<(unsigned)n_keep, 64, smem, (cudaStream_t)stream>>>(q);
  return scv::check_launch("preprocess_kernel");
}

}  // extern "C"
