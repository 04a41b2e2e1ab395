// Fused latent / loss kernels of the SC-VAE step: CholeskyL + reparameterisation (fwd/bwd), KL prior,
// forward-kinematics joint-position loss + root loss (with unit gradients), tanh backward,
// gradient-reversal head losses.  Reductions are accumulated in double.
#include "scv_common.cuh"
#include "scv_fk.h"

namespace {

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

// one block per sample
__global__ void __launch_bounds__(256) reparam_fwd_kernel(const float* __restrict__ ms, int64_t ms_ld,
                                                          const float* __restrict__ eps, const float* __restrict__ var,
                                                          int nvar, float* __restrict__ mu, float* __restrict__ L,
                                                          float* __restrict__ zc, int64_t zc_ld, int z, int rnd) {
  extern __shared__ float sh[];
  const int nsig = z * (z + 1) / 2;
  float* sig = sh;          // nsig
  float* e = sh + nsig;     // z
  const int64_t b = blockIdx.x;
  const float* row = ms + b * ms_ld;
  for (int i = threadIdx.x; i < nsig; i += blockDim.x) sig[i] = row[z + i];
  for (int i = threadIdx.x; i < z; i += blockDim.x) e[i] = eps ? eps[b * z + i] : 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < z; i += blockDim.x) {
    const int tri = i * (i + 1) / 2;
    float m = row[i], acc = m;
    for (int j = 0; j < i; ++j) acc = fmaf(sig[tri + j], e[j], acc);
    acc = fmaf(softplus_f(sig[tri + i]), e[i], acc);
    if (mu) mu[b * z + i] = m;
    if (zc) scv::store_out(zc, b * zc_ld + i, eps ? acc : m, rnd);
  }
  if (zc) {
    for (int t = threadIdx.x; t < (int)zc_ld - z; t += blockDim.x)
      scv::store_out(zc, b * zc_ld + z + t, t < nvar ? var[b * nvar + t] : 0.f, rnd);
  }
  if (L) {
    float* Lb = L + b * (int64_t)z * z;
    for (int idx = threadIdx.x; idx < z * z; idx += blockDim.x) {
      int i = idx / z, j = idx - i * z;
      float v = 0.f;
      if (j < i) v = sig[i * (i + 1) / 2 + j];
      else if (j == i) v = softplus_f(sig[i * (i + 1) / 2 + i]);
      Lb[idx] = v;
    }
  }
}

__global__ void __launch_bounds__(256) reparam_bwd_kernel(const float* __restrict__ ms, int64_t ms_ld,
                                                          const float* __restrict__ eps, const float* __restrict__ dmu,
                                                          const float* __restrict__ dmu2, float s2,
                                                          const float* __restrict__ dz, int64_t dz_ld,
                                                          const float* __restrict__ dL, float* __restrict__ dms,
                                                          int64_t dms_ld, int z, int rnd) {
  extern __shared__ float sh[];
  float* e = sh;       // z
  float* gz = sh + z;  // z
  const int nsig = z * (z + 1) / 2;
  const int64_t b = blockIdx.x;
  for (int i = threadIdx.x; i < z; i += blockDim.x) {
    e[i] = eps ? eps[b * z + i] : 0.f;
    gz[i] = (dz && eps) ? dz[b * dz_ld + i] : 0.f;
    float g = 0.f;
    if (dmu) g += dmu[b * z + i];
    if (dmu2) g += s2 * dmu2[b * z + i];
    if (dz) g += dz[b * dz_ld + i];
    scv::store_out(dms, b * dms_ld + i, g, rnd);
  }
  __syncthreads();
  const float* sraw = ms + b * ms_ld + z;
  const int64_t drow = b * dms_ld + z;
  const float* dLb = dL ? dL + b * (int64_t)z * z : nullptr;
  for (int idx = threadIdx.x; idx < z * z; idx += blockDim.x) {
    int i = idx / z, j = idx - i * z;
    if (j > i) continue;
    float g = gz[i] * e[j];
    if (dLb) g += dLb[idx];
    const int t = i * (i + 1) / 2 + j;
    if (j == i) g *= softplus_grad(sraw[t]);
    scv::store_out(dms, drow + t, g, rnd);
  }
  for (int t = z + nsig + threadIdx.x; t < dms_ld; t += blockDim.x) scv::store_out(dms, b * dms_ld + t, 0.f, rnd);
}

__global__ void __launch_bounds__(256) kl_kernel(const float* __restrict__ mu, const float* __restrict__ L,
                                                 double* loss, const float* gscale, float* __restrict__ dmu,
                                                 float* __restrict__ dL, int B, int z) {
  __shared__ double shd[32];
  const int64_t b = blockIdx.x;
  const float invB = (gscale ? *gscale : 1.f) / (float)B;
  const float* Lb = L + b * (int64_t)z * z;
  float acc = 0.f;
  for (int idx = threadIdx.x; idx < z * z; idx += blockDim.x) {
    int i = idx / z, j = idx - i * z;
    float v = Lb[idx];
    float g = 0.f;
    if (j <= i) {
      acc += 0.5f * v * v;
      g = v * invB;
      if (j == i) {
        float m = mu[b * z + i];
        acc += -0.5f * (1.f + 2.f * logf(v) - m * m);
        g = (v - 1.f / v) * invB;
        if (dmu) dmu[b * z + i] = m * invB;
      }
    }
    if (dL) dL[b * (int64_t)z * z + idx] = g;
  }
  if (loss) {
    double s = scv::block_sum_d((double)acc, shd);
    if (threadIdx.x == 0) atomicAdd(loss, s / (double)B);
  }
}

constexpr int FK_WARPS = 8;

// ---- fast path: one LANE per (frame, chain), 8 lanes per frame, 4 frames per warp ------------------------------------
// The chain products and their gradients run sequentially in registers along each chain (at most CH_LEN joints),
// only chain start positions, attached-chain gradient totals and the root-matrix gradient cross lanes (a few warp
// shuffles inside the frame's 8-lane group).  ~7x fewer instructions per frame than one lane per joint.  Skeletons
// with more than 8 chains or longer chains take recon_loss_kernel below (both kernels evaluate chain_tree_ok).
constexpr int CH_LANES = 8, CH_LEN = 5;

__device__ __forceinline__ bool chain_tree_ok(const int32_t* tree, int J) {
  if (tree[0] > CH_LANES || J > 32) return false;
  int pos = 1;
  for (int ch = 0; ch < tree[0]; ++ch) {
    if (tree[pos] > CH_LEN || tree[pos] < 1) return false;
    pos += 1 + tree[pos];
  }
  return true;
}

struct ChainTables {
  int8_t len[CH_LANES], jt[CH_LANES][CH_LEN];   // joints of each chain (jt[c][0] = start joint)
  int8_t owner[CH_LANES], oidx[CH_LANES], level[CH_LANES];  // chain / index holding the start joint (-1: root joint)
  int8_t covered[32];
  int nch, maxlevel, ok;
};

__global__ void __launch_bounds__(FK_WARPS * 32, 2) recon_loss_chain_kernel(
    const float* __restrict__ xh, int64_t ld, const float* __restrict__ offsets, const float* __restrict__ target,
    const float* __restrict__ root, const float* __restrict__ arena, const int32_t* __restrict__ tree, double* loss,
    float* __restrict__ root_hat, float* __restrict__ dxh, int64_t F, int B, int J) {
  __shared__ ChainTables T;
  __shared__ double shd[32];
  if (threadIdx.x == 0) {
    T.ok = chain_tree_ok(tree, J) ? 1 : 0;
    if (T.ok) {
      T.nch = tree[0];
      for (int j = 0; j < 32; ++j) T.covered[j] = j == 0 ? 1 : 0;
      int pos = 1;
      for (int c = 0; c < T.nch; ++c) {
        T.len[c] = (int8_t)tree[pos];
        for (int i = 0; i < CH_LEN; ++i) T.jt[c][i] = (int8_t)(i < tree[pos] ? tree[pos + 1 + i] : 0);
        for (int i = 1; i < tree[pos]; ++i) T.covered[tree[pos + 1 + i]] = 1;
        pos += 1 + tree[pos];
      }
      int maxl = 0;
      for (int c = 0; c < T.nch; ++c) {  // where does each chain start?  (the LAST earlier definition of that joint)
        T.owner[c] = -1; T.oidx[c] = 0; T.level[c] = 0;
        const int s = T.jt[c][0];
        if (s != 0)
          for (int k = 0; k < c; ++k)
            for (int i = 1; i < T.len[k]; ++i)
              if (T.jt[k][i] == s) { T.owner[c] = (int8_t)k; T.oidx[c] = (int8_t)i; T.level[c] = (int8_t)(T.level[k] + 1); }
        if (s != 0 && T.owner[c] < 0) T.ok = 0;  // starts at a joint nobody placed: not a tree this path handles
        maxl = max(maxl, (int)T.level[c]);
      }
      T.maxlevel = maxl;
    }
  }
  __syncthreads();
  if (!T.ok) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, c = lane & 7, gbase = grp * CH_LANES;
  const bool act = c < T.nch;
  const int len = act ? T.len[c] : 1, owner = act ? T.owner[c] : -1, oidx = act ? T.oidx[c] : 0, level = act ? T.level[c] : -1;
  int jt[CH_LEN];
#pragma unroll
  for (int i = 0; i < CH_LEN; ++i) jt[i] = act ? T.jt[c][i] : 0;
  const int nch = T.nch, maxlevel = T.maxlevel;
  const int nx = J * 6;
  const float scale = 1.f / ((float)B * 3.f * (float)J), invB = 1.f / (float)B;
  double ljpe = 0.0, lroot = 0.0;
  const int64_t fstep = (int64_t)gridDim.x * FK_WARPS * 4;
  for (int64_t f0 = ((int64_t)blockIdx.x * FK_WARPS + warp) * 4; f0 < F; f0 += fstep) {
    const int64_t f = f0 + grp;
    const bool fok = f < F;
    const float* xr = xh + (fok ? f : 0) * ld;
    float* dr = dxh + (fok ? f : 0) * ld;
    // joints no chain places (and the pad columns) get zero gradient
    if (fok) {
      for (int j = c; j < J; j += CH_LANES)
        if (!T.covered[j]) {
#pragma unroll
          for (int q = 0; q < 6; ++q) dr[j * 6 + q] = 0.f;
        }
      for (int q = nx + 3 + c; q < ld; q += CH_LANES) dr[q] = 0.f;
    }
    float c6[CH_LEN][6], off[CH_LEN][3], Racc[CH_LEN][9], lp[CH_LEN][3];
#pragma unroll
    for (int q = 0; q < 6; ++q) c6[0][q] = fok ? xr[q] : (q == 0 || q == 4 ? 1.f : 0.f);  // the root joint's 6-D starts every chain
    scvfk::c6d_to_mat_r(c6[0], 1e-8f, Racc[0]);
#pragma unroll
    for (int q = 0; q < 3; ++q) lp[0][q] = 0.f;
#pragma unroll
    for (int i = 1; i < CH_LEN; ++i) {
      const bool on = fok && act && i < len;
      const int j = jt[i];
#pragma unroll
      for (int q = 0; q < 6; ++q) c6[i][q] = on ? xr[j * 6 + q] : (q == 0 || q == 4 ? 1.f : 0.f);
#pragma unroll
      for (int q = 0; q < 3; ++q) off[i][q] = on ? offsets[f * J * 3 + j * 3 + q] : 0.f;
      float Mi[9];
      scvfk::c6d_to_mat_r(c6[i], 1e-8f, Mi);
      scvfk::mat_mul(Racc[i - 1], Mi, Racc[i]);
#pragma unroll
      for (int r = 0; r < 3; ++r)
        lp[i][r] = lp[i - 1][r] + (Racc[i][r * 3] * off[i][0] + Racc[i][r * 3 + 1] * off[i][1] + Racc[i][r * 3 + 2] * off[i][2]);
    }
    // absolute position of the chain's start joint, level by level
    float ps[3] = {0.f, 0.f, 0.f};
    for (int L = 1; L <= maxlevel; ++L) {
#pragma unroll
      for (int i = 1; i < CH_LEN; ++i) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float v = __shfl_sync(0xffffffffu, ps[r] + lp[i][r], gbase + (owner >= 0 ? owner : 0));
          if (level == L && oidx == i) ps[r] = v;
        }
      }
    }
    // position error, loss, per-joint position gradients
    float S[CH_LEN][3], lsum = 0.f;
#pragma unroll
    for (int i = 1; i < CH_LEN; ++i) {
      const bool on = fok && act && i < len;
      const int j = jt[i];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float dlt = on ? ps[r] + lp[i][r] - target[f * J * 3 + j * 3 + r] : 0.f;
        lsum += dlt * dlt;
        S[i][r] = 2.f * scale * dlt;
      }
    }
    if (fok && c == 0) {  // joint 0 sits at the origin (FK is called with a zero root)
#pragma unroll
      for (int r = 0; r < 3; ++r) { const float t0 = target[f * J * 3 + r]; lsum += t0 * t0; }
    }
    lsum = scv::warp_sum(lsum);
    if (lane == 0) ljpe += (double)lsum * (double)scale;
    // subtree sums: own suffix + totals of the chains attached to own joints, deepest level first
    float tot[3] = {0.f, 0.f, 0.f};
    for (int L = maxlevel; L >= 0; --L) {
      if (level == L) {
#pragma unroll
        for (int i = CH_LEN - 2; i >= 1; --i)
#pragma unroll
          for (int r = 0; r < 3; ++r) S[i][r] += S[i + 1][r];
#pragma unroll
        for (int r = 0; r < 3; ++r) tot[r] = S[1][r];
      }
      if (L > 0) {
        for (int k = 0; k < nch; ++k) {  // chain k (level L) hands its total to the joint it hangs off
          const int ko = T.owner[k], ki = T.oidx[k], kl = T.level[k];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float tv = __shfl_sync(0xffffffffu, tot[r], gbase + k);
            if (kl == L && ko == c) {
#pragma unroll
              for (int i = 1; i < CH_LEN; ++i)
                if (i <= ki) S[i][r] += (i == ki) ? tv : 0.f;  // suffix sums of this lane are formed when ITS level comes up
            }
          }
        }
      }
    }
    // rotation gradients back along the chain
    float gR[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) gR[q] = 0.f;
#pragma unroll
    for (int i = CH_LEN - 1; i >= 1; --i) {
      const bool on = fok && act && i < len;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) gR[r * 3 + cc] += on ? S[i][r] * off[i][cc] : 0.f;
      float gM[9], Mi[9], Tm[9], gc[6];
#pragma unroll
      for (int q = 0; q < 9; ++q) gM[q] = 0.f;
      scvfk::mat_mul_at_acc(Racc[i - 1], gR, gM);
      scvfk::c6d_to_mat_r(c6[i], 1e-8f, Mi);
      scvfk::c6d_to_mat_bwd_r(c6[i], 1e-8f, gM, gc);
      if (on) {
#pragma unroll
        for (int q = 0; q < 6; ++q) dr[jt[i] * 6 + q] = gc[q];
        scvfk::mat_mul_bt(gR, Mi, Tm);
#pragma unroll
        for (int q = 0; q < 9; ++q) gR[q] = Tm[q];
      }
    }
    // the root joint's matrix starts every chain: sum its gradient over the frame's lanes
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      float v = (fok && act) ? gR[q] : 0.f;
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      gR[q] = v;
    }
    if (fok && c == 0) {
      float gc[6];
      scvfk::c6d_to_mat_bwd_r(c6[0], 1e-8f, gR, gc);
#pragma unroll
      for (int q = 0; q < 6; ++q) dr[q] = gc[q];
    }
    if (fok && c < 3) {  // root channels: inverse normalisation, squared error, unit gradient
      const int d = c;
      const float a0 = arena[d], a1 = arena[3 + d];
      const float rh = 0.5f * (xr[nx + d] + 1.f) * (a1 - a0) + a0;
      if (root_hat) root_hat[f * 3 + d] = rh;
      const float diff = rh - root[f * 3 + d];
      lroot += (double)(diff * diff) * (double)invB;
      dr[nx + d] = 2.f * diff * 0.5f * (a1 - a0) * invB;
    }
  }
  double s0 = scv::block_sum_d(ljpe, shd);
  if (threadIdx.x == 0) atomicAdd(loss, s0);
  double s1 = scv::block_sum_d(lroot, shd);
  if (threadIdx.x == 0) atomicAdd(loss + 1, s1);
}

// One WARP per frame, one LANE per joint.  The kinematic tree is turned into per-joint tables once per block
// (chain index, rotation parent, position parent, position children); chain products, positions, subtree sums of
// the position gradients and the rotation gradients then move between lanes with warp shuffles, level by level.
// Everything a lane owns (its joint's 3x3 matrices) stays in registers: no local-memory arrays, coalesced row reads
// and writes.  Semantics: fwd_kin_cont6d_torch (every chain restarts from the ROOT joint's rotation) + mpjpe_loss +
// root MSE, as scvfk::fk_jpe_frame (scv_fk.h, kept as the host-testable statement of the same math).
constexpr int FK_MAXCH = 6;  // position children per joint

struct FkTables {
  int8_t cidx[32], rp[32], pp[32], pdepth[32], rc[32], nchild[32], child[32][FK_MAXCH];
  int maxc, maxd, maxch, bad;
};

__device__ __forceinline__ void shfl9(const float* v, int src, float* o) {
#pragma unroll
  for (int q = 0; q < 9; ++q) o[q] = __shfl_sync(0xffffffffu, v[q], src);
}

__global__ void __launch_bounds__(FK_WARPS * 32) recon_loss_kernel(const float* __restrict__ xh, int64_t ld,
                                                                   const float* __restrict__ offsets,
                                                                   const float* __restrict__ target,
                                                                   const float* __restrict__ root,
                                                                   const float* __restrict__ arena,
                                                                   const int32_t* __restrict__ tree, int n_tree,
                                                                   double* loss, float* __restrict__ root_hat,
                                                                   float* __restrict__ dxh, int64_t F, int B, int J) {
  __shared__ FkTables T;
  __shared__ double shd[32];
  __shared__ int chain_done;
  if (threadIdx.x == 0) {  // the chain kernel (launched first) has done the work when the skeleton fits it
    int okc = chain_tree_ok(tree, J) ? 1 : 0;
    if (okc) {
      int pos = 1;
      for (int c = 0; c < tree[0] && okc; ++c) {
        const int s0 = tree[pos + 1];
        bool found = s0 == 0;
        int p2 = 1;
        for (int k = 0; k < c && !found; ++k) {
          for (int i = 1; i < tree[p2]; ++i) found = found || tree[p2 + 1 + i] == s0;
          p2 += 1 + tree[p2];
        }
        if (!found) okc = 0;
        pos += 1 + tree[pos];
      }
    }
    chain_done = okc;
  }
  __syncthreads();
  if (chain_done) return;
  if (threadIdx.x == 0) {
    for (int j = 0; j < 32; ++j) {
      T.cidx[j] = j == 0 ? 0 : -1; T.rp[j] = 0; T.pp[j] = -1; T.pdepth[j] = 0; T.rc[j] = -1; T.nchild[j] = 0;
      for (int c = 0; c < FK_MAXCH; ++c) T.child[j][c] = -1;
    }
    T.bad = 0;
    int pos = 1;
    for (int ch = 0; ch < tree[0]; ++ch) {
      const int len = tree[pos];
      const int32_t* cj = tree + pos + 1;
      for (int i = 1; i < len; ++i) {
        const int j = cj[i];
        T.cidx[j] = (int8_t)i;
        T.pp[j] = (int8_t)cj[i - 1];
        T.rp[j] = (int8_t)(i == 1 ? 0 : cj[i - 1]);
        if (i >= 2) T.rc[cj[i - 1]] = (int8_t)j;
      }
      pos += 1 + len;
    }
    int maxc = 0, maxd = 0, maxch = 0;
    for (int it = 0; it < J; ++it)  // depths: parents may be defined by any chain
      for (int j = 1; j < J; ++j)
        if (T.pp[j] >= 0) T.pdepth[j] = (int8_t)(T.pdepth[T.pp[j]] + 1);
    for (int j = 1; j < J; ++j) {
      if (T.pp[j] >= 0) {
        const int pj = T.pp[j];
        if (T.nchild[pj] < FK_MAXCH) T.child[pj][T.nchild[pj]++] = (int8_t)j; else T.bad = 1;
      }
      maxc = max(maxc, (int)T.cidx[j]);
      maxd = max(maxd, (int)T.pdepth[j]);
    }
    for (int j = 0; j < J; ++j) maxch = max(maxch, (int)T.nchild[j]);
    T.maxc = maxc; T.maxd = maxd; T.maxch = maxch;
  }
  __syncthreads();
  if (T.bad) {  // a joint with more than FK_MAXCH children: not a skeleton this kernel was built for
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("libscv: recon_loss: kinematic tree has a joint with too many children\n");
    __trap();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane;
  const bool act = j < J;
  const int cidx = act ? T.cidx[j] : -1, rp = act ? T.rp[j] : 0, pp = (act && T.pp[j] >= 0) ? T.pp[j] : 0;
  const int pdepth = act ? T.pdepth[j] : -1, rc = act ? T.rc[j] : -1, nchild = act ? T.nchild[j] : 0;
  int child[FK_MAXCH];
#pragma unroll
  for (int c = 0; c < FK_MAXCH; ++c) child[c] = act ? T.child[j][c] : -1;
  const int maxc = T.maxc, maxd = T.maxd, maxch = T.maxch;
  const int nx = J * 6;
  const float scale = 1.f / ((float)B * 3.f * (float)J), invB = 1.f / (float)B;
  double ljpe = 0.0, lroot = 0.0;
  for (int64_t f = (int64_t)blockIdx.x * FK_WARPS + warp; f < F; f += (int64_t)gridDim.x * FK_WARPS) {
    const float* xr = xh + f * ld;
    float* dr = dxh + f * ld;
    float c6[6] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f}, off[3] = {0.f, 0.f, 0.f}, tgt[3] = {0.f, 0.f, 0.f};
    if (act) {
#pragma unroll
      for (int q = 0; q < 6; ++q) c6[q] = xr[j * 6 + q];
#pragma unroll
      for (int q = 0; q < 3; ++q) { off[q] = offsets[f * J * 3 + j * 3 + q]; tgt[q] = target[f * J * 3 + j * 3 + q]; }
    }
    float M[9], R[9], Rp[9], tmp[9];
    scvfk::c6d_to_mat(c6, 1e-8f, M);
#pragma unroll
    for (int q = 0; q < 9; ++q) { R[q] = M[q]; Rp[q] = 0.f; }
    for (int s = 1; s <= maxc; ++s) {  // chain products, one chain level at a time
      shfl9(R, rp, tmp);
      if (cidx == s) {
#pragma unroll
        for (int q = 0; q < 9; ++q) Rp[q] = tmp[q];
        scvfk::mat_mul(Rp, M, R);
      }
    }
    float v[3] = {0.f, 0.f, 0.f}, pose[3] = {0.f, 0.f, 0.f};
    if (cidx >= 1) {
#pragma unroll
      for (int r = 0; r < 3; ++r) v[r] = R[r * 3] * off[0] + R[r * 3 + 1] * off[1] + R[r * 3 + 2] * off[2];
    }
    for (int d = 1; d <= maxd; ++d) {  // positions, one tree level at a time
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float pv = __shfl_sync(0xffffffffu, pose[r], pp);
        if (pdepth == d) pose[r] = v[r] + pv;
      }
    }
    float S[3], lsum = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float dlt = act ? pose[r] - tgt[r] : 0.f;
      lsum += dlt * dlt;
      S[r] = 2.f * scale * dlt;
    }
    lsum = scv::warp_sum(lsum);
    if (lane == 0) ljpe += (double)lsum * (double)scale;
    for (int d = maxd - 1; d >= 0; --d) {  // subtree sums of the position gradients
#pragma unroll
      for (int ci = 0; ci < FK_MAXCH; ++ci) {
        if (ci < maxch) {  // warp-uniform
          const int c = child[ci];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float cv = __shfl_sync(0xffffffffu, S[r], c >= 0 ? c : 0);
            if (pdepth == d && ci < nchild) S[r] += cv;
          }
        }
      }
    }
    float gR[9], gM[9], Tm[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) { gR[r * 3 + c] = cidx >= 1 ? S[r] * off[c] : 0.f; gM[r * 3 + c] = 0.f; }
    for (int s = maxc; s >= 1; --s) {  // rotation gradients back along the chains
#pragma unroll
      for (int q = 0; q < 9; ++q) Tm[q] = 0.f;
      if (cidx == s) {
        scvfk::mat_mul_at_acc(Rp, gR, gM);  // gM += Rp^T gR
        scvfk::mat_mul_bt(gR, M, Tm);       // to the rotation parent: gR M^T
      }
      if (s >= 2) {
        shfl9(Tm, rc >= 0 ? rc : 0, tmp);
        if (cidx == s - 1 && rc >= 0) {
#pragma unroll
          for (int q = 0; q < 9; ++q) gR[q] += tmp[q];
        }
      } else {  // the root joint's matrix starts every chain: sum over all first chain elements
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          const float tot = scv::warp_sum(Tm[q]);
          if (j == 0) gM[q] += tot;
        }
      }
    }
    if (act) {
      float gc[6];
      scvfk::c6d_to_mat_bwd(c6, 1e-8f, gM, gc);
#pragma unroll
      for (int q = 0; q < 6; ++q) dr[j * 6 + q] = gc[q];
    }
    if (lane < 3) {  // root channels: inverse normalisation, squared error, unit gradient
      const int d = lane;
      const float a0 = arena[d], a1 = arena[3 + d];
      const float rh = 0.5f * (xr[nx + d] + 1.f) * (a1 - a0) + a0;
      if (root_hat) root_hat[f * 3 + d] = rh;
      const float diff = rh - root[f * 3 + d];
      lroot += (double)(diff * diff) * (double)invB;
      dr[nx + d] = 2.f * diff * 0.5f * (a1 - a0) * invB;
    }
    for (int q = nx + 3 + lane; q < ld; q += 32) dr[q] = 0.f;
  }
  double s0 = scv::block_sum_d(ljpe, shd);
  if (threadIdx.x == 0) atomicAdd(loss, s0);
  double s1 = scv::block_sum_d(lroot, shd);
  if (threadIdx.x == 0) atomicAdd(loss + 1, s1);
}

// float4 per thread (ld % 4 == 0, nx % 4 == 0: the 3 root channels + 1 pad share the last group), 32-bit index math
__global__ void __launch_bounds__(256) out_bwd_kernel(const float* __restrict__ xh, const float* __restrict__ dxh,
                                                      int ld, const float* g_jpe, const float* g_root, int nx,
                                                      float* __restrict__ draw, int64_t d_bs, int64_t d_ls,
                                                      int64_t rows, int W, int rnd) {
  const float gj = g_jpe ? *g_jpe : 0.f, gr = g_root ? *g_root : 0.f;
  const uint32_t ld4n = (uint32_t)ld >> 2;
  const int64_t total = rows * ld4n;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const uint32_t r = (uint32_t)(i / ld4n), c = ((uint32_t)i - r * ld4n) * 4;
    const uint32_t b = r / (uint32_t)W, w = r - b * (uint32_t)W;
    const float4 y = *reinterpret_cast<const float4*>(xh + (int64_t)r * ld + c);
    const float4 g = *reinterpret_cast<const float4*>(dxh + (int64_t)r * ld + c);
    const float yy[4] = {y.x, y.y, y.z, y.w}, gg[4] = {g.x, g.y, g.z, g.w};
    float dv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int cc = (int)c + q;
      const float sc = cc < nx ? gj : (cc < nx + 3 ? gr : 0.f);
      dv[q] = gg[q] * sc * (1.f - yy[q] * yy[q]);
    }
    float4 o = make_float4(dv[0], dv[1], dv[2], dv[3]);
    scv::store_out4(draw, (int64_t)b * d_bs + (int64_t)w * d_ls + c, o, rnd);
  }
}

struct GrPtrs {
  const float* pred[8];
  float* dpred[8];
  float w[8];
};
__global__ void __launch_bounds__(256) gr_loss_kernel(const GrPtrs P, int ld, int n_ens, const float* __restrict__ target,
                                                      const int64_t* __restrict__ labels, int B, int d, double* loss,
                                                      const float* gscale) {
  __shared__ double shd[32];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const float gs = gscale ? *gscale : 1.f;
  double acc = 0.0;
  if (b < B) {
    for (int e = 0; e < n_ens; ++e) {
      const float* pr = P.pred[e] + (int64_t)b * ld;
      float* dp = P.dpred[e] ? P.dpred[e] + (int64_t)b * ld : nullptr;
      float le = 0.f;
      if (labels) {  // CrossEntropyLoss(reduction="sum")
        float mx = pr[0];
        for (int q = 1; q < d; ++q) mx = fmaxf(mx, pr[q]);
        float se = 0.f;
        for (int q = 0; q < d; ++q) se += expf(pr[q] - mx);
        const int lab = (int)labels[b];
        le = logf(se) + mx - pr[lab];
        if (dp)
          for (int q = 0; q < d; ++q) dp[q] = gs * P.w[e] * (expf(pr[q] - mx) / se - (q == lab ? 1.f : 0.f));
      } else {  // MSELoss(reduction="sum")
        for (int q = 0; q < d; ++q) {
          float df = pr[q] - target[(int64_t)b * d + q];
          le += df * df;
          if (dp) dp[q] = gs * P.w[e] * 2.f * df;
        }
      }
      if (dp)
        for (int q = d; q < ld; ++q) dp[q] = 0.f;
      acc += (double)le * (double)P.w[e];
    }
  }
  if (loss) {
    double s = scv::block_sum_d(acc, shd);
    if (threadIdx.x == 0) atomicAdd(loss, s);
  }
}

}  // namespace

extern "C" {

int scv_reparam_fwd(const float* ms, int64_t ms_ld, const float* eps, const float* var, int64_t nvar, float* mu,
                    float* L, float* zc, int64_t zc_ld, int64_t B, int64_t z, int64_t flags, void* stream) {
  size_t sh = (size_t)(z * (z + 1) / 2 + z) * sizeof(float);
  SCV_REQUIRE(sh <= 48 * 1024, "scv_reparam_fwd: z_dim %lld too large", (long long)z);
  SCV_REQUIRE(!zc || zc_ld >= z + nvar, "scv_reparam_fwd: zc_ld too small");
  if (B <= 0) return 0;
  reparam_fwd_kernel<<<(unsigned)B, 256, sh, (cudaStream_t)stream>>>(ms, ms_ld, eps, var, (int)nvar, mu, L, zc, zc_ld,
                                                                    (int)z, (flags & SCV_F_OUT_BF16) ? 2 : (int)(flags & SCV_F_ROUND_TF32));
  return scv::check_launch("reparam_fwd_kernel");
}

int scv_reparam_bwd(const float* ms, int64_t ms_ld, const float* eps, const float* dmu, const float* dmu2,
                    double dmu2_scale, const float* dz, int64_t dz_ld, const float* dL, float* dms, int64_t dms_ld,
                    int64_t B, int64_t z, int64_t flags, void* stream) {
  if (B <= 0) return 0;
  reparam_bwd_kernel<<<(unsigned)B, 256, (size_t)(2 * z) * sizeof(float), (cudaStream_t)stream>>>(
      ms, ms_ld, eps, dmu, dmu2, (float)dmu2_scale, dz, dz_ld, dL, dms, dms_ld, (int)z, (flags & SCV_F_OUT_BF16) ? 2 : (int)(flags & SCV_F_ROUND_TF32));
  return scv::check_launch("reparam_bwd_kernel");
}

int scv_kl(const float* mu, const float* L, double* loss, const float* gscale, float* dmu, float* dL, int64_t B,
           int64_t z, void* stream) {
  if (B <= 0) return 0;
  kl_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(mu, L, loss, gscale, dmu, dL, (int)B, (int)z);
  return scv::check_launch("kl_kernel");
}

int scv_recon_loss(const float* xh, int64_t ld, const float* offsets, const float* target, const float* root,
                   const float* arena, const int32_t* tree, int64_t n_tree, double* loss, float* root_hat, float* dxh,
                   int64_t F, int64_t B, int64_t J, int64_t tree_kind, void* stream) {
  SCV_REQUIRE(J <= SCV_MAX_J && n_tree <= SCV_MAX_J * 3 && ld >= J * 6 + 3, "scv_recon_loss: bad J/tree/ld");
  SCV_REQUIRE(tree_kind >= 0 && tree_kind <= 2, "scv_recon_loss: tree_kind is 0 (decide on the device), 1 or 2");
  if (F <= 0) return 0;
  int64_t blocks = (F + FK_WARPS - 1) / FK_WARPS;
  const int64_t cap = (int64_t)scv::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  // fast path first (lane per chain); the generic kernel exits at once when the fast path took the skeleton
  int64_t cblocks = (F + FK_WARPS * 4 - 1) / (FK_WARPS * 4);
  if (cblocks > cap) cblocks = cap;
  // tree_kind: the tree lives on the device and the library never synchronises, so by default BOTH kernels are launched
  // and each decides for itself (the one that does not serve the skeleton exits at once, ~12 us of empty launch); a
  // caller that knows the tree says which one serves it: 1 = lane per chain, 2 = lane per joint
  if (tree_kind != 2) {
    recon_loss_chain_kernel<<<(unsigned)cblocks, FK_WARPS * 32, 0, (cudaStream_t)stream>>>(
        xh, ld, offsets, target, root, arena, tree, loss, root_hat, dxh, F, (int)B, (int)J);
    int rc0 = scv::check_launch("recon_loss_chain_kernel");
    if (rc0) return rc0;
  }
  if (tree_kind == 1) return 0;
  recon_loss_kernel<<<(unsigned)blocks, FK_WARPS * 32, 0, (cudaStream_t)stream>>>(
      xh, ld, offsets, target, root, arena, tree, (int)n_tree, loss, root_hat, dxh, F, (int)B, (int)J);
  return scv::check_launch("recon_loss_kernel");
}

int scv_out_bwd(const float* xh, const float* dxh, int64_t ld, const float* g_jpe, const float* g_root, int64_t nx,
                float* draw, int64_t d_bs, int64_t d_ls, int64_t B, int64_t W, int64_t flags, void* stream) {
  SCV_REQUIRE(ld % 4 == 0 && d_bs % 4 == 0 && d_ls % 4 == 0 && scv::aligned16(xh) && scv::aligned16(dxh) &&
                  scv::aligned16(draw) && B * W < (1LL << 31),
              "scv_out_bwd: rows must be 16-byte aligned groups of 4 columns");
  int64_t total = B * W * (ld / 4);
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)scv::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return 0;
  out_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(xh, dxh, (int)ld, g_jpe, g_root, (int)nx, draw,
                                                                    d_bs, d_ls, B * W, (int)W,
                                                                    (flags & SCV_F_OUT_BF16) ? 2 : (int)(flags & SCV_F_ROUND_TF32));
  return scv::check_launch("out_bwd_kernel");
}

int scv_gr_loss(const float* const* pred, float* const* dpred, int64_t ld, int64_t n_ens, const float* target,
                const int64_t* labels, int64_t B, int64_t d, int64_t num_keys, double* loss, const float* gscale,
                void* stream) {
  SCV_REQUIRE(n_ens >= 1 && n_ens <= 8 && d <= ld, "scv_gr_loss: bad n_ens/d");
  GrPtrs P;
  const double c = (double)n_ens * (double)num_keys * (double)B;
  for (int e = 0; e < 8; ++e) {
    P.pred[e] = e < n_ens ? pred[e] : nullptr;
    P.dpred[e] = (e < n_ens && dpred) ? dpred[e] : nullptr;
    P.w[e] = e < n_ens ? (float)pow(c, -(double)(n_ens - e)) : 0.f;
  }
  if (B <= 0) return 0;
  gr_loss_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, (int)ld, (int)n_ens, target, labels,
                                                                               (int)B, (int)d, loss, gscale);
  return scv::check_launch("gr_loss_kernel");
}

}  // extern "C"
