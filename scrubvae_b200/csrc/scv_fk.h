// Per-frame forward kinematics on continuous-6D rotations + squared joint-position error, with the
// analytic gradient w.r.t. the 6-D inputs.  __host__ __device__ so that the same arithmetic is
// unit-tested on the CPU (tests/test_host_math.py builds it with g++).
//
// Follows (not copies) the reference semantics of fwd_kin_cont6d_torch data/dataset.py:83-116 with
// do_root_R=True, cont6d_to_matrix data/quaternion.py:337-353 (eps added to each norm) and
// mpjpe_loss train/losses.py:148-171.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SCV_HD __host__ __device__ __forceinline__
#else
#define SCV_HD inline
#endif

#define SCV_MAX_J 32
#define SCV_MAX_CHAIN 16

namespace scvfk {

// columns [x y z] of the rotation built from c = (a, b): row-major m[r*3+col]
SCV_HD void c6d_to_mat(const float* c, float eps, float* m) {
  float ax = c[0], ay = c[1], az = c[2], bx = c[3], by = c[4], bz = c[5];
  float na = sqrtf(ax * ax + ay * ay + az * az) + eps;
  float xx = ax / na, xy = ay / na, xz = az / na;
  float wx = xy * bz - xz * by, wy = xz * bx - xx * bz, wz = xx * by - xy * bx;
  float nw = sqrtf(wx * wx + wy * wy + wz * wz) + eps;
  float zx = wx / nw, zy = wy / nw, zz = wz / nw;
  float yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;
  m[0] = xx; m[1] = yx; m[2] = zx;
  m[3] = xy; m[4] = yy; m[5] = zy;
  m[6] = xz; m[7] = yz; m[8] = zz;
}

// gradient w.r.t. c given the gradient gm (row-major 3x3) w.r.t. the matrix
SCV_HD void c6d_to_mat_bwd(const float* c, float eps, const float* gm, float* gc) {
  float ax = c[0], ay = c[1], az = c[2], bx = c[3], by = c[4], bz = c[5];
  float ra = sqrtf(ax * ax + ay * ay + az * az), na = ra + eps;
  float xx = ax / na, xy = ay / na, xz = az / na;
  float wx = xy * bz - xz * by, wy = xz * bx - xx * bz, wz = xx * by - xy * bx;
  float rw = sqrtf(wx * wx + wy * wy + wz * wz), nw = rw + eps;
  float zx = wx / nw, zy = wy / nw, zz = wz / nw;
  // column gradients
  float gxx = gm[0], gxy = gm[3], gxz = gm[6];
  float gyx = gm[1], gyy = gm[4], gyz = gm[7];
  float gzx = gm[2], gzy = gm[5], gzz = gm[8];
  // y = z x x_:  gz += x_ x gy ; gx += gy x z
  gzx += xy * gyz - xz * gyy; gzy += xz * gyx - xx * gyz; gzz += xx * gyy - xy * gyx;
  gxx += gyy * zz - gyz * zy; gxy += gyz * zx - gyx * zz; gxz += gyx * zy - gyy * zx;
  // z = w / (|w| + eps)
  float dzw = gzx * wx + gzy * wy + gzz * wz;
  float k = rw > 0.f ? dzw / (rw * nw * nw) : 0.f;
  float gwx = gzx / nw - wx * k, gwy = gzy / nw - wy * k, gwz = gzz / nw - wz * k;
  // w = x_ x b:  gx += b x gw ; gb = gw x x_
  gxx += by * gwz - bz * gwy; gxy += bz * gwx - bx * gwz; gxz += bx * gwy - by * gwx;
  float gbx = gwy * xz - gwz * xy, gby = gwz * xx - gwx * xz, gbz = gwx * xy - gwy * xx;
  // x_ = a / (|a| + eps)
  float dxa = gxx * ax + gxy * ay + gxz * az;
  float ka = ra > 0.f ? dxa / (ra * na * na) : 0.f;
  gc[0] = gxx / na - ax * ka; gc[1] = gxy / na - ay * ka; gc[2] = gxz / na - az * ka;
  gc[3] = gbx; gc[4] = gby; gc[5] = gbz;
}

// Same two functions with the divisions hoisted: one reciprocal per norm (2 sqrt + 2 div forward, 2 sqrt + 4 div
// backward instead of 6 / 14 divisions).  Results differ from the exact-division forms above by <= 1 ulp per
// element; the loss kernels (instruction-bound on exactly these divisions) use them, the preprocessing kernel keeps
// the exact forms.
SCV_HD void c6d_to_mat_r(const float* c, float eps, float* m) {
  const float ax = c[0], ay = c[1], az = c[2], bx = c[3], by = c[4], bz = c[5];
  const float ina = 1.f / (sqrtf(ax * ax + ay * ay + az * az) + eps);
  const float xx = ax * ina, xy = ay * ina, xz = az * ina;
  const float wx = xy * bz - xz * by, wy = xz * bx - xx * bz, wz = xx * by - xy * bx;
  const float inw = 1.f / (sqrtf(wx * wx + wy * wy + wz * wz) + eps);
  const float zx = wx * inw, zy = wy * inw, zz = wz * inw;
  m[0] = xx; m[1] = zy * xz - zz * xy; m[2] = zx;
  m[3] = xy; m[4] = zz * xx - zx * xz; m[5] = zy;
  m[6] = xz; m[7] = zx * xy - zy * xx; m[8] = zz;
}

SCV_HD void c6d_to_mat_bwd_r(const float* c, float eps, const float* gm, float* gc) {
  const float ax = c[0], ay = c[1], az = c[2], bx = c[3], by = c[4], bz = c[5];
  const float ra = sqrtf(ax * ax + ay * ay + az * az), ina = 1.f / (ra + eps);
  const float xx = ax * ina, xy = ay * ina, xz = az * ina;
  const float wx = xy * bz - xz * by, wy = xz * bx - xx * bz, wz = xx * by - xy * bx;
  const float rw = sqrtf(wx * wx + wy * wy + wz * wz), inw = 1.f / (rw + eps);
  const float zx = wx * inw, zy = wy * inw, zz = wz * inw;
  float gxx = gm[0], gxy = gm[3], gxz = gm[6];
  const float gyx = gm[1], gyy = gm[4], gyz = gm[7];
  float gzx = gm[2], gzy = gm[5], gzz = gm[8];
  gzx += xy * gyz - xz * gyy; gzy += xz * gyx - xx * gyz; gzz += xx * gyy - xy * gyx;
  gxx += gyy * zz - gyz * zy; gxy += gyz * zx - gyx * zz; gxz += gyx * zy - gyy * zx;
  const float dzw = gzx * wx + gzy * wy + gzz * wz;
  const float k = rw > 0.f ? dzw * inw * inw / rw : 0.f;
  const float gwx = gzx * inw - wx * k, gwy = gzy * inw - wy * k, gwz = gzz * inw - wz * k;
  gxx += by * gwz - bz * gwy; gxy += bz * gwx - bx * gwz; gxz += bx * gwy - by * gwx;
  const float dxa = gxx * ax + gxy * ay + gxz * az;
  const float ka = ra > 0.f ? dxa * ina * ina / ra : 0.f;
  gc[0] = gxx * ina - ax * ka; gc[1] = gxy * ina - ay * ka; gc[2] = gxz * ina - az * ka;
  gc[3] = gwy * xz - gwz * xy; gc[4] = gwz * xx - gwx * xz; gc[5] = gwx * xy - gwy * xx;
}

SCV_HD void mat_mul(const float* a, const float* b, float* o) {  // o = a b
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[r * 3 + c] = a[r * 3] * b[c] + a[r * 3 + 1] * b[3 + c] + a[r * 3 + 2] * b[6 + c];
}
SCV_HD void mat_mul_bt(const float* a, const float* b, float* o) {  // o = a b^T
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      o[r * 3 + c] = a[r * 3] * b[c * 3] + a[r * 3 + 1] * b[c * 3 + 1] + a[r * 3 + 2] * b[c * 3 + 2];
}
SCV_HD void mat_mul_at_acc(const float* a, const float* b, float* o) {  // o += a^T b
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[r * 3 + c] += a[r] * b[c] + a[3 + r] * b[3 + c] + a[6 + r] * b[6 + c];
}

// tree: [n_chains, len0, j.., len1, j.., ...].  c6: J x 6 (row stride 6), off/tgt: J x 3.
// Returns sum_j |tgt_j - pose_j|^2 and writes gc (J x 6) = d(scale * that sum)/d c6.
SCV_HD float fk_jpe_frame(const float* c6, const float* off, const float* tgt, const int32_t* tree, int J,
                          float eps, float scale, float* gc) {
  float M[SCV_MAX_J * 9], gM[SCV_MAX_J * 9], pose[SCV_MAX_J * 3], S[SCV_MAX_J * 3];
  for (int j = 0; j < J; ++j) {
    c6d_to_mat(c6 + j * 6, eps, M + j * 9);
    for (int q = 0; q < 9; ++q) gM[j * 9 + q] = 0.f;
    for (int q = 0; q < 3; ++q) pose[j * 3 + q] = 0.f;
  }
  const int nch = tree[0];
  int pos = 1;
  for (int ch = 0; ch < nch; ++ch) {
    int len = tree[pos];
    const int32_t* cj = tree + pos + 1;
    float R[9], T[9];
    for (int q = 0; q < 9; ++q) R[q] = M[q];
    for (int i = 1; i < len; ++i) {
      int j = cj[i], pj = cj[i - 1];
      mat_mul(R, M + j * 9, T);
      for (int q = 0; q < 9; ++q) R[q] = T[q];
      for (int r = 0; r < 3; ++r)
        pose[j * 3 + r] = R[r * 3] * off[j * 3] + R[r * 3 + 1] * off[j * 3 + 1] + R[r * 3 + 2] * off[j * 3 + 2] +
                          pose[pj * 3 + r];
    }
    pos += 1 + len;
  }
  float loss = 0.f;
  for (int q = 0; q < J * 3; ++q) {
    float d = pose[q] - tgt[q];
    loss += d * d;
    S[q] = 2.f * scale * d;
  }
  // chain starts (so that chains can be walked in reverse order)
  int starts[SCV_MAX_J];
  pos = 1;
  for (int ch = 0; ch < nch; ++ch) { starts[ch] = pos; pos += 1 + tree[pos]; }
  // subtree sums of the position gradients
  for (int ch = nch - 1; ch >= 0; --ch) {
    int len = tree[starts[ch]];
    const int32_t* cj = tree + starts[ch] + 1;
    for (int i = len - 1; i >= 1; --i)
      for (int r = 0; r < 3; ++r) S[cj[i - 1] * 3 + r] += S[cj[i] * 3 + r];
  }
  for (int ch = 0; ch < nch; ++ch) {
    int len = tree[starts[ch]];
    const int32_t* cj = tree + starts[ch] + 1;
    float Racc[SCV_MAX_CHAIN * 9];
    for (int q = 0; q < 9; ++q) Racc[q] = M[q];
    for (int i = 1; i < len; ++i) mat_mul(Racc + (i - 1) * 9, M + cj[i] * 9, Racc + i * 9);
    float gR[9], T[9];
    for (int q = 0; q < 9; ++q) gR[q] = 0.f;
    for (int i = len - 1; i >= 1; --i) {
      int j = cj[i];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) gR[r * 3 + c] += S[j * 3 + r] * off[j * 3 + c];
      mat_mul_at_acc(Racc + (i - 1) * 9, gR, gM + j * 9);
      mat_mul_bt(gR, M + j * 9, T);
      for (int q = 0; q < 9; ++q) gR[q] = T[q];
    }
    for (int q = 0; q < 9; ++q) gM[q] += gR[q];
  }
  for (int j = 0; j < J; ++j) c6d_to_mat_bwd(c6 + j * 6, eps, gM + j * 9, gc + j * 6);
  return loss;
}

}  // namespace scvfk
