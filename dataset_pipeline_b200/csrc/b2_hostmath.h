// Host-side pose algebra and the dense symmetric solve of the LM step (product code; independent of oracle/).
//
// Mirrors what the reference gets from Sophus/Eigen on the ICP path:
//   SE3d::exp(-x).cast<float>() * T   icp_point_to_plane_impl.h:235  (sophus se3.hpp:763-784, so3.hpp:585-621,328-342,360-370)
//   so3().matrix()                    icp_point_to_plane_impl.h:133,139 (Eigen quaternion -> rotation matrix)
//   selfadjointView<Upper>().ldlt()   icp_point_to_plane_impl.h:226
//   Affine3f * Affine3f               icp_point_to_plane.cc:324-325
// Compiled without FMA contraction (-fmad=false / -ffp-contract=off) so fp32 roundings match a plain C++ evaluation.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace b2 {

struct Pose {          // Sophus::SE3f: unit quaternion (x,y,z,w) + translation
  float q[4] = {0.f, 0.f, 0.f, 1.f};
  float t[3] = {0.f, 0.f, 0.f};
};

inline void quat_normalize(float q[4]) {
  const float len = std::sqrt((q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]));
  for (int i = 0; i < 4; ++i) q[i] = q[i] / len;
}

// Row-major 3x3 from a unit quaternion (Eigen::QuaternionBase::toRotationMatrix).
template <typename S>
inline void quat_matrix(const S q[4], S R[9]) {
  const S x = q[0], y = q[1], z = q[2], w = q[3];
  const S tx = S(2) * x, ty = S(2) * y, tz = S(2) * z;
  const S twx = tx * w, twy = ty * w, twz = tz * w;
  const S txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = S(1) - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = S(1) - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = S(1) - (txx + tyy);
}

// a * b (group product, result renormalised as Sophus' SO3 constructor does) and t = ta + qa (x) tb.
inline Pose pose_mul(const Pose& a, const Pose& b) {
  Pose r;
  const float ax = a.q[0], ay = a.q[1], az = a.q[2], aw = a.q[3];
  const float bx = b.q[0], by = b.q[1], bz = b.q[2], bw = b.q[3];
  r.q[3] = aw * bw - ax * bx - ay * by - az * bz;
  r.q[0] = aw * bx + ax * bw + ay * bz - az * by;
  r.q[1] = aw * by + ay * bw + az * bx - ax * bz;
  r.q[2] = aw * bz + az * bw + ax * by - ay * bx;
  quat_normalize(r.q);
  // rotate b.t by a.q:  p + w*uv + v x uv,  uv = 2 (v x p)
  const float px = b.t[0], py = b.t[1], pz = b.t[2];
  float ux = ay * pz - az * py, uy = az * px - ax * pz, uz = ax * py - ay * px;
  ux += ux; uy += uy; uz += uz;
  const float cx = ay * uz - az * uy, cy = az * ux - ax * uz, cz = ax * uy - ay * ux;
  r.t[0] = a.t[0] + (px + aw * ux + cx);
  r.t[1] = a.t[1] + (py + aw * uy + cy);
  r.t[2] = a.t[2] + (pz + aw * uz + cz);
  return r;
}

// exp of a twist [v; w] in double, cast to float (quaternion renormalised in float).
inline Pose pose_exp(const double a[6]) {
  const double wx = a[3], wy = a[4], wz = a[5];
  const double th2 = wx * wx + (wy * wy + wz * wz);
  double th, im, re;
  if (th2 < 1e-20) {
    th = 0.0;
    const double th4 = th2 * th2;
    im = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    re = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    th = std::sqrt(th2);
    im = std::sin(0.5 * th) / th;
    re = std::cos(0.5 * th);
  }
  const double qd[4] = {im * wx, im * wy, im * wz, re};
  const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double W2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) W2[3 * i + j] = W[3 * i] * W[j] + (W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j]);
  double V[9];
  if (th < 1e-10) {
    quat_matrix(qd, V);
  } else {
    const double c1 = (1.0 - std::cos(th)) / th2, c2 = (th - std::sin(th)) / (th2 * th);
    for (int k = 0; k < 9; ++k) V[k] = ((k == 0 || k == 4 || k == 8) ? 1.0 : 0.0) + c1 * W[k] + c2 * W2[k];
  }
  Pose r;
  for (int i = 0; i < 4; ++i) r.q[i] = (float)qd[i];
  quat_normalize(r.q);
  for (int i = 0; i < 3; ++i) r.t[i] = (float)(V[3 * i] * a[0] + (V[3 * i + 1] * a[1] + V[3 * i + 2] * a[2]));
  return r;
}

// Column-major 4x4 affine helpers (Eigen::Affine3f storage).
inline void affine_from_pose(const Pose& p, float M[16]) {
  float R[9]; quat_matrix(p.q, R);
  std::memset(M, 0, 16 * sizeof(float));
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) M[4 * c + r] = R[3 * r + c]; M[12 + r] = p.t[r]; }
  M[15] = 1.f;
}
inline void affine_mul(const float A[16], const float B[16], float C[16]) {
  float out[16]; std::memset(out, 0, sizeof(out)); out[15] = 1.f;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) out[4 * c + r] = A[r] * B[4 * c] + (A[4 + r] * B[4 * c + 1] + A[8 + r] * B[4 * c + 2]);
    out[12 + r] = (A[r] * B[12] + (A[4 + r] * B[13] + A[8 + r] * B[14])) + A[12 + r];
  }
  std::memcpy(C, out, sizeof(out));
}

// Solve (sym(A)) x = b, A n x n column-major FULL symmetric. Diagonally pivoted LDL^T (robust for the damped,
// possibly rank-deficient normal equations of the planar case, test_icp.cc:111-172).
inline void sym_solve(std::vector<double> A, int n, const double* b, double* x) {
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  auto a = [&](int r, int c) -> double& { return A[(size_t)c * n + r]; };
  for (int k = 0; k < n; ++k) {
    int p = k; double best = std::fabs(a(k, k));
    for (int i = k + 1; i < n; ++i) { const double v = std::fabs(a(i, i)); if (v > best) { best = v; p = i; } }
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(a(k, j), a(p, j));
      for (int j = 0; j < n; ++j) std::swap(a(j, k), a(j, p));
      std::swap(perm[k], perm[p]);
    }
    const double d = a(k, k);
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; ++i) a(i, k) /= d;
    for (int j = k + 1; j < n; ++j) {
      const double f = a(j, k) * d;
      for (int i = j; i < n; ++i) a(i, j) -= a(i, k) * f;
    }
    for (int j = k + 1; j < n; ++j) for (int i = j + 1; i < n; ++i) a(j, i) = a(i, j);
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
  for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= a(i, j) * y[j];
  for (int i = 0; i < n; ++i) y[i] = (a(i, i) != 0.0) ? y[i] / a(i, i) : 0.0;
  for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[i] -= a(j, i) * y[j];
  for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
}

}  // namespace b2
