// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call anything under oracle/.
//
// Small fixed-size linear algebra restating the third-party arithmetic the reference's
// ICP path relies on (Eigen 3.3.7 + Sophus, neither present in this container):
//   * Eigen fixed-size 3-vector reductions: a0 + (a1 + a2)  (redux_novec_unroller split,
//     from memory of Eigen 3.3 Redux.h — parity of this op ORDER is UNPINNED, no copy of
//     Eigen exists here; it only moves results at the 1e-7 relative level).
//   * Eigen::Quaternion::toRotationMatrix.
//   * Sophus SO3 product + renormalise   (/root/reference/thirdparty/sophus/so3.hpp:328-342, 297-303, 478-489)
//   * Sophus SO3 point action            (so3.hpp:360-370)
//   * Sophus SO3::expAndTheta            (so3.hpp:585-621), eps = 1e-10 (common.hpp:111)
//   * Sophus SE3::exp                    (se3.hpp:763-784), SE3 product (se3.hpp:308-312), cast (se3.hpp:128-131)
//   * dense symmetric solve standing in for Eigen::LDLT (icp_point_to_plane_impl.h:226).
// Compile with -ffp-contract=off so no FMA contraction changes fp32 roundings.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

namespace orc {

struct V3f { float x, y, z; };
struct V3d { double x, y, z; };

static inline float sum3(float a, float b, float c) { return a + (b + c); }
static inline float dot(const V3f& a, const V3f& b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3f sub(const V3f& a, const V3f& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3f add(const V3f& a, const V3f& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3f cross(const V3f& a, const V3f& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Row-major 3x3.
struct M3f { float m[9]; };
static inline V3f mul(const M3f& R, const V3f& p) {
  return {sum3(R.m[0] * p.x, R.m[1] * p.y, R.m[2] * p.z),
          sum3(R.m[3] * p.x, R.m[4] * p.y, R.m[5] * p.z),
          sum3(R.m[6] * p.x, R.m[7] * p.y, R.m[8] * p.z)};
}
static inline M3f mul(const M3f& A, const M3f& B) {
  M3f C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = sum3(A.m[3 * i] * B.m[j], A.m[3 * i + 1] * B.m[3 + j], A.m[3 * i + 2] * B.m[6 + j]);
  return C;
}

// Quaternion (x,y,z,w) — Eigen coefficient order.
template <typename S> struct Quat { S x, y, z, w; };

template <typename S>
static inline void quat_to_matrix(const Quat<S>& q, S* R /*row-major 9*/) {
  const S tx = S(2) * q.x, ty = S(2) * q.y, tz = S(2) * q.z;
  const S twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const S txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const S tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = S(1) - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
  R[3] = txy + twz;          R[4] = S(1) - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = S(1) - (txx + tyy);
}

template <typename S>
static inline Quat<S> quat_normalized(const Quat<S>& q) {
  const S len = std::sqrt((q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w));
  return {q.x / len, q.y / len, q.z / len, q.w / len};
}

// so3.hpp:328-342 — product, then the SO3(quaternion) constructor renormalises (so3.hpp:488).
template <typename S>
static inline Quat<S> quat_mul_normalized(const Quat<S>& a, const Quat<S>& b) {
  Quat<S> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return quat_normalized(r);
}

// so3.hpp:360-370 — p + w*uv + vec x uv, uv = 2 (vec x p).
static inline V3f quat_rotate(const Quat<float>& q, const V3f& p) {
  const V3f v{q.x, q.y, q.z};
  V3f uv = cross(v, p);
  uv = add(uv, uv);
  const V3f c2 = cross(v, uv);
  return {p.x + q.w * uv.x + c2.x, p.y + q.w * uv.y + c2.y, p.z + q.w * uv.z + c2.z};
}

struct SE3f {
  Quat<float> q{0.f, 0.f, 0.f, 1.f};
  V3f t{0.f, 0.f, 0.f};
};

// se3.hpp:308-312.
static inline SE3f se3_mul(const SE3f& a, const SE3f& b) {
  SE3f r;
  r.q = quat_mul_normalized(a.q, b.q);
  r.t = add(a.t, quat_rotate(a.q, b.t));
  return r;
}

// Sophus::SE3d::exp(a).cast<float>()   (se3.hpp:763-784, so3.hpp:585-621, se3.hpp:128-131).
// Tangent order: [upsilon(3); omega(3)].
static inline SE3f se3d_exp_cast_float(const double a[6]) {
  const double wx = a[3], wy = a[4], wz = a[5];
  const double theta_sq = wx * wx + (wy * wy + wz * wz);
  double theta, imag, real;
  if (theta_sq < 1e-10 * 1e-10) {
    theta = 0.0;
    const double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    theta = std::sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  const Quat<double> qd{imag * wx, imag * wy, imag * wz, real};
  // Omega = hat(omega), Omega_sq = Omega*Omega.
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      O2[3 * i + j] = O[3 * i] * O[j] + (O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j]);
  double V[9];
  if (theta < 1e-10) {
    quat_to_matrix(qd, V);
  } else {
    const double c1 = (1.0 - std::cos(theta)) / theta_sq;
    const double c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int k = 0; k < 9; ++k) V[k] = ((k % 4 == 0) ? 1.0 : 0.0) + c1 * O[k] + c2 * O2[k];
  }
  const double td[3] = {V[0] * a[0] + (V[1] * a[1] + V[2] * a[2]),
                        V[3] * a[0] + (V[4] * a[1] + V[5] * a[2]),
                        V[6] * a[0] + (V[7] * a[1] + V[8] * a[2])};
  SE3f r;
  r.q = quat_normalized(Quat<float>{(float)qd.x, (float)qd.y, (float)qd.z, (float)qd.w});
  r.t = {(float)td[0], (float)td[1], (float)td[2]};
  return r;
}

// Column-major 4x4 float affine (Eigen::Affine3f storage order).
struct Affine3f {
  float m[16];
  float& at(int r, int c) { return m[4 * c + r]; }
  float at(int r, int c) const { return m[4 * c + r]; }
};
static inline Affine3f affine_identity() {
  Affine3f a; std::memset(a.m, 0, sizeof(a.m)); a.m[0] = a.m[5] = a.m[10] = a.m[15] = 1.f; return a;
}
// Eigen Transform(Affine) * Transform(Affine): linear = L*L', translation = L*t' + t.
static inline Affine3f affine_mul(const Affine3f& A, const Affine3f& B) {
  Affine3f C = affine_identity();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      C.at(i, j) = sum3(A.at(i, 0) * B.at(0, j), A.at(i, 1) * B.at(1, j), A.at(i, 2) * B.at(2, j));
    C.at(i, 3) = sum3(A.at(i, 0) * B.at(0, 3), A.at(i, 1) * B.at(1, 3), A.at(i, 2) * B.at(2, 3)) + A.at(i, 3);
  }
  return C;
}
static inline Affine3f se3_to_affine(const SE3f& T) {
  float R[9]; quat_to_matrix(T.q, R);
  Affine3f a = affine_identity();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a.at(i, j) = R[3 * i + j];
  a.at(0, 3) = T.t.x; a.at(1, 3) = T.t.y; a.at(2, 3) = T.t.z;
  return a;
}

// Symmetric solve (A + nothing) x = b using the UPPER triangle of A only (n x n, column-major),
// LDL^T with diagonal pivoting (largest remaining |d|), standing in for
// Eigen `selfadjointView<Upper>().ldlt().solve(b)` — any backward-stable solve is equivalent at 1e-5.
static inline bool ldlt_solve_upper(const std::vector<double>& Aupper, int n, const std::vector<double>& b,
                                    std::vector<double>* x) {
  std::vector<double> A((size_t)n * n);
  for (int c = 0; c < n; ++c)
    for (int r = 0; r < n; ++r) A[(size_t)c * n + r] = (r <= c) ? Aupper[(size_t)c * n + r] : Aupper[(size_t)r * n + c];
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  auto at = [&](int r, int c) -> double& { return A[(size_t)c * n + r]; };
  for (int k = 0; k < n; ++k) {
    int p = k; double best = std::fabs(at(k, k));
    for (int i = k + 1; i < n; ++i) if (std::fabs(at(i, i)) > best) { best = std::fabs(at(i, i)); p = i; }
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(at(k, j), at(p, j));
      for (int j = 0; j < n; ++j) std::swap(at(j, k), at(j, p));
      std::swap(perm[k], perm[p]);
    }
    const double d = at(k, k);
    if (d == 0.0) continue;  // semidefinite tail: leave zeros (Eigen does the same cut-off)
    for (int i = k + 1; i < n; ++i) at(i, k) /= d;          // L column
    for (int j = k + 1; j < n; ++j) {
      const double ljk_d = at(j, k) * d;
      for (int i = j; i < n; ++i) at(i, j) -= at(i, k) * ljk_d;
    }
    for (int i = k + 1; i < n; ++i) for (int j = i + 1; j < n; ++j) at(i, j) = at(j, i);  // keep symmetric copy
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
  for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= at(i, j) * y[j];
  for (int i = 0; i < n; ++i) y[i] = (at(i, i) != 0.0) ? y[i] / at(i, i) : 0.0;
  for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[i] -= at(j, i) * y[j];
  x->assign(n, 0.0);
  for (int i = 0; i < n; ++i) (*x)[perm[i]] = y[i];
  return true;
}

}  // namespace orc
