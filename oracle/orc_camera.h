// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see orc_math.h header).
//
// CPU restatement of the reference's camera models as Path B uses them (projection, its derivatives by the 3-D point and
// by the intrinsics, pyramid scaling, and the radius cut-off search that runs every time a camera is constructed):
//   CameraBase ctor (inverse intrinsics)      /root/reference/src/camera/camera_base.cc:81-85
//   ScaledBy / NormalizedToImage / derivatives /root/reference/src/camera/camera_base_impl.h:70-89,155-164,333-408
//   IterativeUndistort / UndistortFromInside / InitCutoff
//                                              /root/reference/src/camera/camera_base_impl.h:214-250,276-328,410-462
//   PinholeCamera   (type 4, 4 parameters)     /root/reference/src/camera/camera_pinhole.h:40-86 (no cut-off search)
//   ThinPrismCamera (type 14, 12 parameters)   /root/reference/src/camera/camera_thin_prism.h:56-139, camera_thin_prism.cc:34-50
//   BenchmarkCamera (type 5, 12 parameters) = FisheyeBase<ThinPrismCamera>
//                                              /root/reference/src/camera/camera_base_impl_fisheye.h:65-146,
//                                              camera_benchmark.cc:36-46 (the INNER thin-prism model runs InitCutoff; the outer
//                                              camera's own radius_cutoff_squared_ stays +inf)
// Pinned by the reference's camera tests ported in tests/test_oracle_camera.py (camera/test/test_camera.cc:40-420,508-515).
// Eigen evaluation order restated by hand: 2-term sums are a*b + c*d, Matrix2f::inverse() is the cofactor form times 1/det
// (Eigen/src/LU/InverseImpl.h, size-2 specialisation), no fused multiply-add (the reference builds without -mfma).
// Defined here where the reference is platform-dependent: atan2f -> correctly rounded (atan_r1 below). Where it is undefined behaviour: float->int of a non-finite value yields INT_MIN (x86 cvttss2si).
#ifndef ORC_CAMERA_H_
#define ORC_CAMERA_H_
#include <algorithm>
#include <cmath>
#include <limits>

#include "orc_math.h"

namespace orc {

enum CameraType { kCamPinhole = 4, kCamBenchmark = 5, kCamThinPrism = 14 };

// atan2(r, 1.f) of camera_base_impl_fisheye.h:68,104,135. The reference gets whatever its libm's atan2f returns: correctly rounded
// with glibc >= 2.41 (CORE-MATH), up to 1 ulp off (and dependent on the CPU's FMA dispatch) with older glibc. The oracle pins the
// correctly rounded value, computed as the rounding of the fp64 arctangent — the same definition the device code uses.
static inline float atan_r1(float r) { return (float)std::atan((double)r); }

static inline int f2i(float v) {   // x86 cvttss2si semantics made explicit
  if (!(v > -2147483904.f && v < 2147483648.f)) return std::numeric_limits<int>::min();
  return (int)v;
}

struct Camera {
  int type = kCamPinhole;
  int w = 0, h = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0, fx_inv = 0, fy_inv = 0, cx_inv = 0, cy_inv = 0;
  float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // k1 k2 p1 p2 k3 k4 sx1 sy1
  float cutoff2 = std::numeric_limits<float>::infinity();         // CameraBaseImpl::radius_cutoff_squared_ of this camera
  float inner_cutoff2 = std::numeric_limits<float>::infinity();   // benchmark: the inner thin-prism model's cut-off

  static int param_count(int type) { return type == kCamPinhole ? 4 : 12; }
  int np() const { return param_count(type); }
  static bool known(int type) { return type == kCamPinhole || type == kCamBenchmark || type == kCamThinPrism; }

  void set(int type_, int w_, int h_, const float* p) {
    type = type_; w = w_; h = h_; fx = p[0]; fy = p[1]; cx = p[2]; cy = p[3];
    fx_inv = (float)(1.0 / fx); fy_inv = (float)(1.0 / fy);                 // camera_base.cc:83
    cx_inv = (float)(-1.0 * cx / fx); cy_inv = (float)(-1.0 * cy / fy);
    for (int i = 0; i < 8; ++i) d[i] = type == kCamPinhole ? 0.f : p[4 + i];
    cutoff2 = inner_cutoff2 = std::numeric_limits<float>::infinity();
    if (type == kCamThinPrism) cutoff2 = thin_prism_cutoff();
    if (type == kCamBenchmark) inner_cutoff2 = thin_prism_cutoff();
  }
  void get_params(float* p) const {
    p[0] = fx; p[1] = fy; p[2] = cx; p[3] = cy;
    if (type != kCamPinhole) for (int i = 0; i < 8; ++i) p[4 + i] = d[i];
  }
  Camera scaled_half() const {                                             // camera_base_impl.h:70-89, factor 0.5
    const float f = 0.5f;
    float p[12]; get_params(p);
    p[0] *= f; p[1] *= f; p[2] = f * (cx + 0.5f) - 0.5f; p[3] = f * (cy + 0.5f) - 0.5f;
    Camera s; s.set(type, (int)(f * w + 0.5f), (int)(f * h + 0.5f), p);
    return s;
  }

  // ---- thin-prism distortion (camera_thin_prism.h:56-139) ----
  void tp_distort(float x, float y, float* ox, float* oy) const {
    const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4], k4 = d[5], sx1 = d[6], sy1 = d[7];
    const float x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2;
    const float radial = 1 + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4)));
    const float dx = 2.f * p1 * xy + p2 * (r2 + 2.f * x2) + sx1 * r2;
    const float dy = 2.f * p2 * xy + p1 * (r2 + 2.f * y2) + sy1 * r2;
    *ox = x * radial + dx; *oy = y * radial + dy;
  }
  void tp_deriv(float nx, float ny, float J[4]) const {                    // row-major [ddx_dnx ddx_dny; ddy_dnx ddy_dny]
    const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4], k4 = d[5], sx1 = d[6], sy1 = d[7];
    const float nx_ny = nx * ny, nx2 = nx * nx, ny2 = ny * ny, r2 = nx2 + ny2;
    const float term1 = 2 * k1 + r2 * (4 * k2 + r2 * (6 * k3 + r2 * 8 * k4));
    const float term2 = 1 + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4)));
    const float term3 = nx_ny * term1 + 2 * (p1 * nx + p2 * ny);
    J[0] = nx2 * term1 + term2 + 6 * p2 * nx + 2 * p1 * ny + 2 * sx1 * nx;
    J[1] = term3 + 2 * sx1 * ny;
    J[2] = term3 + 2 * sy1 * nx;
    J[3] = ny2 * term1 + term2 + 6 * p1 * ny + 2 * p2 * nx + 2 * sy1 * ny;
  }
  static void tp_deriv_params(float nx, float ny, float D[16]) {           // 2 x 8 row-major
    const float nx2 = nx * nx, ny2 = ny * ny, two_nx_ny = 2.f * nx * ny, r2 = nx2 + ny2;
    D[0] = nx * r2; D[1] = D[0] * r2; D[2] = two_nx_ny; D[3] = (r2 + 2.f * nx2); D[4] = D[1] * r2; D[5] = D[4] * r2; D[6] = r2; D[7] = 0;
    D[8] = ny * r2; D[9] = D[8] * r2; D[10] = (r2 + 2.f * ny2); D[11] = two_nx_ny; D[12] = D[9] * r2; D[13] = D[12] * r2; D[14] = 0; D[15] = r2;
  }

  // ---- Child::Distort / DistortedDerivativeByNormalized / ...ByDistortionParameters ----
  void distort(float x, float y, float* ox, float* oy) const {
    if (type == kCamPinhole) { *ox = x; *oy = y; return; }
    if (type == kCamThinPrism) { tp_distort(x, y, ox, oy); return; }
    const float r = std::sqrt(x * x + y * y);                              // camera_base_impl_fisheye.h:65-78
    if (r > 1e-6f) {
      const float atan_r = atan_r1(r);
      if (atan_r * atan_r > inner_cutoff2) { *ox = x * std::numeric_limits<float>::infinity(); *oy = y * std::numeric_limits<float>::infinity(); return; }
      const float theta_by_r = atan_r / r;
      tp_distort(x * theta_by_r, y * theta_by_r, ox, oy);
    } else {
      tp_distort(x, y, ox, oy);
    }
  }
  void distort_deriv(float nx, float ny, float J[4]) const {
    if (type == kCamPinhole) { J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 1; return; }
    if (type == kCamThinPrism) { tp_deriv(nx, ny, J); return; }
    const float nx_ny = nx * ny, nx2 = nx * nx, ny2 = ny * ny, r2 = nx2 + ny2;   // camera_base_impl_fisheye.h:96-126
    const float r = sqrtf(r2);
    if (r > 1e-6f) {
      const float atan_r = atan_r1(r);
      if (atan_r * atan_r > inner_cutoff2) { J[0] = J[1] = J[2] = J[3] = 0; return; }
      const float theta_by_r = atan_r / r;
      const float term1 = r2 * (r2 + 1);
      const float term2 = theta_by_r / r2;
      const float a = ny2 * term2 + nx2 / term1;
      const float b = nx_ny / term1 - nx_ny * term2;
      const float c = b;
      const float dd = nx2 * term2 + ny2 / term1;
      float Jd[4]; tp_deriv(theta_by_r * nx, theta_by_r * ny, Jd);
      J[0] = Jd[0] * a + Jd[1] * c; J[1] = Jd[0] * b + Jd[1] * dd;
      J[2] = Jd[2] * a + Jd[3] * c; J[3] = Jd[2] * b + Jd[3] * dd;
    } else {
      tp_deriv(nx, ny, J);
    }
  }
  void distort_deriv_params(float nx, float ny, float D[16]) const {      // camera_base_impl_fisheye.h:128-146
    if (type == kCamThinPrism) { tp_deriv_params(nx, ny, D); return; }
    const float r = std::sqrt(nx * nx + ny * ny);
    if (r > 1e-6f) {
      const float atan_r = atan_r1(r);
      if (atan_r * atan_r > inner_cutoff2) { for (int i = 0; i < 16; ++i) D[i] = 0; return; }
      const float theta_by_r = atan_r / r;
      tp_deriv_params(theta_by_r * nx, theta_by_r * ny, D);
    } else {
      tp_deriv_params(nx, ny, D);
    }
  }

  // NormalizedToImage (camera_base_impl.h:155-164)
  void project(float nx, float ny, float* ix, float* iy) const {
    const float r2 = nx * nx + ny * ny;
    if (std::isinf(r2) || r2 > cutoff2) { *ix = nx * std::numeric_limits<float>::infinity(); *iy = ny * std::numeric_limits<float>::infinity(); return; }
    float dx, dy; distort(nx, ny, &dx, &dy);
    *ix = fx * dx + cx; *iy = fy * dy + cy;
  }
  // ImageDerivativeByWorld (camera_base_impl.h:333-360): f.asDiagonal() * (Jdist * [I/z | -n/z]); 2x3 row-major
  void d_by_world(const V3f& p, float o[6]) const {
    const float nx = p.x / p.z, ny = p.y / p.z;
    if (nx * nx + ny * ny < cutoff2) {
      const float z_inv = 1.f / p.z;
      float J[4]; distort_deriv(nx, ny, J);
      const float N[6] = {1.f * z_inv, 0.f * z_inv, -1.f * nx * z_inv, 0.f * z_inv, 1.f * z_inv, -1.f * ny * z_inv};
      for (int c = 0; c < 3; ++c) {
        o[c] = fx * (J[0] * N[c] + J[1] * N[3 + c]);
        o[3 + c] = fy * (J[2] * N[c] + J[3] * N[3 + c]);
      }
    } else {
      for (int i = 0; i < 6; ++i) o[i] = fx * 0.f;
    }
  }
  // ImageDerivativeByIntrinsics (camera_base_impl.h:362-408): 2 x np row-major
  void d_by_intrinsics(const V3f& p, float* o) const {
    const int n = np();
    const float nx = p.x / p.z, ny = p.y / p.z;
    if (nx * nx + ny * ny > cutoff2) { for (int i = 0; i < 2 * n; ++i) o[i] = 0; return; }
    float dx, dy; distort(nx, ny, &dx, &dy);
    o[0] = dx; o[1] = 0.f; o[2] = 1.f; o[3] = 0.f;
    o[n + 0] = 0.f; o[n + 1] = dy; o[n + 2] = 0.f; o[n + 3] = 1.f;
    if (n > 4) {
      float D[16]; distort_deriv_params(nx, ny, D);
      for (int i = 0; i < 8; ++i) { o[4 + i] = fx * D[i]; o[n + 4 + i] = fy * D[8 + i]; }
    }
  }

  // Undistort(distorted) (camera_base_impl.h:251-253 = IterativeUndistort started at the distorted point; camera_pinhole.h:65-68 identity;
  // camera_base_impl_fisheye.h:80-91 inner Undistort then r -> tan(r))
  void undistort(float dx, float dy, float* ox, float* oy) const {
    float ux = dx, uy = dy;
    if (type != kCamPinhole) tp_iterative_undistort(dx, dy, dx, dy, &ux, &uy);
    if (type == kCamBenchmark) {
      const float r = std::sqrt(ux * ux + uy * uy);
      const float factor = (r < 1e-6f) ? 1.f : (r > M_PI / 2.f) ? std::numeric_limits<float>::infinity() : tanf(r) / r;
      ux = factor * ux; uy = factor * uy;
    }
    *ox = ux; *oy = uy;
  }
  // InitializeUndistortionLookup (camera_base_impl.h:255-269): Undistort at every integer pixel; w*h x 2 floats
  void undistortion_lookup(float* table) const {
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) undistort(fx_inv * x + cx_inv, fy_inv * y + cy_inv, &table[2 * ((size_t)y * w + x)], &table[2 * ((size_t)y * w + x) + 1]);
  }
  // ImageToNormalized(pixel_position) (camera_base_impl.h:187-212): bilinear filter of the lookup; camera_pinhole.h:60-63: ImageToDistorted.
  // The reference reads row h of the table when the clamped y is exactly h - 1 (weight 0): defined here as a clamped (finite) read.
  void image_to_normalized(const float* table, float px, float py, float* ox, float* oy) const {
    if (type == kCamPinhole) { *ox = fx_inv * px + cx_inv; *oy = fy_inv * py + cy_inv; return; }
    const float cxp = std::max(std::min(px, w - 1.001f), 0.f), cyp = std::max(std::min(py, h - 1.00f), 0.f);
    const int ix = (int)cxp, iy = (int)cyp;
    const float fx_ = cxp - (float)ix, fy_ = cyp - (float)iy;
    const int ix1 = std::min(ix + 1, w - 1), iy1 = std::min(iy + 1, h - 1);
    const float* tl = &table[2 * ((size_t)iy * w + ix)]; const float* tr = &table[2 * ((size_t)iy * w + ix1)];
    const float* bl = &table[2 * ((size_t)iy1 * w + ix)]; const float* br = &table[2 * ((size_t)iy1 * w + ix1)];
    *ox = (1 - fy_) * ((1 - fx_) * tl[0] + fx_ * tr[0]) + fy_ * ((1 - fx_) * bl[0] + fx_ * br[0]);
    *oy = (1 - fy_) * ((1 - fx_) * tl[1] + fx_ * tr[1]) + fy_ * ((1 - fx_) * bl[1] + fx_ * br[1]);
  }

  // ---- cut-off search of the thin-prism model (camera_base_impl.h:214-250, 276-328, 410-462) ----
  bool tp_iterative_undistort(float tx, float ty, float sx, float sy, float* ox, float* oy) const {
    float ux = sx, uy = sy;
    bool converged = false;
    for (int i = 0; i < 100; ++i) {
      float cxd, cyd; tp_distort(ux, uy, &cxd, &cyd);
      const float ex = cxd - tx, ey = cyd - ty;
      if (ex * ex + ey * ey < 1e-10f) { converged = true; break; }
      float J[4]; tp_deriv(ux, uy, J);
      // Jd2 = Jd^T Jd; step = (Jd2^-1 * Jd) * delta   (sic: Jd, not its transpose)
      const float a = J[0] * J[0] + J[2] * J[2], b = J[0] * J[1] + J[2] * J[3], c = J[1] * J[0] + J[3] * J[2], dd = J[1] * J[1] + J[3] * J[3];
      const float invdet = 1.f / (a * dd - c * b);
      const float i00 = dd * invdet, i10 = -c * invdet, i01 = -b * invdet, i11 = a * invdet;
      const float m00 = i00 * J[0] + i01 * J[2], m01 = i00 * J[1] + i01 * J[3];
      const float m10 = i10 * J[0] + i11 * J[2], m11 = i10 * J[1] + i11 * J[3];
      ux -= m00 * ex + m01 * ey;
      uy -= m10 * ex + m11 * ey;
    }
    *ox = ux; *oy = uy;
    return converged;
  }
  // returns converged; best (bx,by); second best squared radius in *second_r2 when *second_available
  bool tp_undistort_from_inside(float tx, float ty, float* bx, float* by, float* second_r2, bool* second_available) const {
    const int kNumGridSteps = 10; const float kGridHalfExtent = 1.5f, kImproveThreshold = 0.99f;
    bool converged = false; *second_available = false;
    float best_radius = std::numeric_limits<float>::infinity(), second_best_radius = std::numeric_limits<float>::infinity();
    float best_x = 0, best_y = 0, sbx = std::numeric_limits<float>::infinity(), sby = std::numeric_limits<float>::infinity();
    for (int y = 0; y < kNumGridSteps; ++y) {
      const float iy = ty + kGridHalfExtent * (y - 0.5f * kNumGridSteps) / (0.5f * kNumGridSteps);
      for (int x = 0; x < kNumGridSteps; ++x) {
        const float ix = tx + kGridHalfExtent * (x - 0.5f * kNumGridSteps) / (0.5f * kNumGridSteps);
        float rx, ry;
        if (tp_iterative_undistort(tx, ty, ix, iy, &rx, &ry)) {
          const float radius = std::sqrt(rx * rx + ry * ry);
          if (radius < kImproveThreshold * best_radius) {
            second_best_radius = best_radius; sbx = best_x; sby = best_y; *second_available = converged;
            best_radius = radius; best_x = rx; best_y = ry; converged = true;
          } else if (radius > 1 / kImproveThreshold * best_radius && radius < kImproveThreshold * second_best_radius) {
            second_best_radius = radius; sbx = rx; sby = ry; *second_available = true;
          }
        }
      }
    }
    *bx = best_x; *by = best_y; *second_r2 = sbx * sbx + sby * sby;
    return converged;
  }
  float thin_prism_cutoff() const {
    const float kIncreaseFactor = 1.01f, inf = std::numeric_limits<float>::infinity();
    float min_candidate = 0, max_candidate = inf;
    auto test = [&](float px, float py) {
      float bx, by, s2; bool sa;
      if (tp_undistort_from_inside(fx_inv * px + cx_inv, fy_inv * py + cy_inv, &bx, &by, &s2, &sa)) {
        min_candidate = std::max(bx * bx + by * by, min_candidate);
        if (sa) max_candidate = std::min(s2, max_candidate);
      }
    };
    for (int x = 0; x < w; ++x) { test((float)x, 0.f); test((float)x, (float)(h - 1)); }
    for (int y = 0; y < h; ++y) { test(0.f, (float)y); test((float)(w - 1), (float)y); }
    return std::min(min_candidate * kIncreaseFactor, max_candidate);
  }
};

}  // namespace orc
#endif
