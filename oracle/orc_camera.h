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
#include <vector>

#include "orc_math.h"

namespace orc {

enum CameraType {   // camera::CameraBase::Type (camera_base.h:67-84)
  kCamFOV = 0, kCamPolynomial = 1, kCamPolynomialTangential = 2, kCamFisheyePolynomialTangential = 3, kCamPinhole = 4, kCamBenchmark = 5,
  kCamFisheyePolynomial4 = 6, kCamSimplePinhole = 7, kCamRadial = 8, kCamSimpleRadial = 9, kCamFullOpenCV = 10, kCamPolynomial4 = 11,
  kCamRadialFisheye = 12, kCamSimpleRadialFisheye = 13, kCamThinPrism = 14
};
// distortion function of the camera itself, or of the inner model of a FisheyeBase<> camera
enum DistKind { kDistNone = 0, kDistRadial1, kDistRadial2, kDistPoly3, kDistPoly4, kDistPolyTan, kDistOpenCV, kDistThinPrism, kDistFOV };
struct ModelInfo { int dist; bool fisheye, unique_focal; int nd; };   // nd = number of distortion parameters
static inline bool model_info(int type, ModelInfo* m) {
  switch (type) {
    case kCamFOV: *m = {kDistFOV, false, false, 1}; return true;                              // camera_fisheye_fov.h
    case kCamPolynomial: *m = {kDistPoly3, false, false, 3}; return true;                     // camera_polynomial.h
    case kCamPolynomialTangential: *m = {kDistPolyTan, false, false, 4}; return true;         // camera_polynomial_tangential.h
    case kCamFisheyePolynomialTangential: *m = {kDistPolyTan, true, false, 4}; return true;   // camera_fisheye_polynomial_tangential.cc
    case kCamPinhole: *m = {kDistNone, false, false, 0}; return true;                         // camera_pinhole.h
    case kCamBenchmark: *m = {kDistThinPrism, true, false, 8}; return true;                   // camera_benchmark.cc
    case kCamFisheyePolynomial4: *m = {kDistPoly4, true, false, 4}; return true;              // camera_fisheye_polynomial_4.cc
    case kCamSimplePinhole: *m = {kDistNone, false, true, 0}; return true;                    // camera_simple_pinhole.h
    case kCamRadial: *m = {kDistRadial2, false, true, 2}; return true;                        // camera_radial.h
    case kCamSimpleRadial: *m = {kDistRadial1, false, true, 1}; return true;                  // camera_simple_radial.h
    case kCamFullOpenCV: *m = {kDistOpenCV, false, false, 8}; return true;                    // camera_full_opencv.h
    case kCamPolynomial4: *m = {kDistPoly4, false, false, 4}; return true;                    // camera_polynomial_4.h
    case kCamRadialFisheye: *m = {kDistRadial2, true, true, 2}; return true;                  // camera_radial_fisheye.cc
    case kCamSimpleRadialFisheye: *m = {kDistRadial1, true, true, 1}; return true;            // camera_simple_radial_fisheye.cc
    case kCamThinPrism: *m = {kDistThinPrism, false, false, 8}; return true;                  // camera_thin_prism.h
  }
  return false;
}

// atan2(r, 1.f) of camera_base_impl_fisheye.h:68,104,135. The reference gets whatever its libm's atan2f returns: correctly rounded
// with glibc >= 2.41 (CORE-MATH), up to 1 ulp off (and dependent on the CPU's FMA dispatch) with older glibc. The oracle pins the
// correctly rounded value, computed as the rounding of the fp64 arctangent — the same definition the device code uses. The same holds
// for the atanf / tanf of the FOV camera (camera_fisheye_fov.h:60-62,80-86, .cc:41-42).
static inline float atan_r1(float r) { return (float)std::atan((double)r); }
static inline float tan_r1(float r) { return (float)std::tan((double)r); }

static inline int f2i(float v) {   // x86 cvttss2si semantics made explicit
  if (!(v > -2147483904.f && v < 2147483648.f)) return std::numeric_limits<int>::min();
  return (int)v;
}

struct Camera {
  int type = kCamPinhole;
  int dist = kDistNone; bool fisheye = false, unique_focal = false; int nd = 0;
  int w = 0, h = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0, fx_inv = 0, fy_inv = 0, cx_inv = 0, cy_inv = 0;
  float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // distortion parameters in GetParameters order (thin prism: k1 k2 p1 p2 k3 k4 sx1 sy1)
  float two_tan = 0, image_radius = 0;     // FOV camera: two_tan_omega_half_, image_radius_
  float cutoff2 = std::numeric_limits<float>::infinity();         // CameraBaseImpl::radius_cutoff_squared_ of this camera
  float inner_cutoff2 = std::numeric_limits<float>::infinity();   // fisheye cameras: the inner model's cut-off

  static int param_count(int type) { ModelInfo m; if (!model_info(type, &m)) return -1; return (m.unique_focal ? 3 : 4) + m.nd; }
  int np() const { return (unique_focal ? 3 : 4) + nd; }
  int nbase() const { return unique_focal ? 3 : 4; }
  static bool known(int type) { ModelInfo m; return model_info(type, &m); }

  void set(int type_, int w_, int h_, const float* p) {
    ModelInfo m; model_info(type_, &m);
    type = type_; dist = m.dist; fisheye = m.fisheye; unique_focal = m.unique_focal; nd = m.nd; w = w_; h = h_;
    if (unique_focal) { fx = fy = p[0]; cx = p[1]; cy = p[2]; } else { fx = p[0]; fy = p[1]; cx = p[2]; cy = p[3]; }
    fx_inv = (float)(1.0 / fx); fy_inv = (float)(1.0 / fy);                 // camera_base.cc:83
    cx_inv = (float)(-1.0 * cx / fx); cy_inv = (float)(-1.0 * cy / fy);
    for (int i = 0; i < 8; ++i) d[i] = i < nd ? p[nbase() + i] : 0.f;
    if (dist == kDistFOV) { two_tan = 2.0f * tan_r1(0.5f * d[0]); image_radius = (float)(M_PI / (double)(2 * d[0])); }   // camera_fisheye_fov.cc:38-43
    cutoff2 = inner_cutoff2 = std::numeric_limits<float>::infinity();
    // the constructors that run InitCutoff: the generic search (camera_base_impl.h:410-462) for the tangential / rational / thin-prism
    // models, the radial one (camera_base_impl_radial.h:143-171) for Radial / Polynomial / Polynomial4, the closed form of
    // SimpleRadialCamera::InitCutoff (camera_simple_radial.cc:53-57). A fisheye camera keeps +inf; its INNER model carries the cut-off.
    float c2 = std::numeric_limits<float>::infinity();
    if (dist == kDistPolyTan || dist == kDistOpenCV || dist == kDistThinPrism) c2 = generic_cutoff();
    else if (dist == kDistRadial2 || dist == kDistPoly3 || dist == kDistPoly4) c2 = radial_cutoff();
    else if (dist == kDistRadial1) { if (d[0] < 0) c2 = -1.f / (3 * d[0]); }
    if (fisheye) inner_cutoff2 = c2; else cutoff2 = c2;
  }
  void get_params(float* p) const {
    if (unique_focal) { p[0] = fx; p[1] = cx; p[2] = cy; } else { p[0] = fx; p[1] = fy; p[2] = cx; p[3] = cy; }
    for (int i = 0; i < nd; ++i) p[nbase() + i] = d[i];
  }
  Camera scaled_half() const {                                             // camera_base_impl.h:70-89, factor 0.5
    const float f = 0.5f;
    float p[12]; get_params(p);
    if (!unique_focal) { p[0] *= f; p[1] *= f; p[2] = f * (cx + 0.5f) - 0.5f; p[3] = f * (cy + 0.5f) - 0.5f; }
    else { p[0] *= f; p[1] = f * (cx + 0.5f) - 0.5f; p[2] = f * (cy + 0.5f) - 0.5f; }
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

  // ---- RadialBase models (camera_base_impl_radial.h:54-58): Distort = p * DistortionFactor(|p|^2) ----
  float radial_factor(float r2) const {
    switch (dist) {
      case kDistRadial1: return 1.0f + r2 * d[0];                                              // camera_simple_radial.h:62-64
      case kDistRadial2: return 1.0f + r2 * (d[0] + r2 * d[1]);                                // camera_radial.h:63-67
      case kDistPoly3: return 1.0f + r2 * (d[0] + r2 * (d[1] + r2 * d[2]));                    // camera_polynomial.h:60-65
      default: return 1.0f + r2 * (d[0] + r2 * (d[1] + r2 * (d[2] + r2 * d[3])));              // camera_polynomial_4.h:60-66
    }
  }
  float radial_deriv_r(float r2) const {                                   // DistortedDerivativeByNormalized(const float r2)
    switch (dist) {
      case kDistRadial1: return 1.f + 3.f * d[0] * r2;
      case kDistRadial2: return 1.f + r2 * (3.f * d[0] + r2 * 5.f * d[1]);
      case kDistPoly3: return 1.0f + r2 * (3.0f * d[0] + r2 * (5.0f * d[1] + r2 * 7.0f * d[2]));
      default: return 1.0f + r2 * (3.0f * d[0] + r2 * (5.0f * d[1] + r2 * (7.0f * d[2] + r2 * (9.0f * d[3]))));
    }
  }
  void radial_deriv(float nx, float ny, float J[4]) const {
    if (dist == kDistRadial1) {                                            // camera_simple_radial.h:75-87
      const float k1 = d[0], nxs = nx * nx, nys = ny * ny, ru2 = nxs + nys;
      J[0] = k1 * (ru2 + 2 * nxs) + 1; J[1] = 2 * nx * ny * k1; J[2] = J[1]; J[3] = k1 * (ru2 + 2 * nys) + 1;
      return;
    }
    const float nx2 = nx * nx, ny2 = ny * ny, nxny = nx * ny, r2 = nx2 + ny2;
    float term1, term2;
    if (dist == kDistRadial2) { const float k1 = d[0], k2 = d[1]; term1 = 2 * k1 + r2 * (4 * k2); term2 = 1 + r2 * (k1 + r2 * (k2)); }
    else if (dist == kDistPoly3) { const float k1 = d[0], k2 = d[1], k3 = d[2]; term1 = 2 * k1 + r2 * (4 * k2 + r2 * 6 * k3); term2 = 1 + r2 * (k1 + r2 * (k2 + r2 * k3)); }
    else { const float k1 = d[0], k2 = d[1], k3 = d[2], k4 = d[3]; term1 = 2 * k1 + r2 * (4 * k2 + r2 * (6 * k3 + r2 * 8 * k4)); term2 = 1 + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4))); }
    J[0] = nx2 * term1 + term2; J[1] = nxny * term1; J[2] = J[1]; J[3] = ny2 * term1 + term2;
  }

  // ---- inner model: Distort / DistortedDerivativeByNormalized / DistortedDerivativeByDistortionParameters (2 x 8 row-major, first nd columns) ----
  void inner_distort(float x, float y, float* ox, float* oy) const {
    switch (dist) {
      case kDistNone: *ox = x; *oy = y; return;
      case kDistThinPrism: tp_distort(x, y, ox, oy); return;
      case kDistPolyTan: {                                                 // camera_polynomial_tangential.h:60-74
        const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3];
        const float x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2;
        const float radial = 1 + r2 * (k1 + r2 * k2);
        *ox = x * radial + (2.f * p1 * xy + p2 * (r2 + 2.f * x2)); *oy = y * radial + (2.f * p2 * xy + p1 * (r2 + 2.f * y2));
        return;
      }
      case kDistOpenCV: {                                                  // camera_full_opencv.h:60-85
        const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4], k4 = d[5], k5 = d[6], k6 = d[7];
        const float x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2, r4 = r2 * r2, r6 = r4 * r2;
        const float radial = (1.f + k1 * r2 + k2 * r4 + k3 * r6) / (1.f + k4 * r2 + k5 * r4 + k6 * r6);
        *ox = radial * x + (2.f * p1 * xy + p2 * (r2 + 2.f * x2)); *oy = radial * y + (2.f * p2 * xy + p1 * (r2 + 2.f * y2));
        return;
      }
      case kDistFOV: {                                                     // camera_fisheye_fov.h:56-64
        const float r = std::sqrt(x * x + y * y);
        const float factor = (r < 1e-6f) ? 1.f : (atan_r1(r * two_tan) / (r * d[0]));
        *ox = x * factor; *oy = y * factor;
        return;
      }
      default: { const float f = radial_factor(x * x + y * y); *ox = x * f; *oy = y * f; return; }
    }
  }
  void inner_deriv(float nx, float ny, float J[4]) const {
    switch (dist) {
      case kDistNone: J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 1; return;
      case kDistThinPrism: tp_deriv(nx, ny, J); return;
      case kDistPolyTan: {                                                 // camera_polynomial_tangential.h:97-117
        const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3];
        const float nx2 = nx * nx, ny2 = ny * ny, r2 = nx2 + ny2;
        const float term1 = 2 * k1 + r2 * 4 * k2;
        const float term2 = 1 + r2 * (k1 + r2 * k2);
        J[0] = nx2 * term1 + term2 + 6 * p2 * nx + 2 * p1 * ny;
        J[1] = nx * ny * term1 + 2 * p1 * nx + 2 * p2 * ny;
        J[2] = J[1];
        J[3] = ny2 * term1 + term2 + 2 * p2 * nx + 6 * p1 * ny;
        return;
      }
      case kDistOpenCV: {                                                  // camera_full_opencv.h:131-171
        const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4], k4 = d[5], k5 = d[6], k6 = d[7];
        const float x2 = nx * nx, y2 = ny * ny, xy = nx * ny, r2 = x2 + y2, r4 = r2 * r2, r6 = r4 * r2;
        const float num = 1.f + k1 * r2 + k2 * r4 + k3 * r6, den = 1.f + k4 * r2 + k5 * r4 + k6 * r6;
        const float radial = num / den;
        const float d_num = 2 * k1 + 4 * k2 * r2 + 6 * k3 * r4, d_den = 2 * k4 + 4 * k5 * r2 + 6 * k6 * r4;
        const float d_radial = (d_num * den - d_den * num) / (den * den);
        const float d_tan_x_nx = 2 * ny * p1 + 6 * p2 * nx, d_tan_y_ny = 2 * nx * p2 + 6 * p1 * ny;
        const float d_tan_y_nx = 2 * ny * p2 + 2 * p1 * nx, d_tan_x_ny = 2 * nx * p1 + 2 * p2 * ny;
        J[0] = radial + x2 * d_radial + d_tan_x_nx; J[1] = xy * d_radial + d_tan_x_ny;
        J[2] = xy * d_radial + d_tan_y_nx; J[3] = radial + y2 * d_radial + d_tan_y_ny;
        return;
      }
      case kDistFOV: {                                                     // camera_fisheye_fov.h:122-149
        const float omega = d[0];
        const float nx_times_ny = nx * ny, nxs = nx * nx, nys = ny * ny, radius_square = nxs + nys, radius = sqrtf(radius_square);
        if (radius < 1e-6f) { J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 1; return; }
        const float rdw = atan_r1(radius * two_tan);
        const float two_tan_sq = two_tan * two_tan;
        const float part1 = omega * radius_square * radius;
        const float part2 = omega * (two_tan_sq * radius_square + 1) * radius_square;
        const float part3 = rdw / (omega * radius);
        J[0] = part3 - (nxs * rdw) / part1 + (nxs * two_tan) / part2;
        J[1] = nx_times_ny * (two_tan / part2 - rdw / part1);
        J[2] = J[1];
        J[3] = part3 - (nys * rdw) / part1 + (nys * two_tan) / part2;
        return;
      }
      default: radial_deriv(nx, ny, J); return;
    }
  }
  void inner_deriv_params(float nx, float ny, float D[16]) const {
    for (int i = 0; i < 16; ++i) D[i] = 0;
    switch (dist) {
      case kDistNone: return;
      case kDistThinPrism: tp_deriv_params(nx, ny, D); return;
      case kDistPolyTan: {                                                 // camera_polynomial_tangential.h:77-94
        const float nx2 = nx * nx, ny2 = ny * ny, two_nx_ny = 2.f * nx * ny, r2 = nx2 + ny2;
        D[0] = nx * r2; D[1] = D[0] * r2; D[2] = two_nx_ny; D[3] = (r2 + 2.f * nx2);
        D[8] = ny * r2; D[9] = D[8] * r2; D[10] = (r2 + 2.f * ny2); D[11] = two_nx_ny;
        return;
      }
      case kDistOpenCV: {                                                  // camera_full_opencv.h:88-128
        const float k1 = d[0], k2 = d[1], k3 = d[4], k4 = d[5], k5 = d[6], k6 = d[7];
        const float x2 = nx * nx, y2 = ny * ny, r2 = x2 + y2, r4 = r2 * r2, r6 = r4 * r2;
        const float num = 1.f + k1 * r2 + k2 * r4 + k3 * r6, den = 1.f + k4 * r2 + k5 * r4 + k6 * r6;
        const float radial = num / den;
        D[0] = nx * r2 / den; D[1] = nx * r4 / den; D[2] = nx * 2.f * ny; D[3] = (r2 + 2 * x2); D[4] = nx * r6 / den;
        D[5] = -nx * r2 * radial / den; D[6] = -nx * r4 * radial / den; D[7] = -nx * r6 * radial / den;
        D[8] = ny * r2 / den; D[9] = ny * r4 / den; D[10] = (r2 + 2 * y2); D[11] = ny * 2.f * nx; D[12] = ny * r6 / den;
        D[13] = -ny * r2 * radial / den; D[14] = -ny * r4 * radial / den; D[15] = -ny * r6 * radial / den;
        return;
      }
      case kDistFOV: {                                                     // camera_fisheye_fov.h:92-119
        const float omega = d[0];
        const float radius_square = nx * nx + ny * ny, radius = sqrtf(radius_square);
        const float four_tan_sq = two_tan * two_tan;
        const float tan_sq_plus_one = 0.25f * four_tan_sq + 1.f;
        const float denominator_1 = omega * (four_tan_sq * radius_square + 1.f);
        const float numerator_2 = atan_r1(two_tan * radius);
        const float denominator_2 = omega * omega * radius;
        D[0] = (radius < 1e-6f) ? 0.f : ((nx * tan_sq_plus_one) / denominator_1 - (nx * numerator_2) / denominator_2);
        D[8] = (radius < 1e-6f) ? 0.f : ((ny * tan_sq_plus_one) / denominator_1 - (ny * numerator_2) / denominator_2);
        return;
      }
      default: {                                                           // camera_radial.h:70-78, camera_polynomial.h:68-78, camera_polynomial_4.h:69-81
        const float r2 = nx * nx + ny * ny;
        D[0] = nx * r2; D[8] = ny * r2;
        for (int i = 1; i < nd; ++i) { D[i] = D[i - 1] * r2; D[8 + i] = D[8 + i - 1] * r2; }
        return;
      }
    }
  }

  // ---- Child::Distort / DistortedDerivativeByNormalized / ...ByDistortionParameters (FisheyeBase wrapper: camera_base_impl_fisheye.h:65-146) ----
  void distort(float x, float y, float* ox, float* oy) const {
    if (!fisheye) { inner_distort(x, y, ox, oy); return; }
    const float r = std::sqrt(x * x + y * y);
    if (r > 1e-6f) {
      const float atan_r = atan_r1(r);
      if (atan_r * atan_r > inner_cutoff2) { *ox = x * std::numeric_limits<float>::infinity(); *oy = y * std::numeric_limits<float>::infinity(); return; }
      const float theta_by_r = atan_r / r;
      inner_distort(x * theta_by_r, y * theta_by_r, ox, oy);
    } else {
      inner_distort(x, y, ox, oy);
    }
  }
  void distort_deriv(float nx, float ny, float J[4]) const {
    if (!fisheye) { inner_deriv(nx, ny, J); return; }
    const float nx_ny = nx * ny, nx2 = nx * nx, ny2 = ny * ny, r2 = nx2 + ny2;
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
      float Jd[4]; inner_deriv(theta_by_r * nx, theta_by_r * ny, Jd);
      J[0] = Jd[0] * a + Jd[1] * c; J[1] = Jd[0] * b + Jd[1] * dd;
      J[2] = Jd[2] * a + Jd[3] * c; J[3] = Jd[2] * b + Jd[3] * dd;
    } else {
      inner_deriv(nx, ny, J);
    }
  }
  void distort_deriv_params(float nx, float ny, float D[16]) const {
    if (!fisheye) { inner_deriv_params(nx, ny, D); return; }
    const float r = std::sqrt(nx * nx + ny * ny);
    if (r > 1e-6f) {
      const float atan_r = atan_r1(r);
      if (atan_r * atan_r > inner_cutoff2) { for (int i = 0; i < 16; ++i) D[i] = 0; return; }
      const float theta_by_r = atan_r / r;
      inner_deriv_params(theta_by_r * nx, theta_by_r * ny, D);
    } else {
      inner_deriv_params(nx, ny, D);
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
    const int n = np(), nb = nbase();
    const float nx = p.x / p.z, ny = p.y / p.z;
    if (nx * nx + ny * ny > cutoff2) { for (int i = 0; i < 2 * n; ++i) o[i] = 0; return; }
    float dx, dy; distort(nx, ny, &dx, &dy);
    if (!unique_focal) {
      o[0] = dx; o[1] = 0.f; o[2] = 1.f; o[3] = 0.f;
      o[n + 0] = 0.f; o[n + 1] = dy; o[n + 2] = 0.f; o[n + 3] = 1.f;
    } else {
      o[0] = dx; o[1] = 1.f; o[2] = 0.f;
      o[n + 0] = dy; o[n + 1] = 0.f; o[n + 2] = 1.f;
    }
    if (nd > 0) {
      float D[16]; distort_deriv_params(nx, ny, D);
      for (int i = 0; i < nd; ++i) { o[nb + i] = fx * D[i]; o[n + nb + i] = fy * D[8 + i]; }
    }
  }

  // Undistort(distorted): camera_base_impl.h:251-253 = the generic IterativeUndistort started at the distorted point (also for the
  // RadialBase models: their scalar IterativeUndistort only serves InitCutoff); identity for Pinhole / SimplePinhole
  // (camera_pinhole.h:65-68); closed form for the FOV camera (camera_fisheye_fov.h:78-88); camera_base_impl_fisheye.h:80-91 for the
  // fisheye cameras: inner Undistort, then r -> tan(r).
  void inner_undistort(float dx, float dy, float* ox, float* oy) const {
    if (dist == kDistNone) { *ox = dx; *oy = dy; return; }
    if (dist == kDistFOV) {
      const float r = std::sqrt(dx * dx + dy * dy);
      const float factor = (r < 1e-6f) ? 1.f : (r > image_radius) ? std::numeric_limits<float>::infinity() : (tan_r1(r * d[0]) / (r * two_tan));
      *ox = factor * dx; *oy = factor * dy;
      return;
    }
    iterative_undistort(dx, dy, dx, dy, ox, oy);
  }
  void undistort(float dx, float dy, float* ox, float* oy) const {
    float ux, uy; inner_undistort(dx, dy, &ux, &uy);
    if (fisheye) {
      const float r = std::sqrt(ux * ux + uy * uy);
      const float factor = (r < 1e-6f) ? 1.f : (r > M_PI / 2.f) ? std::numeric_limits<float>::infinity() : tanf(r) / r;
      ux = factor * ux; uy = factor * uy;
    }
    *ox = ux; *oy = uy;
  }
  bool has_lookup() const { return !(dist == kDistNone && !fisheye) && dist != kDistFOV; }   // models whose ImageToNormalized reads the table
  // InitializeUndistortionLookup (camera_base_impl.h:255-269): Undistort at every integer pixel; w*h x 2 floats
  void undistortion_lookup(float* table) const {
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) undistort(fx_inv * x + cx_inv, fy_inv * y + cy_inv, &table[2 * ((size_t)y * w + x)], &table[2 * ((size_t)y * w + x) + 1]);
  }
  // ImageToNormalized(pixel_position) (camera_base_impl.h:187-212): bilinear filter of the lookup; camera_pinhole.h:60-63 /
  // camera_simple_pinhole.h: ImageToDistorted; camera_fisheye_fov.h:66-76: Undistort(ImageToDistorted).
  // The reference reads row h of the table when the clamped y is exactly h - 1 (weight 0): defined here as a clamped (finite) read.
  void image_to_normalized(const float* table, float px, float py, float* ox, float* oy) const {
    if (dist == kDistNone && !fisheye) { *ox = fx_inv * px + cx_inv; *oy = fy_inv * py + cy_inv; return; }
    if (dist == kDistFOV) { undistort(fx_inv * px + cx_inv, fy_inv * py + cy_inv, ox, oy); return; }
    const float cxp = std::max(std::min(px, w - 1.001f), 0.f), cyp = std::max(std::min(py, h - 1.00f), 0.f);
    const int ix = (int)cxp, iy = (int)cyp;
    const float fx_ = cxp - (float)ix, fy_ = cyp - (float)iy;
    const int ix1 = std::min(ix + 1, w - 1), iy1 = std::min(iy + 1, h - 1);
    const float* tl = &table[2 * ((size_t)iy * w + ix)]; const float* tr = &table[2 * ((size_t)iy * w + ix1)];
    const float* bl = &table[2 * ((size_t)iy1 * w + ix)]; const float* br = &table[2 * ((size_t)iy1 * w + ix1)];
    *ox = (1 - fy_) * ((1 - fx_) * tl[0] + fx_ * tr[0]) + fy_ * ((1 - fx_) * bl[0] + fx_ * br[0]);
    *oy = (1 - fy_) * ((1 - fx_) * tl[1] + fx_ * tr[1]) + fy_ * ((1 - fx_) * bl[1] + fx_ * br[1]);
  }

  // ---- generic cut-off search (camera_base_impl.h:214-250, 276-328, 410-462) on the inner model ----
  bool iterative_undistort(float tx, float ty, float sx, float sy, float* ox, float* oy) const {
    float ux = sx, uy = sy;
    bool converged = false;
    for (int i = 0; i < 100; ++i) {
      float cxd, cyd; inner_distort(ux, uy, &cxd, &cyd);
      const float ex = cxd - tx, ey = cyd - ty;
      if (ex * ex + ey * ey < 1e-10f) { converged = true; break; }
      float J[4]; inner_deriv(ux, uy, J);
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
  bool undistort_from_inside(float tx, float ty, float* bx, float* by, float* second_r2, bool* second_available) const {
    const int kNumGridSteps = 10; const float kGridHalfExtent = 1.5f, kImproveThreshold = 0.99f;
    bool converged = false; *second_available = false;
    float best_radius = std::numeric_limits<float>::infinity(), second_best_radius = std::numeric_limits<float>::infinity();
    float best_x = 0, best_y = 0, sbx = std::numeric_limits<float>::infinity(), sby = std::numeric_limits<float>::infinity();
    for (int y = 0; y < kNumGridSteps; ++y) {
      const float iy = ty + kGridHalfExtent * (y - 0.5f * kNumGridSteps) / (0.5f * kNumGridSteps);
      for (int x = 0; x < kNumGridSteps; ++x) {
        const float ix = tx + kGridHalfExtent * (x - 0.5f * kNumGridSteps) / (0.5f * kNumGridSteps);
        float rx, ry;
        if (iterative_undistort(tx, ty, ix, iy, &rx, &ry)) {
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
  float generic_cutoff() const {
    const float kIncreaseFactor = 1.01f, inf = std::numeric_limits<float>::infinity();
    // border pixels in the reference's order (camera_base_impl.h:418-447); the two results are a max and a min over them, so the pixels
    // can be searched in parallel (test infrastructure: the search is 100 starts x 100 iterations per pixel)
    std::vector<float> bx_, by_;
    for (int x = 0; x < w; ++x) { bx_.push_back((float)x); by_.push_back(0.f); bx_.push_back((float)x); by_.push_back((float)(h - 1)); }
    for (int y = 0; y < h; ++y) { bx_.push_back(0.f); by_.push_back((float)y); bx_.push_back((float)(w - 1)); by_.push_back((float)y); }
    float min_candidate = 0, max_candidate = inf;
    const long n = (long)bx_.size();
#pragma omp parallel for schedule(dynamic, 16) reduction(max : min_candidate) reduction(min : max_candidate)
    for (long i = 0; i < n; ++i) {
      float bx, by, s2; bool sa;
      if (undistort_from_inside(fx_inv * bx_[i] + cx_inv, fy_inv * by_[i] + cy_inv, &bx, &by, &s2, &sa)) {
        min_candidate = std::max(bx * bx + by * by, min_candidate);
        if (sa) max_candidate = std::min(s2, max_candidate);
      }
    }
    return std::min(min_candidate * kIncreaseFactor, max_candidate);
  }

  // ---- RadialBase cut-off search (camera_base_impl_radial.h:60-171): one scalar search from the farthest image corner ----
  float radial_iterative_undistort(float distorted_r, float starting_r, bool* converged) const {
    *converged = false;
    float undistorted_r = starting_r, undistorted_r2 = starting_r * starting_r;
    for (int i = 0; i < 100; ++i) {
      const float r_candidate = undistorted_r * radial_factor(undistorted_r2);
      const float delta_r = r_candidate - distorted_r;
      if (delta_r * delta_r < 1e-10f) { *converged = true; break; }
      const float deriv = radial_deriv_r(undistorted_r2);
      const float step = delta_r / deriv;
      undistorted_r -= step;
      undistorted_r2 = undistorted_r * undistorted_r;
    }
    return undistorted_r;
  }
  float radial_cutoff() const {
    const float kIncreaseFactor = 1.01f, kImproveThreshold = 0.99f, inf = std::numeric_limits<float>::infinity();
    auto corner_r = [&](float px, float py) { const float x = fx_inv * px + cx_inv, y = fy_inv * py + cy_inv; return std::sqrt(x * x + y * y); };
    float test_image_radius = corner_r(0, 0);                                           // (sic) width_ / height_, not the last pixel
    test_image_radius = std::max(test_image_radius, corner_r(0, (float)h));
    test_image_radius = std::max(test_image_radius, corner_r((float)w, 0));
    test_image_radius = std::max(test_image_radius, corner_r((float)w, (float)h));
    bool converged = false, second_available = false;
    float best = inf, second = inf;
    for (int i = 0; i < 10; ++i) {
      const float init_radius = (float)(test_image_radius + 1.5f * (i - 0.5 * 10) / (0.5f * 10));   // (sic) 0.5 is a double here
      bool ok;
      const float result = radial_iterative_undistort(test_image_radius, init_radius, &ok);
      if (ok) {
        if (result < kImproveThreshold * best) { second = best; second_available = converged; best = result; converged = true; }
        else if (result > 1 / kImproveThreshold * best && result < kImproveThreshold * second) { second = result; second_available = true; }
      }
    }
    if (converged && best > 0) {
      if (second_available && second > 0) return std::min(best * best * kIncreaseFactor, second * second);
      return best * best * kIncreaseFactor;
    }
    return inf;
  }
};

}  // namespace orc
#endif
