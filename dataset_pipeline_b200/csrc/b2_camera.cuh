// Camera models of Path B on the device: projection, its derivatives by the 3-D point and by the intrinsics, and the radius
// cut-off search the reference runs in the camera constructors (so: for every pyramid level of every LM trial state).
// All 15 models of /root/reference/src/camera (camera::CameraBase::Type, camera_base.h:67-84), described by three properties:
//   the distortion function (of the camera itself or of the inner model of a FisheyeBase<> camera)
//     none        PinholeCamera (4), SimplePinholeCamera (7)                       camera_pinhole.h:40-86, camera_simple_pinhole.h
//     radial 1    SimpleRadialCamera (9), SimpleRadialFisheyeCamera (13)          camera_simple_radial.h:62-91, .cc:53-57
//     radial 2    RadialCamera (8), RadialFisheyeCamera (12)                      camera_radial.h:63-112
//     poly 3 / 4  PolynomialCamera (1); Polynomial4Camera (11), FisheyePolynomial4Camera (6)   camera_polynomial.h, camera_polynomial_4.h
//     tangential  PolynomialTangentialCamera (2), FisheyePolynomialTangentialCamera (3)        camera_polynomial_tangential.h:60-117
//     rational    FullOpenCVCamera (10)                                           camera_full_opencv.h:60-171
//     thin prism  ThinPrismCamera (14), BenchmarkCamera (5) = ETH3D's THIN_PRISM_FISHEYE       camera_thin_prism.h:56-139
//     FOV         FisheyeFOVCamera (0)                                            camera_fisheye_fov.h:56-149, .cc:38-50
//   the fisheye wrapper (camera_base_impl_fisheye.h:43-161) and the single focal length (UniqueFocalLength(): parameters f cx cy ...).
//   shared machinery: /root/reference/src/camera/camera_base_impl.h:70-89,155-164,214-250,276-328,333-462, camera_base_impl_radial.h:54-171
// Arithmetic is written in the reference's evaluation order and the file is built with -fmad=false, so it is bit-identical to
// the CPU; atan2(r, 1.f) / atanf / tanf are the correctly rounded fp32 values (fp64 function, rounded once) — what glibc >= 2.41
// returns; older glibc differs by <= 1 ulp in ~5 % of the calls, which is why the oracle pins the same definition (orc_camera.h).
// K16 kr_cutoff_starts / kr_cutoff_points / kr_cutoff_final: the generic InitCutoff as three kernels (one thread per border pixel and
// Gauss-Newton start; one thread per border pixel replaying the reference's sequential best / second-best bookkeeping over
// its 100 starts; one block for the max / min over border pixels); kr_cutoff_radial: the RadialBase search, one thread per camera.
#pragma once
#include "b2_common.cuh"

namespace b2 {

enum {
  kCamFOV = 0, kCamPolynomial = 1, kCamPolynomialTangential = 2, kCamFisheyePolynomialTangential = 3, kCamPinhole = 4, kCamBenchmark = 5,
  kCamFisheyePolynomial4 = 6, kCamSimplePinhole = 7, kCamRadial = 8, kCamSimpleRadial = 9, kCamFullOpenCV = 10, kCamPolynomial4 = 11,
  kCamRadialFisheye = 12, kCamSimpleRadialFisheye = 13, kCamThinPrism = 14
};
enum { kDistNone = 0, kDistRadial1, kDistRadial2, kDistPoly3, kDistPoly4, kDistPolyTan, kDistOpenCV, kDistThinPrism, kDistFOV };
static constexpr int kMaxIntrinsics = 12;

struct Cam {
  int w, h; float fx, fy, cx, cy, fx_inv, fy_inv, cx_inv, cy_inv;
  int type; float cutoff2, inner_cutoff2;   // radius_cutoff_squared_ of the camera itself / of a fisheye camera's inner model
  float d[8];                               // distortion parameters in GetParameters order (zero beyond nd)
  int dist, fisheye, unique_focal, nd;      // distortion function, FisheyeBase<> wrapper, UniqueFocalLength(), number of distortion parameters
  float two_tan, image_radius;              // FOV camera: two_tan_omega_half_, image_radius_
};

struct CamModel { int dist, fisheye, unique_focal, nd; };
__host__ __device__ inline bool cam_model(int type, CamModel* m) {
  switch (type) {
    case kCamFOV: *m = {kDistFOV, 0, 0, 1}; return true;
    case kCamPolynomial: *m = {kDistPoly3, 0, 0, 3}; return true;
    case kCamPolynomialTangential: *m = {kDistPolyTan, 0, 0, 4}; return true;
    case kCamFisheyePolynomialTangential: *m = {kDistPolyTan, 1, 0, 4}; return true;
    case kCamPinhole: *m = {kDistNone, 0, 0, 0}; return true;
    case kCamBenchmark: *m = {kDistThinPrism, 1, 0, 8}; return true;
    case kCamFisheyePolynomial4: *m = {kDistPoly4, 1, 0, 4}; return true;
    case kCamSimplePinhole: *m = {kDistNone, 0, 1, 0}; return true;
    case kCamRadial: *m = {kDistRadial2, 0, 1, 2}; return true;
    case kCamSimpleRadial: *m = {kDistRadial1, 0, 1, 1}; return true;
    case kCamFullOpenCV: *m = {kDistOpenCV, 0, 0, 8}; return true;
    case kCamPolynomial4: *m = {kDistPoly4, 0, 0, 4}; return true;
    case kCamRadialFisheye: *m = {kDistRadial2, 1, 1, 2}; return true;
    case kCamSimpleRadialFisheye: *m = {kDistRadial1, 1, 1, 1}; return true;
    case kCamThinPrism: *m = {kDistThinPrism, 0, 0, 8}; return true;
  }
  return false;
}
__host__ __device__ inline int cam_param_count(int type) { CamModel m; return cam_model(type, &m) ? (m.unique_focal ? 3 : 4) + m.nd : -1; }
// Width of the Jacobian rows the kernels carry for a model: 4 for the distortion-free models (SimplePinhole: 3 + one zero column),
// 12 for the others (zero columns beyond the model's parameter count; the host drops them when it assembles H and b).
__host__ __device__ inline int cam_kernel_ni(int type) { CamModel m; return (cam_model(type, &m) && m.dist == kDistNone && !m.fisheye) ? 4 : 12; }
__host__ __device__ inline bool cam_plain(const Cam& c) { return c.dist == kDistNone && !c.fisheye; }   // no distortion at all

// ---- thin-prism distortion (camera_thin_prism.h:56-139) ----
__device__ __forceinline__ void tp_distort(const float* d, float x, float y, float* ox, float* oy) {
  const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4], k4 = d[5], sx1 = d[6], sy1 = d[7];
  const float x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2;
  const float radial = 1 + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4)));
  const float dx = 2.f * p1 * xy + p2 * (r2 + 2.f * x2) + sx1 * r2;
  const float dy = 2.f * p2 * xy + p1 * (r2 + 2.f * y2) + sy1 * r2;
  *ox = x * radial + dx; *oy = y * radial + dy;
}
__device__ __forceinline__ void tp_deriv(const float* d, float nx, float ny, float J[4]) {
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
__device__ __forceinline__ void tp_deriv_params(float nx, float ny, float D[16]) {
  const float nx2 = nx * nx, ny2 = ny * ny, two_nx_ny = 2.f * nx * ny, r2 = nx2 + ny2;
  D[0] = nx * r2; D[1] = D[0] * r2; D[2] = two_nx_ny; D[3] = (r2 + 2.f * nx2); D[4] = D[1] * r2; D[5] = D[4] * r2; D[6] = r2; D[7] = 0;
  D[8] = ny * r2; D[9] = D[8] * r2; D[10] = (r2 + 2.f * ny2); D[11] = two_nx_ny; D[12] = D[9] * r2; D[13] = D[12] * r2; D[14] = 0; D[15] = r2;
}

__device__ __forceinline__ float atan_pos(float r) { return (float)atan((double)r); }   // atan2(r, 1.f) / atanf(r), r >= 0
__device__ __forceinline__ float tan_r1(float r) { return (float)tan((double)r); }

// ---- RadialBase models (camera_base_impl_radial.h:54-58): Distort = p * DistortionFactor(|p|^2) ----
__device__ __forceinline__ float radial_factor(const Cam& c, float r2) {
  const float* d = c.d;
  switch (c.dist) {
    case kDistRadial1: return 1.0f + r2 * d[0];
    case kDistRadial2: return 1.0f + r2 * (d[0] + r2 * d[1]);
    case kDistPoly3: return 1.0f + r2 * (d[0] + r2 * (d[1] + r2 * d[2]));
    default: return 1.0f + r2 * (d[0] + r2 * (d[1] + r2 * (d[2] + r2 * d[3])));
  }
}
__device__ __forceinline__ float radial_deriv_r(const Cam& c, float r2) {      // DistortedDerivativeByNormalized(const float r2)
  const float* d = c.d;
  switch (c.dist) {
    case kDistRadial1: return 1.f + 3.f * d[0] * r2;
    case kDistRadial2: return 1.f + r2 * (3.f * d[0] + r2 * 5.f * d[1]);
    case kDistPoly3: return 1.0f + r2 * (3.0f * d[0] + r2 * (5.0f * d[1] + r2 * 7.0f * d[2]));
    default: return 1.0f + r2 * (3.0f * d[0] + r2 * (5.0f * d[1] + r2 * (7.0f * d[2] + r2 * (9.0f * d[3]))));
  }
}
__device__ __forceinline__ void radial_deriv(const Cam& c, float nx, float ny, float J[4]) {
  const float* d = c.d;
  if (c.dist == kDistRadial1) {
    const float k1 = d[0], nxs = nx * nx, nys = ny * ny, ru2 = nxs + nys;
    J[0] = k1 * (ru2 + 2 * nxs) + 1; J[1] = 2 * nx * ny * k1; J[2] = J[1]; J[3] = k1 * (ru2 + 2 * nys) + 1;
    return;
  }
  const float nx2 = nx * nx, ny2 = ny * ny, nxny = nx * ny, r2 = nx2 + ny2;
  float term1, term2;
  if (c.dist == kDistRadial2) { const float k1 = d[0], k2 = d[1]; term1 = 2 * k1 + r2 * (4 * k2); term2 = 1 + r2 * (k1 + r2 * (k2)); }
  else if (c.dist == kDistPoly3) { const float k1 = d[0], k2 = d[1], k3 = d[2]; term1 = 2 * k1 + r2 * (4 * k2 + r2 * 6 * k3); term2 = 1 + r2 * (k1 + r2 * (k2 + r2 * k3)); }
  else { const float k1 = d[0], k2 = d[1], k3 = d[2], k4 = d[3]; term1 = 2 * k1 + r2 * (4 * k2 + r2 * (6 * k3 + r2 * 8 * k4)); term2 = 1 + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4))); }
  J[0] = nx2 * term1 + term2; J[1] = nxny * term1; J[2] = J[1]; J[3] = ny2 * term1 + term2;
}

// ---- the inner model: Distort / DistortedDerivativeByNormalized (row-major 2x2) / DistortedDerivativeByDistortionParameters (2 x 8) ----
__device__ __forceinline__ void cam_inner_distort(const Cam& c, float x, float y, float* ox, float* oy) {
  const float* d = c.d;
  switch (c.dist) {
    case kDistNone: *ox = x; *oy = y; return;
    case kDistThinPrism: tp_distort(d, x, y, ox, oy); return;
    case kDistPolyTan: {
      const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3];
      const float x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2;
      const float radial = 1 + r2 * (k1 + r2 * k2);
      *ox = x * radial + (2.f * p1 * xy + p2 * (r2 + 2.f * x2)); *oy = y * radial + (2.f * p2 * xy + p1 * (r2 + 2.f * y2));
      return;
    }
    case kDistOpenCV: {
      const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4], k4 = d[5], k5 = d[6], k6 = d[7];
      const float x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2, r4 = r2 * r2, r6 = r4 * r2;
      const float radial = (1.f + k1 * r2 + k2 * r4 + k3 * r6) / (1.f + k4 * r2 + k5 * r4 + k6 * r6);
      *ox = radial * x + (2.f * p1 * xy + p2 * (r2 + 2.f * x2)); *oy = radial * y + (2.f * p2 * xy + p1 * (r2 + 2.f * y2));
      return;
    }
    case kDistFOV: {
      const float r = sqrtf(x * x + y * y);
      const float factor = (r < 1e-6f) ? 1.f : (atan_pos(r * c.two_tan) / (r * d[0]));
      *ox = x * factor; *oy = y * factor;
      return;
    }
    default: { const float f = radial_factor(c, x * x + y * y); *ox = x * f; *oy = y * f; return; }
  }
}
__device__ __forceinline__ void cam_inner_deriv(const Cam& c, float nx, float ny, float J[4]) {
  const float* d = c.d;
  switch (c.dist) {
    case kDistNone: J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 1; return;
    case kDistThinPrism: tp_deriv(d, nx, ny, J); return;
    case kDistPolyTan: {
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
    case kDistOpenCV: {
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
    case kDistFOV: {
      const float omega = d[0], two_tan = c.two_tan;
      const float nx_times_ny = nx * ny, nxs = nx * nx, nys = ny * ny, radius_square = nxs + nys, radius = sqrtf(radius_square);
      if (radius < 1e-6f) { J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 1; return; }
      const float rdw = atan_pos(radius * two_tan);
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
    default: radial_deriv(c, nx, ny, J); return;
  }
}
__device__ __forceinline__ void cam_inner_deriv_params(const Cam& c, float nx, float ny, float D[16]) {
  if (c.dist == kDistThinPrism) { tp_deriv_params(nx, ny, D); return; }
#pragma unroll
  for (int i = 0; i < 16; ++i) D[i] = 0;
  const float* d = c.d;
  switch (c.dist) {
    case kDistNone: return;
    case kDistPolyTan: {
      const float nx2 = nx * nx, ny2 = ny * ny, two_nx_ny = 2.f * nx * ny, r2 = nx2 + ny2;
      D[0] = nx * r2; D[1] = D[0] * r2; D[2] = two_nx_ny; D[3] = (r2 + 2.f * nx2);
      D[8] = ny * r2; D[9] = D[8] * r2; D[10] = (r2 + 2.f * ny2); D[11] = two_nx_ny;
      return;
    }
    case kDistOpenCV: {
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
    case kDistFOV: {
      const float omega = d[0], two_tan = c.two_tan;
      const float radius_square = nx * nx + ny * ny, radius = sqrtf(radius_square);
      const float four_tan_sq = two_tan * two_tan;
      const float tan_sq_plus_one = 0.25f * four_tan_sq + 1.f;
      const float denominator_1 = omega * (four_tan_sq * radius_square + 1.f);
      const float numerator_2 = atan_pos(two_tan * radius);
      const float denominator_2 = omega * omega * radius;
      D[0] = (radius < 1e-6f) ? 0.f : ((nx * tan_sq_plus_one) / denominator_1 - (nx * numerator_2) / denominator_2);
      D[8] = (radius < 1e-6f) ? 0.f : ((ny * tan_sq_plus_one) / denominator_1 - (ny * numerator_2) / denominator_2);
      return;
    }
    default: {
      const float r2 = nx * nx + ny * ny;
      D[0] = nx * r2; D[8] = ny * r2;
      for (int i = 1; i < c.nd; ++i) { D[i] = D[i - 1] * r2; D[8 + i] = D[8 + i - 1] * r2; }
      return;
    }
  }
}

// Child::Distort (FisheyeBase wrapper: camera_base_impl_fisheye.h:65-78)
__device__ __forceinline__ void cam_distort(const Cam& c, float x, float y, float* ox, float* oy) {
  if (!c.fisheye) { cam_inner_distort(c, x, y, ox, oy); return; }
  const float r = sqrtf(x * x + y * y);
  if (r > 1e-6f) {
    const float atan_r = atan_pos(r);
    if (atan_r * atan_r > c.inner_cutoff2) { *ox = x * INFINITY; *oy = y * INFINITY; return; }
    const float theta_by_r = atan_r / r;
    cam_inner_distort(c, x * theta_by_r, y * theta_by_r, ox, oy);
  } else {
    cam_inner_distort(c, x, y, ox, oy);
  }
}
// Child::DistortedDerivativeByNormalized, row-major 2x2 (camera_base_impl_fisheye.h:96-126)
__device__ __forceinline__ void cam_distort_deriv(const Cam& c, float nx, float ny, float J[4]) {
  if (!c.fisheye) { cam_inner_deriv(c, nx, ny, J); return; }
  const float nx_ny = nx * ny, nx2 = nx * nx, ny2 = ny * ny, r2 = nx2 + ny2;
  const float r = sqrtf(r2);
  if (r > 1e-6f) {
    const float atan_r = atan_pos(r);
    if (atan_r * atan_r > c.inner_cutoff2) { J[0] = J[1] = J[2] = J[3] = 0; return; }
    const float theta_by_r = atan_r / r;
    const float term1 = r2 * (r2 + 1);
    const float term2 = theta_by_r / r2;
    const float a = ny2 * term2 + nx2 / term1;
    const float b = nx_ny / term1 - nx_ny * term2;
    const float dd = nx2 * term2 + ny2 / term1;
    float Jd[4]; cam_inner_deriv(c, theta_by_r * nx, theta_by_r * ny, Jd);
    J[0] = Jd[0] * a + Jd[1] * b; J[1] = Jd[0] * b + Jd[1] * dd;
    J[2] = Jd[2] * a + Jd[3] * b; J[3] = Jd[2] * b + Jd[3] * dd;
  } else {
    cam_inner_deriv(c, nx, ny, J);
  }
}
// Child::DistortedDerivativeByDistortionParameters, 2 x 8 row-major, first nd columns (camera_base_impl_fisheye.h:128-146)
__device__ __forceinline__ void cam_distort_deriv_params(const Cam& c, float nx, float ny, float D[16]) {
  if (!c.fisheye) { cam_inner_deriv_params(c, nx, ny, D); return; }
  const float r = sqrtf(nx * nx + ny * ny);
  if (r > 1e-6f) {
    const float atan_r = atan_pos(r);
    if (atan_r * atan_r > c.inner_cutoff2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) D[i] = 0;
      return;
    }
    const float theta_by_r = atan_r / r;
    cam_inner_deriv_params(c, theta_by_r * nx, theta_by_r * ny, D);
  } else {
    cam_inner_deriv_params(c, nx, ny, D);
  }
}

// NormalizedToImage (camera_base_impl.h:155-164)
__device__ __forceinline__ void cam_project(const Cam& c, float nx, float ny, float* ix, float* iy) {
  const float r2 = nx * nx + ny * ny;
  if (isinf(r2) || r2 > c.cutoff2) { *ix = nx * INFINITY; *iy = ny * INFINITY; return; }
  float dx, dy; cam_distort(c, nx, ny, &dx, &dy);
  *ix = c.fx * dx + c.cx; *iy = c.fy * dy + c.cy;
}
// ImageDerivativeByWorld (camera_base_impl.h:333-360), 2x3 row-major
__device__ __forceinline__ void cam_d_by_world(const Cam& c, float px, float py, float pz, float o[6]) {
  const float nx = px / pz, ny = py / pz;
  if (nx * nx + ny * ny < c.cutoff2) {
    const float z_inv = 1.f / pz;
    float J[4]; cam_distort_deriv(c, nx, ny, J);
    const float N[6] = {1.f * z_inv, 0.f * z_inv, -1.f * nx * z_inv, 0.f * z_inv, 1.f * z_inv, -1.f * ny * z_inv};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o[k] = c.fx * (J[0] * N[k] + J[1] * N[3 + k]);
      o[3 + k] = c.fy * (J[2] * N[k] + J[3] * N[3 + k]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = 0.f;
  }
}
// ImageDerivativeByIntrinsics (camera_base_impl.h:362-408): row 0 in ox[NI], row 1 in oy[NI]; NI = cam_kernel_ni(type), columns beyond
// the model's parameter count are zero
template <int NI>
__device__ __forceinline__ void cam_d_by_intrinsics(const Cam& c, float px, float py, float pz, float* ox, float* oy) {
  const float nx = px / pz, ny = py / pz;
#pragma unroll
  for (int i = 0; i < NI; ++i) { ox[i] = 0.f; oy[i] = 0.f; }
  if (nx * nx + ny * ny > c.cutoff2) return;
  float dx, dy; cam_distort(c, nx, ny, &dx, &dy);
  if (!c.unique_focal) { ox[0] = dx; ox[2] = 1.f; oy[1] = dy; oy[3] = 1.f; }
  else { ox[0] = dx; ox[1] = 1.f; oy[0] = dy; oy[2] = 1.f; }
  if (NI > 4) {
    float D[16]; cam_distort_deriv_params(c, nx, ny, D);
    const int nb = c.unique_focal ? 3 : 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < c.nd) { ox[(nb + i) % NI] = c.fx * D[i]; oy[(nb + i) % NI] = c.fy * D[8 + i]; }
  }
}

// The vertex stage of the reference's depth renderer (opengl/renderer.cc:42-131 with the distortion snippets :581-583 pinhole,
// :630-653 benchmark): camera-space (x, y) <- z * distort(x/z, y/z), or (x, y) * 99 beyond the camera's own cut-off (never for the
// benchmark camera: its radius_cutoff_squared() is +inf). Same operations as oracle/orc_mesh.h:vertex_distort.
// The other models (renderer.cc:154-560): every snippet evaluates z * Distort(x/z, y/z) with the (x, y) * 99 push-out beyond the cut-off;
// GLSL arithmetic is driver-defined, so their individual operation orders are not restated: cam_distort is used (as in the oracle).
__device__ __forceinline__ void cam_vertex_distort(const Cam& c, float* x, float* y, float z) {
  if (cam_plain(c)) return;
  float nx = *x / z, ny = *y / z;
  float r2 = nx * nx + ny * ny;
  if (c.dist != kDistThinPrism) {
    float dx = 0.f, dy = 0.f;
    if (r2 <= c.cutoff2) cam_distort(c, nx, ny, &dx, &dy);
    if (r2 <= c.cutoff2 && isfinite(dx) && isfinite(dy)) { *x = z * dx; *y = z * dy; }
    else { *x = *x * 99.0f; *y = *y * 99.0f; }
    return;
  }
  if (r2 <= c.cutoff2) {
    if (c.fisheye) {
      const float r = sqrtf(r2);
      if (r > 1e-6f) { const float theta_by_r = atan_pos(r) / r; nx = theta_by_r * nx; ny = theta_by_r * ny; }
    }
    const float k1 = c.d[0], k2 = c.d[1], p1 = c.d[2], p2 = c.d[3], k3 = c.d[4], k4 = c.d[5], sx1 = c.d[6], sy1 = c.d[7];
    const float x2 = nx * nx, xy = nx * ny, y2 = ny * ny;
    r2 = x2 + y2;
    const float radial = 1.0f + r2 * (k1 + r2 * (k2 + r2 * (k3 + r2 * k4)));
    *x = z * (radial * nx + 2.0f * p1 * xy + p2 * (r2 + 2.0f * x2) + sx1 * r2);
    *y = z * (radial * ny + 2.0f * p2 * xy + p1 * (r2 + 2.0f * y2) + sy1 * r2);
  } else {
    *x = *x * 99.0f; *y = *y * 99.0f;
  }
}

// x86 cvttss2si semantics (INT_MIN for non-finite / out-of-range), which is what the reference's `int ix = f` does on its hosts;
// CUDA's cast would saturate / map NaN to 0 and let a point beyond the cut-off radius land on pixel 0.
__device__ __forceinline__ int f2i_x86(float v) { return (v > -2147483904.f && v < 2147483648.f) ? (int)v : (int)0x80000000; }

// ------------------------------------------------------------------------------------------------------------------
// K16: the thin-prism cut-off search (camera_base_impl.h:214-250 IterativeUndistort, :276-328 UndistortFromInside,
// :410-462 InitCutoff). `cams` lists the cameras (pyramid levels) to search; border pixel t of camera k:
//   t < 2w: (t/2, t odd ? h-1 : 0), else u = t-2w: (u odd ? w-1 : 0, u/2)      -- the reference's test_points order
// ------------------------------------------------------------------------------------------------------------------
struct CutoffStart { float x, y; int converged; };
struct CutoffPoint { float r2, s2; int converged, second; };

__device__ __forceinline__ void cutoff_border_pixel(const Cam& c, int t, float* px, float* py) {
  if (t < 2 * c.w) { *px = (float)(t >> 1); *py = (t & 1) ? (float)(c.h - 1) : 0.f; }
  else { const int u = t - 2 * c.w; *px = (u & 1) ? (float)(c.w - 1) : 0.f; *py = (float)(u >> 1); }
}

__global__ void __launch_bounds__(128) kr_cutoff_starts(const Cam* __restrict__ cams, const int* __restrict__ first_point /* [ncam+1] */,
                                                        int ncam, CutoffStart* __restrict__ out) {
  const int gp = blockIdx.x;                  // global border-pixel index
  const int s = threadIdx.x;                  // Gauss-Newton start (100 used)
  int k = 0;
  while (k + 1 < ncam && gp >= first_point[k + 1]) ++k;
  if (s >= 100) return;
  const Cam c = cams[k];
  float px, py; cutoff_border_pixel(c, gp - first_point[k], &px, &py);
  const float tx = c.fx_inv * px + c.cx_inv, ty = c.fy_inv * py + c.cy_inv;
  const int gy = s / 10, gx = s % 10;
  const float iy = ty + 1.5f * (gy - 0.5f * 10) / (0.5f * 10);
  const float ix = tx + 1.5f * (gx - 0.5f * 10) / (0.5f * 10);
  float ux = ix, uy = iy;
  int converged = 0;
  for (int i = 0; i < 100; ++i) {
    float qx, qy; cam_inner_distort(c, ux, uy, &qx, &qy);
    const float ex = qx - tx, ey = qy - ty;
    if (ex * ex + ey * ey < 1e-10f) { converged = 1; break; }
    float J[4]; cam_inner_deriv(c, ux, uy, J);
    const float a = J[0] * J[0] + J[2] * J[2], b = J[0] * J[1] + J[2] * J[3], cc = J[1] * J[0] + J[3] * J[2], dd = J[1] * J[1] + J[3] * J[3];
    const float invdet = 1.f / (a * dd - cc * b);
    const float i00 = dd * invdet, i10 = -cc * invdet, i01 = -b * invdet, i11 = a * invdet;
    const float m00 = i00 * J[0] + i01 * J[2], m01 = i00 * J[1] + i01 * J[3];
    const float m10 = i10 * J[0] + i11 * J[2], m11 = i10 * J[1] + i11 * J[3];
    ux -= m00 * ex + m01 * ey;
    uy -= m10 * ex + m11 * ey;
  }
  CutoffStart r; r.x = ux; r.y = uy; r.converged = converged;
  out[(size_t)gp * 100 + s] = r;
}

__global__ void __launch_bounds__(128) kr_cutoff_points(const CutoffStart* __restrict__ starts, int npoints, CutoffPoint* __restrict__ out) {
  const int gp = blockIdx.x * blockDim.x + threadIdx.x;
  if (gp >= npoints) return;
  const float kImproveThreshold = 0.99f;
  int converged = 0, second = 0;
  float best_radius = INFINITY, second_best_radius = INFINITY, bx = 0.f, by = 0.f, sx = INFINITY, sy = INFINITY;
  for (int s = 0; s < 100; ++s) {
    const CutoffStart r = starts[(size_t)gp * 100 + s];
    if (!r.converged) continue;
    const float radius = sqrtf(r.x * r.x + r.y * r.y);
    if (radius < kImproveThreshold * best_radius) {
      second_best_radius = best_radius; sx = bx; sy = by; second = converged;
      best_radius = radius; bx = r.x; by = r.y; converged = 1;
    } else if (radius > 1 / kImproveThreshold * best_radius && radius < kImproveThreshold * second_best_radius) {
      second_best_radius = radius; sx = r.x; sy = r.y; second = 1;
    }
  }
  CutoffPoint p; p.converged = converged; p.second = second; p.r2 = bx * bx + by * by; p.s2 = sx * sx + sy * sy;
  out[gp] = p;
}

// one block per camera: radius_cutoff_squared = min(max_p(r2) * 1.01, min_p(s2))
__global__ void __launch_bounds__(256) kr_cutoff_final(const CutoffPoint* __restrict__ pts, const int* __restrict__ first_point, float* __restrict__ out) {
  const int k = blockIdx.x;
  float mx = 0.f, mn = INFINITY;
  for (int i = first_point[k] + threadIdx.x; i < first_point[k + 1]; i += blockDim.x) {
    const CutoffPoint p = pts[i];
    if (p.converged) { mx = fmaxf(p.r2, mx); if (p.second) mn = fminf(p.s2, mn); }
  }
  __shared__ float smx[256], smn[256];
  smx[threadIdx.x] = mx; smn[threadIdx.x] = mn;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + o]); smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + o]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[k] = fminf(smx[0] * 1.01f, smn[0]);
}

// RadialBase::InitCutoff (camera_base_impl_radial.h:60-171): one scalar search from the farthest image corner; one thread per camera.
__global__ void __launch_bounds__(32) kr_cutoff_radial(const Cam* __restrict__ cams, int ncam, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncam) return;
  const Cam c = cams[k];
  const float kIncreaseFactor = 1.01f, kImproveThreshold = 0.99f;
  float test_image_radius = 0.f;
  for (int corner = 0; corner < 4; ++corner) {                    // (sic) width_ / height_, not the last pixel
    const float x = c.fx_inv * ((corner & 2) ? (float)c.w : 0.f) + c.cx_inv, y = c.fy_inv * ((corner & 1) ? (float)c.h : 0.f) + c.cy_inv;
    const float r = sqrtf(x * x + y * y);
    test_image_radius = corner == 0 ? r : fmaxf(test_image_radius, r);
  }
  bool converged = false, second_available = false;
  float best = INFINITY, second = INFINITY;
  for (int i = 0; i < 10; ++i) {
    const float init_radius = (float)((double)test_image_radius + (double)1.5f * ((double)i - 0.5 * 10) / (double)(0.5f * 10));   // (sic) double arithmetic
    float undistorted_r = init_radius, undistorted_r2 = init_radius * init_radius;
    bool ok = false;
    for (int it = 0; it < 100; ++it) {
      const float r_candidate = undistorted_r * radial_factor(c, undistorted_r2);
      const float delta_r = r_candidate - test_image_radius;
      if (delta_r * delta_r < 1e-10f) { ok = true; break; }
      const float step = delta_r / radial_deriv_r(c, undistorted_r2);
      undistorted_r -= step;
      undistorted_r2 = undistorted_r * undistorted_r;
    }
    if (ok) {
      const float result = undistorted_r;
      if (result < kImproveThreshold * best) { second = best; second_available = converged; best = result; converged = true; }
      else if (result > 1 / kImproveThreshold * best && result < kImproveThreshold * second) { second = result; second_available = true; }
    }
  }
  float cut = INFINITY;
  if (converged && best > 0) cut = (second_available && second > 0) ? fminf(best * best * kIncreaseFactor, second * second) : best * best * kIncreaseFactor;
  out[k] = cut;
}

// Undistort(distorted) — camera_base_impl.h:251-253 (IterativeUndistort started at the distorted point itself), camera_pinhole.h:65-68
// (identity), camera_base_impl_fisheye.h:80-91 (inner Undistort, then r -> tan r; tanf is the device's, within an ulp or two of glibc's).
__device__ __forceinline__ void cam_undistort(const Cam& c, float tx, float ty, float* ox, float* oy) {
  float ux = tx, uy = ty;
  if (c.dist == kDistFOV) {                                        // camera_fisheye_fov.h:78-88
    const float r = sqrtf(tx * tx + ty * ty);
    const float factor = (r < 1e-6f) ? 1.f : (r > c.image_radius) ? INFINITY : (tan_r1(r * c.d[0]) / (r * c.two_tan));
    ux = factor * tx; uy = factor * ty;
  } else if (c.dist != kDistNone) {
    for (int i = 0; i < 100; ++i) {
      float qx, qy; cam_inner_distort(c, ux, uy, &qx, &qy);
      const float ex = qx - tx, ey = qy - ty;
      if (ex * ex + ey * ey < 1e-10f) break;
      float J[4]; cam_inner_deriv(c, ux, uy, J);
      const float a = J[0] * J[0] + J[2] * J[2], b = J[0] * J[1] + J[2] * J[3], cc = J[1] * J[0] + J[3] * J[2], dd = J[1] * J[1] + J[3] * J[3];
      const float invdet = 1.f / (a * dd - cc * b);
      const float i00 = dd * invdet, i10 = -cc * invdet, i01 = -b * invdet, i11 = a * invdet;
      const float m00 = i00 * J[0] + i01 * J[2], m01 = i00 * J[1] + i01 * J[3];
      const float m10 = i10 * J[0] + i11 * J[2], m11 = i10 * J[1] + i11 * J[3];
      ux -= m00 * ex + m01 * ey;
      uy -= m10 * ex + m11 * ey;
    }
  }
  if (c.fisheye) {
    const float r = sqrtf(ux * ux + uy * uy);
    const float factor = (r < 1e-6f) ? 1.f : (r > 1.57079637f) ? INFINITY : tanf(r) / r;      // M_PI / 2.f rounds to 1.57079637f
    ux = factor * ux; uy = factor * uy;
  }
  *ox = ux; *oy = uy;
}
// the models whose ImageToNormalized(pixel_position) reads the undistortion lookup (all but Pinhole / SimplePinhole / FOV)
__host__ __device__ inline bool cam_has_lookup(const Cam& c) { return !cam_plain(c) && c.dist != kDistFOV; }
// InitializeUndistortionLookup (camera_base_impl.h:255-269): one thread per pixel, table[y * w + x] = Undistort(k_inv * (x, y)).
__global__ void __launch_bounds__(128) kr_undistortion_lookup(Cam cam, float2* __restrict__ table) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)cam.w * cam.h) return;
  const int x = (int)(i % cam.w), y = (int)(i / cam.w);
  float ux, uy; cam_undistort(cam, cam.fx_inv * x + cam.cx_inv, cam.fy_inv * y + cam.cy_inv, &ux, &uy);
  table[i] = make_float2(ux, uy);
}
// ImageToNormalized(pixel_position) (camera_base_impl.h:187-212; camera_pinhole.h:60-63 for pinhole). The reference reads one row past the
// table when the clamped y is exactly h - 1 (weight 0): read the last row instead.
__device__ __forceinline__ void cam_image_to_normalized(const Cam& c, const float2* __restrict__ table, float px, float py, float* ox, float* oy) {
  if (cam_plain(c)) { *ox = c.fx_inv * px + c.cx_inv; *oy = c.fy_inv * py + c.cy_inv; return; }
  if (c.dist == kDistFOV) { cam_undistort(c, c.fx_inv * px + c.cx_inv, c.fy_inv * py + c.cy_inv, ox, oy); return; }   // camera_fisheye_fov.h:66-76
  const float cxp = fmaxf(fminf(px, c.w - 1.001f), 0.f), cyp = fmaxf(fminf(py, c.h - 1.00f), 0.f);
  const int ix = (int)cxp, iy = (int)cyp;
  const float fx_ = cxp - (float)ix, fy_ = cyp - (float)iy;
  const int ix1 = min(ix + 1, c.w - 1), iy1 = min(iy + 1, c.h - 1);
  const float2 tl = __ldg(&table[(size_t)iy * c.w + ix]), tr = __ldg(&table[(size_t)iy * c.w + ix1]);
  const float2 bl = __ldg(&table[(size_t)iy1 * c.w + ix]), br = __ldg(&table[(size_t)iy1 * c.w + ix1]);
  *ox = (1 - fy_) * ((1 - fx_) * tl.x + fx_ * tr.x) + fy_ * ((1 - fx_) * bl.x + fx_ * br.x);
  *oy = (1 - fy_) * ((1 - fx_) * tl.y + fx_ * tr.y) + fy_ * ((1 - fx_) * bl.y + fx_ * br.y);
}

// op 1: NormalizedToImage (n x 2 -> n x 2), 2: ImageDerivativeByWorld (n x 3 -> n x 6), 3: ImageDerivativeByIntrinsics (n x 3 -> n x 2np)
__global__ void __launch_bounds__(128) kr_camera_eval(Cam cam, int op, const float* __restrict__ in, size_t n, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (op == 1) { cam_project(cam, in[2 * i], in[2 * i + 1], &out[2 * i], &out[2 * i + 1]); return; }
  const float x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
  if (op == 2) { float d[6]; cam_d_by_world(cam, x, y, z, d); for (int k = 0; k < 6; ++k) out[6 * i + k] = d[k]; return; }
  const int np = (cam.unique_focal ? 3 : 4) + cam.nd;
  if (cam_plain(cam)) {
    float a[4], b[4]; cam_d_by_intrinsics<4>(cam, x, y, z, a, b);
    for (int k = 0; k < np; ++k) { out[2 * np * i + k] = a[k]; out[2 * np * i + np + k] = b[k]; }
  } else {
    float a[12], b[12]; cam_d_by_intrinsics<12>(cam, x, y, z, a, b);
    for (int k = 0; k < np; ++k) { out[2 * np * i + k] = a[k]; out[2 * np * i + np + k] = b[k]; }
  }
}

}  // namespace b2
