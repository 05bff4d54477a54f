// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see orc_math.h header).
//
// CPU restatement of the reference's dense photometric image<->scan alignment (Path B), pinhole cameras, no rigs,
// depth residuals off (the reference default, parameters.h:54):
//   interpolation        /root/reference/src/opt/interpolate_bilinear.h:36-74, interpolate_trilinear.h:44-87
//   robust weighting     /root/reference/src/opt/robust_weighting.h:61-106
//   camera (pinhole)     /root/reference/src/camera/camera_base.cc:81-85, camera_base_impl.h:70-89,135-137,155-164,333-408,
//                        camera_pinhole.h:40-86
//   pyramids             /root/reference/src/opt/image.cc:106-154, intrinsics.cc:45-79, intrinsics.h:61-86, problem.cc:478-494
//   splat depth map      /root/reference/src/opt/occlusion_geometry.cc:404-464 (and the all-inf map :271-281)
//   visibility           /root/reference/src/opt/visibility_estimator.cc:61-91,140-168,199-256,258-295,366-532
//   intensity+Jacobians  /root/reference/src/opt/intrinsics_and_pose_optimizer.cc:933-1147
//   accumulate           /root/reference/src/opt/intrinsics_and_pose_optimizer.cc:624-930,1220-1296
//   LM step              /root/reference/src/opt/intrinsics_and_pose_optimizer.cc:48-259,385-473,475-558
//   cost                 /root/reference/src/opt/cost_calculator.cc:44-271, problem.cc:602-631
//   colour update        /root/reference/src/opt/color_optimizer.cc:40-123
//   outer loop           /root/reference/src/opt/optimizer.cc:49-190
// Pinned by the reference's own tests ported in tests/test_oracle_reg.py: test_interpolation.cc:39-185 (exact values),
// test_intrinsics_and_pose_optimizer.cc:101-336 (analytic Jacobians vs finite differences).
// Unpinned / defined here: iteration order over images (the reference iterates unordered_maps: here ascending image id, which
// also fixes the variable layout [intrinsics | image poses]); cv::resize INTER_AREA for even sizes = (a+b+c+d+2)>>2 (verified
// against Python cv2 4.13 in SURVEY.md §7) — odd parent sizes are rejected.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "orc_api.h"
#include "orc_camera.h"
#include "orc_math.h"
#include "orc_mesh.h"

namespace orc {

typedef Camera Pinhole;   // historical name: one pyramid level of a camera model (orc_camera.h)

// Row-major u8 image with pitch w, addressed like cv::Mat_<uint8_t>::operator()(row, col) in a release build: no bounds check, so a
// column index == w reads the first pixel of the next row. That matters: for odd-sized pyramid parents the reference's image level is
// one column narrower than its camera level (image.cc:116 truncates, camera_base_impl.h:72 rounds) while the bounds tests use the
// camera's size (visibility_estimator.cc:479), so the last interpolation column reads across the row end. Past the last pixel the
// reference reads unallocated memory; defined here (and in the CUDA path) as 0.
struct Img8 {
  int w = 0, h = 0; std::vector<uint8_t> d;
  uint8_t at(int y, int x) const { const size_t i = (size_t)y * w + x; return i < d.size() ? d[i] : (uint8_t)0; }
};
struct ImgF { int w = 0, h = 0; std::vector<float> d; float at(int y, int x) const { return d[(size_t)y * w + x]; } };

// interpolate_bilinear.h:36-74
static inline float bilinear(const Img8& im, float x, float y, int ix, int iy) {
  const float fx = x - ix, fx_inv = 1.f - fx, fy = y - iy, fy_inv = 1.f - fy;
  return fy_inv * (fx_inv * im.at(iy, ix) + fx * im.at(iy, ix + 1)) + fy * (fx_inv * im.at(iy + 1, ix) + fx * im.at(iy + 1, ix + 1));
}
static inline void bilinear_d(const Img8& im, float x, float y, int ix, int iy, float* v, float* dx, float* dy) {
  const uint8_t tl = im.at(iy, ix), tr = im.at(iy, ix + 1), bl = im.at(iy + 1, ix), br = im.at(iy + 1, ix + 1);
  const float fx = x - ix, fx_inv = 1.f - fx, fy = y - iy, fy_inv = 1.f - fy;
  const float top = fx_inv * tl + fx * tr, bottom = fx_inv * bl + fx * br;
  *v = fy_inv * top + fy * bottom;
  *dx = fy * (br - bl) + fy_inv * (tr - tl);
  *dy = bottom - top;
}
// interpolate_trilinear.h:44-87
static inline void trilinear(const Img8& i0, const Img8& i1, float x0, float y0, float z, float* v) {
  const float v0 = bilinear(i0, x0, y0, (int)x0, (int)y0);
  const float x1 = 2 * (x0 + 0.5f) - 0.5f, y1 = 2 * (y0 + 0.5f) - 0.5f;
  const float v1 = bilinear(i1, x1, y1, (int)x1, (int)y1);
  *v = (1 - z) * v0 + z * v1;
}
static inline void trilinear_d(const Img8& i0, const Img8& i1, float x0, float y0, float z, float* v, float* dx, float* dy, float* dz) {
  float v0, d0x, d0y, v1, d1x, d1y;
  bilinear_d(i0, x0, y0, (int)x0, (int)y0, &v0, &d0x, &d0y);
  const float x1 = 2 * (x0 + 0.5f) - 0.5f, y1 = 2 * (y0 + 0.5f) - 0.5f;
  bilinear_d(i1, x1, y1, (int)x1, (int)y1, &v1, &d1x, &d1y);
  *v = (1 - z) * v0 + z * v1;
  *dx = (1 - z) * d0x + z * 2 * d1x;
  *dy = (1 - z) * d0y + z * 2 * d1y;
  *dz = v1 - v0;
}

// robust_weighting.h:61-106
struct Robust {
  int type = 1; float p = 0.f;   // 0 none, 1 huber, 2 tukey
  float residual(float r) const {
    if (type == 1) { const float a = std::fabs(r); return a < p ? 0.5f * r * r : p * (a - 0.5f * p); }
    if (type == 2) {
      const float a = std::fabs(r);
      if (a < p) { const float q = r / p; const float t = 1.f - q * q; return (1 / 6.f) * p * p * (1 - t * t * t); }
      return (1 / 6.f) * p * p;
    }
    return 0.5f * r * r;
  }
  float weight(float r) const {
    if (type == 1) { const float a = std::fabs(r); return a < p ? 1.f : p / a; }
    if (type == 2) { const float a = std::fabs(r); if (a < p) { const float q = r / p; const float t = 1.f - q * q; return t * t; } return 0.f; }
    return 1.f;
  }
};

struct Observation { uint64_t point_index; float x, y, scale; };   // point_observation.h:97-116
static inline int smaller_scale(const Observation& o) { return (int)o.scale + 1; }
static inline int larger_scale(const Observation& o) { return (int)o.scale; }

struct Intrinsics {
  std::vector<Pinhole> models;   // index 0 = original resolution
  std::vector<Img8> camera_mask; // intrinsics.h:104: one mask pyramid per camera (empty = none), same semantics as the image masks
  int min_image_scale = -1;
  const Pinhole& model(int image_scale) const { return models[std::max(0, image_scale - min_image_scale)]; }
  int best_available(int image_scale) const { return std::min<int>(min_image_scale + (int)models.size() - 1, std::max<int>(min_image_scale, image_scale)); }
  void build_pyramid() { for (size_t i = 1; i < models.size(); ++i) models[i] = models[i - 1].scaled_half(); }
};

struct Image {
  int intrinsics_id = 0;
  int rig_images_id = -1, rig_camera_index = 0;   // image.h rig_images_id; index of this image in its RigImages::image_ids
  SE3f image_T_global;
  std::vector<Img8> image, mask;   // pyramids (mask levels may be empty)
  ImgF given_depth; bool has_given_depth = false;
};

struct ScalePoints {
  std::vector<float> xyz; float radius = 0;
  std::vector<uint64_t> nbr;              // n * K
  std::vector<float> fixed_desc, var_desc; std::vector<int> obs_count;
  size_t n() const { return xyz.size() / 3; }
};

// rig.h:40-73, rig_images.h:38-64. image_T_rig[0] stays identity; the other extrinsics are optimized.
struct Rig { std::vector<SE3f> image_T_rig; };
struct RigImages { int rig_id = 0; std::vector<int> image_ids; };

struct State { std::vector<Intrinsics> intr; std::vector<Image> images; std::vector<Rig> rigs; };

}  // namespace orc

using namespace orc;

struct orc_reg {
  orc_reg_params prm;
  Robust robust;
  State st;
  std::vector<ScalePoints> pts;
  std::vector<RigImages> rig_images;
  std::vector<float> splat_xyz; bool has_splats = false;
  OccMesh mesh; bool has_mesh = false;
  int image_scale_count = 0, current_image_scale = 0;
  // observations: [image][point_scale]
  std::vector<std::vector<std::vector<Observation>>> obs;
  std::vector<std::vector<std::vector<uint8_t>>> nbr_obs;
  double last_sums[6] = {0, 0, 0, 0, 0, 0};
};

namespace orc {

static int K(const orc_reg* h) { return h->prm.point_neighbor_count; }

// ---- pyramids ----
// cv::resize(src, dst, dsize = (int(0.5 cols), int(0.5 rows)), 0.5, 0.5, INTER_AREA) as image.cc:115-118 calls it. OpenCV (imgproc
// resize.cpp; not vendored: restated from its published algorithm, pinned against cv2 in tests/test_oracle_reg.py):
//   scale = 1 / ((double)dsize / ssize) per axis (dsize wins over fx, fy); both scales == 2 -> integer 2x2 mean (a+b+c+d+2)>>2;
//   otherwise the general area filter: per axis a table of (dst, src, alpha) with fractional coverage at the cell borders
//   (computeResizeAreaTab), float accumulation in table order, beta * row sums, saturate_cast<uchar> = round half to even.
struct AreaTap { int di, si; float alpha; };
static std::vector<AreaTap> resize_area_tab(int ssize, int dsize, double scale) {
  std::vector<AreaTap> tab;
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale, cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1); sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) tab.push_back({dx, sx1 - 1, (float)((sx1 - fsx1) / cell)});
    for (int sx = sx1; sx < sx2; ++sx) tab.push_back({dx, sx, float(1.0 / cell)});
    if (fsx2 - sx2 > 1e-3) tab.push_back({dx, sx2, (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)});
  }
  return tab;
}
static void resize_area_half(const Img8& s, Img8& d) {
  d.w = (int)(0.5 * s.w); d.h = (int)(0.5 * s.h);
  d.d.assign((size_t)d.w * d.h, 0);
  if (d.w == 0 || d.h == 0) return;
  const double scale_x = 1. / ((double)d.w / s.w), scale_y = 1. / ((double)d.h / s.h);
  if (scale_x == 2.0 && scale_y == 2.0) {
    for (int y = 0; y < d.h; ++y) for (int x = 0; x < d.w; ++x)
      d.d[(size_t)y * d.w + x] = (uint8_t)((s.at(2 * y, 2 * x) + s.at(2 * y, 2 * x + 1) + s.at(2 * y + 1, 2 * x) + s.at(2 * y + 1, 2 * x + 1) + 2) >> 2);
    return;
  }
  const std::vector<AreaTap> xt = resize_area_tab(s.w, d.w, scale_x), yt = resize_area_tab(s.h, d.h, scale_y);
  std::vector<float> buf(d.w), sum(d.w, 0.f);
  auto flush = [&](int dy) {
    for (int x = 0; x < d.w; ++x) { const long r = lrintf(sum[x]); d.d[(size_t)dy * d.w + x] = (uint8_t)std::min(255l, std::max(0l, r)); }
  };
  int prev = -1;
  for (const AreaTap& ty : yt) {
    std::fill(buf.begin(), buf.end(), 0.f);
    for (const AreaTap& tx : xt) buf[tx.di] += s.at(ty.si, tx.si) * tx.alpha;
    if (ty.di != prev) {
      if (prev >= 0) flush(prev);
      for (int x = 0; x < d.w; ++x) sum[x] = ty.alpha * buf[x];
      prev = ty.di;
    } else {
      for (int x = 0; x < d.w; ++x) sum[x] += ty.alpha * buf[x];
    }
  }
  if (prev >= 0) flush(prev);
}
static bool build_image_pyramid(std::vector<Img8>& pyr) {   // image.cc:106-131
  for (size_t i = 1; i < pyr.size(); ++i) resize_area_half(pyr[i - 1], pyr[i]);
  return true;
}
static void build_mask_pyramid(std::vector<Img8>& pyr) {    // image.cc:133-154
  for (size_t i = 1; i < pyr.size(); ++i) {
    const Img8& s = pyr[i - 1];
    Img8& d = pyr[i];
    d.h = (int)(0.5 * s.h); d.w = (int)(0.5 * s.w);
    d.d.resize((size_t)d.w * d.h);
    for (int y = 0; y < d.h; ++y) for (int x = 0; x < d.w; ++x)
      d.d[(size_t)y * d.w + x] = s.at(2 * y, 2 * x) | s.at(2 * y, 2 * x + 1) | s.at(2 * y + 1, 2 * x) | s.at(2 * y + 1, 2 * x + 1);
  }
}

// ---- occlusion depth map (occlusion_geometry.cc:185-282: splats, given, or all-inf) ----
static ImgF render_depth(const orc_reg* h, const Intrinsics& intr, const Image& im, int image_scale) {
  const Pinhole& cam = intr.model(image_scale);
  ImgF out; out.w = cam.w; out.h = cam.h; out.d.assign((size_t)cam.w * cam.h, std::numeric_limits<float>::infinity());
  if (im.has_given_depth) return im.given_depth;
  float R[9]; quat_to_matrix(im.image_T_global.q, R);
  if (!h->has_splats && h->has_mesh) {
    // mesh path (occlusion_geometry.cc:211-270): depth pass + MaskOutOcclusionBoundaries; see orc_mesh.h for the rasteriser definition
    const RasterCam& rc = cam;
    std::vector<float> raw;
    raster_mesh(h->mesh, rc, R, im.image_T_global.t, h->prm.min_occlusion_depth, h->prm.max_occlusion_depth, &raw);
    // image position = global_T_image.translation() = inv(q) * (-t)   (sophus se3.hpp:208-211)
    const Quat<float> qi{-im.image_T_global.q.x, -im.image_T_global.q.y, -im.image_T_global.q.z, im.image_T_global.q.w};
    const V3f pos = quat_rotate(qi, V3f{im.image_T_global.t.x * -1.f, im.image_T_global.t.y * -1.f, im.image_T_global.t.z * -1.f});
    if (h->prm.mask_occlusion_boundaries) mask_boundaries(h->mesh, rc, R, im.image_T_global.t, pos, h->prm.splat_radius, raw, &out.d);
    else out.d = raw;
    return out;
  }
  if (!h->has_splats) return out;
  const M3f& M = *reinterpret_cast<const M3f*>(R);
  const float max_splat_radius = 10;
  for (size_t i = 0; i < h->splat_xyz.size() / 3; ++i) {
    const V3f v = mul(M, V3f{h->splat_xyz[3 * i], h->splat_xyz[3 * i + 1], h->splat_xyz[3 * i + 2]});
    const V3f pp{v.x + im.image_T_global.t.x, v.y + im.image_T_global.t.y, v.z + im.image_T_global.t.z};
    if (!(pp.z > 0.f)) continue;
    float px, py; cam.project(pp.x / pp.z, pp.y / pp.z, &px, &py);
    float d[6]; cam.d_by_world(pp, d);
    float rx = std::sqrt(sum3(d[0] * d[0], d[1] * d[1], d[2] * d[2])) * h->prm.splat_radius;
    float ry = std::sqrt(sum3(d[3] * d[3], d[4] * d[4], d[5] * d[5])) * h->prm.splat_radius;
    rx = std::min(rx, max_splat_radius); ry = std::min(ry, max_splat_radius);
    const int ix = f2i(px + 0.5f), iy = f2i(py + 0.5f);
    if (ix == std::numeric_limits<int>::min() || iy == std::numeric_limits<int>::min()) continue;   // beyond the cut-off radius
    const int min_x = std::max(0, int(ix - rx + 0.5)), min_y = std::max(0, int(iy - ry + 0.5));
    const int end_x = std::min(cam.w, int(ix + rx + 1.5)), end_y = std::min(cam.h, int(iy + ry + 1.5));
    if (min_y < end_y && min_x < end_x)
      for (int y = min_y; y < end_y; ++y) for (int x = min_x; x < end_x; ++x)
        if (out.d[(size_t)y * out.w + x] > pp.z) out.d[(size_t)y * out.w + x] = pp.z;
  }
  return out;
}

// ---- CreateObservationIfScaleFits (visibility_estimator.cc:405-532) ----
static void create_observation_if_scale_fits(const orc_reg* h, const Intrinsics& intr, const Image& im, const Pinhole& cam, int image_scale,
                                             uint64_t point_index, const V3f& pp, float point_radius, float ixx, float ixy, int border,
                                             bool check_masks, std::vector<Observation>* out) {
  const V3f ppr{pp.x + point_radius, pp.y + 0, pp.z + 0};
  float rx, ry; cam.project(ppr.x / ppr.z, ppr.y / ppr.z, &rx, &ry);
  const float dx = rx - ixx, dy = ry - ixy;
  const float radius_pixels = std::sqrt(dx * dx + dy * dy);
  const float observation_scale = image_scale + log2((double)(2 * radius_pixels));   // ::log2(double); float-vs-double overload UNPINNED (differs by <= 1 ulp)
  if (!std::isfinite(observation_scale)) return;   // offset point beyond the cut-off radius: the reference's int cast is UB there
  if (observation_scale >= std::max(intr.min_image_scale, h->current_image_scale) &&
      static_cast<int>(observation_scale) < h->image_scale_count - 1) {
    const int small = static_cast<int>(observation_scale) + 1;
    const Pinhole& ic = intr.model(small);
    const float nx = cam.fx_inv * ixx + cam.cx_inv, ny = cam.fy_inv * ixy + cam.cy_inv;
    const float jx = ic.fx * nx + ic.cx, jy = ic.fy * ny + ic.cy;
    const int ix = jx + 0.5f, iy = jy + 0.5f;
    if (jx + 0.5f >= border && jy + 0.5f >= border && ix >= border && iy >= border && ix < ic.w - border && iy < ic.h - border) {
      if (check_masks) {
        const int level = small - intr.min_image_scale;
        if (level < (int)im.mask.size() && !im.mask[level].d.empty() && im.mask[level].at(iy, ix) != 0) return;
        if (level < (int)intr.camera_mask.size() && !intr.camera_mask[level].d.empty() && intr.camera_mask[level].at(iy, ix) != 0) return;   // :492-501
        if (im.image[level].at(iy, ix) > h->prm.maximum_valid_intensity) return;
      }
      out->push_back(Observation{point_index, jx, jy, observation_scale});
    }
  }
}

static void rigid_pp(const float R[9], const V3f& t, const float* p, V3f* out) {
  const M3f& M = *reinterpret_cast<const M3f*>(R);
  const V3f v = mul(M, V3f{p[0], p[1], p[2]});
  *out = {v.x + t.x, v.y + t.y, v.z + t.z};
}

// AppendObservationsForImage (visibility_estimator.cc:61-91 + 258-295) or, with `lists`, the indexed variant (:140-168, 366-403).
static void append_observations(const orc_reg* h, const Image& im, const Intrinsics& intr, int border,
                                const std::vector<std::vector<uint64_t>>* lists, std::vector<std::vector<Observation>>* out) {
  const int best = intr.best_available(std::max(h->prm.min_occlusion_check_image_scale, h->current_image_scale));
  const Pinhole& cam = intr.model(best);
  ImgF depth;
  if (!lists) depth = render_depth(h, intr, im, best);
  float R[9]; quat_to_matrix(im.image_T_global.q, R);
  out->assign(h->pts.size(), {});
  bool had_many = false;
  for (int ps = (int)h->pts.size() - 1; ps >= 0; --ps) {
    const ScalePoints& P = h->pts[ps];
    std::vector<Observation>& o = (*out)[ps];
    const size_t count = lists ? (*lists)[ps].size() : P.n();
    for (size_t i = 0; i < count; ++i) {
      const uint64_t pi = lists ? (*lists)[ps][i] : i;
      V3f pp; rigid_pp(R, im.image_T_global.t, &P.xyz[3 * pi], &pp);
      if (pp.z > 0.f) {
        float ixx, ixy; cam.project(pp.x / pp.z, pp.y / pp.z, &ixx, &ixy);
        const int ix = f2i(ixx + 0.5f), iy = f2i(ixy + 0.5f);
        if (ix >= 0 && iy >= 0 && ix < cam.w && iy < cam.h &&
            (lists || depth.d[(size_t)iy * depth.w + ix] + h->prm.occlusion_depth_threshold >= pp.z))
          create_observation_if_scale_fits(h, intr, im, cam, best, pi, pp, P.radius, ixx, ixy, border, !lists, &o);
      }
    }
    if (o.size() > 100) had_many = true;                 // kManyObservationsCount (visibility_estimator.cc:44,75-90)
    else if (o.size() == 0 && had_many) break;
  }
}

// DetermineIfAllNeighborsAreObserved (visibility_estimator.cc:199-256)
static void neighbors_observed(const orc_reg* h, int ps, const std::vector<Observation>& o, std::vector<uint8_t>* out) {
  const ScalePoints& P = h->pts[ps];
  std::vector<bool> seen(P.n(), false);
  for (const Observation& ob : o) seen[ob.point_index] = true;
  out->resize(o.size());
  for (size_t i = 0; i < o.size(); ++i) {
    bool all = true;
    for (int k = 0; k < K(h); ++k) if (!seen[P.nbr[o[i].point_index * K(h) + k]]) { all = false; break; }
    (*out)[i] = all;
  }
}

// What a dependent rig image (camera index > 0) adds to ComputePointIntensityAndJacobians (intrinsics_and_pose_optimizer.cc:651-670).
struct RigCtx { bool dependent = false; SE3f image_T_rig, rig_T_global; };

// ComputePointIntensityAndJacobians (intrinsics_and_pose_optimizer.cc:933-1147), no depth residual. For a dependent rig image jP is
// the derivative by the REFERENCE image's pose and jR (6) the derivative by the rig extrinsics of this camera.
static void point_intensity_and_jacobians(const orc_reg* h, const Intrinsics& intr, const Image& im, const float R[9], float point_radius,
                                          const float* point, const Observation& ob, float* intensity, float* jK /* np */, float jP[6],
                                          const RigCtx* rig = nullptr, float* jR = nullptr) {
  const Pinhole& cam = intr.model(0);
  const int np = cam.np();
  V3f tp; rigid_pp(R, im.image_T_global.t, point, &tp);
  float ji[3];
  const Img8& i0 = im.image[smaller_scale(ob) - intr.min_image_scale];
  const Img8& i1 = im.image[larger_scale(ob) - intr.min_image_scale];
  trilinear_d(i0, i1, ob.x, ob.y, 1 - (ob.scale - static_cast<int>(ob.scale)), intensity, &ji[0], &ji[1], &ji[2]);
  ji[2] = -1 * ji[2];
  const float scale_factor = (float)pow(2, intr.min_image_scale - smaller_scale(ob));
  const float inv_scale_factor = 1.f / scale_factor;
  ji[0] *= scale_factor; ji[1] *= scale_factor;
  const float mx = inv_scale_factor * (ob.x + 0.5f) - 0.5f, my = inv_scale_factor * (ob.y + 0.5f) - 0.5f;
  const V3f tpo{tp.x + point_radius, tp.y, tp.z};
  float ox, oy; cam.project(tpo.x / tpo.z, tpo.y / tpo.z, &ox, &oy);
  const float rdx = ox - mx, rdy = oy - my;
  const float denom = std::max(1e-6f, 0.693147180559945f * (rdx * rdx + rdy * rdy));
  float jpi[36], jpoi[24];   // 3 x np row-major, 2 x np
  cam.d_by_intrinsics(tp, jpi); cam.d_by_intrinsics(tpo, jpoi);
  for (int i = 0; i < np; ++i) jpi[2 * np + i] = ((jpoi[i] - jpi[i]) * rdx + (jpoi[np + i] - jpi[np + i]) * rdy) / denom;
  for (int i = 0; i < np; ++i) jK[i] = sum3(ji[0] * jpi[i], ji[1] * jpi[np + i], ji[2] * jpi[2 * np + i]);
  float jpp[9], jpop[6];    // 3x3 row-major, 2x3
  cam.d_by_world(tp, jpp); cam.d_by_world(tpo, jpop);
  for (int i = 0; i < 3; ++i) jpp[6 + i] = ((jpop[i] - jpp[i]) * rdx + (jpop[3 + i] - jpp[3 + i]) * rdy) / denom;
  float a[3];
  for (int i = 0; i < 3; ++i) a[i] = sum3(ji[0] * jpp[i], ji[1] * jpp[3 + i], ji[2] * jpp[6 + i]);
  // [I | -[p]x] rows: (1,0,0,0,z,-y), (0,1,0,-z,0,x), (0,0,1,y,-x,0)
  const float C[18] = {1, 0, 0, 0, tp.z, -1 * tp.y, 0, 1, 0, -1 * tp.z, 0, tp.x, 0, 0, 1, tp.y, -1 * tp.x, 0};
  if (rig && rig->dependent) {
    // :1107-1143: j_pose = ((ji * jpp) * image_T_rig.rotationMatrix()) * [I | -[rig_point]x], rig_point = rig_T_global * point (Sophus
    // SE3 * point: quaternion rotation + translation); j_rig_extrinsics = (ji * jpp) * [I | -[transformed_point]x]
    const V3f rp0 = quat_rotate(rig->rig_T_global.q, V3f{point[0], point[1], point[2]});
    const V3f rp{rp0.x + rig->rig_T_global.t.x, rp0.y + rig->rig_T_global.t.y, rp0.z + rig->rig_T_global.t.z};
    float Rr[9]; quat_to_matrix(rig->image_T_rig.q, Rr);
    float ar[3];
    for (int i = 0; i < 3; ++i) ar[i] = sum3(a[0] * Rr[i], a[1] * Rr[3 + i], a[2] * Rr[6 + i]);
    const float Cr[18] = {1, 0, 0, 0, rp.z, -1 * rp.y, 0, 1, 0, -1 * rp.z, 0, rp.x, 0, 0, 1, rp.y, -1 * rp.x, 0};
    for (int c = 0; c < 6; ++c) jP[c] = sum3(ar[0] * Cr[c], ar[1] * Cr[6 + c], ar[2] * Cr[12 + c]);
    for (int c = 0; c < 6; ++c) jR[c] = sum3(a[0] * C[c], a[1] * C[6 + c], a[2] * C[12 + c]);
    return;
  }
  for (int c = 0; c < 6; ++c) jP[c] = sum3(a[0] * C[c], a[1] * C[6 + c], a[2] * C[12 + c]);
}

struct Sums { double fixed_sum = 0, var_sum = 0; uint64_t nf = 0, nv = 0; };

// Problem::ComputeCost (problem.cc:602-631), depth weight 0.
static double compute_cost(const orc_reg* h, const Sums& s) {
  const bool uf = h->prm.fixed_residuals_weight > 0, uv = h->prm.variable_residuals_weight > 0;
  double r = 0;
  if (uf && s.nf > 0) r += h->prm.fixed_residuals_weight * s.fixed_sum / s.nf;
  if (uv && s.nv > 0) r += h->prm.variable_residuals_weight * s.var_sum / s.nv;
  if ((!uf && !uv) || (s.nf == 0 && s.nv == 0)) r = std::numeric_limits<float>::infinity();
  return r;
}

// cost_calculator.cc:102-271
static void accumulate_residuals(const orc_reg* h, const Intrinsics& intr, const Image& im, int ps, const std::vector<Observation>& o,
                                 const std::vector<uint8_t>& nb, Sums* s) {
  const ScalePoints& P = h->pts[ps];
  std::vector<float> inten(P.n(), -1.f);
  for (const Observation& ob : o)
    trilinear(im.image[smaller_scale(ob) - intr.min_image_scale], im.image[larger_scale(ob) - intr.min_image_scale], ob.x, ob.y,
              1 - (ob.scale - static_cast<int>(ob.scale)), &inten[ob.point_index]);
  auto residual = [&](uint64_t pi, const std::vector<float>& desc) {
    float pr = 0.f;
    for (int k = 0; k < K(h); ++k) {
      const float c = (inten[P.nbr[pi * K(h) + k]] - inten[pi]) - desc[pi * K(h) + k];
      pr += c * c;
    }
    return h->robust.residual(sqrtf(pr));
  };
  for (size_t i = 0; i < o.size(); ++i) {
    if (!nb[i]) continue;
    if (h->prm.fixed_residuals_weight > 0) { s->fixed_sum += residual(o[i].point_index, P.fixed_desc); ++s->nf; }
    if (h->prm.variable_residuals_weight > 0 && P.obs_count[o[i].point_index] >= 2) { s->var_sum += residual(o[i].point_index, P.var_desc); ++s->nv; }
  }
}

// AccumulateOnHAndB (intrinsics_and_pose_optimizer.cc:1220-1296): fp32 products cast to double, upper triangles + full cross block.
static void accumulate_on_H_b(float w, float res, int iv, int pv, int np, const float* jK, const float jP[6], std::vector<double>* H, std::vector<double>* b, int nv,
                              int rv = -1, const float* jR = nullptr) {
  if (w == 0) return;
  auto Hh = [&](int r, int c) -> double& { return (*H)[(size_t)c * nv + r]; };
  for (int c = 0; c < np; ++c) for (int r = 0; r <= c; ++r) Hh(iv + r, iv + c) += (double)(w * jK[r] * jK[c]);
  for (int r = 0; r < np; ++r) for (int c = 0; c < 6; ++c) Hh(iv + r, pv + c) += (double)(w * jK[r] * jP[c]);
  for (int c = 0; c < 6; ++c) for (int r = 0; r <= c; ++r) Hh(pv + r, pv + c) += (double)(w * jP[r] * jP[c]);
  if (rv >= 0) {   // dependent rig image (:1262-1283): top middle, middle (upper), middle right
    for (int r = 0; r < np; ++r) for (int c = 0; c < 6; ++c) Hh(iv + r, rv + c) += (double)(w * jK[r] * jR[c]);
    for (int c = 0; c < 6; ++c) for (int r = 0; r <= c; ++r) Hh(rv + r, rv + c) += (double)(w * jR[r] * jR[c]);
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Hh(rv + r, pv + c) += (double)(w * jR[r] * jP[c]);
  }
  const float wr = w * res;
  for (int i = 0; i < np; ++i) (*b)[iv + i] += (double)(wr * jK[i]);
  for (int i = 0; i < 6; ++i) (*b)[pv + i] += (double)(wr * jP[i]);
  if (rv >= 0) for (int i = 0; i < 6; ++i) (*b)[rv + i] += (double)(wr * jR[i]);
}

// AccumulateHAndBAndResidualsForObservations (intrinsics_and_pose_optimizer.cc:624-837) + ...ForColorObservation (:840-930)
static void accumulate_H_b(const orc_reg* h, const Intrinsics& intr, const Image& im, int ps, const std::vector<Observation>& o,
                           const std::vector<uint8_t>& nb, int iv, int pv, Sums* s, std::vector<double>* H, std::vector<double>* b, int nv,
                           const RigCtx* rig = nullptr, int rv = -1) {
  const ScalePoints& P = h->pts[ps];
  float R[9]; quat_to_matrix(im.image_T_global.q, R);
  const int np = intr.model(0).np();
  const bool dep = rig && rig->dependent;
  std::vector<float> inten(o.size()), jK((size_t)np * o.size()), jP(6 * o.size()), jR(dep ? 6 * o.size() : 0);
  std::vector<int64_t> jac_of_point(P.n(), -1);
  for (size_t i = 0; i < o.size(); ++i) {
    point_intensity_and_jacobians(h, intr, im, R, P.radius, &P.xyz[3 * o[i].point_index], o[i], &inten[i], &jK[(size_t)np * i], &jP[6 * i],
                                  rig, dep ? &jR[6 * i] : nullptr);
    jac_of_point[o[i].point_index] = (int64_t)i;
  }
  if (h->prm.fixed_residuals_weight == 0 && h->prm.variable_residuals_weight == 0) return;
  std::vector<float> comp(K(h));
  auto color_obs = [&](uint64_t pi, size_t oj, const std::vector<float>& desc, float static_w, double* sum, uint64_t* cnt) {
    float pr = 0.f;
    for (int k = 0; k < K(h); ++k) {
      const size_t nj = (size_t)jac_of_point[P.nbr[pi * K(h) + k]];
      const float c = (inten[nj] - inten[oj]) - desc[pi * K(h) + k];
      comp[k] = c; pr += c * c;
    }
    pr = sqrtf(pr);
    ++(*cnt);
    (*sum) += h->robust.residual(pr);
    const float w = static_w * h->robust.weight(pr);
    if (w != 0) {
      for (int k = 0; k < K(h); ++k) {
        const size_t nj = (size_t)jac_of_point[P.nbr[pi * K(h) + k]];
        float dK[12], dP[6], dR[6];
        for (int i = 0; i < np; ++i) dK[i] = jK[(size_t)np * nj + i] - jK[(size_t)np * oj + i];
        for (int i = 0; i < 6; ++i) dP[i] = jP[6 * nj + i] - jP[6 * oj + i];
        if (dep) for (int i = 0; i < 6; ++i) dR[i] = jR[6 * nj + i] - jR[6 * oj + i];
        accumulate_on_H_b(w, comp[k], iv, pv, np, dK, dP, H, b, nv, dep ? rv : -1, dR);
      }
    }
  };
  for (size_t i = 0; i < o.size(); ++i) {
    if (!nb[i]) continue;
    const uint64_t pi = o[i].point_index;
    if (h->prm.fixed_residuals_weight > 0) color_obs(pi, i, P.fixed_desc, h->prm.fixed_residuals_weight, &s->fixed_sum, &s->nf);
    if (h->prm.variable_residuals_weight > 0 && P.obs_count[pi] >= 2) color_obs(pi, i, P.var_desc, h->prm.variable_residuals_weight, &s->var_sum, &s->nv);
  }
}

// Variable layout (CountAndIndexVariables, intrinsics_and_pose_optimizer.cc:442-473): [intrinsics 0 (np0) | intrinsics 1 | ... |
// rig 0 extrinsics (6 per camera after the first) | rig 1 | ... | poses (6 each) of the images that own one: non-rig images and rig
// reference images, ascending image id]. A dependent rig image uses its reference image's pose block.
static int intr_var(const orc_reg* h, int id) { int v = 0; for (int i = 0; i < id; ++i) v += h->st.intr[i].models[0].np(); return v; }
static int rig_var(const orc_reg* h, int rig_id) {
  int v = intr_var(h, (int)h->st.intr.size());
  for (int r = 0; r < rig_id; ++r) v += 6 * ((int)h->st.rigs[r].image_T_rig.size() - 1);
  return v;
}
static bool is_dependent(const orc_reg* h, int image_id) { const Image& im = h->st.images[image_id]; return im.rig_images_id >= 0 && im.rig_camera_index > 0; }
static int ref_image(const orc_reg* h, int image_id) {
  const Image& im = h->st.images[image_id];
  return is_dependent(h, image_id) ? h->rig_images[im.rig_images_id].image_ids[0] : image_id;
}
static int pose_var(const orc_reg* h, int image_id) {
  const int owner = image_id < (int)h->st.images.size() ? ref_image(h, image_id) : image_id;
  int v = rig_var(h, (int)h->st.rigs.size());
  for (int i = 0; i < owner; ++i) if (!is_dependent(h, i)) v += 6;
  return v;
}
static int num_variables(const orc_reg* h) { return pose_var(h, (int)h->st.images.size()); }
static RigCtx rig_ctx(const orc_reg* h, const State& st, int image_id, int* rv) {
  RigCtx c; *rv = -1;
  if (!is_dependent(h, image_id)) return c;
  const Image& im = st.images[image_id];
  const RigImages& ri = h->rig_images[im.rig_images_id];
  c.dependent = true;
  c.image_T_rig = st.rigs[ri.rig_id].image_T_rig[im.rig_camera_index];
  c.rig_T_global = st.images[ri.image_ids[0]].image_T_global;
  *rv = rig_var(h, ri.rig_id) + 6 * (im.rig_camera_index - 1);
  return c;
}

static void accumulate_all(const orc_reg* h, std::vector<double>* H, std::vector<double>* b, Sums* s) {
  const int nv = num_variables(h);
  H->assign((size_t)nv * nv, 0.0); b->assign(nv, 0.0);
  for (size_t im = 0; im < h->st.images.size(); ++im) {
    int rv; const RigCtx rc = rig_ctx(h, h->st, (int)im, &rv);
    for (size_t ps = 0; ps < h->pts.size(); ++ps)
      accumulate_H_b(h, h->st.intr[h->st.images[im].intrinsics_id], h->st.images[im], (int)ps, h->obs[im][ps], h->nbr_obs[im][ps],
                     intr_var(h, h->st.images[im].intrinsics_id), pose_var(h, (int)im), s, H, b, nv, &rc, rv);
  }
}

// CreateDeltaState (intrinsics_and_pose_optimizer.cc:475-558): params += delta (fp32 += double), pose <- exp(delta).cast<float>() * pose.
static State delta_state(const orc_reg* h, const std::vector<double>& delta) {
  State n = h->st;
  for (size_t i = 0; i < n.intr.size(); ++i) {
    Pinhole& m = n.intr[i].models[0];
    float p[12]; m.get_params(p);
    for (int k = 0; k < m.np(); ++k) p[k] += delta[intr_var(h, (int)i) + k];
    m.set(m.type, m.w, m.h, p);
    n.intr[i].build_pyramid();
  }
  for (size_t r = 0; r < n.rigs.size(); ++r)           // Rig::Update (rig.cc:9-23)
    for (size_t c = 1; c < n.rigs[r].image_T_rig.size(); ++c) {
      double d[6]; for (int k = 0; k < 6; ++k) d[k] = delta[rig_var(h, (int)r) + 6 * ((int)c - 1) + k];
      n.rigs[r].image_T_rig[c] = se3_mul(se3d_exp_cast_float(d), h->st.rigs[r].image_T_rig[c]);
    }
  for (size_t i = 0; i < n.images.size(); ++i) {        // pose owners first (:517-534, 556-563)
    if (is_dependent(h, (int)i)) continue;
    double d[6]; for (int k = 0; k < 6; ++k) d[k] = delta[pose_var(h, (int)i) + k];
    n.images[i].image_T_global = se3_mul(se3d_exp_cast_float(d), h->st.images[i].image_T_global);
  }
  for (size_t i = 0; i < n.images.size(); ++i) {        // dependent rig images from the UPDATED rig pose and extrinsics (:546-555)
    if (!is_dependent(h, (int)i)) continue;
    const RigImages& ri = h->rig_images[n.images[i].rig_images_id];
    n.images[i].image_T_global = se3_mul(n.rigs[ri.rig_id].image_T_rig[n.images[i].rig_camera_index], n.images[ri.image_ids[0]].image_T_global);
  }
  return n;
}

// ComputeResidualForState (intrinsics_and_pose_optimizer.cc:385-440) with frozen visibility lists.
static double residual_for_state(orc_reg* h, const State& s, const std::vector<std::vector<std::vector<uint64_t>>>& lists) {
  Sums sums;
  const State saved = h->st;
  h->st = s;   // descriptors / counts / points are not part of the state
  for (size_t im = 0; im < s.images.size(); ++im) {
    std::vector<std::vector<Observation>> o;
    append_observations(h, s.images[im], s.intr[s.images[im].intrinsics_id], 1, &lists[im], &o);
    for (size_t ps = 0; ps < h->pts.size(); ++ps) {
      std::vector<uint8_t> nb; neighbors_observed(h, (int)ps, o[ps], &nb);
      accumulate_residuals(h, s.intr[s.images[im].intrinsics_id], s.images[im], (int)ps, o[ps], nb, &sums);
    }
  }
  h->st = saved;
  return compute_cost(h, sums);
}

}  // namespace orc

extern "C" {

void orc_reg_default_params(orc_reg_params* p) {
  p->point_neighbor_count = 5; p->fixed_residuals_weight = 1.f; p->variable_residuals_weight = 1.f;
  p->robust_weighting_type = 1; p->robust_weighting_parameter = (float)(30 * sqrt(5) / sqrt(2));
  p->maximum_valid_intensity = 252; p->occlusion_depth_threshold = 0.01f; p->min_occlusion_check_image_scale = 0;
  p->max_initial_image_area_in_pixels = 200 * 160; p->splat_radius = 0.03f; p->image_scale_count_override = 0;
  p->min_occlusion_depth = 0.05f; p->max_occlusion_depth = 100.f; p->mask_occlusion_boundaries = 1;
}
orc_reg* orc_reg_create(const orc_reg_params* p) {
  orc_reg* h = new orc_reg(); h->prm = *p; h->robust.type = p->robust_weighting_type; h->robust.p = p->robust_weighting_parameter; return h;
}
void orc_reg_destroy(orc_reg* h) { delete h; }

int orc_reg_add_intrinsics(orc_reg* h, int w, int hh, const float p[4]) { return orc_reg_add_intrinsics_model(h, kCamPinhole, w, hh, p); }
int orc_reg_add_intrinsics_model(orc_reg* h, int type, int w, int hh, const float* p) {
  if (!Camera::known(type)) return -1;
  Intrinsics in; in.models.resize(1); in.models[0].set(type, w, hh, p);
  h->st.intr.push_back(in); return (int)h->st.intr.size() - 1;
}
int orc_reg_add_image(orc_reg* h, int intr_id, const uint8_t* gray, const uint8_t* mask, const float T[7]) {
  Image im; im.intrinsics_id = intr_id;
  im.image_T_global.q = {T[0], T[1], T[2], T[3]}; im.image_T_global.t = {T[4], T[5], T[6]};
  const Pinhole& c = h->st.intr[intr_id].models[0];
  im.image.resize(1); im.image[0].w = c.w; im.image[0].h = c.h; im.image[0].d.assign(gray, gray + (size_t)c.w * c.h);
  if (mask) { im.mask.resize(1); im.mask[0].w = c.w; im.mask[0].h = c.h; im.mask[0].d.assign(mask, mask + (size_t)c.w * c.h); }
  h->st.images.push_back(std::move(im)); return (int)h->st.images.size() - 1;
}
// Camera mask of an intrinsics (image.cc:62-72, intrinsics.h:104): full-resolution uint8, values as the image masks; before initialize.
int orc_reg_set_camera_mask(orc_reg* h, int intr_id, const uint8_t* mask) {
  if (intr_id < 0 || intr_id >= (int)h->st.intr.size()) return -1;
  Intrinsics& in = h->st.intr[intr_id];
  in.camera_mask.resize(1); in.camera_mask[0].w = in.models[0].w; in.camera_mask[0].h = in.models[0].h;
  in.camera_mask[0].d.assign(mask, mask + (size_t)in.models[0].w * in.models[0].h);
  return 0;
}
// Problem::InitializeImages (problem.cc:478-494) + pyramids. Returns image_scale_count, or -1 on odd pyramid parents.
int orc_reg_initialize(orc_reg* h) {
  int count = 1;
  auto scale_count = [&](const Intrinsics& in) {
    const int px = in.models[0].w * in.models[0].h;
    const double af = px * 1.0 / h->prm.max_initial_image_area_in_pixels;
    return std::max<int>(2, 1 + (int)std::ceil(log(af) / log(4)));
  };
  for (const Intrinsics& in : h->st.intr) count = std::max(count, scale_count(in));
  if (h->prm.image_scale_count_override > 0) count = h->prm.image_scale_count_override;
  h->image_scale_count = count;
  for (Intrinsics& in : h->st.intr) {
    const int c = h->prm.image_scale_count_override > 0 ? count : scale_count(in);
    in.min_image_scale = count - c;
    in.models.resize(count - in.min_image_scale);
    in.build_pyramid();
  }
  for (Intrinsics& in : h->st.intr) if (!in.camera_mask.empty()) { in.camera_mask.resize(in.models.size()); build_mask_pyramid(in.camera_mask); }
  for (Image& im : h->st.images) {
    const size_t levels = h->st.intr[im.intrinsics_id].models.size();
    im.image.resize(levels);
    if (!build_image_pyramid(im.image)) return -1;
    if (!im.mask.empty()) { im.mask.resize(levels); build_mask_pyramid(im.mask); }
  }
  return count;
}
// One point scale: xyz n*3, radius, neighbour indices n*K, grey colours n (fixed descriptors d = c_nbr - c_ctr, problem.cc:550-572).
int orc_reg_add_point_scale(orc_reg* h, const float* xyz, size_t n, float radius, const uint64_t* nbr, const float* colors) {
  ScalePoints P; P.xyz.assign(xyz, xyz + 3 * n); P.radius = radius; P.nbr.assign(nbr, nbr + n * K(h));
  P.fixed_desc.assign(n * K(h), 0.f); P.var_desc.assign(n * K(h), 0.f); P.obs_count.assign(n, 0);
  if (h->prm.fixed_residuals_weight > 0)
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < K(h); ++k) P.fixed_desc[i * K(h) + k] = colors[P.nbr[i * K(h) + k]] - colors[i];
      P.obs_count[i] = 99999;
    }
  h->pts.push_back(std::move(P)); return (int)h->pts.size() - 1;
}
// OcclusionGeometry::AddMesh (occlusion_geometry.cc:87-130) with compute_edges = true. Vertices in the global frame.
void orc_reg_set_mesh(orc_reg* h, const float* v, size_t nv, const uint32_t* f, size_t nf) {
  h->mesh.v.assign(v, v + 3 * nv); h->mesh.f.assign(f, f + 3 * nf);
  build_mesh_edges(&h->mesh);
  h->has_mesh = nf > 0;
}
uint64_t orc_reg_mesh_edges(orc_reg* h, uint32_t* v1, uint32_t* v2, uint32_t* f1, uint32_t* f2, uint8_t* flags) {
  const auto& E = h->mesh.edges;
  if (v1) for (size_t i = 0; i < E.size(); ++i) { v1[i] = E[i].v1; v2[i] = E[i].v2; f1[i] = E[i].f1; f2[i] = E[i].f2; flags[i] = (E[i].open ? 1 : 0) | (E[i].opposite ? 2 : 0); }
  return E.size();
}
void orc_reg_set_splat_points(orc_reg* h, const float* xyz, size_t n) { h->splat_xyz.assign(xyz, xyz + 3 * n); h->has_splats = n > 0; }
// A rig with `ncam` cameras (rig.h:40-73); image_T_rig: 7 floats per camera (qx qy qz qw tx ty tz), camera 0 = reference (identity).
int orc_reg_add_rig(orc_reg* h, int ncam, const float* image_T_rig) {
  if (ncam < 2) return -1;
  Rig r; r.image_T_rig.resize(ncam);
  for (int c = 0; c < ncam; ++c) { const float* T = image_T_rig + 7 * c; r.image_T_rig[c].q = {T[0], T[1], T[2], T[3]}; r.image_T_rig[c].t = {T[4], T[5], T[6]}; }
  h->st.rigs.push_back(r); return (int)h->st.rigs.size() - 1;
}
// One set of images recorded together by a rig (rig_images.h:38-64); image_ids[c] = image of camera c, all present. The dependent
// images' poses are set to image_T_rig[c] * image_T_global(reference), as AssignRigs leaves them (rig.cc:216-250).
int orc_reg_add_rig_images(orc_reg* h, int rig_id, const int* image_ids) {
  if (rig_id < 0 || rig_id >= (int)h->st.rigs.size()) return -1;
  const Rig& rig = h->st.rigs[rig_id];
  RigImages ri; ri.rig_id = rig_id;
  for (size_t c = 0; c < rig.image_T_rig.size(); ++c) {
    const int id = image_ids[c];
    if (id < 0 || id >= (int)h->st.images.size() || h->st.images[id].rig_images_id >= 0) return -1;
    ri.image_ids.push_back(id);
  }
  h->rig_images.push_back(ri);
  const int rid = (int)h->rig_images.size() - 1;
  for (size_t c = 0; c < ri.image_ids.size(); ++c) {
    Image& im = h->st.images[ri.image_ids[c]];
    im.rig_images_id = rid; im.rig_camera_index = (int)c;
    if (c > 0) im.image_T_global = se3_mul(rig.image_T_rig[c], h->st.images[ri.image_ids[0]].image_T_global);
  }
  return rid;
}
void orc_reg_get_rigs(orc_reg* h, float* out) {
  for (const Rig& r : h->st.rigs) for (const SE3f& T : r.image_T_rig) { out[0] = T.q.x; out[1] = T.q.y; out[2] = T.q.z; out[3] = T.q.w; out[4] = T.t.x; out[5] = T.t.y; out[6] = T.t.z; out += 7; }
}
void orc_reg_set_rigs(orc_reg* h, const float* in) {
  for (Rig& r : h->st.rigs) for (SE3f& T : r.image_T_rig) { T.q = {in[0], in[1], in[2], in[3]}; T.t = {in[4], in[5], in[6]}; in += 7; }
}
// variable index of an intrinsics block (kind 0), a rig's extrinsics block (1), the pose block an image uses (2)
int orc_reg_variable_index(orc_reg* h, int kind, int id) { return kind == 0 ? intr_var(h, id) : kind == 1 ? rig_var(h, id) : pose_var(h, id); }

int orc_reg_set_depth_map(orc_reg* h, int image, int w, int hh, const float* d) {
  Image& im = h->st.images[image]; im.given_depth.w = w; im.given_depth.h = hh; im.given_depth.d.assign(d, d + (size_t)w * hh); im.has_given_depth = true; return 0;
}
void orc_reg_set_image_scale(orc_reg* h, int s) { h->current_image_scale = s; }
int orc_reg_image_scale_count(orc_reg* h) { return h->image_scale_count; }
int orc_reg_num_variables(orc_reg* h) { return num_variables(h); }

int orc_reg_render_depth(orc_reg* h, int image, int* w, int* hh, float* out) {
  const Image& im = h->st.images[image]; const Intrinsics& in = h->st.intr[im.intrinsics_id];
  const int best = in.best_available(std::max(h->prm.min_occlusion_check_image_scale, h->current_image_scale));
  const ImgF d = render_depth(h, in, im, best);
  *w = d.w; *hh = d.h;
  if (out) std::memcpy(out, d.d.data(), d.d.size() * sizeof(float));
  return best;
}

// ComputeMinMaxPointRadius over all images (multi_scale_point_cloud.cc:126-184 called from :232-255) with the visibility test of
// _AppendObservationsForImageNoScale (visibility_estimator.cc:296-364). min_radius / max_radius: in/out, initialised by the caller
// (+inf / -inf in CreateMultiScalePointCloud). Needs initialize() (pyramids) and the occlusion geometry of the handle.
void orc_reg_min_max_point_radius(orc_reg* h, const float* xyz, size_t n, double min_scaling_factor, float* min_radius, float* max_radius) {
  for (size_t ii = 0; ii < h->st.images.size(); ++ii) {
    const Image& im = h->st.images[ii]; const Intrinsics& intr = h->st.intr[im.intrinsics_id];
    const int image_scale = intr.best_available(std::max(h->prm.min_occlusion_check_image_scale, h->current_image_scale));
    const ImgF depth = render_depth(h, intr, im, image_scale);
    const Pinhole& cam = intr.model(image_scale);
    const Pinhole& cam0 = intr.model(0);                 // min_image_scale_camera = *intrinsics.model(0) (:240)
    std::vector<float> table;
    if (cam0.has_lookup()) { table.resize((size_t)2 * cam0.w * cam0.h); cam0.undistortion_lookup(table.data()); }
    float R[9]; quat_to_matrix(im.image_T_global.q, R);
    const int level = image_scale - intr.min_image_scale;
    for (size_t pi = 0; pi < n; ++pi) {
      V3f pp; rigid_pp(R, im.image_T_global.t, &xyz[3 * pi], &pp);
      if (!(pp.z > 0.f)) continue;
      float ixx, ixy; cam.project(pp.x / pp.z, pp.y / pp.z, &ixx, &ixy);
      const int ix = f2i(ixx + 0.5f), iy = f2i(ixy + 0.5f);
      if (!(ixx + 0.5f >= 0 && ixy + 0.5f >= 0 && ix >= 0 && iy >= 0 && ix < cam.w && iy < cam.h &&
            depth.d[(size_t)iy * depth.w + ix] + h->prm.occlusion_depth_threshold >= pp.z)) continue;
      if (level < (int)im.mask.size() && !im.mask[level].d.empty() && im.mask[level].at(iy, ix) != 0) continue;
      if (level < (int)intr.camera_mask.size() && !intr.camera_mask[level].d.empty() && intr.camera_mask[level].at(iy, ix) != 0) continue;
      if (im.image[level].at(iy, ix) > h->prm.maximum_valid_intensity) continue;
      float returned_scale = image_scale - 1e-6f;
      float ox = ixx, oy = ixy;
      if (returned_scale < 0.f) { returned_scale = 0.f; ox = 0.5f * (ixx + 0.5f) - 0.5f; oy = 0.5f * (ixy + 0.5f) - 0.5f; }
      const Observation o{pi, ox, oy, returned_scale};
      // image_x_at_scale(intrinsics.min_image_scale) (point_observation.h:84-93): double pow, float result
      const double p2 = std::pow(2, smaller_scale(o) - intr.min_image_scale);
      const float x0 = (float)(p2 * (o.x + 0.5f) - 0.5f), y0 = (float)(p2 * (o.y + 0.5f) - 0.5f);
      const float kPixelDistance = 0.5f;
      const float offx = (x0 - kPixelDistance < 0) ? (x0 + kPixelDistance) : (x0 - kPixelDistance);
      float nx, ny; cam0.image_to_normalized(table.data(), offx, y0, &nx, &ny);
      const V3f off{pp.z * nx, pp.z * ny, pp.z * 1.f};
      const float dx = pp.x - off.x, dy = pp.y - off.y, dz = pp.z - off.z;
      const float point_radius = std::sqrt(sum3(dx * dx, dy * dy, dz * dz));
      min_radius[pi] = std::min<float>(min_radius[pi], point_radius);
      max_radius[pi] = std::max<float>(max_radius[pi], (float)(point_radius / min_scaling_factor));
    }
  }
}

// ---- GroundTruthCreator (src/exe/ground_truth_creator.cc:44-215) ----
// The visibility test both passes share (:66-79, :163-174): in front of the camera, inside the highest-resolution image, not behind
// the occlusion depth map rendered at intrinsics.min_image_scale, not on a kEvalObs (= 2) pixel of the image mask.
static bool gt_visible(const orc_reg* h, const Image& im, const Intrinsics& intr, const Pinhole& cam, const ImgF& depth, const float R[9], const float* p,
                       int* ox, int* oy, float* oz) {
  V3f pp; rigid_pp(R, im.image_T_global.t, p, &pp);
  if (!(pp.z > 0)) return false;
  float px, py; cam.project(pp.x / pp.z, pp.y / pp.z, &px, &py);
  const int ix = f2i(px + 0.5f), iy = f2i(py + 0.5f);
  if (!(ix >= 0 && iy >= 0 && ix < cam.w && iy < cam.h && depth.d[(size_t)iy * depth.w + ix] + h->prm.occlusion_depth_threshold >= pp.z)) return false;
  if (!im.mask.empty() && !im.mask[0].d.empty() && im.mask[0].at(iy, ix) == 2) return false;
  *ox = ix; *oy = iy; *oz = pp.z;
  return true;
}
// AccumulateScanObservationsForImage (:44-82): counts[i] += 1 for every scan point visible in `image`.
void orc_reg_gt_accumulate_observations(orc_reg* h, int image, const float* xyz, size_t n, int32_t* counts) {
  const Image& im = h->st.images[image]; const Intrinsics& intr = h->st.intr[im.intrinsics_id];
  const Pinhole& cam = intr.model(0);
  const ImgF depth = render_depth(h, intr, im, intr.min_image_scale);
  float R[9]; quat_to_matrix(im.image_T_global.q, R);
  for (size_t i = 0; i < n; ++i) { int ix, iy; float z; if (gt_visible(h, im, intr, cam, depth, R, &xyz[3 * i], &ix, &iy, &z)) counts[i] += 1; }
}
// CreateGroundTruthForImage (:84-215) without the file I/O: occlusion depth (nullable), ground-truth depth = per-pixel minimum depth of
// the visible points observed in >= 2 images (+inf elsewhere), scan rendering (in/out BGR image, nullable): squares of
// 2 * scan_point_radius + 1 pixels painted in scan order, later points over earlier ones.
void orc_reg_gt_create(orc_reg* h, int image, const float* xyz, const uint8_t* rgb, size_t n, const int32_t* counts, int scan_point_radius,
                       float* occlusion_depth, float* gt_depth, uint8_t* rendering_bgr) {
  const Image& im = h->st.images[image]; const Intrinsics& intr = h->st.intr[im.intrinsics_id];
  const Pinhole& cam = intr.model(0);
  const ImgF depth = render_depth(h, intr, im, intr.min_image_scale);
  if (occlusion_depth) std::memcpy(occlusion_depth, depth.d.data(), depth.d.size() * sizeof(float));
  if (gt_depth) for (size_t i = 0; i < (size_t)cam.w * cam.h; ++i) gt_depth[i] = std::numeric_limits<float>::infinity();
  float R[9]; quat_to_matrix(im.image_T_global.q, R);
  for (size_t i = 0; i < n; ++i) {
    if (counts[i] < 2) continue;
    int ix, iy; float z;
    if (!gt_visible(h, im, intr, cam, depth, R, &xyz[3 * i], &ix, &iy, &z)) continue;
    if (rendering_bgr && rgb) {
      const int min_x = std::max(0, ix - scan_point_radius), min_y = std::max(0, iy - scan_point_radius);
      const int end_x = std::min(cam.w, ix + scan_point_radius + 1), end_y = std::min(cam.h, iy + scan_point_radius + 1);
      for (int y = min_y; y < end_y; ++y) for (int x = min_x; x < end_x; ++x) {
        uint8_t* px = rendering_bgr + 3 * ((size_t)y * cam.w + x);
        px[0] = rgb[3 * i + 2]; px[1] = rgb[3 * i + 1]; px[2] = rgb[3 * i];
      }
    }
    if (gt_depth) gt_depth[(size_t)iy * cam.w + ix] = std::min(gt_depth[(size_t)iy * cam.w + ix], z);
  }
}

// CreateObservationsForAllImages + DetermineIfAllNeighborsAreObserved (optimizer.cc:123-128)
void orc_reg_create_observations(orc_reg* h, int border) {
  h->obs.assign(h->st.images.size(), {}); h->nbr_obs.assign(h->st.images.size(), {});
  for (size_t im = 0; im < h->st.images.size(); ++im) {
    append_observations(h, h->st.images[im], h->st.intr[h->st.images[im].intrinsics_id], border, nullptr, &h->obs[im]);
    h->nbr_obs[im].resize(h->pts.size());
    for (size_t ps = 0; ps < h->pts.size(); ++ps) neighbors_observed(h, (int)ps, h->obs[im][ps], &h->nbr_obs[im][ps]);
  }
}
uint64_t orc_reg_num_observations(orc_reg* h, int image, int ps) { return h->obs[image][ps].size(); }
void orc_reg_get_observations(orc_reg* h, int image, int ps, uint64_t* idx, float* x, float* y, float* s, uint8_t* nb) {
  const auto& o = h->obs[image][ps];
  for (size_t i = 0; i < o.size(); ++i) { idx[i] = o[i].point_index; x[i] = o[i].x; y[i] = o[i].y; s[i] = o[i].scale; if (nb) nb[i] = h->nbr_obs[image][ps][i]; }
}

// ColorOptimizer::Apply (color_optimizer.cc:40-123); images in ascending id.
void orc_reg_color_update(orc_reg* h) {
  for (size_t ps = 0; ps < h->pts.size(); ++ps) {
    ScalePoints& P = h->pts[ps];
    std::fill(P.obs_count.begin(), P.obs_count.end(), 0);
    std::fill(P.var_desc.begin(), P.var_desc.end(), 0.f);
    for (size_t im = 0; im < h->st.images.size(); ++im) {
      const Image& I = h->st.images[im]; const Intrinsics& in = h->st.intr[I.intrinsics_id];
      const auto& o = h->obs[im][ps];
      std::vector<float> inten(P.n(), -1);
      for (const Observation& ob : o)
        trilinear(I.image[smaller_scale(ob) - in.min_image_scale], I.image[larger_scale(ob) - in.min_image_scale], ob.x, ob.y,
                  1 - (ob.scale - static_cast<int>(ob.scale)), &inten[ob.point_index]);
      for (size_t i = 0; i < o.size(); ++i) if (h->nbr_obs[im][ps][i]) {
        const uint64_t pi = o[i].point_index;
        P.obs_count[pi] += 1;
        for (int k = 0; k < K(h); ++k) P.var_desc[pi * K(h) + k] += inten[P.nbr[pi * K(h) + k]] - inten[pi];
      }
    }
    for (size_t i = 0; i < P.n(); ++i) if (P.obs_count[i] > 1) for (int k = 0; k < K(h); ++k) P.var_desc[i * K(h) + k] /= P.obs_count[i];
  }
}
void orc_reg_get_descriptors(orc_reg* h, int ps, float* fixed, float* variable, int* counts) {
  const ScalePoints& P = h->pts[ps];
  if (fixed) std::memcpy(fixed, P.fixed_desc.data(), P.fixed_desc.size() * 4);
  if (variable) std::memcpy(variable, P.var_desc.data(), P.var_desc.size() * 4);
  if (counts) std::memcpy(counts, P.obs_count.data(), P.obs_count.size() * 4);
}

// CostCalculator::ComputeCost (cost_calculator.cc:44-100). sums = [fixed_sum, n_fixed, var_sum, n_var, 0, 0]
double orc_reg_cost(orc_reg* h, double sums[6]) {
  Sums s;
  for (size_t im = 0; im < h->st.images.size(); ++im)
    for (size_t ps = 0; ps < h->pts.size(); ++ps)
      accumulate_residuals(h, h->st.intr[h->st.images[im].intrinsics_id], h->st.images[im], (int)ps, h->obs[im][ps], h->nbr_obs[im][ps], &s);
  if (sums) { sums[0] = s.fixed_sum; sums[1] = (double)s.nf; sums[2] = s.var_sum; sums[3] = (double)s.nv; sums[4] = sums[5] = 0; }
  if (s.nf == 0 && s.nv == 0) return std::numeric_limits<double>::infinity();
  return compute_cost(h, s);
}

// H (nv*nv col-major, the UPPER triangle the solver reads, mirrored), b, sums; returns the cost ("initial residual").
double orc_reg_accumulate(orc_reg* h, double* H, double* b, double sums[6]) {
  std::vector<double> Hv, bv; Sums s;
  accumulate_all(h, &Hv, &bv, &s);
  const int nv = num_variables(h);
  if (H) for (int c = 0; c < nv; ++c) for (int r = 0; r < nv; ++r) H[(size_t)c * nv + r] = r <= c ? Hv[(size_t)c * nv + r] : Hv[(size_t)r * nv + c];
  if (b) std::memcpy(b, bv.data(), sizeof(double) * nv);
  if (sums) { sums[0] = s.fixed_sum; sums[1] = (double)s.nf; sums[2] = s.var_sum; sums[3] = (double)s.nv; sums[4] = sums[5] = 0; }
  return compute_cost(h, s);
}

void orc_reg_get_state(orc_reg* h, float* intr_params, float* poses) {
  for (size_t i = 0; i < h->st.intr.size(); ++i) h->st.intr[i].models[0].get_params(intr_params + intr_var(h, (int)i));
  for (size_t i = 0; i < h->st.images.size(); ++i) {
    const SE3f& T = h->st.images[i].image_T_global; float* p = poses + 7 * i;
    p[0] = T.q.x; p[1] = T.q.y; p[2] = T.q.z; p[3] = T.q.w; p[4] = T.t.x; p[5] = T.t.y; p[6] = T.t.z;
  }
}
void orc_reg_set_state(orc_reg* h, const float* intr_params, const float* poses) {
  for (size_t i = 0; i < h->st.intr.size(); ++i) { Pinhole& m = h->st.intr[i].models[0]; m.set(m.type, m.w, m.h, intr_params + intr_var(h, (int)i)); h->st.intr[i].build_pyramid(); }
  for (size_t i = 0; i < h->st.images.size(); ++i) {
    SE3f& T = h->st.images[i].image_T_global; const float* p = poses + 7 * i;
    T.q = {p[0], p[1], p[2], p[3]}; T.t = {p[4], p[5], p[6]};
  }
}

// Cost of the state (current state + delta) with the CURRENT observations' visibility lists frozen (what each LM try evaluates).
double orc_reg_cost_for_delta(orc_reg* h, const double* delta) {
  std::vector<std::vector<std::vector<uint64_t>>> lists(h->st.images.size());
  for (size_t im = 0; im < h->st.images.size(); ++im) {
    lists[im].resize(h->pts.size());
    for (size_t ps = 0; ps < h->pts.size(); ++ps) for (const Observation& ob : h->obs[im][ps]) lists[im][ps].push_back(ob.point_index);
  }
  std::vector<double> d(delta, delta + num_variables(h));
  return residual_for_state(h, delta_state(h, d), lists);
}

// IntrinsicsAndPoseOptimizer::Apply (intrinsics_and_pose_optimizer.cc:48-259). Returns the number of LM tries used.
int orc_reg_apply(orc_reg* h, float* lambda, float* max_change, int* applied) {
  std::vector<double> H, b; Sums s;
  accumulate_all(h, &H, &b, &s);
  const int nv = num_variables(h);
  const double initial = compute_cost(h, s);
  std::vector<std::vector<std::vector<uint64_t>>> lists(h->st.images.size());
  for (size_t im = 0; im < h->st.images.size(); ++im) {
    lists[im].resize(h->pts.size());
    for (size_t ps = 0; ps < h->pts.size(); ++ps) for (const Observation& ob : h->obs[im][ps]) lists[im][ps].push_back(ob.point_index);
  }
  *applied = 0;
  int tries = 0;
  for (int lm = 0; lm < 10; ++lm) {
    ++tries;
    std::vector<double> HL = H;
    for (int i = 0; i < nv; ++i) HL[(size_t)i * nv + i] *= (1 + (*lambda));
    std::vector<double> x; ldlt_solve_upper(HL, nv, b, &x);
    std::vector<double> neg(nv); for (int i = 0; i < nv; ++i) neg[i] = -1 * x[i];
    const State ns = delta_state(h, neg);
    const double nr = residual_for_state(h, ns, lists);
    if (nr < initial || lm == 9) {
      double mx = -std::numeric_limits<double>::infinity(); for (double v : x) mx = std::max(mx, v);   // signed maxCoeff (:245)
      *max_change = (float)mx;
      h->st = ns;
      *lambda = 0.5f * (*lambda);
      *applied = 1;
      break;
    } else {
      *lambda = 2.f * (*lambda);
    }
  }
  return tries;
}

// Optimizer::RunOnCurrentScale (optimizer.cc:49-182). Returns iterations executed; *converged set.
int orc_reg_run_on_current_scale(orc_reg* h, int max_it, float max_change_thr, int no_opt_thr, double* optimum_cost, int* converged) {
  h->current_image_scale = std::min(h->current_image_scale, h->image_scale_count - 1 - 1);   // max_image_scale() - 1
  float lambda = 64.0f;
  int without = 0, it_done = 0;
  *optimum_cost = std::numeric_limits<double>::infinity();
  *converged = 0;
  State best = h->st;
  for (int it = 0; it < max_it; ++it) {
    ++it_done;
    int applied = 1; float max_change = std::numeric_limits<float>::infinity();
    if (it > 0) { applied = 0; max_change = 0; orc_reg_apply(h, &lambda, &max_change, &applied); }
    orc_reg_create_observations(h, 1);
    if (h->prm.variable_residuals_weight > 0) orc_reg_color_update(h);
    const double cost = orc_reg_cost(h, nullptr);
    if (cost < *optimum_cost) { *optimum_cost = cost; without = 0; best = h->st; } else { ++without; }
    if (!applied || max_change < max_change_thr || without >= no_opt_thr) { *converged = 1; break; }
  }
  h->st = best;
  return it_done;
}

// Unit hooks for the reference's own unit tests.
void orc_reg_point_jacobians(orc_reg* h, int image, int ps, uint64_t obs_index, float* intensity, float* jK, float jP[6]) {
  orc_reg_point_jacobians_rig(h, image, ps, obs_index, intensity, jK, jP, nullptr);
}
// jR (6, may be null): derivative by the rig extrinsics for a dependent rig image (zeros otherwise)
void orc_reg_point_jacobians_rig(orc_reg* h, int image, int ps, uint64_t obs_index, float* intensity, float* jK, float jP[6], float* jR) {
  const Image& im = h->st.images[image]; const Intrinsics& in = h->st.intr[im.intrinsics_id];
  float R[9]; quat_to_matrix(im.image_T_global.q, R);
  const Observation& ob = h->obs[image][ps][obs_index];
  int rv; const RigCtx rc = rig_ctx(h, h->st, image, &rv);
  float tmp[6] = {0, 0, 0, 0, 0, 0};
  point_intensity_and_jacobians(h, in, im, R, h->pts[ps].radius, &h->pts[ps].xyz[3 * ob.point_index], ob, intensity, jK, jP, &rc, tmp);
  if (jR) for (int i = 0; i < 6; ++i) jR[i] = tmp[i];
}
int orc_interp_bilinear(const uint8_t* img, int w, int hh, float x, float y, float* v, float* dx, float* dy) {
  Img8 im; im.w = w; im.h = hh; im.d.assign(img, img + (size_t)w * hh);
  const int ix = (int)x, iy = (int)y;
  if (x < 0.f || y < 0.f || ix >= w - 1 || iy >= hh - 1) return 0;     // interpolate_bilinear.h:79-112
  float a, b, c; bilinear_d(im, x, y, ix, iy, &a, &b, &c);
  const float plain = bilinear(im, x, y, ix, iy);
  if (v) *v = plain; if (dx) *dx = b; if (dy) *dy = c;
  return a == plain ? 1 : 2;
}
void orc_interp_trilinear(const uint8_t* img0, int w0, int h0, const uint8_t* img1, float x, float y, float z, float* v, float* dx, float* dy, float* dz) {
  Img8 a, b; a.w = w0; a.h = h0; a.d.assign(img0, img0 + (size_t)w0 * h0); b.w = 2 * w0; b.h = 2 * h0; b.d.assign(img1, img1 + (size_t)4 * w0 * h0);
  float pv; trilinear(a, b, x, y, z, &pv);
  trilinear_d(a, b, x, y, z, v, dx, dy, dz);
  if (pv != *v) *v = std::numeric_limits<float>::quiet_NaN();
}
int orc_cam_param_count(int type) { return Camera::known(type) ? Camera::param_count(type) : -1; }
int orc_cam_cutoff(int type, int w, int hh, const float* params, float out[2]) {
  if (!Camera::known(type)) return -1;
  Camera c; c.set(type, w, hh, params); out[0] = c.cutoff2; out[1] = c.inner_cutoff2; return 0;
}
int orc_cam_eval(int type, int w, int hh, const float* params, int op, const float* in, size_t n, float* out) {
  if (!Camera::known(type)) return -1;
  Camera c; c.set(type, w, hh, params);
  const int np = c.np();
  for (size_t i = 0; i < n; ++i) {
    if (op == 0) c.distort(in[2 * i], in[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
    else if (op == 1) c.project(in[2 * i], in[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
    else if (op == 2) c.d_by_world(V3f{in[3 * i], in[3 * i + 1], in[3 * i + 2]}, &out[6 * i]);
    else if (op == 3) c.d_by_intrinsics(V3f{in[3 * i], in[3 * i + 1], in[3 * i + 2]}, &out[(size_t)2 * np * i]);
    else if (op == 4) {
      c.undistort(in[2 * i], in[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
    } else if (op == 5) c.distort_deriv(in[2 * i], in[2 * i + 1], &out[4 * i]);
    else return -2;
  }
  return 0;
}
float orc_robust(int type, float p, float r, int weight) { Robust R; R.type = type; R.p = p; return weight ? R.weight(r) : R.residual(r); }
void orc_image_pyramid_level(const uint8_t* src, int w, int hh, uint8_t* dst) {
  std::vector<Img8> p(2); p[0].w = w; p[0].h = hh; p[0].d.assign(src, src + (size_t)w * hh);
  build_image_pyramid(p); std::memcpy(dst, p[1].d.data(), p[1].d.size());
}

}  // extern "C"
