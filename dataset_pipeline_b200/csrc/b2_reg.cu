// Host side of Path B behind the C ABI (include/eth3d_b200.h, b2_reg_*): owns the problem state in HBM (image / mask / intrinsics
// pyramids, multi-resolution points, descriptors, per-(image, point-scale) observation sets) and runs the optimizer loop of
// the reference on top of the kernels in b2_reg_kernels.cuh.
//
// Reference seams replaced:
//   Problem::InitializeImages / LoadImages pyramids            /root/reference/src/opt/problem.cc:478-503, image.cc:106-154, intrinsics.cc:45-79
//   OcclusionGeometry::RenderDepthMap (splats / none)          /root/reference/src/opt/occlusion_geometry.cc:185-282,404-464
//   VisibilityEstimator::CreateObservationsForAllImages        /root/reference/src/opt/visibility_estimator.cc:49-91
//   IntrinsicsAndPoseOptimizer::Apply / ComputeResidualForState /root/reference/src/opt/intrinsics_and_pose_optimizer.cc:48-259,385-440
//   CostCalculator::ComputeCost, ColorOptimizer::Apply         /root/reference/src/opt/cost_calculator.cc:44-100, color_optimizer.cc:40-123
//   Optimizer::RunOnCurrentScale                               /root/reference/src/opt/optimizer.cc:49-182
// Images are visited in ascending id (the reference iterates unordered_maps); per-image fp64 partial sums are combined on the
// host in that fixed order, so results are reproducible run to run.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <vector>

#include "b2_common.cuh"
#include "b2_hostmath.h"
#include "b2_mesh_edges.h"
#include "b2_reg_kernels.cuh"

namespace b2 {

// One camera (pyramid level) from the reference's parameter vector (GetParameters order). Cut-offs are filled in by
// compute_cutoffs() — the constructors' InitCutoff (camera_thin_prism.cc:40,50).
static void cam_set(Cam* c, int type, int w, int h, const float* p) {
  CamModel m{kDistNone, 0, 0, 0};
  cam_model(type, &m);
  c->type = type; c->w = w; c->h = h; c->dist = m.dist; c->fisheye = m.fisheye; c->unique_focal = m.unique_focal; c->nd = m.nd;
  const int nb = m.unique_focal ? 3 : 4;
  if (m.unique_focal) { c->fx = c->fy = p[0]; c->cx = p[1]; c->cy = p[2]; } else { c->fx = p[0]; c->fy = p[1]; c->cx = p[2]; c->cy = p[3]; }
  c->fx_inv = (float)(1.0 / c->fx); c->fy_inv = (float)(1.0 / c->fy);            // camera_base.cc:83
  c->cx_inv = (float)(-1.0 * c->cx / c->fx); c->cy_inv = (float)(-1.0 * c->cy / c->fy);
  for (int i = 0; i < 8; ++i) c->d[i] = i < m.nd ? p[nb + i] : 0.f;
  c->two_tan = c->image_radius = 0.f;
  if (m.dist == kDistFOV) {                                                       // camera_fisheye_fov.cc:38-43 (tan correctly rounded, see b2_camera.cuh)
    c->two_tan = 2.0f * (float)std::tan((double)(0.5f * c->d[0]));
    c->image_radius = (float)(M_PI / (double)(2 * c->d[0]));
  }
  c->cutoff2 = c->inner_cutoff2 = std::numeric_limits<float>::infinity();
}
static void cam_get(const Cam& c, float* p) {
  const int nb = c.unique_focal ? 3 : 4;
  if (c.unique_focal) { p[0] = c.fx; p[1] = c.cx; p[2] = c.cy; } else { p[0] = c.fx; p[1] = c.fy; p[2] = c.cx; p[3] = c.cy; }
  for (int i = 0; i < c.nd; ++i) p[nb + i] = c.d[i];
}
static Cam cam_half(const Cam& s) {                                          // CameraBaseImpl::ScaledBy(0.5), camera_base_impl.h:70-89
  const float f = 0.5f;
  float p[kMaxIntrinsics] = {0}; cam_get(s, p);
  if (!s.unique_focal) { p[0] *= f; p[1] *= f; p[2] = f * (s.cx + 0.5f) - 0.5f; p[3] = f * (s.cy + 0.5f) - 0.5f; }
  else { p[0] *= f; p[1] = f * (s.cx + 0.5f) - 0.5f; p[2] = f * (s.cy + 0.5f) - 0.5f; }
  Cam d; cam_set(&d, s.type, (int)(f * s.w + 0.5f), (int)(f * s.h + 0.5f), p);
  return d;
}

struct IntrinsicsB {
  std::vector<Cam> models;   // [0] = original resolution
  int min_image_scale = -1;
  int np() const { return cam_param_count(models[0].type); }
  int kni() const { return cam_kernel_ni(models[0].type); }   // width of the Jacobian rows the kernels carry (>= np, zero padded)
  const Cam& model(int image_scale) const { return models[std::max(0, image_scale - min_image_scale)]; }
  int best_available(int image_scale) const { return std::min<int>(min_image_scale + (int)models.size() - 1, std::max<int>(min_image_scale, image_scale)); }
  void build_pyramid() { for (size_t i = 1; i < models.size(); ++i) models[i] = cam_half(models[i - 1]); }
};

struct RigB { std::vector<Pose> image_T_rig; };                      // rig.h:40-73 ([0] = reference camera, identity)
struct RigImagesB { int rig_id = 0; std::vector<int> image_ids; };   // rig_images.h:38-64

struct ImageB {
  int intrinsics_id = 0;
  int rig_images_id = -1, rig_camera_index = 0;
  Pose pose;                               // image_T_global
  std::vector<DevBuf> img, mask;           // pyramids in HBM (mask empty = none)
  std::vector<int> lw, lh;                 // level sizes (image pyramid: truncating halving)
  bool has_mask = false;
  DevBuf given_depth; int gd_w = 0, gd_h = 0; bool has_given_depth = false;
};

struct ScaleB {
  size_t n = 0; float radius = 0.f;
  DevBuf xyz, nbr, fixed_desc, var_desc, obs_count;
};

struct ObsSet {       // observations of one (image, point scale)
  size_t count = 0;
  DevBuf idx, x, y, s, nb, inten, jK, jP, jR, slot;   // slot: per POINT (n), -1 = unobserved; jR only for dependent rig images
};

struct StateB { std::vector<IntrinsicsB> intr; std::vector<Pose> poses; std::vector<RigB> rigs; };

static inline unsigned int divup(size_t a, size_t b) { return (unsigned int)((a + b - 1) / b); }

}  // namespace b2

using namespace b2;

struct b2_reg {
  b2_reg_params prm;
  int device = 0, sms = 148;
  cudaStream_t stream = nullptr;
  std::vector<IntrinsicsB> intr;
  std::vector<ImageB> images;
  std::vector<RigB> rigs;
  std::vector<RigImagesB> rig_images;
  std::map<int, std::vector<DevBuf>> cam_masks;      // camera-mask pyramids by intrinsics id (not part of the optimised state)
  std::vector<ScaleB> pts;
  DevBuf splats; size_t nsplats = 0;
  DevBuf mesh_v, mesh_f, mesh_fn, mesh_edges, big_list, big_count, warp_list, splat_queue, depth_masked; size_t mesh_nv = 0, mesh_nf = 0, mesh_ne = 0;
  int image_scale_count = 0, current_image_scale = 0;
  bool initialized = false;
  b2_comm* comm = nullptr; int rank = 0, world = 1;     // multi-GPU: images dealt round-robin, sums allreduced (b2_reg_set_comm)
  DevBuf xchg;                                           // device staging of host-side partial sums for the allreduce
  std::vector<std::vector<ObsSet>> obs;      // [image][scale]
  std::vector<ObsSet> trial;                 // scratch sets for trial states, [scale]
  // scratch
  DevBuf flags, offs, cx, cy, cs, cub_tmp, depth, partials, results, cut_cams, cut_first, cut_starts, cut_points, cut_out;
  DevBuf w_nj, w_ws, w_wr, w_part;           // K12b: per-observation slots / merged weights / residual factors + their residual-sum partials
  int k12_mode = 0;                          // B2_K12: 0 = fp64 products, pre-pass (kr_residual_weights) + kr_accumulate_weighted (pinhole) / block-pair
                                             // kernels (wide systems) (default); 1 ("f32") = fp32 products, thread / shuffle kernels (reference op order);
                                             // 2 ("thread") = pinhole in one thread-per-observation kernel; 3 ("blocks") = block-pair kernel for pinhole too
  PinnedBuf pin;
  b2_reg_stats stats;
  int launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evj0 = nullptr, evj1 = nullptr, eva0 = nullptr, eva1 = nullptr;
  std::vector<cudaEvent_t> ev_pool;          // three events per (image, scale) of an accumulate call: K11 start, K11 end = K12 start, K12 end
};

namespace b2 {

static int K(const b2_reg* h) { return h->prm.point_neighbor_count; }
// Variable layout (CountAndIndexVariables, intrinsics_and_pose_optimizer.cc:442-473): [intrinsics 0 | intrinsics 1 | ... | rig 0
// extrinsics (6 per camera after the first) | rig 1 | ... | poses (6 each) of the images that own one — non-rig images and rig
// reference images, ascending id]. An intrinsics block has the camera model's ParameterCount(); a dependent rig image uses its
// reference image's pose block.
static int intr_var(const b2_reg* h, int id) { int v = 0; for (int i = 0; i < id; ++i) v += h->intr[i].np(); return v; }
static int rig_var(const b2_reg* h, int rig_id) {
  int v = intr_var(h, (int)h->intr.size());
  for (int r = 0; r < rig_id; ++r) v += 6 * ((int)h->rigs[r].image_T_rig.size() - 1);
  return v;
}
static bool is_dependent(const b2_reg* h, int im) { return h->images[im].rig_images_id >= 0 && h->images[im].rig_camera_index > 0; }
static int ref_image(const b2_reg* h, int im) { return is_dependent(h, im) ? h->rig_images[h->images[im].rig_images_id].image_ids[0] : im; }
static int pose_var(const b2_reg* h, int im) {
  const int owner = im < (int)h->images.size() ? ref_image(h, im) : im;
  int v = rig_var(h, (int)h->rigs.size());
  for (int i = 0; i < owner; ++i) if (!is_dependent(h, i)) v += 6;
  return v;
}
static int nvars(const b2_reg* h) { return pose_var(h, (int)h->images.size()); }

// K16: radius cut-offs of every pyramid level of the given intrinsics (the reference re-runs InitCutoff in each camera
// constructor: CreateUpdatedCamera + ScaledBy per level, intrinsics.cc:66-79). One batched search, one readback.
static int compute_cutoffs(b2_reg* h, std::vector<IntrinsicsB>* intr) {
  std::vector<Cam> cams; std::vector<int> first(1, 0); std::vector<std::pair<int, int>> where;      // generic search (camera_base_impl.h:410-462)
  std::vector<Cam> rcams; std::vector<std::pair<int, int>> rwhere;                                  // RadialBase search (camera_base_impl_radial.h:143-171)
  auto assign = [](Cam& c, float v) { if (c.fisheye) c.inner_cutoff2 = v; else c.cutoff2 = v; };     // a fisheye camera keeps +inf; its inner model carries the cut-off
  for (size_t i = 0; i < intr->size(); ++i)
    for (size_t l = 0; l < (*intr)[i].models.size(); ++l) {
      Cam& c = (*intr)[i].models[l];
      if (c.dist == kDistPolyTan || c.dist == kDistOpenCV || c.dist == kDistThinPrism) {
        cams.push_back(c); where.emplace_back((int)i, (int)l);
        first.push_back(first.back() + 2 * (c.w + c.h));
      } else if (c.dist == kDistRadial2 || c.dist == kDistPoly3 || c.dist == kDistPoly4) {
        rcams.push_back(c); rwhere.emplace_back((int)i, (int)l);
      } else if (c.dist == kDistRadial1) {
        if (c.d[0] < 0) assign(c, -1.f / (3 * c.d[0]));                                              // SimpleRadialCamera::InitCutoff (camera_simple_radial.cc:53-57)
      }
    }
  if (!rcams.empty()) {
    const int n = (int)rcams.size();
    B2_TRY(h->cut_cams.ensure(sizeof(Cam) * n)); B2_TRY(h->cut_out.ensure(sizeof(float) * n));
    B2_CUDA(cudaMemcpyAsync(h->cut_cams.p, rcams.data(), sizeof(Cam) * n, cudaMemcpyHostToDevice, h->stream));
    kr_cutoff_radial<<<divup(n, 32), 32, 0, h->stream>>>(h->cut_cams.as<Cam>(), n, h->cut_out.as<float>());
    ++h->launches;
    std::vector<float> out(n);
    B2_CUDA(cudaMemcpyAsync(out.data(), h->cut_out.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    B2_CUDA(cudaGetLastError());
    for (int k = 0; k < n; ++k) assign((*intr)[rwhere[k].first].models[rwhere[k].second], out[k]);
  }
  if (cams.empty()) return B2_OK;
  const int ncam = (int)cams.size(), npoints = first.back();
  B2_TRY(h->cut_cams.ensure(sizeof(Cam) * ncam)); B2_TRY(h->cut_first.ensure(sizeof(int) * (ncam + 1)));
  B2_TRY(h->cut_starts.ensure(sizeof(CutoffStart) * (size_t)npoints * 100)); B2_TRY(h->cut_points.ensure(sizeof(CutoffPoint) * (size_t)npoints));
  B2_TRY(h->cut_out.ensure(sizeof(float) * ncam));
  B2_CUDA(cudaMemcpyAsync(h->cut_cams.p, cams.data(), sizeof(Cam) * ncam, cudaMemcpyHostToDevice, h->stream));
  B2_CUDA(cudaMemcpyAsync(h->cut_first.p, first.data(), sizeof(int) * (ncam + 1), cudaMemcpyHostToDevice, h->stream));
  kr_cutoff_starts<<<npoints, 128, 0, h->stream>>>(h->cut_cams.as<Cam>(), h->cut_first.as<int>(), ncam, h->cut_starts.as<CutoffStart>());
  kr_cutoff_points<<<divup(npoints, 128), 128, 0, h->stream>>>(h->cut_starts.as<CutoffStart>(), npoints, h->cut_points.as<CutoffPoint>());
  kr_cutoff_final<<<ncam, 256, 0, h->stream>>>(h->cut_points.as<CutoffPoint>(), h->cut_first.as<int>(), h->cut_out.as<float>());
  h->launches += 3;
  std::vector<float> out(ncam);
  B2_CUDA(cudaMemcpyAsync(out.data(), h->cut_out.p, sizeof(float) * ncam, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  B2_CUDA(cudaGetLastError());
  for (int k = 0; k < ncam; ++k) assign((*intr)[where[k].first].models[where[k].second], out[k]);
  return B2_OK;
}

static bool owned(const b2_reg* h, size_t im) { return (int)(im % (size_t)h->world) == h->rank; }

// Sum-allreduce of `n` host doubles across the ranks (no-op on one GPU): staged through HBM, NCCL on the handle's stream.
static int allreduce_host(b2_reg* h, double* v, size_t n) {
  if (h->world == 1 || n == 0) return B2_OK;
  B2_TRY(h->xchg.ensure(n * 8));
  B2_CUDA(cudaMemcpyAsync(h->xchg.p, v, n * 8, cudaMemcpyHostToDevice, h->stream));
  B2_TRY(b2_comm_allreduce(h->comm, h->xchg.p, n, B2_F64, (void*)h->stream));
  B2_CUDA(cudaMemcpyAsync(v, h->xchg.p, n * 8, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  return B2_OK;
}
static int allreduce_sums(b2_reg* h, struct Sums* s);

static Pose3 pose3_of(const Pose& p) { Pose3 o; quat_matrix(p.q, o.R); for (int k = 0; k < 3; ++k) o.t[k] = p.t[k]; return o; }

static Levels levels_of(const b2_reg* h, const ImageB& im, const IntrinsicsB& in) {
  Levels L; std::memset(&L, 0, sizeof(L));
  L.nlevels = (int)in.models.size(); L.min_image_scale = in.min_image_scale;
  for (int l = 0; l < L.nlevels; ++l) {
    L.cam[l] = in.models[l];
    L.iw[l] = im.lw[l];
    L.img[l] = im.img[l].as<unsigned char>();
    L.mask[l] = im.has_mask ? im.mask[l].as<unsigned char>() : nullptr;
  }
  const auto cm = h->cam_masks.find(im.intrinsics_id);
  if (cm != h->cam_masks.end()) for (int l = 0; l < L.nlevels && l < (int)cm->second.size(); ++l) L.cmask[l] = cm->second[l].as<unsigned char>();
  return L;
}

static void begin_call(b2_reg* h) { h->launches = 0; cudaEventRecord(h->ev0, h->stream); }
static void end_call(b2_reg* h) {
  cudaEventRecord(h->ev1, h->stream); cudaEventSynchronize(h->ev1);
  float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  h->stats.ms_last_call = ms; h->stats.kernel_launches = h->launches;
}

// RenderDepthMap at `image_scale`: caller-supplied map, splats, or nullptr (= all-inf map, occlusion_geometry.cc:271-281: the
// test `inf + thr >= z` always passes, so the kernel simply skips the tap).
static int render_depth(b2_reg* h, const ImageB& im, const IntrinsicsB& in, const Pose& pose, int image_scale, const float** out) {
  const Cam& cam = in.model(image_scale);
  if (im.has_given_depth) {
    if (im.gd_w != cam.w || im.gd_h != cam.h)
      return set_error(B2_ERR_ARG, "given depth map is %dx%d but the occlusion-check scale %d is %dx%d", im.gd_w, im.gd_h, image_scale, cam.w, cam.h);
    *out = im.given_depth.as<float>(); return B2_OK;
  }
  if (h->nsplats == 0 && h->mesh_nf > 0) {
    // mesh path (occlusion_geometry.cc:211-270): K8 depth pass + K9 boundary masking
    const size_t px = (size_t)cam.w * cam.h;
    B2_TRY(h->depth.ensure(px * 4)); B2_TRY(h->depth_masked.ensure(px * 4));
    B2_TRY(h->big_list.ensure(std::max<size_t>(h->mesh_nf, 1) * 4)); B2_TRY(h->big_count.ensure(16));
    B2_TRY(h->warp_list.ensure(std::max<size_t>(h->mesh_nf, 1) * 4));
    const Pose3 P3 = pose3_of(pose);
    kr_fill_u32<<<divup(px, 256), 256, 0, h->stream>>>(h->depth.as<unsigned int>(), px, 0x7f800000u);
    B2_CUDA(cudaMemsetAsync(h->big_count.p, 0, 16, h->stream));       // [0] block queue, [1] warp queue, [2] splat queue
    unsigned int* counts = h->big_count.as<unsigned int>();
    kr_raster_small<<<divup(h->mesh_nf, 128), 128, 0, h->stream>>>(h->mesh_v.as<float>(), h->mesh_f.as<unsigned int>(), h->mesh_nf, P3, cam,
                                                                   h->prm.min_occlusion_depth, h->prm.max_occlusion_depth, h->depth.as<unsigned int>(),
                                                                   h->big_list.as<unsigned int>(), counts, h->warp_list.as<unsigned int>(), counts + 1);
    kr_raster_warp<<<h->sms * 8, 256, 0, h->stream>>>(h->mesh_v.as<float>(), h->mesh_f.as<unsigned int>(), P3, cam, h->prm.min_occlusion_depth,
                                                      h->prm.max_occlusion_depth, h->depth.as<unsigned int>(), h->warp_list.as<unsigned int>(), counts + 1);
    kr_raster_big<<<h->sms * 4, 256, 0, h->stream>>>(h->mesh_v.as<float>(), h->mesh_f.as<unsigned int>(), P3, cam, h->prm.min_occlusion_depth,
                                                     h->prm.max_occlusion_depth, h->depth.as<unsigned int>(), h->big_list.as<unsigned int>(),
                                                     h->big_count.as<unsigned int>());
    const bool mask = h->prm.mask_occlusion_boundaries != 0 && h->mesh_ne > 0;
    kr_depth_background<<<divup(px, 256), 256, 0, h->stream>>>(h->depth.as<float>(), mask ? h->depth_masked.as<float>() : nullptr, px);
    h->launches += 5;
    if (mask) {
      // image position = global_T_image.translation() = inv(q) * (-t)   (sophus se3.hpp:208-211)
      Pose inv; inv.q[0] = -pose.q[0]; inv.q[1] = -pose.q[1]; inv.q[2] = -pose.q[2]; inv.q[3] = pose.q[3];
      Pose neg; neg.t[0] = pose.t[0] * -1.f; neg.t[1] = pose.t[1] * -1.f; neg.t[2] = pose.t[2] * -1.f;
      const Pose ip = pose_mul(inv, neg);     // ip.t = 0 + inv.q (x) (-t)
      const unsigned int splat_cap = 4u << 20;                     // 4M splats (80 MB); beyond that the edge threads draw themselves
      B2_TRY(h->splat_queue.ensure((size_t)splat_cap * sizeof(EdgeSplat)));
      kr_mask_edges<<<divup(h->mesh_ne, 128), 128, 0, h->stream>>>(h->mesh_edges.as<MeshEdgeDev>(), h->mesh_ne, h->mesh_v.as<float>(), h->mesh_fn.as<float>(),
                                                                   P3, ip.t[0], ip.t[1], ip.t[2], cam, h->prm.splat_radius, h->depth.as<float>(),
                                                                   h->depth_masked.as<float>(), h->splat_queue.as<EdgeSplat>(), counts + 2, splat_cap);
      kr_draw_splats<<<h->sms * 8, 256, 0, h->stream>>>(h->splat_queue.as<EdgeSplat>(), counts + 2, splat_cap, cam.w, h->depth.as<float>(),
                                                        h->depth_masked.as<float>());
      h->launches += 2;
      *out = h->depth_masked.as<float>();
    } else {
      *out = h->depth.as<float>();
    }
    return B2_OK;
  }
  if (h->nsplats == 0) { *out = nullptr; return B2_OK; }
  const size_t px = (size_t)cam.w * cam.h;
  B2_TRY(h->depth.ensure(px * 4));
  kr_fill_u32<<<divup(px, 256), 256, 0, h->stream>>>(h->depth.as<unsigned int>(), px, 0x7f800000u);
  kr_splat_depth<<<divup(h->nsplats, 256), 256, 0, h->stream>>>(h->splats.as<float>(), h->nsplats, pose3_of(pose), cam, h->prm.splat_radius,
                                                                h->depth.as<unsigned int>());
  h->launches += 2;
  *out = h->depth.as<float>();
  return B2_OK;
}

static int ensure_obs_set(ObsSet* o, size_t cap, size_t npoints, bool with_jac) {
  cap = std::max<size_t>(cap, 1);
  B2_TRY(o->idx.ensure(cap * 4)); B2_TRY(o->x.ensure(cap * 4)); B2_TRY(o->y.ensure(cap * 4)); B2_TRY(o->s.ensure(cap * 4));
  B2_TRY(o->nb.ensure(cap)); B2_TRY(o->inten.ensure(cap * 4));
  if (with_jac) { B2_TRY(o->jK.ensure(cap * 4 * kMaxIntrinsics)); B2_TRY(o->jP.ensure(cap * 24)); B2_TRY(o->jR.ensure(cap * 24)); }
  B2_TRY(o->slot.ensure(std::max<size_t>(npoints, 1) * 4));
  return B2_OK;
}

// AppendObservationsForImage (visibility_estimator.cc:61-91) or, with `lists`, AppendObservationsForIndexedPointsVisibleInImage
// (:140-168): observation sets for every point scale of one image under `state`, plus DetermineIfAllNeighborsAreObserved.
static int observations_for_image(b2_reg* h, const StateB& st, int im_id, int border, const std::vector<ObsSet>* lists, std::vector<ObsSet>* out) {
  const ImageB& im = h->images[im_id];
  const IntrinsicsB& in = st.intr[im.intrinsics_id];
  const Pose& pose = st.poses[im_id];
  const int best = in.best_available(std::max(h->prm.min_occlusion_check_image_scale, h->current_image_scale));
  const float* depth = nullptr;
  if (!lists) B2_TRY(render_depth(h, im, in, pose, best, &depth));
  Levels L = levels_of(h, im, in);
  for (int l = 0; l < L.nlevels; ++l) L.cam[l] = in.models[l];
  out->resize(h->pts.size());
  bool had_many = false, stopped = false;
  for (int ps = (int)h->pts.size() - 1; ps >= 0; --ps) {
    const ScaleB& P = h->pts[ps];
    ObsSet& o = (*out)[ps];
    const size_t cand = lists ? (*lists)[ps].count : P.n;
    B2_TRY(ensure_obs_set(&o, cand, P.n, !lists));
    B2_CUDA(cudaMemsetAsync(o.slot.p, 0xFF, std::max<size_t>(P.n, 1) * 4, h->stream));
    o.count = 0;
    if (stopped || cand == 0) continue;
    B2_TRY(h->flags.ensure(cand * 4)); B2_TRY(h->offs.ensure(cand * 4));
    B2_TRY(h->cx.ensure(cand * 4)); B2_TRY(h->cy.ensure(cand * 4)); B2_TRY(h->cs.ensure(cand * 4));
    VisParams V;
    V.P = pose3_of(pose); V.cam = in.model(best); V.image_scale = best; V.depth = lists ? nullptr : depth;
    V.occlusion_threshold = h->prm.occlusion_depth_threshold; V.point_radius = P.radius; V.max_valid_intensity = h->prm.maximum_valid_intensity;
    V.border = border; V.check_masks = lists ? 0 : 1; V.current_image_scale = h->current_image_scale; V.image_scale_count = h->image_scale_count;
    const unsigned int* list = lists ? (*lists)[ps].idx.as<unsigned int>() : nullptr;
    kr_visibility<<<divup(cand, 256), 256, 0, h->stream>>>(P.xyz.as<float>(), list, cand, V, L, h->flags.as<unsigned int>(), h->cx.as<float>(),
                                                           h->cy.as<float>(), h->cs.as<float>());
    size_t tmp = 0;
    B2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, h->flags.as<unsigned int>(), h->offs.as<unsigned int>(), (long long)cand, h->stream));
    B2_TRY(h->cub_tmp.ensure(tmp));
    B2_CUDA(cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, tmp, h->flags.as<unsigned int>(), h->offs.as<unsigned int>(), (long long)cand, h->stream));
    unsigned int* tail = h->pin.as<unsigned int>();
    B2_CUDA(cudaMemcpyAsync(&tail[0], h->offs.as<unsigned int>() + (cand - 1), 4, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaMemcpyAsync(&tail[1], h->flags.as<unsigned int>() + (cand - 1), 4, cudaMemcpyDeviceToHost, h->stream));
    kr_compact<<<divup(cand, 256), 256, 0, h->stream>>>(h->flags.as<unsigned int>(), h->offs.as<unsigned int>(), list, cand, h->cx.as<float>(),
                                                        h->cy.as<float>(), h->cs.as<float>(), o.idx.as<unsigned int>(), o.x.as<float>(),
                                                        o.y.as<float>(), o.s.as<float>(), o.slot.as<int>());
    h->launches += 2;
    B2_CUDA(cudaStreamSynchronize(h->stream));
    o.count = (size_t)tail[0] + tail[1];
    if (o.count > 0) {
      kr_neighbors_observed<<<divup(o.count, 256), 256, 0, h->stream>>>(o.idx.as<unsigned int>(), o.count, P.nbr.as<unsigned int>(), K(h),
                                                                       o.slot.as<int>(), o.nb.as<unsigned char>());
      ++h->launches;
    }
    if (o.count > 100) had_many = true;                    // kManyObservationsCount (visibility_estimator.cc:44,75-90)
    else if (o.count == 0 && had_many) stopped = true;
  }
  B2_CUDA(cudaGetLastError());
  return B2_OK;
}

static ResidualArgs residual_args(const b2_reg* h, int ps, const ObsSet& o) {
  const ScaleB& P = h->pts[ps];
  ResidualArgs A;
  A.count = o.count; A.idx = o.idx.as<unsigned int>(); A.nb = o.nb.as<unsigned char>(); A.nbr = P.nbr.as<unsigned int>(); A.K = K(h);
  A.slot = o.slot.as<int>(); A.inten = o.inten.as<float>(); A.fixed_desc = P.fixed_desc.as<float>(); A.var_desc = P.var_desc.as<float>();
  A.obs_count = P.obs_count.as<int>(); A.robust.type = h->prm.robust_weighting_type; A.robust.p = h->prm.robust_weighting_parameter;
  A.fixed_w = h->prm.fixed_residuals_weight; A.var_w = h->prm.variable_residuals_weight;
  return A;
}

struct Sums { double fixed_sum = 0, var_sum = 0, nf = 0, nv = 0; };
static int allreduce_sums(b2_reg* h, Sums* s) {
  double v[4] = {s->fixed_sum, s->nf, s->var_sum, s->nv};
  B2_TRY(allreduce_host(h, v, 4));
  s->fixed_sum = v[0]; s->nf = v[1]; s->var_sum = v[2]; s->nv = v[3];
  return B2_OK;
}

// Problem::ComputeCost (problem.cc:602-631), depth weight 0.
static double compute_cost(const b2_reg* h, const Sums& s) {
  const bool uf = h->prm.fixed_residuals_weight > 0, uv = h->prm.variable_residuals_weight > 0;
  double r = 0;
  if (uf && s.nf > 0) r += h->prm.fixed_residuals_weight * s.fixed_sum / s.nf;
  if (uv && s.nv > 0) r += h->prm.variable_residuals_weight * s.var_sum / s.nv;
  if ((!uf && !uv) || (s.nf == 0 && s.nv == 0)) r = std::numeric_limits<float>::infinity();
  return r;
}

// Residual sums of ONE image over its observation sets (kr_intensity + kr_residual_sums per point scale, one readback), added to *acc.
static int residual_sums_image(b2_reg* h, const StateB& st, int im, std::vector<ObsSet>& sets, Sums* acc) {
  const int grid = h->sms * 2;
  const size_t S = h->pts.size();
  B2_TRY(h->partials.ensure(sizeof(double) * 4 * grid));
  B2_TRY(h->results.ensure(sizeof(double) * 4 * std::max<size_t>(S, 1)));
  B2_TRY(h->pin.ensure(std::max<size_t>(sizeof(double) * 4 * S, 64)));
  B2_CUDA(cudaMemsetAsync(h->results.p, 0, sizeof(double) * 4 * std::max<size_t>(S, 1), h->stream));
  const Levels L = levels_of(h, h->images[im], st.intr[h->images[im].intrinsics_id]);
  bool any = false;
  for (size_t ps = 0; ps < S; ++ps) {
    ObsSet& o = sets[ps];
    if (o.count == 0) continue;
    any = true;
    kr_intensity<<<divup(o.count, 256), 256, 0, h->stream>>>(o.count, o.x.as<float>(), o.y.as<float>(), o.s.as<float>(), L, o.inten.as<float>());
    kr_residual_sums<<<grid, 256, 0, h->stream>>>(residual_args(h, (int)ps, o), h->partials.as<double>());
    kr_reduce_partials<<<4, 64, 0, h->stream>>>(h->partials.as<double>(), grid, 4, h->results.as<double>() + 4 * ps);
    h->launches += 3;
  }
  if (!any) return B2_OK;
  B2_CUDA(cudaMemcpyAsync(h->pin.p, h->results.p, sizeof(double) * 4 * S, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  B2_CUDA(cudaGetLastError());
  const double* r = h->pin.as<double>();
  for (size_t k = 0; k < S; ++k) { acc->fixed_sum += r[4 * k]; acc->nf += r[4 * k + 1]; acc->var_sum += r[4 * k + 2]; acc->nv += r[4 * k + 3]; }
  return B2_OK;
}

static StateB current_state(const b2_reg* h) {
  StateB s; s.intr = h->intr; s.rigs = h->rigs;
  for (const ImageB& im : h->images) s.poses.push_back(im.pose);
  return s;
}
static void set_current_state(b2_reg* h, const StateB& s) {
  h->intr = s.intr; h->rigs = s.rigs;
  for (size_t i = 0; i < h->images.size(); ++i) h->images[i].pose = s.poses[i];
}

// Rig context of an image under `st` (zero / not dependent for non-rig and reference images); *rv = its extrinsics variable block.
static RigDev rig_dev(const b2_reg* h, const StateB& st, int im, int* rv) {
  RigDev r; std::memset(&r, 0, sizeof(r)); *rv = -1;
  if (!is_dependent(h, im)) return r;
  const RigImagesB& ri = h->rig_images[h->images[im].rig_images_id];
  const int c = h->images[im].rig_camera_index;
  const Pose& T = st.rigs[ri.rig_id].image_T_rig[c];
  const Pose& G = st.poses[ri.image_ids[0]];
  r.dependent = 1;
  quat_matrix(T.q, r.Rr);
  for (int k = 0; k < 4; ++k) r.q[k] = G.q[k];
  for (int k = 0; k < 3; ++k) r.t[k] = G.t[k];
  *rv = rig_var(h, ri.rig_id) + 6 * (c - 1);
  return r;
}

// CreateDeltaState (intrinsics_and_pose_optimizer.cc:475-558).
static int delta_state(b2_reg* h, const StateB& base, const double* delta, StateB* out) {
  StateB& n = *out;
  n = base;
  for (size_t i = 0; i < n.intr.size(); ++i) {
    Cam& m = n.intr[i].models[0];
    float p[kMaxIntrinsics] = {0}; cam_get(m, p);
    for (int k = 0; k < n.intr[i].np(); ++k) p[k] += delta[intr_var(h, (int)i) + k];      // float += double (intrinsics.cc:72-74)
    cam_set(&m, m.type, m.w, m.h, p);
    n.intr[i].build_pyramid();
  }
  B2_TRY(compute_cutoffs(h, &n.intr));
  for (size_t r = 0; r < n.rigs.size(); ++r)                                   // Rig::Update (rig.cc:9-23)
    for (size_t c = 1; c < n.rigs[r].image_T_rig.size(); ++c)
      n.rigs[r].image_T_rig[c] = pose_mul(pose_exp(delta + rig_var(h, (int)r) + 6 * ((int)c - 1)), base.rigs[r].image_T_rig[c]);
  for (size_t i = 0; i < n.poses.size(); ++i)                                  // pose owners (image.cc:161; :517-534, 556-563)
    if (!is_dependent(h, (int)i)) n.poses[i] = pose_mul(pose_exp(delta + pose_var(h, (int)i)), base.poses[i]);
  for (size_t i = 0; i < n.poses.size(); ++i)                                  // dependent rig images from the UPDATED rig pose (:546-555)
    if (is_dependent(h, (int)i)) {
      const RigImagesB& ri = h->rig_images[h->images[i].rig_images_id];
      n.poses[i] = pose_mul(n.rigs[ri.rig_id].image_T_rig[h->images[i].rig_camera_index], n.poses[ri.image_ids[0]]);
    }
  return B2_OK;
}

// ComputeResidualForState with frozen visibility lists (intrinsics_and_pose_optimizer.cc:385-440).
static int residual_for_state(b2_reg* h, const StateB& st, double* cost) {
  Sums total;
  for (size_t im = 0; im < h->images.size(); ++im) {
    if (!owned(h, im)) continue;
    B2_TRY(observations_for_image(h, st, (int)im, 1, &h->obs[im], &h->trial));
    B2_TRY(residual_sums_image(h, st, (int)im, h->trial, &total));
  }
  B2_TRY(allreduce_sums(h, &total));
  *cost = compute_cost(h, total);
  return B2_OK;
}

static void launch_jacobians(b2_reg* h, int kni, ObsSet& o, const ScaleB& P, const Pose3& P3, const Levels& L, const RigDev& rig) {
  if (kni == 4)
    kr_jacobians<4><<<divup(o.count, 256), 256, 0, h->stream>>>(o.count, o.idx.as<unsigned int>(), o.x.as<float>(), o.y.as<float>(), o.s.as<float>(),
                                                                 P.xyz.as<float>(), P3, P.radius, L, o.inten.as<float>(), o.jK.as<float>(), o.jP.as<float>(),
                                                                 rig, o.jR.as<float>());
  else
    kr_jacobians<12><<<divup(o.count, 256), 256, 0, h->stream>>>(o.count, o.idx.as<unsigned int>(), o.x.as<float>(), o.y.as<float>(), o.s.as<float>(),
                                                                  P.xyz.as<float>(), P3, P.radius, L, o.inten.as<float>(), o.jK.as<float>(), o.jP.as<float>(),
                                                                  rig, o.jR.as<float>());
  ++h->launches;
}

// AccumulateHAndBAndResidualsForObservations over all images and point scales (intrinsics_and_pose_optimizer.cc:102-185).
static int accumulate_all(b2_reg* h, std::vector<double>* H, std::vector<double>* b, Sums* sums) {
  const int nv = nvars(h);
  const int grid = h->sms * 2, grid_wide = h->sms * 4;
  const size_t S = h->pts.size(), NI = h->images.size();
  constexpr int kAccW = 24 * 25 / 2 + 24 + 4;      // result stride: the widest local system (12 intrinsics + 6 rig extrinsics + 6 pose)
  H->assign((size_t)nv * nv, 0.0); b->assign(nv, 0.0);
  B2_TRY(h->partials.ensure(sizeof(double) * kAccW * grid_wide));
  B2_TRY(h->results.ensure(sizeof(double) * kAccW * NI * S));
  B2_TRY(h->pin.ensure(std::max<size_t>(sizeof(double) * kAccW * NI * S, 64)));
  B2_CUDA(cudaMemsetAsync(h->results.p, 0, sizeof(double) * kAccW * NI * S, h->stream));
  const StateB st = current_state(h);
  uint64_t evals = 0;
  float ms_j = 0.f, ms_a = 0.f;
  size_t nsets = 0;                          // (image, scale) sets launched; their kernels are queued back to back, timed by events read at the end
  size_t max_count = 1;                      // the pre-pass buffers are sized once, so no set waits for a re-allocation
  for (size_t im = 0; im < NI; ++im) if (owned(h, im)) for (size_t ps = 0; ps < S; ++ps) max_count = std::max(max_count, h->obs[im][ps].count);
  if (K(h) == 5 && h->k12_mode != 1) {
    B2_TRY(h->w_nj.ensure(max_count * 5 * 4)); B2_TRY(h->w_ws.ensure(max_count * 8)); B2_TRY(h->w_wr.ensure(max_count * 5 * 8));
    B2_TRY(h->w_part.ensure(sizeof(double) * 4 * h->sms * 4));
  }
  for (size_t im = 0; im < NI; ++im) {
    if (!owned(h, im)) continue;
    const ImageB& I = h->images[im];
    const Levels L = levels_of(h, I, st.intr[I.intrinsics_id]);
    const Pose3 P3 = pose3_of(I.pose);
    int rv; const RigDev rig = rig_dev(h, st, (int)im, &rv);
    for (size_t ps = 0; ps < S; ++ps) {
      ObsSet& o = h->obs[im][ps];
      if (o.count == 0) continue;
      evals += o.count;
      const int np = st.intr[I.intrinsics_id].kni();      // width of the local system's intrinsics block (zero columns beyond the model's np)
      while (h->ev_pool.size() < 3 * (nsets + 1)) { cudaEvent_t e; B2_CUDA(cudaEventCreate(&e)); h->ev_pool.push_back(e); }
      cudaEvent_t* ev = &h->ev_pool[3 * nsets];
      ++nsets;
      cudaEventRecord(ev[0], h->stream);
      launch_jacobians(h, np, o, h->pts[ps], P3, L, rig);
      cudaEventRecord(ev[1], h->stream);
      const ResidualArgs A = residual_args(h, (int)ps, o);
      const float *pK = o.jK.as<float>(), *pP = o.jP.as<float>(), *pR = o.jR.as<float>();
      double* part = h->partials.as<double>();
      const int lv = np + 6 + (rig.dependent ? 6 : 0), nout = lv * (lv + 1) / 2 + lv + 4;
      const bool pin = np == 4 && !rig.dependent;
      const bool blocks = K(h) == 5 && (pin ? (h->k12_mode == 0 || h->k12_mode == 3) : h->k12_mode != 1);   // pre-pass + second kernel
      const int gridb = h->sms * (pin ? (h->k12_mode == 0 ? 2 : BlockCfg<4, false>::CTAS) : (np == 12 && rig.dependent) ? BlockCfg<12, true>::CTAS : 2);   // persistent CTAs of kr_accumulate_blocks
      if (pin && !blocks) {
        if (h->k12_mode != 1) kr_accumulate<false><<<grid, 128, 0, h->stream>>>(A, pK, pP, part);
        else kr_accumulate<true><<<grid, 128, 0, h->stream>>>(A, pK, pP, part);
      } else if (blocks) {
        const int gridw = h->sms * 4;
        kr_residual_weights<5><<<gridw, 256, 0, h->stream>>>(A, h->w_nj.as<int>(), h->w_ws.as<double>(), h->w_wr.as<double>(), h->w_part.as<double>());
        const int* pn = h->w_nj.as<int>(); const double *pws = h->w_ws.as<double>(), *pwr = h->w_wr.as<double>();
        if (pin && h->k12_mode == 0) kr_accumulate_weighted<5><<<gridb, 128, 0, h->stream>>>(o.count, pn, pws, pwr, pK, pP, part);
        else if (pin) kr_accumulate_blocks<4, false, 5><<<gridb, BlockCfg<4, false>::T, BlockCfg<4, false>::smem(5), h->stream>>>(o.count, pn, pws, pwr, pK, pP, pR, part);
        else if (np == 4) kr_accumulate_blocks<4, true, 5><<<gridb, BlockCfg<4, true>::T, BlockCfg<4, true>::smem(5), h->stream>>>(o.count, pn, pws, pwr, pK, pP, pR, part);
        else if (!rig.dependent) kr_accumulate_blocks<12, false, 5><<<gridb, BlockCfg<12, false>::T, BlockCfg<12, false>::smem(5), h->stream>>>(o.count, pn, pws, pwr, pK, pP, pR, part);
        else kr_accumulate_blocks<12, true, 5><<<gridb, BlockCfg<12, true>::T, BlockCfg<12, true>::smem(5), h->stream>>>(o.count, pn, pws, pwr, pK, pP, pR, part);
        ++h->launches;
      }
      else if (np == 4) kr_accumulate_wide<4, true><<<grid_wide, 128, 0, h->stream>>>(A, pK, pP, pR, part);
      else if (!rig.dependent) kr_accumulate_wide<12, false><<<grid_wide, 128, 0, h->stream>>>(A, pK, pP, pR, part);
      else kr_accumulate_wide<12, true><<<grid_wide, 128, 0, h->stream>>>(A, pK, pP, pR, part);
      cudaEventRecord(ev[2], h->stream);
      double* res = h->results.as<double>() + kAccW * (im * S + ps);
      if (blocks) {
        kr_reduce_partials<<<nout - 4, 64, 0, h->stream>>>(part, gridb, nout - 4, res);
        kr_reduce_partials<<<4, 64, 0, h->stream>>>(h->w_part.as<double>(), h->sms * 4, 4, res + (nout - 4));
        ++h->launches;
      } else {
        kr_reduce_partials<<<nout, 64, 0, h->stream>>>(part, pin ? grid : grid_wide, nout, res);
      }
      h->launches += 2;
    }
  }
  B2_CUDA(cudaMemcpyAsync(h->pin.p, h->results.p, sizeof(double) * kAccW * NI * S, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  B2_CUDA(cudaGetLastError());
  for (size_t k = 0; k < nsets; ++k) {
    float a = 0, c = 0; cudaEventElapsedTime(&a, h->ev_pool[3 * k], h->ev_pool[3 * k + 1]); cudaEventElapsedTime(&c, h->ev_pool[3 * k + 1], h->ev_pool[3 * k + 2]);
    ms_j += a; ms_a += c;
  }
  const double* r = h->pin.as<double>();
  *sums = Sums();
  for (size_t im = 0; im < NI; ++im) {
    if (!owned(h, im)) continue;
    const int iv = intr_var(h, h->images[im].intrinsics_id), pv = pose_var(h, (int)im);
    int rv; const bool dep = rig_dev(h, st, (int)im, &rv).dependent != 0;
    const int ni = st.intr[h->images[im].intrinsics_id].kni(), ni_real = st.intr[h->images[im].intrinsics_id].np();
    const int nr = dep ? 6 : 0, lv = ni + nr + 6, lh = lv * (lv + 1) / 2;
    auto g = [&](int l) { return l < ni ? (l < ni_real ? iv + l : -1) : l < ni + nr ? rv + (l - ni) : pv + (l - ni - nr); };   // -1: padding column
    for (size_t ps = 0; ps < S; ++ps) {
      const double* v = r + kAccW * (im * S + ps);
      int e = 0;
      for (int c = 0; c < lv; ++c) for (int rr = 0; rr <= c; ++rr) {
        if (g(c) >= 0 && g(rr) >= 0) (*H)[(size_t)std::max(g(c), g(rr)) * nv + std::min(g(c), g(rr))] += v[e];
        ++e;
      }
      for (int k = 0; k < lv; ++k) if (g(k) >= 0) (*b)[g(k)] += v[lh + k];
      sums->fixed_sum += v[lh + lv]; sums->nf += v[lh + lv + 1]; sums->var_sum += v[lh + lv + 2]; sums->nv += v[lh + lv + 3];
    }
  }
  if (h->world > 1) {   // data-parallel exchange: one sum-allreduce of [H | b | sums | evals] (nv^2 + nv + 5 doubles)
    std::vector<double> pack(H->begin(), H->end());
    pack.insert(pack.end(), b->begin(), b->end());
    const double tail[5] = {sums->fixed_sum, sums->nf, sums->var_sum, sums->nv, (double)evals};
    pack.insert(pack.end(), tail, tail + 5);
    B2_TRY(allreduce_host(h, pack.data(), pack.size()));
    std::copy(pack.begin(), pack.begin() + (size_t)nv * nv, H->begin());
    std::copy(pack.begin() + (size_t)nv * nv, pack.begin() + (size_t)nv * nv + nv, b->begin());
    const double* t = pack.data() + (size_t)nv * nv + nv;
    sums->fixed_sum = t[0]; sums->nf = t[1]; sums->var_sum = t[2]; sums->nv = t[3]; evals = (uint64_t)t[4];
  }
  for (int c = 0; c < nv; ++c) for (int rr = 0; rr < c; ++rr) (*H)[(size_t)rr * nv + c] = (*H)[(size_t)c * nv + rr];   // mirror the Upper view
  h->stats.residual_evaluations = evals; h->stats.ms_jacobian_kernel = ms_j; h->stats.ms_accumulate_kernel = ms_a;
  return B2_OK;
}

static int create_observations(b2_reg* h, int border) {
  const StateB st = current_state(h);
  h->obs.resize(h->images.size());
  uint64_t total = 0;
  for (size_t im = 0; im < h->images.size(); ++im) {
    if (!owned(h, im)) { h->obs[im].resize(h->pts.size()); for (ObsSet& o : h->obs[im]) o.count = 0; continue; }
    B2_TRY(observations_for_image(h, st, (int)im, border, nullptr, &h->obs[im]));
    for (const ObsSet& o : h->obs[im]) total += o.count;
  }
  double t = (double)total;
  B2_TRY(allreduce_host(h, &t, 1));
  h->stats.observations = (uint64_t)t;
  return B2_OK;
}

static int color_update(b2_reg* h) {
  const StateB st = current_state(h);
  for (size_t ps = 0; ps < h->pts.size(); ++ps) {
    ScaleB& P = h->pts[ps];
    if (P.n == 0) continue;
    B2_CUDA(cudaMemsetAsync(P.obs_count.p, 0, P.n * 4, h->stream));
    B2_CUDA(cudaMemsetAsync(P.var_desc.p, 0, P.n * K(h) * 4, h->stream));
    for (size_t im = 0; im < h->images.size(); ++im) {
      ObsSet& o = h->obs[im][ps];
      if (o.count == 0) continue;
      const Levels L = levels_of(h, h->images[im], st.intr[h->images[im].intrinsics_id]);
      kr_intensity<<<divup(o.count, 256), 256, 0, h->stream>>>(o.count, o.x.as<float>(), o.y.as<float>(), o.s.as<float>(), L, o.inten.as<float>());
      kr_color_accumulate<<<divup(o.count, 256), 256, 0, h->stream>>>(o.count, o.idx.as<unsigned int>(), o.nb.as<unsigned char>(), P.nbr.as<unsigned int>(),
                                                                     K(h), o.slot.as<int>(), o.inten.as<float>(), P.var_desc.as<float>(), P.obs_count.as<int>());
      h->launches += 2;
    }
    if (h->world > 1) {   // descriptor sums and observation counts over ALL images: 5 floats + 1 int per point
      B2_TRY(b2_comm_allreduce(h->comm, P.var_desc.p, P.n * K(h), B2_F32, (void*)h->stream));
      B2_TRY(b2_comm_allreduce(h->comm, P.obs_count.p, P.n, B2_I32, (void*)h->stream));
    }
    kr_color_mean<<<divup(P.n, 256), 256, 0, h->stream>>>(P.n, K(h), P.var_desc.as<float>(), P.obs_count.as<int>());
    ++h->launches;
  }
  B2_CUDA(cudaGetLastError());
  return B2_OK;
}

static int current_cost(b2_reg* h, double* cost, Sums* s) {
  const StateB st = current_state(h);
  *s = Sums();
  for (size_t im = 0; im < h->images.size(); ++im) if (owned(h, im)) B2_TRY(residual_sums_image(h, st, (int)im, h->obs[im], s));
  B2_TRY(allreduce_sums(h, s));
  if (s->nf == 0 && s->nv == 0) { *cost = std::numeric_limits<double>::infinity(); return B2_OK; }   // cost_calculator.cc:88-93
  *cost = compute_cost(h, *s);
  return B2_OK;
}

// IntrinsicsAndPoseOptimizer::Apply (intrinsics_and_pose_optimizer.cc:48-259).
static int apply_lm(b2_reg* h, float* lambda, float* max_change, int* applied, int* tries_out) {
  std::vector<double> H, b; Sums s;
  B2_TRY(accumulate_all(h, &H, &b, &s));
  const int nv = nvars(h);
  const double initial = compute_cost(h, s);
  const StateB base = current_state(h);
  *applied = 0;
  int tries = 0;
  std::vector<double> x(nv), neg(nv);
  for (int lm = 0; lm < 10; ++lm) {
    ++tries;
    std::vector<double> HL = H;
    for (int i = 0; i < nv; ++i) HL[(size_t)i * nv + i] *= (1 + (*lambda));
    sym_solve(HL, nv, b.data(), x.data());
    for (int i = 0; i < nv; ++i) neg[i] = -1 * x[i];
    StateB ns; B2_TRY(delta_state(h, base, neg.data(), &ns));
    double nr = 0;
    B2_TRY(residual_for_state(h, ns, &nr));
    if (nr < initial || lm == 9) {      // kAlwaysApplyLastUpdate (:199,239-240)
      double mx = -std::numeric_limits<double>::infinity();
      for (double v : x) mx = std::max(mx, v);
      *max_change = (float)mx;         // signed maxCoeff (:245)
      set_current_state(h, ns);
      *lambda = 0.5f * (*lambda);
      *applied = 1;
      break;
    } else {
      *lambda = 2.f * (*lambda);
    }
  }
  if (tries_out) *tries_out = tries;
  return B2_OK;
}

}  // namespace b2

#define REG_ENTER(h)                                                            \
  if (!(h)) return set_error(B2_ERR_ARG, "null handle");                         \
  B2_CUDA(cudaSetDevice((h)->device));

extern "C" {

void b2_reg_default_params(b2_reg_params* p) {
  if (!p) return;
  p->point_neighbor_count = 5; p->fixed_residuals_weight = 1.f; p->variable_residuals_weight = 1.f;
  p->robust_weighting_type = 1; p->robust_weighting_parameter = (float)(30 * std::sqrt(5.0) / std::sqrt(2.0));
  p->maximum_valid_intensity = 252; p->occlusion_depth_threshold = 0.01f; p->min_occlusion_check_image_scale = 0;
  p->max_initial_image_area_in_pixels = 200 * 160; p->splat_radius = 0.03f; p->image_scale_count_override = 0; p->device = -1;
  p->min_occlusion_depth = 0.05f; p->max_occlusion_depth = 100.f; p->mask_occlusion_boundaries = 1;
}

int b2_reg_create(const b2_reg_params* p, b2_reg** out) {
  if (!out) return set_error(B2_ERR_ARG, "out is null");
  *out = nullptr;
  std::unique_ptr<b2_reg> h(new b2_reg());
  if (p) h->prm = *p; else b2_reg_default_params(&h->prm);
  if (h->prm.point_neighbor_count < 1 || h->prm.point_neighbor_count > kMaxNbr) return set_error(B2_ERR_ARG, "point_neighbor_count must be in [1,%d]", kMaxNbr);
  B2_TRY(select_device(h->prm.device, &h->device, &h->sms));
  B2_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&h->ev0, &h->ev1, &h->evj0, &h->evj1, &h->eva0, &h->eva1}) B2_CUDA(cudaEventCreate(e));
  B2_TRY(h->pin.ensure(4096));
  std::memset(&h->stats, 0, sizeof(h->stats));
  if (const char* e = std::getenv("B2_K12")) h->k12_mode = std::strcmp(e, "f32") == 0 ? 1 : std::strcmp(e, "thread") == 0 ? 2 : std::strcmp(e, "blocks") == 0 ? 3 : 0;   // A/B switch, see kr_accumulate
  *out = h.release();
  return B2_OK;
}

int b2_reg_destroy(b2_reg* h) {
  if (!h) return B2_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  auto free_set = [](ObsSet& o) { for (DevBuf* b : {&o.idx, &o.x, &o.y, &o.s, &o.nb, &o.inten, &o.jK, &o.jP, &o.jR, &o.slot}) b->release(); };
  for (auto& v : h->obs) for (auto& o : v) free_set(o);
  for (auto& o : h->trial) free_set(o);
  for (auto& im : h->images) { for (auto& b : im.img) b.release(); for (auto& b : im.mask) b.release(); im.given_depth.release(); }
  for (auto& P : h->pts) for (DevBuf* b : {&P.xyz, &P.nbr, &P.fixed_desc, &P.var_desc, &P.obs_count}) b->release();
  for (DevBuf* b : {&h->mesh_v, &h->mesh_f, &h->mesh_fn, &h->mesh_edges, &h->big_list, &h->big_count, &h->warp_list, &h->splat_queue, &h->depth_masked}) b->release();
  for (DevBuf* b : {&h->splats, &h->flags, &h->offs, &h->cx, &h->cy, &h->cs, &h->cub_tmp, &h->depth, &h->partials, &h->results}) b->release();
  for (DevBuf* b : {&h->cut_cams, &h->cut_first, &h->cut_starts, &h->cut_points, &h->cut_out, &h->xchg, &h->w_nj, &h->w_ws, &h->w_wr, &h->w_part}) b->release();
  for (auto& kv : h->cam_masks) for (auto& b : kv.second) b.release();
  h->pin.release();
  for (cudaEvent_t e : {h->ev0, h->ev1, h->evj0, h->evj1, h->eva0, h->eva1}) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  cudaStreamDestroy(h->stream);
  delete h;
  return B2_OK;
}

int b2_reg_add_intrinsics(b2_reg* h, int camera_model, int width, int height, const float* params, int num_params, int* out_id) {
  REG_ENTER(h);
  const int np = cam_param_count(camera_model);
  if (np < 0) return set_error(B2_ERR_ARG, "camera model %d is not a camera::CameraBase::Type (0..14)", camera_model);
  if (!params || num_params != np || width < 2 || height < 2) return set_error(B2_ERR_ARG, "camera model %d needs %d parameters and a size >= 2x2", camera_model, np);
  if (h->initialized) return set_error(B2_ERR_STATE, "add_intrinsics after initialize");
  IntrinsicsB in; in.models.resize(1); cam_set(&in.models[0], camera_model, width, height, params);
  h->intr.push_back(in);
  if (out_id) *out_id = (int)h->intr.size() - 1;
  return B2_OK;
}

int b2_reg_add_image(b2_reg* h, int intrinsics_id, const uint8_t* gray, const uint8_t* mask, const float T[7], int* out_id) {
  REG_ENTER(h);
  const bool mine = owned(h, h->images.size());
  if (intrinsics_id < 0 || intrinsics_id >= (int)h->intr.size() || (mine && !gray) || !T) return set_error(B2_ERR_ARG, "bad argument");
  if (h->initialized) return set_error(B2_ERR_STATE, "add_image after initialize");
  ImageB im; im.intrinsics_id = intrinsics_id;
  for (int k = 0; k < 4; ++k) im.pose.q[k] = T[k];
  for (int k = 0; k < 3; ++k) im.pose.t[k] = T[4 + k];
  const Cam& c = h->intr[intrinsics_id].models[0];
  const size_t px = (size_t)c.w * c.h;
  if (!mine) { im.has_mask = mask != nullptr; h->images.push_back(std::move(im)); if (out_id) *out_id = (int)h->images.size() - 1; return B2_OK; }
  im.img.resize(1); B2_TRY(im.img[0].ensure(px));
  B2_CUDA(cudaMemcpyAsync(im.img[0].p, gray, px, cudaMemcpyHostToDevice, h->stream));
  if (mask) { im.has_mask = true; im.mask.resize(1); B2_TRY(im.mask[0].ensure(px)); B2_CUDA(cudaMemcpyAsync(im.mask[0].p, mask, px, cudaMemcpyHostToDevice, h->stream)); }
  B2_CUDA(cudaStreamSynchronize(h->stream));
  h->images.push_back(std::move(im));
  if (out_id) *out_id = (int)h->images.size() - 1;
  return B2_OK;
}

int b2_reg_add_rig(b2_reg* h, int num_cameras, const float* image_T_rig, int* out_id) {
  REG_ENTER(h);
  if (num_cameras < 2 || !image_T_rig) return set_error(B2_ERR_ARG, "a rig needs at least two cameras (single cameras get no rig, rig.cc:31-34)");
  RigB r; r.image_T_rig.resize(num_cameras);
  for (int c = 0; c < num_cameras; ++c) {
    for (int k = 0; k < 4; ++k) r.image_T_rig[c].q[k] = image_T_rig[7 * c + k];
    for (int k = 0; k < 3; ++k) r.image_T_rig[c].t[k] = image_T_rig[7 * c + 4 + k];
  }
  h->rigs.push_back(r);
  if (out_id) *out_id = (int)h->rigs.size() - 1;
  return B2_OK;
}

int b2_reg_add_rig_images(b2_reg* h, int rig_id, const int32_t* image_ids, int* out_id) {
  REG_ENTER(h);
  if (rig_id < 0 || rig_id >= (int)h->rigs.size() || !image_ids) return set_error(B2_ERR_ARG, "bad rig id");
  const RigB& rig = h->rigs[rig_id];
  RigImagesB ri; ri.rig_id = rig_id;
  for (size_t c = 0; c < rig.image_T_rig.size(); ++c) {
    const int id = image_ids[c];
    if (id < 0 || id >= (int)h->images.size() || h->images[id].rig_images_id >= 0)
      return set_error(B2_ERR_ARG, "rig image set needs one registered, not yet assigned image per camera (camera %d)", (int)c);
    for (int prev : ri.image_ids) if (prev == id) return set_error(B2_ERR_ARG, "image %d listed twice", id);
    ri.image_ids.push_back(id);
  }
  h->rig_images.push_back(ri);
  const int rid = (int)h->rig_images.size() - 1;
  for (size_t c = 0; c < ri.image_ids.size(); ++c) {
    ImageB& im = h->images[ri.image_ids[c]];
    im.rig_images_id = rid; im.rig_camera_index = (int)c;
    if (c > 0) im.pose = pose_mul(rig.image_T_rig[c], h->images[ri.image_ids[0]].pose);   // as AssignRigs leaves them (rig.cc:216-250)
  }
  if (out_id) *out_id = rid;
  return B2_OK;
}

int b2_reg_get_rigs(b2_reg* h, float* out) {
  REG_ENTER(h);
  if (!out) return set_error(B2_ERR_ARG, "null");
  for (const RigB& r : h->rigs) for (const Pose& T : r.image_T_rig) { for (int k = 0; k < 4; ++k) out[k] = T.q[k]; for (int k = 0; k < 3; ++k) out[4 + k] = T.t[k]; out += 7; }
  return B2_OK;
}

int b2_reg_set_rigs(b2_reg* h, const float* in) {
  REG_ENTER(h);
  if (!in) return set_error(B2_ERR_ARG, "null");
  for (RigB& r : h->rigs) for (Pose& T : r.image_T_rig) { for (int k = 0; k < 4; ++k) T.q[k] = in[k]; for (int k = 0; k < 3; ++k) T.t[k] = in[4 + k]; in += 7; }
  return B2_OK;
}

int b2_reg_variable_index(b2_reg* h, int kind, int id, int* out) {
  REG_ENTER(h);
  if (!out) return set_error(B2_ERR_ARG, "null");
  if (kind == 0 && id >= 0 && id <= (int)h->intr.size()) *out = intr_var(h, id);
  else if (kind == 1 && id >= 0 && id <= (int)h->rigs.size()) *out = rig_var(h, id);
  else if (kind == 2 && id >= 0 && id <= (int)h->images.size()) *out = pose_var(h, id);
  else return set_error(B2_ERR_ARG, "bad kind / id");
  return B2_OK;
}

int b2_reg_set_camera_mask(b2_reg* h, int intrinsics_id, const uint8_t* mask) {
  REG_ENTER(h);
  if (intrinsics_id < 0 || intrinsics_id >= (int)h->intr.size() || !mask) return set_error(B2_ERR_ARG, "bad argument");
  if (h->initialized) return set_error(B2_ERR_STATE, "set_camera_mask after initialize");
  const Cam& c = h->intr[intrinsics_id].models[0];
  std::vector<DevBuf>& pyr = h->cam_masks[intrinsics_id];
  pyr.resize(1);
  B2_TRY(pyr[0].ensure((size_t)c.w * c.h));
  B2_CUDA(cudaMemcpy(pyr[0].p, mask, (size_t)c.w * c.h, cudaMemcpyHostToDevice));
  return B2_OK;
}

int b2_reg_image_owner(int image_id, int world_size) { return world_size > 0 && image_id >= 0 ? image_id % world_size : -1; }

int b2_reg_set_comm(b2_reg* h, b2_comm* comm) {
  REG_ENTER(h);
  if (!h->images.empty()) return set_error(B2_ERR_STATE, "b2_reg_set_comm must precede b2_reg_add_image (ownership decides which pixels are uploaded)");
  h->comm = comm; h->rank = 0; h->world = 1;
  if (comm) B2_TRY(b2_comm_info(comm, &h->rank, &h->world));
  return B2_OK;
}

// One pyramid level (image.cc:106-154). Masks: OR of the 2x2 block (the last row / column of an odd parent is ignored, :139-150).
// Images: cv::resize INTER_AREA with dsize given — integer 2x2 mean when both parents are even, the general area filter otherwise
// (OpenCV resize.cpp computeResizeAreaTab / resizeArea_, restated; pinned against cv2 in tests/test_oracle_reg.py and, on the device,
// tests/test_gpu_reg.py). The level is zero-padded by one row past its last pixel (Levels::iw).
static void area_taps(int ssize, int dsize, double scale, std::vector<AreaTapDev>* taps, std::vector<int>* ofs) {
  taps->clear(); ofs->assign(1, 0);
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale, cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1); sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) taps->push_back({sx1 - 1, (float)((sx1 - fsx1) / cell)});
    for (int sx = sx1; sx < sx2; ++sx) taps->push_back({sx, float(1.0 / cell)});
    if (fsx2 - sx2 > 1e-3) taps->push_back({sx2, (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)});
    ofs->push_back((int)taps->size());
  }
}
static int pyr_down_level(b2_reg* h, const DevBuf& src, int sw, int sh, DevBuf* dst, int dw, int dh, bool is_mask) {
  const size_t px = (size_t)dw * dh, pad = (size_t)dw + 2;
  B2_TRY(dst->ensure(px + pad));
  B2_CUDA(cudaMemsetAsync((unsigned char*)dst->p + px, 0, pad, h->stream));
  dim3 g(divup(dw, 256), dh);
  const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
  if (is_mask || (scale_x == 2.0 && scale_y == 2.0)) {
    kr_pyr_down<<<g, 256, 0, h->stream>>>(src.as<unsigned char>(), sw, dst->as<unsigned char>(), dw, dh, is_mask ? 1 : 0);
    ++h->launches;
    return B2_OK;
  }
  std::vector<AreaTapDev> xt, yt; std::vector<int> xo, yo;
  area_taps(sw, dw, scale_x, &xt, &xo); area_taps(sh, dh, scale_y, &yt, &yo);
  DevBuf tabs;       // [xt | yt | xofs | yofs]
  const size_t b_xt = xt.size() * sizeof(AreaTapDev), b_yt = yt.size() * sizeof(AreaTapDev), b_xo = xo.size() * 4, b_yo = yo.size() * 4;
  B2_TRY(tabs.ensure(b_xt + b_yt + b_xo + b_yo));
  unsigned char* t = (unsigned char*)tabs.p;
  cudaError_t e = cudaMemcpyAsync(t, xt.data(), b_xt, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(t + b_xt, yt.data(), b_yt, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(t + b_xt + b_yt, xo.data(), b_xo, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(t + b_xt + b_yt + b_xo, yo.data(), b_yo, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    kr_pyr_area<<<g, 256, 0, h->stream>>>(src.as<unsigned char>(), sw, dst->as<unsigned char>(), dw, dh, (const AreaTapDev*)t,
                                          (const int*)(t + b_xt + b_yt), (const AreaTapDev*)(t + b_xt), (const int*)(t + b_xt + b_yt + b_xo));
    ++h->launches;
    e = cudaStreamSynchronize(h->stream);      // the pageable host tables and `tabs` go out of scope
  }
  tabs.release();
  if (e != cudaSuccess) return set_error(B2_ERR_CUDA, "pyramid level failed: %s", cudaGetErrorString(e));
  return B2_OK;
}

int b2_reg_initialize(b2_reg* h, int* image_scale_count) {
  REG_ENTER(h);
  if (h->intr.empty() || h->images.empty()) return set_error(B2_ERR_STATE, "initialize needs at least one intrinsics and one image");
  auto scale_count = [&](const IntrinsicsB& in) {                       // Intrinsics::ComputeImageScaleCount (intrinsics.h:82-86)
    const int px = in.models[0].w * in.models[0].h;
    const double af = px * 1.0 / h->prm.max_initial_image_area_in_pixels;
    return std::max<int>(2, 1 + (int)std::ceil(std::log(af) / std::log(4)));
  };
  int count = 1;
  for (const IntrinsicsB& in : h->intr) count = std::max(count, scale_count(in));
  if (h->prm.image_scale_count_override > 0) count = h->prm.image_scale_count_override;
  if (count > kMaxLevels) return set_error(B2_ERR_ARG, "more than %d image scales", kMaxLevels);
  h->image_scale_count = count;
  for (IntrinsicsB& in : h->intr) {
    const int c = h->prm.image_scale_count_override > 0 ? count : scale_count(in);
    in.min_image_scale = count - c;
    in.models.resize(count - in.min_image_scale);
    in.build_pyramid();
  }
  begin_call(h);
  B2_TRY(compute_cutoffs(h, &h->intr));
  for (size_t ii = 0; ii < h->images.size(); ++ii) {
    ImageB& im = h->images[ii];
    if (!owned(h, ii)) continue;
    const IntrinsicsB& in = h->intr[im.intrinsics_id];
    const size_t levels = in.models.size();
    im.img.resize(levels); if (im.has_mask) im.mask.resize(levels);
    im.lw.assign(levels, 0); im.lh.assign(levels, 0);
    im.lw[0] = in.models[0].w; im.lh[0] = in.models[0].h;
    for (size_t l = 1; l < levels; ++l) {
      // image.cc:115-118: dsize = (int(0.5 cols), int(0.5 rows)); the camera level may be one pixel larger (see Levels::iw)
      im.lw[l] = (int)(0.5 * im.lw[l - 1]); im.lh[l] = (int)(0.5 * im.lh[l - 1]);
      if (im.lw[l] < 1 || im.lh[l] < 1) return set_error(B2_ERR_ARG, "image pyramid level %zu is empty (%dx%d parent)", l, im.lw[l - 1], im.lh[l - 1]);
      B2_TRY(pyr_down_level(h, im.img[l - 1], im.lw[l - 1], im.lh[l - 1], &im.img[l], im.lw[l], im.lh[l], false));
      if (im.has_mask) B2_TRY(pyr_down_level(h, im.mask[l - 1], im.lw[l - 1], im.lh[l - 1], &im.mask[l], im.lw[l], im.lh[l], true));
    }
  }
  for (auto& kv : h->cam_masks) {                                         // camera-mask pyramids (image.cc:62-72 + BuildMaskPyramid)
    const IntrinsicsB& in = h->intr[kv.first];
    std::vector<DevBuf>& pyr = kv.second;
    pyr.resize(in.models.size());
    int w = in.models[0].w, hh = in.models[0].h;
    for (size_t l = 1; l < in.models.size(); ++l) {
      const int dw = (int)(0.5 * w), dh = (int)(0.5 * hh);                // the sizes of the image pyramid, not of the camera pyramid
      if (dw < 1 || dh < 1) return set_error(B2_ERR_ARG, "camera mask pyramid level %zu is empty", l);
      B2_TRY(pyr_down_level(h, pyr[l - 1], w, hh, &pyr[l], dw, dh, true));
      w = dw; hh = dh;
    }
  }
  B2_CUDA(cudaGetLastError());
  end_call(h);
  h->initialized = true;
  if (image_scale_count) *image_scale_count = count;
  return B2_OK;
}

int b2_reg_add_point_scale(b2_reg* h, const float* xyz, size_t n, float radius, const uint64_t* nbr, const float* colors, int* out_scale) {
  REG_ENTER(h);
  if (n && (!xyz || !nbr || !colors)) return set_error(B2_ERR_ARG, "null argument");
  if (n >= (1ull << 31)) return set_error(B2_ERR_ARG, "point scales above 2^31 points are not supported");
  const int k = K(h);
  // 64-bit caller indices -> 32-bit device indices, range-checked; a few host threads (the 15.8 M-point cloud of the benchmark has
  // 7.9 * 10^7 of them) into an uninitialised buffer
  std::unique_ptr<unsigned int[]> nb32(new unsigned int[std::max<size_t>(n * k, 1)]);
  const int threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  std::vector<char> bad((size_t)threads, 0);
  host_parallel_ranges(n * k, threads, [&](int part, size_t i0, size_t i1) {
    unsigned int* out = nb32.get();
    bool any = false;
    for (size_t i = i0; i < i1; ++i) { any |= nbr[i] >= n; out[i] = (unsigned int)nbr[i]; }
    bad[(size_t)part] = any;
  });
  for (char b : bad) if (b) return set_error(B2_ERR_ARG, "neighbour index out of range");
  ScaleB P; P.n = n; P.radius = radius;
  const size_t m = std::max<size_t>(n, 1);
  B2_TRY(P.xyz.ensure(m * 12)); B2_TRY(P.nbr.ensure(m * k * 4)); B2_TRY(P.fixed_desc.ensure(m * k * 4)); B2_TRY(P.var_desc.ensure(m * k * 4));
  B2_TRY(P.obs_count.ensure(m * 4));
  if (n) {
    DevBuf col; B2_TRY(col.ensure(n * 4));
    B2_CUDA(cudaMemcpyAsync(P.xyz.p, xyz, n * 12, cudaMemcpyHostToDevice, h->stream));
    B2_CUDA(cudaMemcpyAsync(P.nbr.p, nb32.get(), n * k * 4, cudaMemcpyHostToDevice, h->stream));
    B2_CUDA(cudaMemcpyAsync(col.p, colors, n * 4, cudaMemcpyHostToDevice, h->stream));
    B2_CUDA(cudaMemsetAsync(P.fixed_desc.p, 0, n * k * 4, h->stream));
    B2_CUDA(cudaMemsetAsync(P.var_desc.p, 0, n * k * 4, h->stream));
    B2_CUDA(cudaMemsetAsync(P.obs_count.p, 0, n * 4, h->stream));
    if (h->prm.fixed_residuals_weight > 0)
      kr_fixed_descriptors<<<divup(n, 256), 256, 0, h->stream>>>(n, k, P.nbr.as<unsigned int>(), col.as<float>(), P.fixed_desc.as<float>(), P.obs_count.as<int>());
    B2_CUDA(cudaStreamSynchronize(h->stream));
    col.release();
  }
  h->pts.push_back(P);
  if (out_scale) *out_scale = (int)h->pts.size() - 1;
  return B2_OK;
}

// OcclusionGeometry::AddMesh with compute_edges (occlusion_geometry.cc:87-130) = ComputeEdgeNormalsList + FilterEdgeList (:466-645):
// face normals, half edges keyed by the sorted vertex pair, and per edge the two outermost faces (or "open").
int b2_reg_set_mesh(b2_reg* h, const float* vertices, size_t nv, const uint32_t* faces, size_t nf) {
  REG_ENTER(h);
  if ((nv && !vertices) || (nf && !faces)) return set_error(B2_ERR_ARG, "null argument");
  if (nv >= (1ull << 32) || nf >= (1ull << 32)) return set_error(B2_ERR_ARG, "mesh too large");
  for (size_t i = 0; i < 3 * nf; ++i) if (faces[i] >= nv) return set_error(B2_ERR_ARG, "face index out of range");
  h->mesh_nv = nv; h->mesh_nf = nf; h->mesh_ne = 0;
  if (nf == 0) return B2_OK;
  static_assert(sizeof(MeshEdgeHost) == sizeof(MeshEdgeDev), "host and device edge records must have one layout");
  std::vector<float> fn;
  std::vector<MeshEdgeHost> edges;
  build_mesh_edges(vertices, nv, faces, nf, &fn, &edges);
  h->mesh_ne = edges.size();
  B2_TRY(h->mesh_v.ensure(nv * 12)); B2_TRY(h->mesh_f.ensure(nf * 12)); B2_TRY(h->mesh_fn.ensure(nf * 12));
  B2_TRY(h->mesh_edges.ensure(std::max<size_t>(edges.size(), 1) * sizeof(MeshEdgeDev)));
  B2_CUDA(cudaMemcpy(h->mesh_v.p, vertices, nv * 12, cudaMemcpyHostToDevice));
  B2_CUDA(cudaMemcpy(h->mesh_f.p, faces, nf * 12, cudaMemcpyHostToDevice));
  B2_CUDA(cudaMemcpy(h->mesh_fn.p, fn.data(), nf * 12, cudaMemcpyHostToDevice));
  if (!edges.empty()) B2_CUDA(cudaMemcpy(h->mesh_edges.p, edges.data(), edges.size() * sizeof(MeshEdgeDev), cudaMemcpyHostToDevice));
  return B2_OK;
}

int b2_reg_set_splat_points(b2_reg* h, const float* xyz, size_t n) {
  REG_ENTER(h);
  h->nsplats = n;
  if (n) { if (!xyz) return set_error(B2_ERR_ARG, "null"); B2_TRY(h->splats.ensure(n * 12)); B2_CUDA(cudaMemcpy(h->splats.p, xyz, n * 12, cudaMemcpyHostToDevice)); }
  return B2_OK;
}

int b2_reg_set_depth_map(b2_reg* h, int image_id, int width, int height, const float* depth) {
  REG_ENTER(h);
  if (image_id < 0 || image_id >= (int)h->images.size() || !depth || width < 1 || height < 1) return set_error(B2_ERR_ARG, "bad argument");
  ImageB& im = h->images[image_id];
  if (!owned(h, (size_t)image_id)) return B2_OK;   // only the owner renders / taps this image's depth map
  B2_TRY(im.given_depth.ensure((size_t)width * height * 4));
  B2_CUDA(cudaMemcpy(im.given_depth.p, depth, (size_t)width * height * 4, cudaMemcpyHostToDevice));
  im.gd_w = width; im.gd_h = height; im.has_given_depth = true;
  return B2_OK;
}

int b2_reg_set_image_scale(b2_reg* h, int s) { REG_ENTER(h); h->current_image_scale = s; return B2_OK; }
int b2_reg_num_variables(b2_reg* h, int* nv) { REG_ENTER(h); if (nv) *nv = nvars(h); return B2_OK; }

int b2_reg_render_depth(b2_reg* h, int image_id, int* width, int* height, int* image_scale, float* out) {
  REG_ENTER(h);
  if (!h->initialized) return set_error(B2_ERR_STATE, "not initialized");
  if (image_id < 0 || image_id >= (int)h->images.size()) return set_error(B2_ERR_ARG, "bad image id");
  if (!owned(h, (size_t)image_id)) return set_error(B2_ERR_STATE, "image %d is owned by rank %d", image_id, image_id % h->world);
  const ImageB& im = h->images[image_id]; const IntrinsicsB& in = h->intr[im.intrinsics_id];
  const int best = in.best_available(std::max(h->prm.min_occlusion_check_image_scale, h->current_image_scale));
  const Cam& cam = in.model(best);
  if (width) *width = cam.w; if (height) *height = cam.h; if (image_scale) *image_scale = best;
  if (!out) return B2_OK;
  begin_call(h);
  const float* d = nullptr;
  B2_TRY(render_depth(h, im, in, im.pose, best, &d));
  const size_t px = (size_t)cam.w * cam.h;
  if (d) { B2_CUDA(cudaMemcpyAsync(out, d, px * 4, cudaMemcpyDeviceToHost, h->stream)); }
  end_call(h);
  if (!d) for (size_t i = 0; i < px; ++i) out[i] = std::numeric_limits<float>::infinity();
  return B2_OK;
}

// ComputeMinMaxPointRadius over all images (multi_scale_point_cloud.cc:126-184 as called from CreateMultiScalePointCloud, :232-255).
int b2_reg_min_max_point_radius(b2_reg* h, const float* xyz, size_t n, double min_scaling_factor, float* min_radius, float* max_radius) {
  REG_ENTER(h);
  if (!h->initialized) return set_error(B2_ERR_STATE, "not initialized");
  if (n && (!xyz || !min_radius || !max_radius)) return set_error(B2_ERR_ARG, "null argument");
  if (!(min_scaling_factor > 0)) return set_error(B2_ERR_ARG, "min_scaling_factor must be positive");
  if (h->world > 1) return set_error(B2_ERR_STATE, "b2_reg_min_max_point_radius is single-GPU (every image must be resident)");
  if (n == 0) return B2_OK;
  begin_call(h);
  DevBuf d_xyz, d_min, d_max, d_table;
  struct Rel { DevBuf* b[4]; ~Rel() { for (DevBuf* x : b) x->release(); } } rel{{&d_xyz, &d_min, &d_max, &d_table}};
  B2_TRY(d_xyz.ensure(n * 12)); B2_TRY(d_min.ensure(n * 4)); B2_TRY(d_max.ensure(n * 4));
  B2_CUDA(cudaMemcpyAsync(d_xyz.p, xyz, n * 12, cudaMemcpyHostToDevice, h->stream));
  B2_CUDA(cudaMemcpyAsync(d_min.p, min_radius, n * 4, cudaMemcpyHostToDevice, h->stream));
  B2_CUDA(cudaMemcpyAsync(d_max.p, max_radius, n * 4, cudaMemcpyHostToDevice, h->stream));
  int table_for = -1;                                    // intrinsics id the undistortion lookup currently holds
  for (size_t ii = 0; ii < h->images.size(); ++ii) {
    const ImageB& im = h->images[ii]; const IntrinsicsB& in = h->intr[im.intrinsics_id];
    const int best = in.best_available(std::max(h->prm.min_occlusion_check_image_scale, h->current_image_scale));
    const float* depth = nullptr;
    B2_TRY(render_depth(h, im, in, im.pose, best, &depth));
    const Levels L = levels_of(h, im, in);
    RadiusParams V;
    V.P = pose3_of(im.pose); V.cam = in.model(best); V.cam0 = in.model(0); V.image_scale = best; V.min_image_scale = in.min_image_scale;
    V.level = best - in.min_image_scale;
    V.depth = depth; V.mask = L.mask[V.level]; V.cmask = L.cmask[V.level]; V.img = L.img[V.level]; V.iw = L.iw[V.level];
    V.occlusion_threshold = h->prm.occlusion_depth_threshold; V.max_valid_intensity = h->prm.maximum_valid_intensity;
    V.min_scaling_factor = min_scaling_factor;
    V.table = nullptr;
    if (cam_has_lookup(V.cam0)) {
      const size_t px = (size_t)V.cam0.w * V.cam0.h;
      if (table_for != im.intrinsics_id) {
        B2_TRY(d_table.ensure(px * 8));
        kr_undistortion_lookup<<<divup(px, 128), 128, 0, h->stream>>>(V.cam0, d_table.as<float2>());
        ++h->launches; table_for = im.intrinsics_id;
      }
      V.table = d_table.as<float2>();
    }
    kr_min_max_radius<<<divup(n, 256), 256, 0, h->stream>>>(d_xyz.as<float>(), n, V, d_min.as<float>(), d_max.as<float>());
    ++h->launches;
  }
  B2_CUDA(cudaMemcpyAsync(min_radius, d_min.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaMemcpyAsync(max_radius, d_max.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaGetLastError());
  end_call(h);
  return B2_OK;
}

// ---- GroundTruthCreator (src/exe/ground_truth_creator.cc:44-215) ----
static int gt_params(b2_reg* h, int image_id, GtParams* V) {
  const ImageB& im = h->images[image_id]; const IntrinsicsB& in = h->intr[im.intrinsics_id];
  const float* depth = nullptr;
  B2_TRY(render_depth(h, im, in, im.pose, in.min_image_scale, &depth));     // RenderDepthMap(intrinsics, image, intrinsics.min_image_scale, ...)
  const Levels L = levels_of(h, im, in);
  V->P = pose3_of(im.pose); V->cam = in.model(0); V->depth = depth; V->mask = L.mask[0];
  V->occlusion_threshold = h->prm.occlusion_depth_threshold;
  return B2_OK;
}
static int gt_check(b2_reg* h, int image_id) {
  if (!h->initialized) return set_error(B2_ERR_STATE, "not initialized");
  if (image_id < 0 || image_id >= (int)h->images.size()) return set_error(B2_ERR_ARG, "bad image id");
  if (!owned(h, (size_t)image_id)) return set_error(B2_ERR_STATE, "image %d is owned by rank %d", image_id, image_id % h->world);
  return B2_OK;
}
int b2_reg_gt_accumulate_observations(b2_reg* h, int image_id, const float* xyz, size_t n, int32_t* observation_counts) {
  REG_ENTER(h);
  B2_TRY(gt_check(h, image_id));
  if (n && (!xyz || !observation_counts)) return set_error(B2_ERR_ARG, "null argument");
  if (n == 0) return B2_OK;
  begin_call(h);
  DevBuf d_xyz, d_cnt;
  struct Rel { DevBuf* b[2]; ~Rel() { for (DevBuf* x : b) x->release(); } } rel{{&d_xyz, &d_cnt}};
  B2_TRY(d_xyz.ensure(n * 12)); B2_TRY(d_cnt.ensure(n * 4));
  B2_CUDA(cudaMemcpyAsync(d_xyz.p, xyz, n * 12, cudaMemcpyHostToDevice, h->stream));
  B2_CUDA(cudaMemcpyAsync(d_cnt.p, observation_counts, n * 4, cudaMemcpyHostToDevice, h->stream));
  GtParams V; B2_TRY(gt_params(h, image_id, &V));
  kg_count<<<divup(n, 256), 256, 0, h->stream>>>(d_xyz.as<float>(), n, V, d_cnt.as<int>());
  ++h->launches;
  B2_CUDA(cudaMemcpyAsync(observation_counts, d_cnt.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaGetLastError());
  end_call(h);
  return B2_OK;
}
int b2_reg_gt_create(b2_reg* h, int image_id, const float* xyz, const uint8_t* rgb, size_t n, const int32_t* observation_counts, int scan_point_radius,
                     float* out_occlusion_depth, float* out_gt_depth, uint8_t* inout_scan_rendering_bgr) {
  REG_ENTER(h);
  B2_TRY(gt_check(h, image_id));
  if (n && (!xyz || !observation_counts)) return set_error(B2_ERR_ARG, "null argument");
  if (inout_scan_rendering_bgr && !rgb) return set_error(B2_ERR_ARG, "scan rendering needs the point colours");
  if (scan_point_radius < 0) return set_error(B2_ERR_ARG, "scan_point_radius must be >= 0");
  if (n >= 0xFFFFFFFFull) return set_error(B2_ERR_ARG, "more than 2^32 - 2 points");
  begin_call(h);
  GtParams V; B2_TRY(gt_params(h, image_id, &V));
  const size_t npix = (size_t)V.cam.w * V.cam.h;
  DevBuf d_xyz, d_cnt, d_rgb, d_depth, d_owner, d_img;
  struct Rel { DevBuf* b[6]; ~Rel() { for (DevBuf* x : b) x->release(); } } rel{{&d_xyz, &d_cnt, &d_rgb, &d_depth, &d_owner, &d_img}};
  if (out_occlusion_depth) {
    if (V.depth) B2_CUDA(cudaMemcpyAsync(out_occlusion_depth, V.depth, npix * 4, cudaMemcpyDeviceToHost, h->stream));
    else for (size_t i = 0; i < npix; ++i) out_occlusion_depth[i] = std::numeric_limits<float>::infinity();
  }
  if (n && (out_gt_depth || inout_scan_rendering_bgr)) {
    B2_TRY(d_xyz.ensure(n * 12)); B2_TRY(d_cnt.ensure(n * 4));
    B2_CUDA(cudaMemcpyAsync(d_xyz.p, xyz, n * 12, cudaMemcpyHostToDevice, h->stream));
    B2_CUDA(cudaMemcpyAsync(d_cnt.p, observation_counts, n * 4, cudaMemcpyHostToDevice, h->stream));
    if (out_gt_depth) { B2_TRY(d_depth.ensure(npix * 4)); kr_fill_u32<<<divup(npix, 256), 256, 0, h->stream>>>(d_depth.as<unsigned int>(), npix, 0x7f800000u); }
    if (inout_scan_rendering_bgr) {
      B2_TRY(d_owner.ensure(npix * 4)); B2_TRY(d_rgb.ensure(n * 3)); B2_TRY(d_img.ensure(npix * 3));
      B2_CUDA(cudaMemsetAsync(d_owner.p, 0, npix * 4, h->stream));
      B2_CUDA(cudaMemcpyAsync(d_rgb.p, rgb, n * 3, cudaMemcpyHostToDevice, h->stream));
      B2_CUDA(cudaMemcpyAsync(d_img.p, inout_scan_rendering_bgr, npix * 3, cudaMemcpyHostToDevice, h->stream));
    }
    kg_splat<<<divup(n, 256), 256, 0, h->stream>>>(d_xyz.as<float>(), n, V, d_cnt.as<int>(), scan_point_radius,
                                                   out_gt_depth ? d_depth.as<unsigned int>() : nullptr, inout_scan_rendering_bgr ? d_owner.as<unsigned int>() : nullptr);
    ++h->launches;
    if (out_gt_depth) B2_CUDA(cudaMemcpyAsync(out_gt_depth, d_depth.p, npix * 4, cudaMemcpyDeviceToHost, h->stream));
    if (inout_scan_rendering_bgr) {
      kg_paint<<<divup(npix, 256), 256, 0, h->stream>>>(npix, d_owner.as<unsigned int>(), d_rgb.as<unsigned char>(), d_img.as<unsigned char>());
      ++h->launches;
      B2_CUDA(cudaMemcpyAsync(inout_scan_rendering_bgr, d_img.p, npix * 3, cudaMemcpyDeviceToHost, h->stream));
    }
  } else if (out_gt_depth) {
    for (size_t i = 0; i < npix; ++i) out_gt_depth[i] = std::numeric_limits<float>::infinity();
  }
  B2_CUDA(cudaGetLastError());
  end_call(h);
  return B2_OK;
}

int b2_reg_create_observations(b2_reg* h, int border) {
  REG_ENTER(h);
  if (!h->initialized) return set_error(B2_ERR_STATE, "not initialized");
  begin_call(h);
  B2_TRY(create_observations(h, border));
  end_call(h);
  return B2_OK;
}

static int check_obs(b2_reg* h, int image_id, int ps) {
  if (image_id < 0 || image_id >= (int)h->obs.size() || ps < 0 || ps >= (int)h->pts.size() || ps >= (int)h->obs[image_id].size())
    return set_error(B2_ERR_ARG, "no observations for image %d, point scale %d", image_id, ps);
  return B2_OK;
}

int b2_reg_num_observations(b2_reg* h, int image_id, int ps, uint64_t* count) {
  REG_ENTER(h); B2_TRY(check_obs(h, image_id, ps));
  if (count) *count = h->obs[image_id][ps].count;
  return B2_OK;
}

int b2_reg_get_observations(b2_reg* h, int image_id, int ps, uint64_t* pidx, float* x, float* y, float* s, uint8_t* nb) {
  REG_ENTER(h); B2_TRY(check_obs(h, image_id, ps));
  const ObsSet& o = h->obs[image_id][ps];
  if (o.count == 0) return B2_OK;
  std::vector<unsigned int> idx(o.count);
  B2_CUDA(cudaMemcpy(idx.data(), o.idx.p, o.count * 4, cudaMemcpyDeviceToHost));
  if (pidx) for (size_t i = 0; i < o.count; ++i) pidx[i] = idx[i];
  if (x) B2_CUDA(cudaMemcpy(x, o.x.p, o.count * 4, cudaMemcpyDeviceToHost));
  if (y) B2_CUDA(cudaMemcpy(y, o.y.p, o.count * 4, cudaMemcpyDeviceToHost));
  if (s) B2_CUDA(cudaMemcpy(s, o.s.p, o.count * 4, cudaMemcpyDeviceToHost));
  if (nb) B2_CUDA(cudaMemcpy(nb, o.nb.p, o.count, cudaMemcpyDeviceToHost));
  return B2_OK;
}

int b2_reg_get_point_jacobians(b2_reg* h, int image_id, int ps, float* inten, float* jK, float* jP) {
  return b2_reg_get_point_jacobians_rig(h, image_id, ps, inten, jK, jP, nullptr);
}

int b2_reg_get_point_jacobians_rig(b2_reg* h, int image_id, int ps, float* inten, float* jK, float* jP, float* jR) {
  REG_ENTER(h); B2_TRY(check_obs(h, image_id, ps));
  ObsSet& o = h->obs[image_id][ps];
  if (o.count == 0) return B2_OK;
  const ImageB& I = h->images[image_id];
  const Levels L = levels_of(h, I, h->intr[I.intrinsics_id]);
  begin_call(h);
  const int np = h->intr[I.intrinsics_id].np(), kni = h->intr[I.intrinsics_id].kni();
  int rv; const RigDev rig = rig_dev(h, current_state(h), image_id, &rv);
  launch_jacobians(h, kni, o, h->pts[ps], pose3_of(I.pose), L, rig);
  B2_CUDA(cudaGetLastError());
  if (jR) {
    if (rig.dependent) B2_CUDA(cudaMemcpyAsync(jR, o.jR.p, o.count * 24, cudaMemcpyDeviceToHost, h->stream));
    else std::memset(jR, 0, o.count * 24);
  }
  if (inten) B2_CUDA(cudaMemcpyAsync(inten, o.inten.p, o.count * 4, cudaMemcpyDeviceToHost, h->stream));
  if (jK) B2_CUDA(cudaMemcpy2DAsync(jK, (size_t)np * 4, o.jK.p, (size_t)kni * 4, (size_t)np * 4, o.count, cudaMemcpyDeviceToHost, h->stream));   // rows of kni floats, the first np are the model's
  if (jP) B2_CUDA(cudaMemcpyAsync(jP, o.jP.p, o.count * 24, cudaMemcpyDeviceToHost, h->stream));
  end_call(h);
  return B2_OK;
}

int b2_reg_color_update(b2_reg* h) {
  REG_ENTER(h);
  if (h->obs.size() != h->images.size()) return set_error(B2_ERR_STATE, "create_observations first");
  begin_call(h); B2_TRY(color_update(h)); end_call(h);
  return B2_OK;
}

int b2_reg_get_descriptors(b2_reg* h, int ps, float* fixed, float* variable, int32_t* counts) {
  REG_ENTER(h);
  if (ps < 0 || ps >= (int)h->pts.size()) return set_error(B2_ERR_ARG, "bad point scale");
  const ScaleB& P = h->pts[ps];
  B2_CUDA(cudaStreamSynchronize(h->stream));
  if (P.n == 0) return B2_OK;
  if (fixed) B2_CUDA(cudaMemcpy(fixed, P.fixed_desc.p, P.n * K(h) * 4, cudaMemcpyDeviceToHost));
  if (variable) B2_CUDA(cudaMemcpy(variable, P.var_desc.p, P.n * K(h) * 4, cudaMemcpyDeviceToHost));
  if (counts) B2_CUDA(cudaMemcpy(counts, P.obs_count.p, P.n * 4, cudaMemcpyDeviceToHost));
  return B2_OK;
}

static void put_sums(const Sums& s, double out[6]) { if (out) { out[0] = s.fixed_sum; out[1] = s.nf; out[2] = s.var_sum; out[3] = s.nv; out[4] = out[5] = 0; } }

int b2_reg_cost(b2_reg* h, double* cost, double sums[6]) {
  REG_ENTER(h);
  if (h->obs.size() != h->images.size()) return set_error(B2_ERR_STATE, "create_observations first");
  begin_call(h);
  Sums s; double c = 0;
  B2_TRY(current_cost(h, &c, &s));
  end_call(h);
  if (cost) *cost = c;
  put_sums(s, sums);
  return B2_OK;
}

int b2_reg_accumulate(b2_reg* h, double* H, double* b, double sums[6], double* cost) {
  REG_ENTER(h);
  if (h->obs.size() != h->images.size()) return set_error(B2_ERR_STATE, "create_observations first");
  begin_call(h);
  std::vector<double> Hv, bv; Sums s;
  B2_TRY(accumulate_all(h, &Hv, &bv, &s));
  end_call(h);
  if (H) std::copy(Hv.begin(), Hv.end(), H);
  if (b) std::copy(bv.begin(), bv.end(), b);
  put_sums(s, sums);
  if (cost) *cost = compute_cost(h, s);
  return B2_OK;
}

int b2_reg_get_state(b2_reg* h, float* ip, float* poses) {
  REG_ENTER(h);
  if (ip) for (size_t i = 0; i < h->intr.size(); ++i) cam_get(h->intr[i].models[0], ip + intr_var(h, (int)i));
  if (poses) for (size_t i = 0; i < h->images.size(); ++i) { for (int k = 0; k < 4; ++k) poses[7 * i + k] = h->images[i].pose.q[k]; for (int k = 0; k < 3; ++k) poses[7 * i + 4 + k] = h->images[i].pose.t[k]; }
  return B2_OK;
}

int b2_reg_set_state(b2_reg* h, const float* ip, const float* poses) {
  REG_ENTER(h);
  if (ip) {
    for (size_t i = 0; i < h->intr.size(); ++i) { Cam& m = h->intr[i].models[0]; cam_set(&m, m.type, m.w, m.h, ip + intr_var(h, (int)i)); h->intr[i].build_pyramid(); }
    B2_TRY(compute_cutoffs(h, &h->intr));
  }
  if (poses) for (size_t i = 0; i < h->images.size(); ++i) { for (int k = 0; k < 4; ++k) h->images[i].pose.q[k] = poses[7 * i + k]; for (int k = 0; k < 3; ++k) h->images[i].pose.t[k] = poses[7 * i + 4 + k]; }
  return B2_OK;
}

int b2_reg_cost_for_delta(b2_reg* h, const double* delta, double* cost) {
  REG_ENTER(h);
  if (!delta || !cost) return set_error(B2_ERR_ARG, "null argument");
  if (h->obs.size() != h->images.size()) return set_error(B2_ERR_STATE, "create_observations first");
  begin_call(h);
  StateB ns; B2_TRY(delta_state(h, current_state(h), delta, &ns));
  B2_TRY(residual_for_state(h, ns, cost));
  end_call(h);
  return B2_OK;
}

int b2_reg_apply(b2_reg* h, float* lambda, float* max_change, int* applied, int* tries) {
  REG_ENTER(h);
  if (!lambda || !max_change || !applied) return set_error(B2_ERR_ARG, "null argument");
  if (h->obs.size() != h->images.size()) return set_error(B2_ERR_STATE, "create_observations first");
  begin_call(h);
  B2_TRY(apply_lm(h, lambda, max_change, applied, tries));
  end_call(h);
  return B2_OK;
}

int b2_reg_run_on_current_scale(b2_reg* h, int max_it, float max_change_thr, int no_opt_thr, int print, double* optimum_cost, int* converged,
                                int* iterations) {
  REG_ENTER(h);
  if (!h->initialized) return set_error(B2_ERR_STATE, "not initialized");
  if (!optimum_cost || !converged) return set_error(B2_ERR_ARG, "null argument");
  h->current_image_scale = std::min(h->current_image_scale, h->image_scale_count - 1 - 1);   // optimizer.cc:61
  float lambda = 64.0f;
  int without = 0, done = 0;
  *optimum_cost = std::numeric_limits<double>::infinity();
  *converged = 0;
  StateB best = current_state(h);
  for (int it = 0; it < max_it; ++it) {
    ++done;
    if (print) printf("Iteration %d\n", it + 1);
    int applied = 1; float max_change = std::numeric_limits<float>::infinity();
    if (it > 0) { applied = 0; max_change = 0; B2_TRY(apply_lm(h, &lambda, &max_change, &applied, nullptr)); }
    B2_TRY(create_observations(h, 1));
    if (h->prm.variable_residuals_weight > 0) B2_TRY(color_update(h));
    Sums s; double cost = 0;
    B2_TRY(current_cost(h, &cost, &s));
    if (print) printf("  Cost (considering occlusions) is: %.9g\n", cost);
    if (cost < *optimum_cost) { *optimum_cost = cost; without = 0; best = current_state(h); } else { ++without; }
    if (!applied || max_change < max_change_thr || without >= no_opt_thr) { *converged = 1; break; }
  }
  set_current_state(h, best);
  if (iterations) *iterations = done;
  return B2_OK;
}

int b2_reg_last_stats(b2_reg* h, b2_reg_stats* out) {
  REG_ENTER(h);
  if (!out) return set_error(B2_ERR_ARG, "null");
  *out = h->stats;
  return B2_OK;
}

// Stand-alone camera evaluation (the camera::CameraBase calls Path B makes), mainly for parity tests of the camera models.
int b2_camera_eval(int camera_model, int width, int height, const float* params, int num_params, int op, const float* in, size_t n, float* out,
                   float cutoffs[2]) {
  const int np = cam_param_count(camera_model);
  if (np < 0 || !params || num_params != np || width < 2 || height < 2) return set_error(B2_ERR_ARG, "unsupported camera model or parameter count");
  if (op < 0 || op > 3 || (op != 0 && n != 0 && (!in || !out))) return set_error(B2_ERR_ARG, "bad op / null buffers");
  b2_reg_params prm; b2_reg_default_params(&prm);
  b2_reg* h = nullptr;
  B2_TRY(b2_reg_create(&prm, &h));
  int rc = B2_OK;
  do {
    std::vector<IntrinsicsB> intr(1); intr[0].models.resize(1); cam_set(&intr[0].models[0], camera_model, width, height, params);
    if ((rc = compute_cutoffs(h, &intr)) != B2_OK) break;
    const Cam cam = intr[0].models[0];
    if (cutoffs) { cutoffs[0] = cam.cutoff2; cutoffs[1] = cam.inner_cutoff2; }
    if (op == 0 || n == 0) break;
    const int nin = op == 1 ? 2 : 3, nout = op == 1 ? 2 : op == 2 ? 6 : 2 * np;
    DevBuf din, dout;
    if ((rc = din.ensure(n * nin * 4)) != B2_OK || (rc = dout.ensure(n * nout * 4)) != B2_OK) break;
    if (cudaMemcpyAsync(din.p, in, n * nin * 4, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { rc = set_error(B2_ERR_CUDA, "H2D copy failed"); break; }
    kr_camera_eval<<<divup(n, 128), 128, 0, h->stream>>>(cam, op, din.as<float>(), n, dout.as<float>());
    if (cudaMemcpyAsync(out, dout.p, n * nout * 4, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) {
      rc = set_error(B2_ERR_CUDA, "camera_eval failed: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
  } while (false);
  b2_reg_destroy(h);
  return rc;
}

}  // extern "C"
