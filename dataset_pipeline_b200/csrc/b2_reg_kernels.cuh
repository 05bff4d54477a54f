// Device kernels of Path B (dense photometric image<->scan alignment) for sm_100a — pinhole, thin-prism and thin-prism-fisheye
// ("benchmark") cameras (b2_camera.cuh), no rigs.
//
//   K15 kr_pyr_down          u8 2x2 area mean ((a+b+c+d+2)>>2 = cv::resize INTER_AREA at factor 1/2) / mask OR    image.cc:106-154
//   B1  kr_splat_depth       point-splat depth map, atomicMin on float bits                                        occlusion_geometry.cc:404-464
//   K10 kr_visibility        R p + t, project, occlusion test, scale selection, border / mask / saturation tests  visibility_estimator.cc:258-295,366-532
//   B6  kr_neighbors_observed  all 5 neighbours observed? via the point->observation slot map                      visibility_estimator.cc:199-256
//   K11 kr_jacobians         trilinear taps on two pyramid levels, dI/d(K) (1x4), dI/d(pose) (1x6)               intrinsics_and_pose_optimizer.cc:933-1147
//       kr_intensity         trilinear taps only                                                                 cost_calculator.cc:128-142
//   K12 kr_accumulate        5-neighbour descriptor residual, robust weight, (4+6)^2 outer products in fp64     intrinsics_and_pose_optimizer.cc:770-930,1220-1296
//   K12b kr_residual_weights + kr_accumulate_blocks   the same for 16/18/24-column local systems (block pairs per warp, staged in smem)
//   K13 kr_residual_sums     residual sums of a (trial) state                                                  cost_calculator.cc:170-271
//   K14 kr_color_accumulate / kr_color_mean   variable descriptors = mean over images                              color_optimizer.cc:40-123
// Arithmetic follows the oracle's fp32 evaluation order; the file is built with -fmad=false so plain operators are never
// contracted. All of it is gather-bound byte / fp32 work: no tensor cores.
#pragma once
#include "b2_camera.cuh"
#include "b2_common.cuh"

namespace b2 {

static constexpr int kMaxLevels = 16;
static constexpr int kMaxNbr = 8;

// Pyramid of one image (+ its intrinsics) as seen by the kernels. Level l is image scale (min_image_scale + l).
struct Levels {
  Cam cam[kMaxLevels];
  const unsigned char* img[kMaxLevels];
  const unsigned char* mask[kMaxLevels];   // null = no mask at that level
  const unsigned char* cmask[kMaxLevels];  // camera mask of the intrinsics (intrinsics.h:104), null = none
  // Row pitch of img / mask / cmask at each level. Not always cam[l].w: the image pyramid TRUNCATES the halved size (image.cc:116),
  // the camera pyramid ROUNDS it (camera_base_impl.h:72), and the reference's bounds tests use the camera's size — so below an
  // odd-sized parent the last column tapped lies one past the image row and the (unchecked) cv::Mat access reads the first pixel of
  // the next row. Indexing with the image's own pitch reproduces exactly that; the buffers are zero-padded past their last pixel.
  int iw[kMaxLevels];
  int nlevels, min_image_scale;
};

struct Pose3 { float R[9]; float t[3]; };   // image_T_global (row-major R)

struct Robust { int type; float p; };
__device__ __forceinline__ float robust_residual(const Robust& r, float x) {      // robust_weighting.h:61-86
  if (r.type == 1) { const float a = fabsf(x); return a < r.p ? 0.5f * x * x : r.p * (a - 0.5f * r.p); }
  if (r.type == 2) {
    const float a = fabsf(x);
    if (a < r.p) { const float q = x / r.p; const float t = 1.f - q * q; return (1 / 6.f) * r.p * r.p * (1 - t * t * t); }
    return (1 / 6.f) * r.p * r.p;
  }
  return 0.5f * x * x;
}
__device__ __forceinline__ float robust_weight(const Robust& r, float x) {        // robust_weighting.h:90-106
  if (r.type == 1) { const float a = fabsf(x); return a < r.p ? 1.f : r.p / a; }
  if (r.type == 2) { const float a = fabsf(x); if (a < r.p) { const float q = x / r.p; const float t = 1.f - q * q; return t * t; } return 0.f; }
  return 1.f;
}

__device__ __forceinline__ float sum3p(float a, float b, float c) { return a + (b + c); }   // Eigen 3-term reduction order

__device__ __forceinline__ void rigid(const Pose3& P, float x, float y, float z, float* ox, float* oy, float* oz) {
  *ox = sum3p(P.R[0] * x, P.R[1] * y, P.R[2] * z) + P.t[0];
  *oy = sum3p(P.R[3] * x, P.R[4] * y, P.R[5] * z) + P.t[1];
  *oz = sum3p(P.R[6] * x, P.R[7] * y, P.R[8] * z) + P.t[2];
}

// interpolate_bilinear.h:36-74 on a u8 image with row pitch w.
__device__ __forceinline__ float bilinear(const unsigned char* __restrict__ im, int w, float x, float y) {
  const int ix = (int)x, iy = (int)y;
  const float fx = x - ix, fx_inv = 1.f - fx, fy = y - iy, fy_inv = 1.f - fy;
  const unsigned char* r0 = im + (size_t)iy * w + ix; const unsigned char* r1 = r0 + w;
  return fy_inv * (fx_inv * __ldg(r0) + fx * __ldg(r0 + 1)) + fy * (fx_inv * __ldg(r1) + fx * __ldg(r1 + 1));
}
__device__ __forceinline__ void bilinear_d(const unsigned char* __restrict__ im, int w, float x, float y, float* v, float* dx, float* dy) {
  const int ix = (int)x, iy = (int)y;
  const unsigned char* r0 = im + (size_t)iy * w + ix; const unsigned char* r1 = r0 + w;
  const unsigned char tl = __ldg(r0), tr = __ldg(r0 + 1), bl = __ldg(r1), br = __ldg(r1 + 1);
  const float fx = x - ix, fx_inv = 1.f - fx, fy = y - iy, fy_inv = 1.f - fy;
  const float top = fx_inv * tl + fx * tr, bottom = fx_inv * bl + fx * br;
  *v = fy_inv * top + fy * bottom;
  *dx = fy * (br - bl) + fy_inv * (tr - tl);
  *dy = bottom - top;
}
// interpolate_trilinear.h:44-87: image0 = smaller level, image1 = twice the size.
__device__ __forceinline__ float trilinear(const Levels& L, int small_level, float x0, float y0, float z) {
  const float v0 = bilinear(L.img[small_level], L.iw[small_level], x0, y0);
  const float x1 = 2 * (x0 + 0.5f) - 0.5f, y1 = 2 * (y0 + 0.5f) - 0.5f;
  const float v1 = bilinear(L.img[small_level - 1], L.iw[small_level - 1], x1, y1);
  return (1 - z) * v0 + z * v1;
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kr_pyr_down(const unsigned char* __restrict__ src, int sw, unsigned char* __restrict__ dst, int dw, int dh,
                                                   int is_mask) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw || y >= dh) return;
  const unsigned char* r0 = src + (size_t)(2 * y) * sw + 2 * x; const unsigned char* r1 = r0 + sw;
  dst[(size_t)y * dw + x] = is_mask ? (unsigned char)(r0[0] | r0[1] | r1[0] | r1[1]) : (unsigned char)((r0[0] + r0[1] + r1[0] + r1[1] + 2) >> 2);
}

// General cv::resize INTER_AREA (the parent is odd in x or y, so a scale is not the integer 2): per axis a table of (source index, alpha)
// taps with fractional coverage at the cell borders, `ofs[d] .. ofs[d+1]` the taps of destination index d. One thread per destination
// pixel runs OpenCV's loop nest for that pixel in the same order (resizeArea_: per source row buf = sum S * alpha over the x taps, then
// sum (+)= beta * buf), fp32 without contraction, saturate_cast<uchar> = round half to even. Tables: host, area_taps() in b2_reg.cu.
struct AreaTapDev { int si; float alpha; };
__global__ void __launch_bounds__(256) kr_pyr_area(const unsigned char* __restrict__ src, int sw, unsigned char* __restrict__ dst, int dw, int dh,
                                                   const AreaTapDev* __restrict__ xt, const int* __restrict__ xofs,
                                                   const AreaTapDev* __restrict__ yt, const int* __restrict__ yofs) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw || y >= dh) return;
  const int xb = xofs[x], xe = xofs[x + 1], yb = yofs[y], ye = yofs[y + 1];
  float sum = 0.f;
  for (int j = yb; j < ye; ++j) {
    const unsigned char* S = src + (size_t)yt[j].si * sw;
    float buf = 0.f;
    for (int k = xb; k < xe; ++k) buf = buf + (float)__ldg(S + xt[k].si) * xt[k].alpha;
    sum = j == yb ? yt[j].alpha * buf : sum + yt[j].alpha * buf;
  }
  dst[(size_t)y * dw + x] = (unsigned char)min(255, max(0, __float2int_rn(sum)));
}

__global__ void __launch_bounds__(256) kr_fill_u32(unsigned int* __restrict__ p, size_t n, unsigned int v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// occlusion_geometry.cc:404-464: min over splats is order-independent, so atomicMin on the (positive) float bits is exact.
__global__ void __launch_bounds__(256) kr_splat_depth(const float* __restrict__ xyz, size_t n, Pose3 P, Cam cam, float point_radius,
                                                      unsigned int* __restrict__ depth_bits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px, py, pz; rigid(P, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &px, &py, &pz);
  if (!(pz > 0.f)) return;
  float ux, uy; cam_project(cam, px / pz, py / pz, &ux, &uy);
  float d[6]; cam_d_by_world(cam, px, py, pz, d);
  float rx = sqrtf(sum3p(d[0] * d[0], d[1] * d[1], d[2] * d[2])) * point_radius;
  float ry = sqrtf(sum3p(d[3] * d[3], d[4] * d[4], d[5] * d[5])) * point_radius;
  rx = fminf(rx, 10.f); ry = fminf(ry, 10.f);
  const int ix = f2i_x86(ux + 0.5f), iy = f2i_x86(uy + 0.5f);
  if (ix == (int)0x80000000 || iy == (int)0x80000000) return;   // beyond the cut-off radius
  const int min_x = max(0, (int)(ix - rx + 0.5)), min_y = max(0, (int)(iy - ry + 0.5));
  const int end_x = min(cam.w, (int)(ix + rx + 1.5)), end_y = min(cam.h, (int)(iy + ry + 1.5));
  const unsigned int zb = __float_as_uint(pz);
  for (int y = min_y; y < end_y; ++y) for (int x = min_x; x < end_x; ++x) atomicMin(&depth_bits[(size_t)y * cam.w + x], zb);
}

// ------------------------------------------------------------------------------------------------------------------
// K8: triangle-mesh depth pass — the function of the reference's OpenGL renderer (occlusion_geometry.cc:213-245,
// opengl/renderer.cc:42-131,745-847,913-974: linear camera z, GL_LEQUAL, no culling, near/far clip, background 0) as a CUDA
// z-buffer. Rasterisation rule (shared with the oracle, oracle/orc_mesh.h): polygon clipped at z = min_depth in camera space,
// window coordinates X = fx x/z + cx + 0.5, pixel centres at i + 0.5, edge functions in double with an ownership rule for
// ties, perspective-correct depth 1 / sum(lambda_i / z_i), minimum kept by atomicMin on the float bits (order independent).
// Small triangles are rasterised by one thread each; triangles whose pixel bounding box exceeds kBigTriPixels are queued and
// rasterised by one block each.
// ------------------------------------------------------------------------------------------------------------------
static constexpr int kBigTriPixels = 4096;

struct ClippedTri { float x[4], y[4], z[4]; int n; };   // camera-space polygon after near clipping (n = 0, 3 or 4)

__device__ __forceinline__ ClippedTri clip_triangle(const float* __restrict__ v, const unsigned int* __restrict__ f, size_t fi, const Pose3& P,
                                                    const Cam& cam, float min_depth) {
  float px[3], py[3], pz[3];
  bool finite = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const size_t vi = f[3 * fi + k];
    rigid(P, v[3 * vi], v[3 * vi + 1], v[3 * vi + 2], &px[k], &py[k], &pz[k]);
    finite = finite && isfinite(px[k]) && isfinite(py[k]) && isfinite(pz[k]);
    if (finite) cam_vertex_distort(cam, &px[k], &py[k], pz[k]);      // the renderer's vertex stage; clipping follows it, as in GL
  }
  ClippedTri c; c.n = 0;
  if (!finite) return c;      // a triangle with a non-finite vertex draws nothing (test_renderer.cc:77-82,204-206)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int k2 = (k + 1) % 3;
    const bool ain = pz[k] >= min_depth, bin = pz[k2] >= min_depth;
    if (ain) { c.x[c.n] = px[k]; c.y[c.n] = py[k]; c.z[c.n] = pz[k]; ++c.n; }
    if (ain != bin) {
      const float tt = (min_depth - pz[k]) / (pz[k2] - pz[k]);
      c.x[c.n] = px[k] + tt * (px[k2] - px[k]); c.y[c.n] = py[k] + tt * (py[k2] - py[k]); c.z[c.n] = min_depth; ++c.n;
    }
  }
  if (c.n < 3) c.n = 0;
  return c;
}

struct TriSetup { double x0, y0, x1, y1, x2, y2, iz0, iz1, iz2, area; int ix0, ix1, iy0, iy1; bool o0, o1, o2, valid; };

__device__ __forceinline__ bool edge_owns(double dx, double dy) { return dy < 0.0 || (dy == 0.0 && dx > 0.0); }

__device__ __forceinline__ TriSetup setup_triangle(const Cam& c, const ClippedTri& t, int a, int b, int cc) {
  TriSetup s; s.valid = false;
  const float X0 = c.fx * (t.x[a] / t.z[a]) + c.cx + 0.5f, Y0 = c.fy * (t.y[a] / t.z[a]) + c.cy + 0.5f;
  const float X1 = c.fx * (t.x[b] / t.z[b]) + c.cx + 0.5f, Y1 = c.fy * (t.y[b] / t.z[b]) + c.cy + 0.5f;
  const float X2 = c.fx * (t.x[cc] / t.z[cc]) + c.cx + 0.5f, Y2 = c.fy * (t.y[cc] / t.z[cc]) + c.cy + 0.5f;
  double z0 = t.z[a], z1 = t.z[b], z2 = t.z[cc];
  s.x0 = X0; s.y0 = Y0; s.x1 = X1; s.y1 = Y1; s.x2 = X2; s.y2 = Y2;
  s.area = (s.x1 - s.x0) * (s.y2 - s.y0) - (s.y1 - s.y0) * (s.x2 - s.x0);
  if (s.area == 0.0 || !(s.area == s.area)) return s;
  if (s.area < 0.0) { double q; q = s.x1; s.x1 = s.x2; s.x2 = q; q = s.y1; s.y1 = s.y2; s.y2 = q; q = z1; z1 = z2; z2 = q; s.area = -s.area; }
  s.iz0 = 1.0 / z0; s.iz1 = 1.0 / z1; s.iz2 = 1.0 / z2;
  const double minx = fmin(s.x0, fmin(s.x1, s.x2)), maxx = fmax(s.x0, fmax(s.x1, s.x2));
  const double miny = fmin(s.y0, fmin(s.y1, s.y2)), maxy = fmax(s.y0, fmax(s.y1, s.y2));
  if (!(maxx >= 0.0 && maxy >= 0.0 && minx <= (double)c.w && miny <= (double)c.h)) return s;
  s.ix0 = max(0, (int)floor(minx - 0.5)); s.ix1 = min(c.w - 1, (int)ceil(maxx - 0.5));
  s.iy0 = max(0, (int)floor(miny - 0.5)); s.iy1 = min(c.h - 1, (int)ceil(maxy - 0.5));
  if (s.ix1 < s.ix0 || s.iy1 < s.iy0) return s;
  s.o0 = edge_owns(s.x2 - s.x1, s.y2 - s.y1); s.o1 = edge_owns(s.x0 - s.x2, s.y0 - s.y2); s.o2 = edge_owns(s.x1 - s.x0, s.y1 - s.y0);
  s.valid = true;
  return s;
}

__device__ __forceinline__ void raster_pixels(const Cam& c, const TriSetup& s, float max_depth, unsigned int* __restrict__ depth_bits, int first, int step) {
  const int bw = s.ix1 - s.ix0 + 1, total = bw * (s.iy1 - s.iy0 + 1);
  for (int k = first; k < total; k += step) {
    const int ix = s.ix0 + k % bw, iy = s.iy0 + k / bw;
    const double px = ix + 0.5, py = iy + 0.5;
    const double w0 = (s.x2 - s.x1) * (py - s.y1) - (s.y2 - s.y1) * (px - s.x1);
    const double w1 = (s.x0 - s.x2) * (py - s.y2) - (s.y0 - s.y2) * (px - s.x2);
    const double w2 = (s.x1 - s.x0) * (py - s.y0) - (s.y1 - s.y0) * (px - s.x0);
    if (!((w0 > 0.0 || (w0 == 0.0 && s.o0)) && (w1 > 0.0 || (w1 == 0.0 && s.o1)) && (w2 > 0.0 || (w2 == 0.0 && s.o2)))) continue;
    const float z = (float)(s.area / (w0 * s.iz0 + w1 * s.iz1 + w2 * s.iz2));   // one division per pixel (orc_mesh.h: raster_triangle)
    if (!(z <= max_depth)) continue;
    atomicMin(&depth_bits[(size_t)iy * c.w + ix], __float_as_uint(z));
  }
}

// Pass 1, one thread per triangle: clip + set-up; triangles that touch the image go to the warp queue (pixel bounding box up to
// kBigTriPixels) or to the block queue (above). Nothing is drawn here: a 2 cm triangle seen from 3 m at 4400 px focal length covers
// ~450 pixels (bounding box ~900), and one thread looping over them while the rest of its warp holds culled triangles was 1/3 of a
// whole Path B iteration (ncu r02u).
__global__ void __launch_bounds__(128) kr_raster_small(const float* __restrict__ v, const unsigned int* __restrict__ f, size_t nf, Pose3 P, Cam cam,
                                                       float min_depth, float max_depth, unsigned int* __restrict__ depth_bits,
                                                       unsigned int* __restrict__ big_list, unsigned int* __restrict__ big_count,
                                                       unsigned int* __restrict__ warp_list, unsigned int* __restrict__ warp_count) {
  const size_t fi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool queue = false, big = false;
  if (fi < nf) {
    const ClippedTri t = clip_triangle(v, f, fi, P, cam, min_depth);
    if (t.n != 0) {
      const TriSetup s0 = setup_triangle(cam, t, 0, 1, 2);
      TriSetup s1; s1.valid = false;
      if (t.n == 4) s1 = setup_triangle(cam, t, 0, 2, 3);
      long long px = 0;
      if (s0.valid) px += (long long)(s0.ix1 - s0.ix0 + 1) * (s0.iy1 - s0.iy0 + 1);
      if (s1.valid) px += (long long)(s1.ix1 - s1.ix0 + 1) * (s1.iy1 - s1.iy0 + 1);
      big = px > kBigTriPixels;
      queue = px > 0 && !big;
    }
  }
  // warp-aggregated appends
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int qm = __ballot_sync(0xffffffffu, queue), bm = __ballot_sync(0xffffffffu, big);
  unsigned int qbase = 0, bbase = 0;
  if (lane == 0) { if (qm) qbase = atomicAdd(warp_count, (unsigned int)__popc(qm)); if (bm) bbase = atomicAdd(big_count, (unsigned int)__popc(bm)); }
  qbase = __shfl_sync(0xffffffffu, qbase, 0); bbase = __shfl_sync(0xffffffffu, bbase, 0);
  if (queue) warp_list[qbase + __popc(qm & ((1u << lane) - 1u))] = (unsigned int)fi;
  if (big) big_list[bbase + __popc(bm & ((1u << lane) - 1u))] = (unsigned int)fi;
}

// Pass 2: one WARP per queued triangle, the lanes stride over its bounding-box pixels (same set-up, same per-pixel arithmetic, atomicMin
// on the depth bits: the map does not depend on who draws what).
__global__ void __launch_bounds__(256) kr_raster_warp(const float* __restrict__ v, const unsigned int* __restrict__ f, Pose3 P, Cam cam, float min_depth,
                                                      float max_depth, unsigned int* __restrict__ depth_bits, const unsigned int* __restrict__ warp_list,
                                                      const unsigned int* __restrict__ warp_count) {
  const unsigned int lane = threadIdx.x & 31u, nwarps = gridDim.x * (blockDim.x >> 5), n = *warp_count;
  for (unsigned int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n; b += nwarps) {
    const ClippedTri t = clip_triangle(v, f, warp_list[b], P, cam, min_depth);
    if (t.n == 0) continue;
    const TriSetup s0 = setup_triangle(cam, t, 0, 1, 2);
    if (s0.valid) raster_pixels(cam, s0, max_depth, depth_bits, (int)lane, 32);
    if (t.n == 4) { const TriSetup s1 = setup_triangle(cam, t, 0, 2, 3); if (s1.valid) raster_pixels(cam, s1, max_depth, depth_bits, (int)lane, 32); }
  }
}

__global__ void __launch_bounds__(256) kr_raster_big(const float* __restrict__ v, const unsigned int* __restrict__ f, Pose3 P, Cam cam, float min_depth,
                                                     float max_depth, unsigned int* __restrict__ depth_bits, const unsigned int* __restrict__ big_list,
                                                     const unsigned int* __restrict__ big_count) {
  for (unsigned int b = blockIdx.x; b < *big_count; b += gridDim.x) {
    const ClippedTri t = clip_triangle(v, f, big_list[b], P, cam, min_depth);
    if (t.n == 0) continue;
    const TriSetup s0 = setup_triangle(cam, t, 0, 1, 2);
    if (s0.valid) raster_pixels(cam, s0, max_depth, depth_bits, threadIdx.x, blockDim.x);
    if (t.n == 4) { const TriSetup s1 = setup_triangle(cam, t, 0, 2, 3); if (s1.valid) raster_pixels(cam, s1, max_depth, depth_bits, threadIdx.x, blockDim.x); }
  }
}

// inf (nothing drawn) -> 0, the GL clear colour (renderer.cc:766); also seeds the masked copy.
__global__ void __launch_bounds__(256) kr_depth_background(float* __restrict__ depth, float* __restrict__ copy, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d = depth[i];
  if (isinf(d)) { d = 0.f; depth[i] = d; }
  if (copy) copy[i] = d;
}

// K9: MaskOutOcclusionBoundaries (occlusion_geometry.cc:284-402). Pass 1, one thread per mesh edge: silhouette test by face-normal
// signs, then the splats along the edge (centre visibility test against the UNMASKED map `in`) are appended to a queue as pixel
// rectangles + depth. Pass 2, one warp per queued splat: writes -1 where the unmasked map is empty or not more than 0.05 in front of
// the splat (write-write races all store -1: benign; every test reads `in`, so the result does not depend on the drawing order).
// A splat of radius 0.03 m seen from 3 m at 4400 px focal length is an 89 x 89 pixel rectangle: drawn by the edge's own thread
// (as before) the pass was half of a whole Path B iteration (ncu r02u). Queue overflow: the edge's thread draws the splat itself.
struct MeshEdgeDev { unsigned int v1, v2, f1, f2, flags; };   // flags: bit0 open, bit1 opposite_normals
struct EdgeSplat { int min_x, min_y, end_x, end_y; float pz; };
__device__ __forceinline__ void draw_splat(const EdgeSplat& s, int w, const float* __restrict__ in, float* __restrict__ out, int first, int step) {
  const int bw = s.end_x - s.min_x, total = bw * (s.end_y - s.min_y);
  for (int k = first; k < total; k += step) {
    const size_t pix = (size_t)(s.min_y + k / bw) * w + (s.min_x + k % bw);
    const float old = in[pix];
    if (old == 0 || old + 0.05f > s.pz) out[pix] = -1.f;
  }
}
__global__ void __launch_bounds__(128) kr_mask_edges(const MeshEdgeDev* __restrict__ edges, size_t ne, const float* __restrict__ v,
                                                     const float* __restrict__ fn, Pose3 P, float ipx, float ipy, float ipz, Cam cam,
                                                     float splat_radius, const float* __restrict__ in, float* __restrict__ out,
                                                     EdgeSplat* __restrict__ queue, unsigned int* __restrict__ queue_count, unsigned int queue_cap) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  const MeshEdgeDev e = edges[i];
  const float e1x = v[3 * (size_t)e.v1], e1y = v[3 * (size_t)e.v1 + 1], e1z = v[3 * (size_t)e.v1 + 2];
  if (!(e.flags & 1u)) {
    const float tx = ipx - e1x, ty = ipy - e1y, tz = ipz - e1z;
    const bool face1 = sum3p(fn[3 * (size_t)e.f1] * tx, fn[3 * (size_t)e.f1 + 1] * ty, fn[3 * (size_t)e.f1 + 2] * tz) > 0;
    const bool face2 = sum3p(fn[3 * (size_t)e.f2] * tx, fn[3 * (size_t)e.f2 + 1] * ty, fn[3 * (size_t)e.f2 + 2] * tz) > 0;
    const bool opp = (e.flags & 2u) != 0;
    if (!((opp && (face1 == face2)) || (face1 != face2 && !opp))) return;
  }
  float ax, ay, az; rigid(P, e1x, e1y, e1z, &ax, &ay, &az);
  if (az <= 0) return;
  float bx, by, bz; rigid(P, v[3 * (size_t)e.v2], v[3 * (size_t)e.v2 + 1], v[3 * (size_t)e.v2 + 2], &bx, &by, &bz);
  if (bz <= 0) return;
  const float dx = bx - ax, dy = by - ay, dz = bz - az;
  const int count = 1 + min((int)(sqrtf(sum3p(dx * dx, dy * dy, dz * dz)) / splat_radius + 0.5f), 150);
  for (int k = 0; k < count; ++k) {
    const float factor = k / (count - 1.0f);
    const float px = ax + factor * dx, py = ay + factor * dy, pz = az + factor * dz;
    if (!(pz > 0)) continue;
    const float nx = px / pz, ny = py / pz;
    float ux, uy; cam_project(cam, nx, ny, &ux, &uy);
    const int ix = f2i_x86(ux + 0.5f), iy = f2i_x86(uy + 0.5f);
    if (!(ux + 0.5f >= 0 && uy + 0.5f >= 0 && ix >= 0 && iy >= 0 && ix < cam.w && iy < cam.h && in[(size_t)iy * cam.w + ix] + 0.05f >= pz)) continue;
    float dd[6]; cam_d_by_world(cam, px, py, pz, dd);
    const float rx = sqrtf(sum3p(dd[0] * dd[0], dd[1] * dd[1], dd[2] * dd[2])) * splat_radius, ry = sqrtf(sum3p(dd[3] * dd[3], dd[4] * dd[4], dd[5] * dd[5])) * splat_radius;
    EdgeSplat s;
    s.min_x = max(0, (int)(ix - rx + 0.5)); s.min_y = max(0, (int)(iy - ry + 0.5));
    s.end_x = min(cam.w, (int)(ix + rx + 1.5)); s.end_y = min(cam.h, (int)(iy + ry + 1.5));
    s.pz = pz;
    if (s.end_x <= s.min_x || s.end_y <= s.min_y) continue;
    const unsigned int slot = atomicAdd(queue_count, 1u);
    if (slot < queue_cap) queue[slot] = s;
    else draw_splat(s, cam.w, in, out, 0, 1);
  }
}
__global__ void __launch_bounds__(256) kr_draw_splats(const EdgeSplat* __restrict__ queue, const unsigned int* __restrict__ queue_count, unsigned int queue_cap,
                                                      int w, const float* __restrict__ in, float* __restrict__ out) {
  const unsigned int lane = threadIdx.x & 31u, nwarps = gridDim.x * (blockDim.x >> 5), n = min(*queue_count, queue_cap);
  for (unsigned int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n; b += nwarps) draw_splat(queue[b], w, in, out, (int)lane, 32);
}

// K10. One thread per candidate point (all points of the scale, or the i-th entry of a visibility list).
struct VisParams {
  Pose3 P; Cam cam; int image_scale;            // camera of the occlusion-check scale
  const float* depth;                           // null = no occlusion test (indexed visibility lists)
  float occlusion_threshold, point_radius, max_valid_intensity;
  int border, check_masks, current_image_scale, image_scale_count;
};
__global__ void __launch_bounds__(256) kr_visibility(const float* __restrict__ xyz, const unsigned int* __restrict__ list, size_t count, VisParams V,
                                                     Levels L, unsigned int* __restrict__ flags, float* __restrict__ ox, float* __restrict__ oy,
                                                     float* __restrict__ os) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const size_t pi = list ? list[i] : i;
  unsigned int ok = 0;
  float rx_ = 0.f, ry_ = 0.f, rs_ = 0.f;
  float px, py, pz; rigid(V.P, xyz[3 * pi], xyz[3 * pi + 1], xyz[3 * pi + 2], &px, &py, &pz);
  if (pz > 0.f) {
    float ixx, ixy; cam_project(V.cam, px / pz, py / pz, &ixx, &ixy);
    const int ix = f2i_x86(ixx + 0.5f), iy = f2i_x86(ixy + 0.5f);
    if (ix >= 0 && iy >= 0 && ix < V.cam.w && iy < V.cam.h &&
        (V.depth == nullptr || __ldg(V.depth + (size_t)iy * V.cam.w + ix) + V.occlusion_threshold >= pz)) {
      // CreateObservationIfScaleFits (visibility_estimator.cc:405-532)
      const float qx = px + V.point_radius, qy = py + 0, qz = pz + 0;
      float rx, ry; cam_project(V.cam, qx / qz, qy / qz, &rx, &ry);
      const float dx = rx - ixx, dy = ry - ixy;
      const float radius_pixels = sqrtf(dx * dx + dy * dy);
      const float observation_scale = (float)((double)V.image_scale + log2((double)(2 * radius_pixels)));
      // a non-finite scale (offset point beyond the cut-off radius) is undefined behaviour in the reference: no observation
      if (isfinite(observation_scale) && observation_scale >= (float)max(L.min_image_scale, V.current_image_scale) &&
          (int)observation_scale < V.image_scale_count - 1) {
        const int small = (int)observation_scale + 1;
        const int lvl = max(0, small - L.min_image_scale);
        const Cam& ic = L.cam[lvl];
        const float nx = V.cam.fx_inv * ixx + V.cam.cx_inv, ny = V.cam.fy_inv * ixy + V.cam.cy_inv;
        const float jx = ic.fx * nx + ic.cx, jy = ic.fy * ny + ic.cy;
        const int jix = (int)(jx + 0.5f), jiy = (int)(jy + 0.5f);
        if (jx + 0.5f >= (float)V.border && jy + 0.5f >= (float)V.border && jix >= V.border && jiy >= V.border && jix < ic.w - V.border &&
            jiy < ic.h - V.border) {
          bool keep = true;
          if (V.check_masks) {
            const int level = small - L.min_image_scale;
            const size_t jpix = (size_t)jiy * L.iw[level] + jix;
            if (L.mask[level] != nullptr && __ldg(L.mask[level] + jpix) != 0) keep = false;
            else if (L.cmask[level] != nullptr && __ldg(L.cmask[level] + jpix) != 0) keep = false;
            else if ((float)__ldg(L.img[level] + jpix) > V.max_valid_intensity) keep = false;
          }
          if (keep) { ok = 1; rx_ = jx; ry_ = jy; rs_ = observation_scale; }
        }
      }
    }
  }
  flags[i] = ok; ox[i] = rx_; oy[i] = ry_; os[i] = rs_;
}

// ComputeMinMaxPointRadius (multi_scale_point_cloud.cc:126-184) fused with the visibility test of _AppendObservationsForImageNoScale
// (visibility_estimator.cc:296-364): one thread per point, one launch per image (images in sequence on one stream, so the per-point
// min / max need no atomics). cam0 / table: the camera of image scale `min_image_scale` and its undistortion lookup (null for pinhole).
struct RadiusParams {
  Pose3 P; Cam cam; Cam cam0; int image_scale, min_image_scale, level, iw;   // iw: row pitch of mask / cmask / img (Levels::iw)
  const float* depth; const unsigned char* mask; const unsigned char* cmask; const unsigned char* img;
  const float2* table;
  float occlusion_threshold, max_valid_intensity; double min_scaling_factor;
};
__global__ void __launch_bounds__(256) kr_min_max_radius(const float* __restrict__ xyz, size_t n, RadiusParams V, float* __restrict__ min_radius,
                                                         float* __restrict__ max_radius) {
  const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= n) return;
  float px, py, pz; rigid(V.P, xyz[3 * pi], xyz[3 * pi + 1], xyz[3 * pi + 2], &px, &py, &pz);
  if (!(pz > 0.f)) return;
  float ixx, ixy; cam_project(V.cam, px / pz, py / pz, &ixx, &ixy);
  const int ix = f2i_x86(ixx + 0.5f), iy = f2i_x86(ixy + 0.5f);
  if (!(ixx + 0.5f >= 0 && ixy + 0.5f >= 0 && ix >= 0 && iy >= 0 && ix < V.cam.w && iy < V.cam.h &&
        (V.depth == nullptr || __ldg(V.depth + (size_t)iy * V.cam.w + ix) + V.occlusion_threshold >= pz))) return;
  const size_t pix = (size_t)iy * V.iw + ix;
  if (V.mask != nullptr && __ldg(V.mask + pix) != 0) return;
  if (V.cmask != nullptr && __ldg(V.cmask + pix) != 0) return;
  if ((float)__ldg(V.img + pix) > V.max_valid_intensity) return;
  float returned_scale = (float)V.image_scale - 1e-6f;
  float ox = ixx, oy = ixy;
  if (returned_scale < 0.f) { returned_scale = 0.f; ox = 0.5f * (ixx + 0.5f) - 0.5f; oy = 0.5f * (ixy + 0.5f) - 0.5f; }
  // image_x_at_scale(min_image_scale) (point_observation.h:84-93): 2^(smaller scale - desired) in double (exact power of two)
  const double p2 = ldexp(1.0, ((int)returned_scale + 1) - V.min_image_scale);
  const float x0 = (float)(p2 * (double)(ox + 0.5f) - 0.5), y0 = (float)(p2 * (double)(oy + 0.5f) - 0.5);
  const float offx = (x0 - 0.5f < 0) ? (x0 + 0.5f) : (x0 - 0.5f);
  float nx, ny; cam_image_to_normalized(V.cam0, V.table, offx, y0, &nx, &ny);
  const float dx = px - pz * nx, dy = py - pz * ny, dz = pz - pz * 1.f;
  const float point_radius = sqrtf(sum3p(dx * dx, dy * dy, dz * dz));
  min_radius[pi] = fminf(min_radius[pi], point_radius);
  max_radius[pi] = fmaxf(max_radius[pi], (float)((double)point_radius / V.min_scaling_factor));
}

// GroundTruthCreator (src/exe/ground_truth_creator.cc:44-215). The visibility test both passes share (:66-79, :163-174).
struct GtParams { Pose3 P; Cam cam; const float* depth; const unsigned char* mask; float occlusion_threshold; };
__device__ __forceinline__ bool gt_visible(const GtParams& V, const float* __restrict__ xyz, size_t i, int* ox, int* oy, float* oz) {
  float px, py, pz; rigid(V.P, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &px, &py, &pz);
  if (!(pz > 0)) return false;
  float ixx, ixy; cam_project(V.cam, px / pz, py / pz, &ixx, &ixy);
  const int ix = f2i_x86(ixx + 0.5f), iy = f2i_x86(ixy + 0.5f);
  if (!(ix >= 0 && iy >= 0 && ix < V.cam.w && iy < V.cam.h)) return false;
  const size_t pix = (size_t)iy * V.cam.w + ix;
  if (V.depth != nullptr && !(__ldg(V.depth + pix) + V.occlusion_threshold >= pz)) return false;
  if (V.mask != nullptr && __ldg(V.mask + pix) == 2) return false;      // opt::MaskType::kEvalObs
  *ox = ix; *oy = iy; *oz = pz;
  return true;
}
// AccumulateScanObservationsForImage (:44-82)
__global__ void __launch_bounds__(256) kg_count(const float* __restrict__ xyz, size_t n, GtParams V, int* __restrict__ counts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ix, iy; float z;
  if (gt_visible(V, xyz, i, &ix, &iy, &z)) counts[i] += 1;
}
// CreateGroundTruthForImage (:150-190): depth = per-pixel minimum (order independent: atomicMin on the bits of positive floats); the
// rendering paints squares in scan order, later points over earlier ones = per pixel the HIGHEST point index wins (atomicMax on
// index + 1), resolved to colours by kg_paint.
__global__ void __launch_bounds__(256) kg_splat(const float* __restrict__ xyz, size_t n, GtParams V, const int* __restrict__ counts, int radius,
                                                unsigned int* __restrict__ depth_bits /* nullable */, unsigned int* __restrict__ owner /* nullable */) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || counts[i] < 2) return;
  int ix, iy; float z;
  if (!gt_visible(V, xyz, i, &ix, &iy, &z)) return;
  if (owner != nullptr) {
    const int min_x = max(0, ix - radius), min_y = max(0, iy - radius), end_x = min(V.cam.w, ix + radius + 1), end_y = min(V.cam.h, iy + radius + 1);
    for (int y = min_y; y < end_y; ++y) for (int x = min_x; x < end_x; ++x) atomicMax(&owner[(size_t)y * V.cam.w + x], (unsigned int)i + 1u);
  }
  if (depth_bits != nullptr) atomicMin(&depth_bits[(size_t)iy * V.cam.w + ix], __float_as_uint(z));
}
__global__ void __launch_bounds__(256) kg_paint(size_t npix, const unsigned int* __restrict__ owner, const unsigned char* __restrict__ rgb,
                                                unsigned char* __restrict__ bgr) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const unsigned int o = owner[p];
  if (o == 0u) return;
  const size_t i = (size_t)o - 1;
  bgr[3 * p] = rgb[3 * i + 2]; bgr[3 * p + 1] = rgb[3 * i + 1]; bgr[3 * p + 2] = rgb[3 * i];
}

__global__ void __launch_bounds__(256) kr_compact(const unsigned int* __restrict__ flags, const unsigned int* __restrict__ offs,
                                                  const unsigned int* __restrict__ list, size_t count, const float* __restrict__ cx,
                                                  const float* __restrict__ cy, const float* __restrict__ cs, unsigned int* __restrict__ idx,
                                                  float* __restrict__ ox, float* __restrict__ oy, float* __restrict__ os, int* __restrict__ slot) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count || !flags[i]) return;
  const unsigned int o = offs[i];
  const unsigned int pi = list ? list[i] : (unsigned int)i;
  idx[o] = pi; ox[o] = cx[i]; oy[o] = cy[i]; os[o] = cs[i];
  slot[pi] = (int)o;
}

__global__ void __launch_bounds__(256) kr_neighbors_observed(const unsigned int* __restrict__ idx, size_t count, const unsigned int* __restrict__ nbr,
                                                             int K, const int* __restrict__ slot, unsigned char* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const size_t p = idx[i];
  bool all = true;
  for (int k = 0; k < K; ++k) if (slot[nbr[p * K + k]] < 0) { all = false; break; }
  out[i] = all ? 1 : 0;
}

__global__ void __launch_bounds__(256) kr_intensity(size_t count, const float* __restrict__ ox, const float* __restrict__ oy,
                                                    const float* __restrict__ os, Levels L, float* __restrict__ inten) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float s = os[i];
  const int small = (int)s + 1 - L.min_image_scale;
  inten[i] = trilinear(L, small, ox[i], oy[i], 1 - (s - (int)s));
}

// What a dependent rig image (camera index > 0) adds to K11 (intrinsics_and_pose_optimizer.cc:651-670, 1107-1143).
struct RigDev { int dependent; float Rr[9]; float q[4]; float t[3]; };   // image_T_rig rotation matrix; rig_T_global as quaternion + translation
__device__ __forceinline__ void quat_rotate_dev(const float q[4], float px, float py, float pz, float* ox, float* oy, float* oz) {   // so3.hpp:360-370
  float ux = q[1] * pz - q[2] * py, uy = q[2] * px - q[0] * pz, uz = q[0] * py - q[1] * px;
  ux = ux + ux; uy = uy + uy; uz = uz + uz;
  const float cx = q[1] * uz - q[2] * uy, cy = q[2] * ux - q[0] * uz, cz = q[0] * uy - q[1] * ux;
  *ox = px + q[3] * ux + cx; *oy = py + q[3] * uy + cy; *oz = pz + q[3] * uz + cz;
}

// K11 (intrinsics_and_pose_optimizer.cc:933-1147), depth residuals off. NI = intrinsics parameter count of the camera model. For a
// dependent rig image jP is the derivative by the rig REFERENCE image's pose and jR (6 per observation) by this camera's extrinsics.
template <int NI>
__global__ void __launch_bounds__(256) kr_jacobians(size_t count, const unsigned int* __restrict__ idx, const float* __restrict__ ox,
                                                    const float* __restrict__ oy, const float* __restrict__ os, const float* __restrict__ xyz,
                                                    Pose3 P, float point_radius, Levels L, float* __restrict__ inten, float* __restrict__ jK,
                                                    float* __restrict__ jP, RigDev rig, float* __restrict__ jR) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const Cam& cam = L.cam[0];
  const size_t p = idx[i];
  float tx, ty, tz; rigid(P, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], &tx, &ty, &tz);
  const float s = os[i], x0 = ox[i], y0 = oy[i];
  const int smaller = (int)s + 1;
  const int sl = smaller - L.min_image_scale;
  const float z = 1 - (s - (int)s);
  float v0, d0x, d0y, v1, d1x, d1y;
  bilinear_d(L.img[sl], L.iw[sl], x0, y0, &v0, &d0x, &d0y);
  const float x1 = 2 * (x0 + 0.5f) - 0.5f, y1 = 2 * (y0 + 0.5f) - 0.5f;
  bilinear_d(L.img[sl - 1], L.iw[sl - 1], x1, y1, &v1, &d1x, &d1y);
  inten[i] = (1 - z) * v0 + z * v1;
  float ji0 = (1 - z) * d0x + z * 2 * d1x;
  float ji1 = (1 - z) * d0y + z * 2 * d1y;
  const float ji2 = -1 * (v1 - v0);
  const float scale_factor = ldexpf(1.f, L.min_image_scale - smaller);      // pow(2, min_image_scale - smaller_interpolation_scale)
  const float inv_scale_factor = 1.f / scale_factor;
  ji0 *= scale_factor; ji1 *= scale_factor;
  const float mx = inv_scale_factor * (x0 + 0.5f) - 0.5f, my = inv_scale_factor * (y0 + 0.5f) - 0.5f;
  const float qx = tx + point_radius;
  float oxp, oyp; cam_project(cam, qx / tz, ty / tz, &oxp, &oyp);
  const float rdx = oxp - mx, rdy = oyp - my;
  const float denom = fmaxf(1e-6f, 0.693147180559945f * (rdx * rdx + rdy * rdy));
  // d(project)/d(intrinsics) rows (pinhole: [x 0 1 0], [0 y 0 1]), scale row from the offset point
  float ax[NI], ay[NI], bx[NI], by[NI];
  cam_d_by_intrinsics<NI>(cam, tx, ty, tz, ax, ay);
  cam_d_by_intrinsics<NI>(cam, qx, ty, tz, bx, by);
#pragma unroll
  for (int k = 0; k < NI; ++k) {
    const float row2 = ((bx[k] - ax[k]) * rdx + (by[k] - ay[k]) * rdy) / denom;
    jK[(size_t)NI * i + k] = sum3p(ji0 * ax[k], ji1 * ay[k], ji2 * row2);
  }
  float dw[6], dq[6];
  cam_d_by_world(cam, tx, ty, tz, dw);
  cam_d_by_world(cam, qx, ty, tz, dq);
  float g[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float row2 = ((dq[k] - dw[k]) * rdx + (dq[3 + k] - dw[3 + k]) * rdy) / denom;
    g[k] = sum3p(ji0 * dw[k], ji1 * dw[3 + k], ji2 * row2);
  }
  // [I | -[p]x]: rows (1,0,0,0,z,-y), (0,1,0,-z,0,x), (0,0,1,y,-x,0)
  const float C[18] = {1, 0, 0, 0, tz, -1 * ty, 0, 1, 0, -1 * tz, 0, tx, 0, 0, 1, ty, -1 * tx, 0};
  if (rig.dependent) {
    float rx, ry, rz; quat_rotate_dev(rig.q, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], &rx, &ry, &rz);
    rx += rig.t[0]; ry += rig.t[1]; rz += rig.t[2];
    float gr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) gr[k] = sum3p(g[0] * rig.Rr[k], g[1] * rig.Rr[3 + k], g[2] * rig.Rr[6 + k]);
    const float Cr[18] = {1, 0, 0, 0, rz, -1 * ry, 0, 1, 0, -1 * rz, 0, rx, 0, 0, 1, ry, -1 * rx, 0};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      jP[6 * i + c] = sum3p(gr[0] * Cr[c], gr[1] * Cr[6 + c], gr[2] * Cr[12 + c]);
      jR[6 * i + c] = sum3p(g[0] * C[c], g[1] * C[6 + c], g[2] * C[12 + c]);
    }
    return;
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) jP[6 * i + c] = sum3p(g[0] * C[c], g[1] * C[6 + c], g[2] * C[12 + c]);
}

// Common inputs of the residual kernels for one (image, point scale).
struct ResidualArgs {
  size_t count;
  const unsigned int* idx; const unsigned char* nb; const unsigned int* nbr; int K; const int* slot; const float* inten;
  const float* fixed_desc; const float* var_desc; const int* obs_count;
  Robust robust; float fixed_w, var_w;
};

template <int NV>
__device__ __forceinline__ void block_reduce_d(double* acc, double (*sm)[NV], int nthreads, double* out) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[k] = v;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sm[w][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double v = sm[0][threadIdx.x];
    for (int i = 1; i < nthreads / 32; ++i) v += sm[i][threadIdx.x];
    out[threadIdx.x] = v;
  }
}

// K13: residual sums [fixed_sum, n_fixed, var_sum, n_var] per block (cost_calculator.cc:170-271).
__global__ void __launch_bounds__(256) kr_residual_sums(ResidualArgs A, double* __restrict__ partials /* [grid][4] */) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.count; i += (size_t)gridDim.x * blockDim.x) {
    if (!A.nb[i]) continue;
    const size_t p = A.idx[i];
    const float Ic = A.inten[i];
    float In[kMaxNbr];
#pragma unroll
    for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) In[k] = A.inten[A.slot[A.nbr[p * A.K + k]]];
    if (A.fixed_w > 0) {
      float pr = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) { const float c = (In[k] - Ic) - A.fixed_desc[p * A.K + k]; pr += c * c; }
      acc[0] += (double)robust_residual(A.robust, sqrtf(pr)); acc[1] += 1.0;
    }
    if (A.var_w > 0 && A.obs_count[p] >= 2) {
      float pr = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) { const float c = (In[k] - Ic) - A.var_desc[p * A.K + k]; pr += c * c; }
      acc[2] += (double)robust_residual(A.robust, sqrtf(pr)); acc[3] += 1.0;
    }
  }
  __shared__ double sm[8][4];
  double out;
  block_reduce_d<4>(acc, sm, 256, &out);
  if (threadIdx.x < 4) partials[(size_t)blockIdx.x * 4 + threadIdx.x] = out;
}

// K12: per block [55 upper entries of the local (4+6)^2 system | 10 of b | fixed_sum n_fixed var_sum n_var] = 69 doubles.
// F32PROD = true: every product (w * dj[r]) * dj[c] is formed in fp32 and converted to fp64 before the add, exactly like
// AccumulateOnHAndB (:1262-1293) — 650 fp32->fp64 conversions per observation, which made the kernel conversion-pipe bound.
// F32PROD = false (default): the fixed and the variable descriptor residual share their Jacobian differences, so their weights are
// merged (ws = w_f + w_v, wr_k = w_f c_f[k] + w_v c_v[k]) and the products are formed in fp64 from the fp32 differences
// (10 conversions + 10 DMUL + 65 DFMA per neighbour). The sums differ from the reference's only by the fp32 rounding of its
// individual products (<= 2 ulp_fp32 each, random sign): far inside the 1e-5 parity tolerance, and closer to the real-number value.
static constexpr int kNI = 4, kNV = kNI + 6, kNH = kNV * (kNV + 1) / 2, kAccB = kNH + kNV + 4;
template <bool F32PROD>
__global__ void __launch_bounds__(128) kr_accumulate(ResidualArgs A, const float* __restrict__ jK, const float* __restrict__ jP,
                                                     double* __restrict__ partials /* [grid][kAccB] */) {
  double acc[kAccB];
#pragma unroll
  for (int k = 0; k < kAccB; ++k) acc[k] = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.count; i += (size_t)gridDim.x * blockDim.x) {
    if (!A.nb[i]) continue;
    const size_t p = A.idx[i];
    const float Ic = A.inten[i];
    float jc[kNV];
#pragma unroll
    for (int v = 0; v < kNI; ++v) jc[v] = jK[kNI * i + v];
#pragma unroll
    for (int v = 0; v < 6; ++v) jc[kNI + v] = jP[6 * i + v];
    int nj[kMaxNbr]; float In[kMaxNbr];
#pragma unroll
    for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) { nj[k] = A.slot[A.nbr[p * A.K + k]]; In[k] = A.inten[nj[k]]; }
    float wt[2] = {0.f, 0.f}; float comp[2][kMaxNbr];
#pragma unroll
    for (int type = 0; type < 2; ++type) {
#pragma unroll
      for (int k = 0; k < kMaxNbr; ++k) comp[type][k] = 0.f;
      const float sw = type == 0 ? A.fixed_w : A.var_w;
      if (!(sw > 0)) continue;
      if (type == 1 && A.obs_count[p] < 2) continue;
      const float* desc = type == 0 ? A.fixed_desc : A.var_desc;
      float pr = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) { const float c = (In[k] - Ic) - desc[p * A.K + k]; comp[type][k] = c; pr += c * c; }
      pr = sqrtf(pr);
      acc[kNH + kNV + 2 * type] += (double)robust_residual(A.robust, pr);
      acc[kNH + kNV + 2 * type + 1] += 1.0;
      wt[type] = sw * robust_weight(A.robust, pr);
    }
    if (!(wt[0] != 0 || wt[1] != 0)) continue;
    const double ws = (double)wt[0] + (double)wt[1];
#pragma unroll
    for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) {
      float dj[kNV];
#pragma unroll
      for (int v = 0; v < kNI; ++v) dj[v] = jK[kNI * (size_t)nj[k] + v] - jc[v];
#pragma unroll
      for (int v = 0; v < 6; ++v) dj[kNI + v] = jP[6 * (size_t)nj[k] + v] - jc[kNI + v];
      if (F32PROD) {
#pragma unroll
        for (int type = 0; type < 2; ++type) {
          const float w = wt[type];
          if (w != 0) {
            int e = 0;
#pragma unroll
            for (int c = 0; c < kNV; ++c)
#pragma unroll
              for (int r = 0; r <= c; ++r) { acc[e] += (double)((w * dj[r]) * dj[c]); ++e; }
            const float wr = w * comp[type][k];
#pragma unroll
            for (int v = 0; v < kNV; ++v) acc[kNH + v] += (double)(wr * dj[v]);
          }
        }
      } else {
        double d[kNV];
#pragma unroll
        for (int v = 0; v < kNV; ++v) d[v] = (double)dj[v];
        const double wr = (double)wt[0] * (double)comp[0][k] + (double)wt[1] * (double)comp[1][k];
        int e = 0;
#pragma unroll
        for (int c = 0; c < kNV; ++c) {
          const double t = ws * d[c];
#pragma unroll
          for (int r = 0; r <= c; ++r) { acc[e] = fma(t, d[r], acc[e]); ++e; }
          acc[kNH + c] = fma(wr, d[c], acc[kNH + c]);
        }
      }
    }
  }
  __shared__ double sm[4][kAccB];
  double out;
  block_reduce_d<kAccB>(acc, sm, 128, &out);
  if (threadIdx.x < kAccB) partials[(size_t)blockIdx.x * kAccB + threadIdx.x] = out;
}

// K12w: the same accumulation for local systems that no longer fit one thread's registers (12 intrinsics: (12+6)^2 = 171 upper
// entries + 18 of b; with rig extrinsics up to (12+6+6)^2 = 300 + 24). A warp takes 32 observations at a time:
//   1. lane-parallel: lane l does observation l's scalar work (neighbour slots, descriptor residuals, robust weights, residual sums);
//   2. serial over the 32 observations: lane l < NV holds column l of the centre row and of the K neighbour rows (all K+1 row loads
//      issued together), the NH + NV accumulators are spread over the lanes (<= 11 each) and every (r, c) product fetches its two
//      factors with shuffles.
// Same fp32 products and fp64 accumulation as the per-thread kernel; per-block output [NH | NV | 4 sums]. The kernel is bound by
// the fp32->fp64 conversion of every product (quarter-rate pipe) and by the shuffles, not by memory.
// With RIG (dependent rig image) the local system has 6 more columns, ordered [intrinsics | rig extrinsics | reference pose] like the
// global variable vector, so that every upper-triangle product has the factor order of AccumulateOnHAndB (:1262-1283).
template <int NI, bool RIG>
__global__ void __launch_bounds__(128) kr_accumulate_wide(ResidualArgs A, const float* __restrict__ jK, const float* __restrict__ jP,
                                                          const float* __restrict__ jR, double* __restrict__ partials /* [grid][NH + NV + 4] */) {
  constexpr int NR = RIG ? 6 : 0, NV = NI + NR + 6, NH = NV * (NV + 1) / 2, NE = NH + NV, SL = (NE + 31) / 32, NOUT = NE + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int er[SL], ec[SL];   // entry e = s*32 + lane: H(r, c) for e < NH (e = c(c+1)/2 + r), b(c) with r = -1 for NH <= e < NE, unused r = -2
#pragma unroll
  for (int s = 0; s < SL; ++s) {
    const int e = s * 32 + lane;
    if (e < NH) { int c = 0; while ((c + 1) * (c + 2) / 2 <= e) ++c; ec[s] = c; er[s] = e - c * (c + 1) / 2; }
    else if (e < NE) { er[s] = -1; ec[s] = e - NH; }
    else { er[s] = -2; ec[s] = 0; }
  }
  double acc[SL];
#pragma unroll
  for (int s = 0; s < SL; ++s) acc[s] = 0.0;
  double sums[4] = {0.0, 0.0, 0.0, 0.0};   // per lane: this lane's observations
  // Jacobian row `slot`, column `lane` of the local system
  auto row = [&](size_t slot) -> float {
    if (lane < NI) return jK[(size_t)NI * slot + lane];
    if (RIG && lane < NI + NR) return jR[6 * slot + (lane - NI)];
    if (lane < NV) return jP[6 * slot + (lane - NI - NR)];
    return 0.f;
  };
  const size_t nw = (size_t)gridDim.x * 4;
  for (size_t base = ((size_t)blockIdx.x * 4 + warp) * 32; base < A.count; base += nw * 32) {
    // ---- 1. lane-parallel scalar work of observation base + lane ----
    const size_t i = base + lane;
    int nj[kMaxNbr]; float cf[kMaxNbr], cv[kMaxNbr];
    float wf = 0.f, wv = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxNbr; ++k) { nj[k] = 0; cf[k] = 0.f; cv[k] = 0.f; }
    if (i < A.count && A.nb[i]) {
      const size_t p = A.idx[i];
      const float Ic = A.inten[i];
      float In[kMaxNbr];
#pragma unroll
      for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) { nj[k] = A.slot[A.nbr[p * A.K + k]]; In[k] = A.inten[nj[k]]; }
      if (A.fixed_w > 0) {
        float pr = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) { const float c = (In[k] - Ic) - A.fixed_desc[p * A.K + k]; cf[k] = c; pr += c * c; }
        pr = sqrtf(pr);
        sums[0] += (double)robust_residual(A.robust, pr); sums[1] += 1.0;
        wf = A.fixed_w * robust_weight(A.robust, pr);
      }
      if (A.var_w > 0 && A.obs_count[p] >= 2) {
        float pr = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) { const float c = (In[k] - Ic) - A.var_desc[p * A.K + k]; cv[k] = c; pr += c * c; }
        pr = sqrtf(pr);
        sums[2] += (double)robust_residual(A.robust, pr); sums[3] += 1.0;
        wv = A.var_w * robust_weight(A.robust, pr);
      }
    }
    // ---- 2. the outer products, one observation at a time, all lanes cooperating ----
    unsigned int todo = __ballot_sync(0xffffffffu, wf != 0 || wv != 0);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const float jc = row(base + src);
      float dj[kMaxNbr];
#pragma unroll
      for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) dj[k] = row((size_t)__shfl_sync(0xffffffffu, nj[k], src));
#pragma unroll
      for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) dj[k] -= jc;
      // the two factors of an entry are the same for the fixed and the variable descriptor residual: fetch once, use twice
      const float w0 = __shfl_sync(0xffffffffu, wf, src), w1 = __shfl_sync(0xffffffffu, wv, src);
#pragma unroll
      for (int k = 0; k < kMaxNbr; ++k) if (k < A.K) {
        const float wr0 = w0 * __shfl_sync(0xffffffffu, cf[k], src), wr1 = w1 * __shfl_sync(0xffffffffu, cv[k], src);
#pragma unroll
        for (int s = 0; s < SL; ++s) {
          const float a = __shfl_sync(0xffffffffu, dj[k], max(er[s], 0)), b = __shfl_sync(0xffffffffu, dj[k], ec[s]);
          if (er[s] >= 0) {
            if (w0 != 0) acc[s] += (double)((w0 * a) * b);
            if (w1 != 0) acc[s] += (double)((w1 * a) * b);
          } else if (er[s] == -1) {
            if (w0 != 0) acc[s] += (double)(wr0 * b);
            if (w1 != 0) acc[s] += (double)(wr1 * b);
          }
        }
      }
    }
  }
  // residual sums: lanes -> warp (fixed shuffle tree), then the four warps in order
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sums[k] += __shfl_xor_sync(0xffffffffu, sums[k], o);
  __shared__ double sm[4][SL * 32 + 4];
#pragma unroll
  for (int s = 0; s < SL; ++s) sm[warp][s * 32 + lane] = acc[s];
  if (lane == 0) { for (int k = 0; k < 4; ++k) sm[warp][SL * 32 + k] = sums[k]; }
  __syncthreads();
  for (int t = threadIdx.x; t < NOUT; t += blockDim.x) {
    const int src = t < NE ? t : SL * 32 + (t - NE);
    partials[(size_t)blockIdx.x * NOUT + t] = ((sm[0][src] + sm[1][src]) + sm[2][src]) + sm[3][src];
  }
}

// K12b = kr_residual_weights + kr_accumulate_blocks: the wide local systems (16, 18 or 24 columns) without shuffles.
//
// kr_residual_weights (thread per observation) does the scalar part once: neighbour slots nj[k] (-1 = observation contributes nothing),
// the merged weight ws = w_f + w_v, the residual factors wr_k = w_f c_f[k] + w_v c_v[k] (fp64; see kr_accumulate) and the residual sums.
//
// kr_accumulate_blocks: the local system is cut into column blocks of <= 6 ([4|6], [4|6|6], [6|6|6] or [6|6|6|6] = intrinsics | rig | pose);
// a CTA has one warp per block pair (bi <= bj): 3, 6 or 10 warps. Per chunk of 32 observations
//   1. all threads stage the Jacobian differences dj[k][col] = row(nj[k])[col] - row(centre)[col] (fp32 subtraction as the reference,
//      :880-905), converted ONCE to fp64, into shared memory as [k][col][observation] — coalesced row reads, every row read once per CTA;
//   2. warp (bi, bj), lane = observation: acc[r][c] += (ws * dj[bi,r]) * dj[bj,c] (36 DFMA per neighbour), diagonal warps also
//      b[c] += wr_k * dj[bj,c]. The accumulators (42 doubles) stay in registers for the whole kernel.
// Slots, weights and rows of the next chunk(s) are prefetched while the current chunk is multiplied, so the gather latency hides
// behind the DFMAs. Per-CTA partials [NH | NV]; fixed shuffle tree + fixed CTA order => deterministic.
template <int KN>
__global__ void __launch_bounds__(256) kr_residual_weights(ResidualArgs A, int* __restrict__ nj_out /* [KN][count] */, double* __restrict__ ws_out,
                                                           double* __restrict__ wr_out /* [KN][count] */, double* __restrict__ partials /* [grid][4] */) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.count; i += (size_t)gridDim.x * blockDim.x) {
    int nj[KN]; double wr[KN]; double ws = 0.0;
#pragma unroll
    for (int k = 0; k < KN; ++k) { nj[k] = -1; wr[k] = 0.0; }
    if (A.nb[i]) {
      const size_t p = A.idx[i];
      const float Ic = A.inten[i];
      int sl[KN]; float In[KN], cf[KN], cv[KN];
#pragma unroll
      for (int k = 0; k < KN; ++k) { sl[k] = A.slot[A.nbr[p * KN + k]]; In[k] = A.inten[sl[k]]; cf[k] = 0.f; cv[k] = 0.f; }
      float wf = 0.f, wv = 0.f;
      if (A.fixed_w > 0) {
        float pr = 0.f;
#pragma unroll
        for (int k = 0; k < KN; ++k) { const float c = (In[k] - Ic) - A.fixed_desc[p * KN + k]; cf[k] = c; pr += c * c; }
        pr = sqrtf(pr);
        acc[0] += (double)robust_residual(A.robust, pr); acc[1] += 1.0;
        wf = A.fixed_w * robust_weight(A.robust, pr);
      }
      if (A.var_w > 0 && A.obs_count[p] >= 2) {
        float pr = 0.f;
#pragma unroll
        for (int k = 0; k < KN; ++k) { const float c = (In[k] - Ic) - A.var_desc[p * KN + k]; cv[k] = c; pr += c * c; }
        pr = sqrtf(pr);
        acc[2] += (double)robust_residual(A.robust, pr); acc[3] += 1.0;
        wv = A.var_w * robust_weight(A.robust, pr);
      }
      if (wf != 0 || wv != 0) {
        ws = (double)wf + (double)wv;
#pragma unroll
        for (int k = 0; k < KN; ++k) { nj[k] = sl[k]; wr[k] = (double)wf * (double)cf[k] + (double)wv * (double)cv[k]; }
      }
    }
    ws_out[i] = ws;
#pragma unroll
    for (int k = 0; k < KN; ++k) { nj_out[(size_t)k * A.count + i] = nj[k]; wr_out[(size_t)k * A.count + i] = wr[k]; }
  }
  __shared__ double sm[8][4];
  double out;
  block_reduce_d<4>(acc, sm, 256, &out);
  if (threadIdx.x < 4) partials[(size_t)blockIdx.x * 4 + threadIdx.x] = out;
}

// K12 for pinhole on top of the pre-pass: thread per observation, the scalar chain (point -> neighbour indices -> observation slots ->
// intensities, descriptors, robust weights) already done by kr_residual_weights at full occupancy, so what is left per thread is one
// coalesced read of (slots, ws, wr_k), six row reads and the 5 x (10 DMUL + 65 DFMA). Per-block output [55 | 10].
template <int KN>
__global__ void __launch_bounds__(128) kr_accumulate_weighted(size_t count, const int* __restrict__ nj_in, const double* __restrict__ ws_in,
                                                              const double* __restrict__ wr_in, const float* __restrict__ jK, const float* __restrict__ jP,
                                                              double* __restrict__ partials /* [grid][kNH + kNV] */) {
  constexpr int NE = kNH + kNV;
  double acc[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) acc[k] = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    int nj[KN];
#pragma unroll
    for (int k = 0; k < KN; ++k) nj[k] = nj_in[(size_t)k * count + i];
    if (nj[0] < 0) continue;
    const double ws = ws_in[i];
    double wr[KN];
#pragma unroll
    for (int k = 0; k < KN; ++k) wr[k] = wr_in[(size_t)k * count + i];
    float jc[kNV];
    {
      const float4 a = *reinterpret_cast<const float4*>(jK + kNI * i);
      const float2 b0 = *reinterpret_cast<const float2*>(jP + 6 * i), b1 = *reinterpret_cast<const float2*>(jP + 6 * i + 2), b2 = *reinterpret_cast<const float2*>(jP + 6 * i + 4);
      jc[0] = a.x; jc[1] = a.y; jc[2] = a.z; jc[3] = a.w; jc[4] = b0.x; jc[5] = b0.y; jc[6] = b1.x; jc[7] = b1.y; jc[8] = b2.x; jc[9] = b2.y;
    }
    float rows[KN][kNV];
#pragma unroll
    for (int k = 0; k < KN; ++k) {
      const size_t s = (size_t)nj[k];
      const float4 a = *reinterpret_cast<const float4*>(jK + kNI * s);
      const float2 b0 = *reinterpret_cast<const float2*>(jP + 6 * s), b1 = *reinterpret_cast<const float2*>(jP + 6 * s + 2), b2 = *reinterpret_cast<const float2*>(jP + 6 * s + 4);
      rows[k][0] = a.x; rows[k][1] = a.y; rows[k][2] = a.z; rows[k][3] = a.w;
      rows[k][4] = b0.x; rows[k][5] = b0.y; rows[k][6] = b1.x; rows[k][7] = b1.y; rows[k][8] = b2.x; rows[k][9] = b2.y;
    }
#pragma unroll
    for (int k = 0; k < KN; ++k) {
      double d[kNV];
#pragma unroll
      for (int v = 0; v < kNV; ++v) d[v] = (double)(rows[k][v] - jc[v]);
      int e = 0;
#pragma unroll
      for (int c = 0; c < kNV; ++c) {
        const double t = ws * d[c];
#pragma unroll
        for (int r = 0; r <= c; ++r) { acc[e] = fma(t, d[r], acc[e]); ++e; }
        acc[kNH + c] = fma(wr[k], d[c], acc[kNH + c]);
      }
    }
  }
  __shared__ double sm[4][NE];
  double out;
  block_reduce_d<NE>(acc, sm, 128, &out);
  if (threadIdx.x < NE) partials[(size_t)blockIdx.x * NE + threadIdx.x] = out;
}

template <int NI, bool RIG>
struct BlockCfg {
  static constexpr int NR = RIG ? 6 : 0, NV = NI + NR + 6, NH = NV * (NV + 1) / 2, NE = NH + NV;
  static constexpr int B0 = NI == 4 ? 4 : 6;              // length of the first column block; all others are 6
  static constexpr int NBLK = (NV - B0) / 6 + 1;           // 3 or 4
  static constexpr int G = NBLK * (NBLK + 1) / 2;          // block pairs = warps per CTA: 6 or 10
  static constexpr int T = 32 * G;
  static constexpr int CTAS = G <= 3 ? 4 : G <= 6 ? 2 : 1; // resident CTAs per SM the register budget is set for
  static constexpr int EPT = (32 * NV + T - 1) / T;        // staged (observation, column) elements per thread
  static constexpr int SDS = 33;                           // padded observation stride of the staged differences
  static constexpr size_t smem(int kn) { return sizeof(double) * (size_t)kn * NV * SDS; }
};

template <int NI, bool RIG, int KN>
__global__ void __launch_bounds__(BlockCfg<NI, RIG>::T, BlockCfg<NI, RIG>::CTAS) kr_accumulate_blocks(size_t count, const int* __restrict__ nj_in, const double* __restrict__ ws_in,
                                                                             const double* __restrict__ wr_in, const float* __restrict__ jK,
                                                                             const float* __restrict__ jP, const float* __restrict__ jR,
                                                                             double* __restrict__ partials /* [grid][NE] */) {
  using C = BlockCfg<NI, RIG>;
  constexpr int NR = C::NR, NV = C::NV, NH = C::NH, NE = C::NE, B0 = C::B0, T = C::T, EPT = C::EPT, SDS = C::SDS;
  extern __shared__ double sd[];                           // [KN][NV][SDS]
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  int bj = 0;
  while ((bj + 1) * (bj + 2) / 2 <= g) ++bj;
  const int bi = g - bj * (bj + 1) / 2;                    // pairs in column-major upper order: (0,0) (0,1) (1,1) (0,2) ...
  const int ca = bi == 0 ? 0 : B0 + 6 * (bi - 1), cb = bj == 0 ? 0 : B0 + 6 * (bj - 1);
  const int la = (B0 != 6 && bi == 0) ? B0 : 6, lb = (B0 != 6 && bj == 0) ? B0 : 6;
  const bool diag = bi == bj;
  double acc[36], accb[6];
#pragma unroll
  for (int e = 0; e < 36; ++e) acc[e] = 0.0;
#pragma unroll
  for (int e = 0; e < 6; ++e) accb[e] = 0.0;

  int eo[EPT], ecol[EPT];                                  // this thread's staged elements: observation within the chunk, local column
#pragma unroll
  for (int j = 0; j < EPT; ++j) { const int e = tid + j * T; eo[j] = e < 32 * NV ? e / NV : -1; ecol[j] = e % NV; }
  auto row = [&](int col, size_t slot) -> float {
    if (col < NI) return jK[(size_t)NI * slot + col];
    if (RIG && col < NI + NR) return jR[6 * slot + (col - NI)];
    return jP[6 * slot + (col - NI - NR)];
  };
  const size_t nchunks = (count + 31) / 32, stride = gridDim.x;
  // Software pipeline over this CTA's chunks c, c + stride, ...: slots travel global -> register (one per thread) -> s_nj two chunks
  // ahead, weights global -> register -> s_w one chunk ahead, rows global -> registers (issued right after the first barrier, consumed
  // at the top of the next iteration), so every global load has a whole multiply phase to land.
  __shared__ int s_nj[KN][32];
  __shared__ double s_w[KN + 1][32];                       // [0] = ws, [1 + k] = wr_k
  float rw[EPT][KN + 1]; bool rv[EPT];
  // slot planes k = sk, sk + G, ... and weight planes (0 = ws, 1 + k = wr_k) sk, sk + G, ... belong to warp sk; lane = observation
  const int sk = tid >> 5;
  constexpr int G = C::G, NJP = (KN + G - 1) / G, NWP = (KN + 1 + G - 1) / G;
  auto load_nj = [&](size_t chunk, int (&r)[NJP]) {
    const size_t i = chunk * 32 + lane;
#pragma unroll
    for (int a = 0; a < NJP; ++a) {
      const int k = sk + a * G;
      r[a] = (k < KN && chunk < nchunks && i < count) ? nj_in[(size_t)k * count + i] : -1;
    }
  };
  auto load_w = [&](size_t chunk, double (&r)[NWP]) {
    const size_t i = chunk * 32 + lane;
#pragma unroll
    for (int a = 0; a < NWP; ++a) {
      const int k = sk + a * G;
      r[a] = !(k <= KN && chunk < nchunks && i < count) ? 0.0 : k == 0 ? ws_in[i] : wr_in[(size_t)(k - 1) * count + i];
    }
  };
  auto publish_nj = [&](const int (&r)[NJP]) {
#pragma unroll
    for (int a = 0; a < NJP; ++a) if (sk + a * G < KN) s_nj[sk + a * G][lane] = r[a];
  };
  auto publish_w = [&](const double (&r)[NWP]) {
#pragma unroll
    for (int a = 0; a < NWP; ++a) if (sk + a * G <= KN) s_w[sk + a * G][lane] = r[a];
  };
  auto load_rows = [&](size_t chunk) {
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      rv[j] = false;
      if (eo[j] >= 0 && chunk < nchunks) {
        const int n0 = s_nj[0][eo[j]];
        if (n0 >= 0) {
          rv[j] = true;
          rw[j][KN] = row(ecol[j], chunk * 32 + (size_t)eo[j]);
          rw[j][0] = row(ecol[j], (size_t)n0);
#pragma unroll
          for (int k = 1; k < KN; ++k) rw[j][k] = row(ecol[j], (size_t)s_nj[k][eo[j]]);
        }
      }
    }
  };

  size_t c = blockIdx.x;
  int njreg[NJP]; double wreg[NWP];
  load_nj(c, njreg); publish_nj(njreg);
  __syncthreads();
  load_rows(c);
  load_nj(c + stride, njreg);
  load_w(c, wreg);
  __syncthreads();
  for (; c < nchunks; c += stride) {
    // stage chunk c: fp32 difference (as the reference), one conversion per element; publish slots(c + stride) and weights(c)
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      if (eo[j] >= 0) {
#pragma unroll
        for (int k = 0; k < KN; ++k) sd[(k * NV + ecol[j]) * SDS + eo[j]] = rv[j] ? (double)(rw[j][k] - rw[j][KN]) : 0.0;
      }
    }
    publish_nj(njreg); publish_w(wreg);
    __syncthreads();
    load_rows(c + stride);
    load_nj(c + 2 * stride, njreg);
    load_w(c + stride, wreg);
    // multiply chunk c
    const double ws = s_w[0][lane];
    if (__any_sync(0xffffffffu, ws != 0.0)) {
#pragma unroll
      for (int k = 0; k < KN; ++k) {
        double a[6], b[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) a[r] = r < la ? sd[(k * NV + ca + r) * SDS + lane] : 0.0;
        if (diag) {
#pragma unroll
          for (int r = 0; r < 6; ++r) b[r] = a[r];
        } else {
#pragma unroll
          for (int r = 0; r < 6; ++r) b[r] = r < lb ? sd[(k * NV + cb + r) * SDS + lane] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const double t = ws * a[r];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) acc[r * 6 + cc] = fma(t, b[cc], acc[r * 6 + cc]);
        }
        if (diag) {
          const double wrk = s_w[1 + k][lane];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) accb[cc] = fma(wrk, b[cc], accb[cc]);
        }
      }
    }
    __syncthreads();
  }
  // lanes -> warp with a fixed shuffle tree; lane 0 of warp (bi, bj) owns the entries of its block pair
#pragma unroll
  for (int e = 0; e < 36; ++e)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
#pragma unroll
  for (int e = 0; e < 6; ++e)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accb[e] += __shfl_xor_sync(0xffffffffu, accb[e], o);
  if (lane == 0) {
    double* out = partials + (size_t)blockIdx.x * NE;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        const int R = ca + r, Cc = cb + cc;
        if (r < la && cc < lb && R <= Cc) out[Cc * (Cc + 1) / 2 + R] = acc[r * 6 + cc];
      }
    if (diag) {
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) if (cc < lb) out[NH + cb + cc] = accb[cc];
    }
  }
}

// Fixed-order sum of the per-block partials: out[v] = sum_b partials[b][v]. One CTA of 64 threads per value: thread t adds the blocks
// t, t + 64, ... in order, then a fixed shared-memory tree combines the 64 partial sums (deterministic; a single serial loop over
// ~600 blocks per value cost 20-40 us per launch and showed up 240 times per accumulate call).
__global__ void __launch_bounds__(64) kr_reduce_partials(const double* __restrict__ partials, int nblocks, int nvals, double* __restrict__ out) {
  const int v = blockIdx.x;
  if (v >= nvals) return;
  double s = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 64) s += partials[(size_t)b * nvals + v];
  __shared__ double sm[64];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[v] = sm[0];
}

// K14 (color_optimizer.cc:84-108): one image at a time; a point has at most one observation per image and scale, so the
// read-modify-write needs no atomics and images applied in ascending order reproduce the oracle's fp32 sums bit for bit.
__global__ void __launch_bounds__(256) kr_color_accumulate(size_t count, const unsigned int* __restrict__ idx, const unsigned char* __restrict__ nb,
                                                           const unsigned int* __restrict__ nbr, int K, const int* __restrict__ slot,
                                                           const float* __restrict__ inten, float* __restrict__ var_desc, int* __restrict__ obs_count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count || !nb[i]) return;
  const size_t p = idx[i];
  const float Ic = inten[i];
  obs_count[p] += 1;
  for (int k = 0; k < K; ++k) var_desc[p * K + k] += inten[slot[nbr[p * K + k]]] - Ic;
}
__global__ void __launch_bounds__(256) kr_color_mean(size_t n, int K, float* __restrict__ var_desc, const int* __restrict__ obs_count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = obs_count[i];
  if (c > 1) for (int k = 0; k < K; ++k) var_desc[i * K + k] /= c;
}
__global__ void __launch_bounds__(256) kr_fixed_descriptors(size_t n, int K, const unsigned int* __restrict__ nbr, const float* __restrict__ colors,
                                                            float* __restrict__ fixed_desc, int* __restrict__ obs_count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < K; ++k) fixed_desc[i * K + k] = colors[nbr[i * K + k]] - colors[i];
  obs_count[i] = 99999;    // problem.cc:570
}

}  // namespace b2
