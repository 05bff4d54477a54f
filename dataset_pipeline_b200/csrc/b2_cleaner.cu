// Point-cloud tools next to the hot paths, on sm_100a, on top of K7's exact kNN search:
//
//   b2_lsor_filter            pcl::LocalStatisticalOutlierRemoval<PointT>::applyFilterIndices
//                             (/root/reference/src/geometry/local_statistical_outlier_removal.hpp:72-176) as PointCloudCleaner drives it
//                             (/root/reference/src/exe/point_cloud_cleaner.cc:80-96)
//   b2_mesh_squared_distance  igl::AABB::squared_distance (/root/reference/thirdparty/igl/AABB.cpp, point_simplex_squared_distance.cpp:44-115):
//                             exact minimum over the triangles of the fp32 point-triangle distance, evaluated in libigl's operation order
//   b2_splat_create           the per-point body of SplatCreator (/root/reference/src/exe/splat_creator.cc:146-215)
//
// Kernels: K7 kn_knn_normals in statistic mode (mean neighbour distance / squared distance of the k-th neighbour, b2_normals.cu);
// kc_lsor_classify (second pass of the filter); kt_* (implicit BVH over Morton-sorted triangles, 4 per leaf); kc_mesh_distance;
// kc_splats. The box lower bound used for pruning is reduced by the worst-case rounding of the closest-point computation, so a
// triangle whose fp32 distance could undercut the current minimum is never skipped.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/eth3d_b200.h"
#include "b2_bvh.cuh"
#include "b2_common.cuh"
#include "b2_knn.h"

namespace b2 {

// ---- LocalStatisticalOutlierRemoval, second pass (:122-170) ----------------------------------------------------------------------------
// keep[i] = 1 for an inlier. distances[] are the first-pass means (float); sums in double, in neighbour order, as the reference.
__global__ void __launch_bounds__(256) kc_lsor_classify(const int* __restrict__ idx, const float* __restrict__ distances, size_t n, int k,
                                                        double factor, int negative, unsigned char* __restrict__ keep) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int valid = 0;
  double sum = 0.0;
  for (int a = 1; a < k; ++a) {            // a = 0 is the query point
    const int j = __ldg(&idx[i * (size_t)k + a]);
    if (j < 0) continue;
    const double distance = (double)__ldg(&distances[j]);
    if (distance > 0) { ++valid; sum += distance; }
  }
  const double mean = sum / (double)valid;
  const double distance_threshold = mean * factor;
  const double own = (double)distances[i];
  const bool removed = (!negative && own > distance_threshold) || (negative && own <= distance_threshold);
  keep[i] = removed ? 0 : 1;
}

// ---- triangles: implicit BVH over Morton-sorted triangles -----------------------------------------------------------------------------
static constexpr int kTriLeaf = 4;

__global__ void __launch_bounds__(256) kt_centroids(const float* __restrict__ v, const unsigned int* __restrict__ f, size_t nf, float* __restrict__ c) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nf) return;
  const unsigned int a = f[3 * t], b = f[3 * t + 1], d = f[3 * t + 2];
  for (int k = 0; k < 3; ++k) c[3 * t + k] = (v[3 * (size_t)a + k] + v[3 * (size_t)b + k] + v[3 * (size_t)d + k]) * (1.f / 3.f);
}
// sorted triangle j: three float4 (a.xyz, .w = bits(original face index)), (b.xyz, 0), (c.xyz, 0)
__global__ void __launch_bounds__(256) kt_gather(const float* __restrict__ v, const unsigned int* __restrict__ f, size_t nf,
                                                 const unsigned int* __restrict__ perm, float4* __restrict__ tris) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nf) return;
  const unsigned int t = perm[j];
  const unsigned int a = f[3 * (size_t)t], b = f[3 * (size_t)t + 1], c = f[3 * (size_t)t + 2];
  tris[3 * j] = make_float4(v[3 * (size_t)a], v[3 * (size_t)a + 1], v[3 * (size_t)a + 2], __uint_as_float(t));
  tris[3 * j + 1] = make_float4(v[3 * (size_t)b], v[3 * (size_t)b + 1], v[3 * (size_t)b + 2], 0.f);
  tris[3 * j + 2] = make_float4(v[3 * (size_t)c], v[3 * (size_t)c + 1], v[3 * (size_t)c + 2], 0.f);
}
__global__ void __launch_bounds__(256) kt_leaf_aabb(const float4* __restrict__ tris, size_t nf, unsigned int nleaf, Aabb* __restrict__ nodes) {
  const unsigned int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nleaf) return;
  Aabb b; for (int d = 0; d < 3; ++d) { b.lo[d] = INFINITY; b.hi[d] = -INFINITY; }
  const size_t e = min(nf, (size_t)(l + 1) * kTriLeaf);
  for (size_t p = (size_t)l * kTriLeaf * 3; p < e * 3; ++p) {
    const float4 v = tris[p];
    b.lo[0] = fminf(b.lo[0], v.x); b.lo[1] = fminf(b.lo[1], v.y); b.lo[2] = fminf(b.lo[2], v.z);
    b.hi[0] = fmaxf(b.hi[0], v.x); b.hi[1] = fmaxf(b.hi[1], v.y); b.hi[2] = fmaxf(b.hi[2], v.z);
  }
  nodes[l] = b;
}

__device__ __forceinline__ float dot3f(const float* a, const float* b) { return fadd(fadd(fmul(a[0], b[0]), fmul(a[1], b[1])), fmul(a[2], b[2])); }

// point_simplex_squared_distance<3> for a triangle (point_simplex_squared_distance.cpp:44-135; Ericson's closest point, Scalar = float,
// the literal 1.0 of `denom` makes that one division a double division).
__device__ __forceinline__ float point_triangle_sqr(const float p[3], const float4 A, const float4 B, const float4 C) {
  const float a[3] = {A.x, A.y, A.z}, b[3] = {B.x, B.y, B.z}, c[3] = {C.x, C.y, C.z};
  float ab[3], ac[3], ap[3], q[3];
  for (int k = 0; k < 3; ++k) { ab[k] = fsub(b[k], a[k]); ac[k] = fsub(c[k], a[k]); ap[k] = fsub(p[k], a[k]); }
  const float d1 = dot3f(ab, ap), d2 = dot3f(ac, ap);
  bool done = false;
  if (d1 <= 0.f && d2 <= 0.f) { for (int k = 0; k < 3; ++k) q[k] = a[k]; done = true; }
  float d3 = 0.f, d4 = 0.f, d5 = 0.f, d6 = 0.f, vc = 0.f, vb = 0.f;
  if (!done) {
    float bp[3]; for (int k = 0; k < 3; ++k) bp[k] = fsub(p[k], b[k]);
    d3 = dot3f(ab, bp); d4 = dot3f(ac, bp);
    if (d3 >= 0.f && d4 <= d3) { for (int k = 0; k < 3; ++k) q[k] = b[k]; done = true; }
  }
  if (!done) {
    vc = fsub(fmul(d1, d4), fmul(d3, d2));
    const bool a_ne_b = a[0] != b[0] || a[1] != b[1] || a[2] != b[2];
    if (a_ne_b && vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
      const float v = d1 / fsub(d1, d3);
      for (int k = 0; k < 3; ++k) q[k] = fadd(a[k], fmul(v, ab[k]));
      done = true;
    }
  }
  if (!done) {
    float cp[3]; for (int k = 0; k < 3; ++k) cp[k] = fsub(p[k], c[k]);
    d5 = dot3f(ab, cp); d6 = dot3f(ac, cp);
    if (d6 >= 0.f && d5 <= d6) { for (int k = 0; k < 3; ++k) q[k] = c[k]; done = true; }
  }
  if (!done) {
    vb = fsub(fmul(d5, d2), fmul(d1, d6));
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
      const float w = d2 / fsub(d2, d6);
      for (int k = 0; k < 3; ++k) q[k] = fadd(a[k], fmul(w, ac[k]));
      done = true;
    }
  }
  if (!done) {
    const float va = fsub(fmul(d3, d6), fmul(d5, d4));
    const float e43 = fsub(d4, d3), e56 = fsub(d5, d6);
    if (va <= 0.f && e43 >= 0.f && e56 >= 0.f) {
      const float w = e43 / fadd(e43, e56);
      for (int k = 0; k < 3; ++k) q[k] = fadd(b[k], fmul(w, fsub(c[k], b[k])));
    } else {
      const float denom = (float)(1.0 / (double)fadd(fadd(va, vb), vc));
      const float v = fmul(vb, denom), w = fmul(vc, denom);
      for (int k = 0; k < 3; ++k) q[k] = fadd(fadd(a[k], fmul(ab[k], v)), fmul(ac[k], w));
    }
  }
  float d[3]; for (int k = 0; k < 3; ++k) d[k] = fsub(p[k], q[k]);
  return dot3f(d, d);
}

struct TriBvhView { const float4* tris; size_t nf; const Aabb* nodes; BvhLevels lv; float slack; };

// Minimum of point_triangle_sqr over all triangles, or (stop_at >= 0) any value <= stop_at as soon as one is found.
// Pruning: a node is skipped only when (sqrt(box bound) - slack)^2 > best; slack bounds the rounding error of the closest point.
__device__ __forceinline__ float mesh_sqr_distance(const TriBvhView& m, const float p[3], float stop_at) {
  float best = INFINITY;
  unsigned int stack[2 * kBvhMaxLevels + 2];
  int sp = 0;
  stack[sp++] = ((unsigned int)(m.lv.nlevels - 1) << 27);
  while (sp > 0) {
    const unsigned int e = stack[--sp];
    const int level = (int)(e >> 27);
    const unsigned int i = e & 0x7FFFFFFu;
    {
      const float lb = dist2_box(p[0], p[1], p[2], m.nodes[m.lv.offset[level] + i]);
      const float s = sqrtf(lb) - m.slack;
      if (s > 0.f && s * s > best) continue;
    }
    if (level == 0) {
      const size_t b = (size_t)i * kTriLeaf, e2 = min(m.nf, b + kTriLeaf);
      for (size_t t = b; t < e2; ++t) {
        const float d = point_triangle_sqr(p, __ldg(&m.tris[3 * t]), __ldg(&m.tris[3 * t + 1]), __ldg(&m.tris[3 * t + 2]));
        best = fminf(best, d);
      }
      if (best <= stop_at) return best;
      continue;
    }
    const unsigned int c0 = 2 * i, c1 = 2 * i + 1;
    if (c1 >= m.lv.count[level - 1]) { stack[sp++] = ((unsigned int)(level - 1) << 27) | c0; continue; }
    const float d0 = dist2_box(p[0], p[1], p[2], m.nodes[m.lv.offset[level - 1] + c0]);
    const float d1 = dist2_box(p[0], p[1], p[2], m.nodes[m.lv.offset[level - 1] + c1]);
    if (d0 <= d1) { stack[sp++] = ((unsigned int)(level - 1) << 27) | c1; stack[sp++] = ((unsigned int)(level - 1) << 27) | c0; }   // near child on top
    else { stack[sp++] = ((unsigned int)(level - 1) << 27) | c0; stack[sp++] = ((unsigned int)(level - 1) << 27) | c1; }
  }
  return best;
}

__global__ void __launch_bounds__(128) kc_mesh_distance(TriBvhView m, const float* __restrict__ pts, size_t n, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  out[i] = mesh_sqr_distance(m, p, -1.f);
}

// SplatCreator per-point body (splat_creator.cc:146-215). radius2: squared distance to the 4th nearest other point (K7, statistic mode).
// corners: 4 x 3 floats per point (top right, bottom right, bottom left, top left); added[i] = 1 when the centre or one of the corners
// is farther than sqrt(thr2) from the mesh.
__global__ void __launch_bounds__(128) kc_splats(TriBvhView m, const float* __restrict__ xyz, const float* __restrict__ nrm, size_t n,
                                                 const float* __restrict__ radius2, float max_splat_size, float thr2, float* __restrict__ corners,
                                                 unsigned char* __restrict__ added, float* __restrict__ out_radius) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  const float nx = nrm[3 * i], ny = nrm[3 * i + 1], nz = nrm[3 * i + 2];
  if (out_radius) out_radius[i] = 0.f;
  if (isnan(nx) || isnan(ny) || isnan(nz)) {
    added[i] = 0;
    for (int k = 0; k < 12; ++k) corners[i * 12 + k] = 0.f;
    return;
  }
  const float splat_radius = fminf(sqrtf(radius2[i]), max_splat_size);
  if (out_radius) out_radius[i] = splat_radius;
  // Eigen::MatrixBase::unitOrthogonal() for 3-vectors (Eigen/src/Geometry/OrthoMethods.h): isMuchSmallerThan(a, b) = |a| <= |b| * 1e-5f
  float right[3];
  if (!(fabsf(nx) <= fmul(fabsf(nz), 1e-5f)) || !(fabsf(ny) <= fmul(fabsf(nz), 1e-5f))) {
    const float invnm = 1.f / sqrtf(fadd(fmul(nx, nx), fmul(ny, ny)));
    right[0] = fmul(-ny, invnm); right[1] = fmul(nx, invnm); right[2] = 0.f;
  } else {
    const float invnm = 1.f / sqrtf(fadd(fmul(ny, ny), fmul(nz, nz)));
    right[0] = 0.f; right[1] = fmul(-nz, invnm); right[2] = fmul(ny, invnm);
  }
  // up = normal x right
  const float up[3] = {fsub(fmul(ny, right[2]), fmul(nz, right[1])), fsub(fmul(nz, right[0]), fmul(nx, right[2])), fsub(fmul(nx, right[1]), fmul(ny, right[0]))};
  float c[4][3];
  for (int k = 0; k < 3; ++k) {
    c[0][k] = fadd(p[k], fmul(splat_radius, fadd(right[k], up[k])));
    c[1][k] = fadd(p[k], fmul(splat_radius, fsub(right[k], up[k])));
    c[2][k] = fadd(p[k], fmul(splat_radius, fsub(-right[k], up[k])));
    c[3][k] = fadd(p[k], fmul(splat_radius, fadd(-right[k], up[k])));
  }
  for (int v = 0; v < 4; ++v) for (int k = 0; k < 3; ++k) corners[(i * 4 + v) * 3 + k] = c[v][k];
  bool add = mesh_sqr_distance(m, p, thr2) > thr2;
  for (int v = 0; v < 4 && !add; ++v) add = mesh_sqr_distance(m, c[v], thr2) > thr2;
  added[i] = add ? 1 : 0;
}

struct TriBvh {
  DevBuf verts, faces, cent, part, keys, keys2, idx, perm, tris, nodes, tmp;
  PinnedBuf ppart;
  BvhLevels lv;
  size_t nf = 0;
  float slack = 0.f;
  int build(const float* vertices, size_t nv, const uint32_t* face_idx, size_t num_faces, int sms, cudaStream_t st) {
    nf = num_faces;
    for (size_t t = 0; t < 3 * nf; ++t) if (face_idx[t] >= nv) return set_error(B2_ERR_ARG, "face %zu references vertex %u of %zu", t / 3, face_idx[t], nv);
    B2_TRY(verts.ensure(nv * 12)); B2_TRY(faces.ensure(nf * 12)); B2_TRY(cent.ensure(nf * 12));
    B2_CUDA(cudaMemcpyAsync(verts.p, vertices, nv * 12, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(faces.p, face_idx, nf * 12, cudaMemcpyHostToDevice, st));
    kt_centroids<<<bvh_div_up(nf, 256), 256, 0, st>>>(verts.as<float>(), faces.as<unsigned int>(), nf, cent.as<float>());
    const int bb = sms * 2;
    B2_TRY(part.ensure(sizeof(float) * 6 * bb)); B2_TRY(ppart.ensure(sizeof(float) * 6 * bb));
    kn_bbox<<<bb, 256, 0, st>>>(verts.as<float>(), nv, part.as<float>());
    B2_CUDA(cudaMemcpyAsync(ppart.p, part.p, sizeof(float) * 6 * bb, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int b = 0; b < bb; ++b) for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], ppart.as<float>()[6 * b + d]); mx[d] = std::max(mx[d], ppart.as<float>()[6 * b + 3 + d]);
    }
    float amax = 0.f;
    for (int d = 0; d < 3; ++d) {
      if (!std::isfinite(mn[d]) || !std::isfinite(mx[d])) return set_error(B2_ERR_ARG, "non-finite mesh vertices");
      amax = std::max({amax, std::fabs(mn[d]), std::fabs(mx[d])});
    }
    // closest point q = a + v*ab + w*ac: a handful of roundings at the magnitude of the coordinates / edge lengths
    slack = 64.f * 1.1920929e-7f * std::max(amax, 1e-30f);
    const float ext = std::max({mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2], 1e-30f});
    const float scale = 2097151.f / ext;
    B2_TRY(keys.ensure(nf * 8)); B2_TRY(keys2.ensure(nf * 8)); B2_TRY(idx.ensure(nf * 4)); B2_TRY(perm.ensure(nf * 4)); B2_TRY(tris.ensure(nf * 48));
    kn_morton<<<bvh_div_up(nf, 256), 256, 0, st>>>(cent.as<float>(), nf, mn[0], mn[1], mn[2], scale, keys.as<unsigned long long>(), idx.as<unsigned int>());
    size_t t = 0;
    B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), idx.as<unsigned int>(),
                                            perm.as<unsigned int>(), (long long)nf, 0, 63, st));
    B2_TRY(tmp.ensure(t));
    B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, t, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), idx.as<unsigned int>(),
                                            perm.as<unsigned int>(), (long long)nf, 0, 63, st));
    kt_gather<<<bvh_div_up(nf, 256), 256, 0, st>>>(verts.as<float>(), faces.as<unsigned int>(), nf, perm.as<unsigned int>(), tris.as<float4>());
    std::memset(&lv, 0, sizeof(lv));
    unsigned int cnt = bvh_div_up(nf, kTriLeaf), off = 0; int L = 0;
    while (true) { lv.offset[L] = off; lv.count[L] = cnt; off += cnt; ++L; if (cnt == 1) break; cnt = (cnt + 1) / 2; }
    lv.nlevels = L;
    B2_TRY(nodes.ensure(sizeof(Aabb) * (size_t)off));
    kt_leaf_aabb<<<bvh_div_up(lv.count[0], 256), 256, 0, st>>>(tris.as<float4>(), nf, lv.count[0], nodes.as<Aabb>());
    for (int l = 1; l < L; ++l)
      kn_merge_level<<<bvh_div_up(lv.count[l], 256), 256, 0, st>>>(nodes.as<Aabb>() + lv.offset[l - 1], lv.count[l - 1], nodes.as<Aabb>() + lv.offset[l], lv.count[l]);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
  }
  TriBvhView view(float query_amax) const {
    // the query's own magnitude enters p - q as well
    return TriBvhView{tris.as<float4>(), nf, nodes.as<Aabb>(), lv, std::max(slack, 64.f * 1.1920929e-7f * query_amax)};
  }
  void release() { for (DevBuf* b : {&verts, &faces, &cent, &part, &keys, &keys2, &idx, &perm, &tris, &nodes, &tmp}) b->release(); ppart.release(); }
};

static float abs_max(const float* xyz, size_t n, size_t stride_floats) {
  float m = 0.f;
  for (size_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) { const float v = std::fabs(xyz[i * stride_floats + d]); if (v > m) m = v; }
  return m;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_lsor_filter(const float* xyz, size_t n, size_t stride_bytes, int mean_k, double distance_factor_threshold, int negative,
                              int32_t* out_indices, size_t* out_count, int32_t* out_removed_indices, size_t* out_removed_count,
                              float* out_mean_distances) {
  if ((n && !xyz) || !out_indices || !out_count) return set_error(B2_ERR_ARG, "null argument");
  if (stride_bytes < 12 || stride_bytes % 4) return set_error(B2_ERR_ARG, "stride_bytes must be a multiple of 4, >= 12");
  if (mean_k < 1 || mean_k > 2047) return set_error(B2_ERR_ARG, "mean_k must be in [1,2047]");
  if (n >= (1ull << 30)) return set_error(B2_ERR_ARG, "clouds above 2^30 points are not supported");
  *out_count = 0;
  if (out_removed_count) *out_removed_count = 0;
  if (n == 0) return B2_OK;
  const size_t sf = stride_bytes / 4;
  const int k = mean_k + 1;
  // KdTreeFLANN indexes the finite points only; the common (dense) case passes the caller's buffer straight through
  std::vector<float> fin; std::vector<int32_t> orig;
  size_t first_bad = n;
  for (size_t i = 0; i < n; ++i) {
    const float* p = xyz + i * sf;
    if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) { first_bad = i; break; }
  }
  const float* pts = xyz; size_t m = n, pts_stride = stride_bytes;
  if (first_bad < n) {
    fin.reserve(3 * n); orig.reserve(n);
    for (size_t i = 0; i < n; ++i) {
      const float* p = xyz + i * sf;
      if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) { fin.insert(fin.end(), p, p + 3); orig.push_back((int32_t)i); }
    }
    pts = fin.data(); m = orig.size(); pts_stride = 12;
  }
  if (m > 0 && m < (size_t)k) return set_error(B2_ERR_STATE, "cloud has %zu finite points, fewer than mean_k + 1 = %d", m, k);
  std::vector<unsigned char> keep(m);
  std::vector<float> dist(out_mean_distances ? m : 0);
  if (m > 0) {
    DevBuf d_keep;
    KnnHook hook;
    hook.stat_mode = kKnnStatMeanDistance;
    hook.need_idx = true;
    const bool trace = std::getenv("B2_CLEAN_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    hook.run = [&](cudaStream_t st, const float*, const int* idx_dev, const float* stat_dev, size_t cnt, int kk) -> int {
      if (trace) {
        B2_CUDA(cudaStreamSynchronize(st));
        fprintf(stderr, "[b2_lsor_filter] %zu points, k = %d: upload + index + kNN %.1f ms\n", cnt, kk,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
      const auto t1 = std::chrono::steady_clock::now();
      struct Tail { bool on; std::chrono::steady_clock::time_point t; ~Tail() { if (on) fprintf(stderr, "[b2_lsor_filter] classify + download %.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count()); } } tail{trace, t1};
      B2_TRY(d_keep.ensure(cnt));
      kc_lsor_classify<<<bvh_div_up(cnt, 256), 256, 0, st>>>(idx_dev, stat_dev, cnt, kk, distance_factor_threshold, negative, d_keep.as<unsigned char>());
      B2_CUDA(cudaGetLastError());
      B2_CUDA(cudaMemcpyAsync(keep.data(), d_keep.p, cnt, cudaMemcpyDeviceToHost, st));
      if (out_mean_distances) B2_CUDA(cudaMemcpyAsync(dist.data(), stat_dev, cnt * 4, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      return B2_OK;
    };
    const int rc = knn_with_hook(pts, m, pts_stride, k, hook);
    d_keep.release();
    if (rc != B2_OK) return rc;
  }
  size_t oii = 0, rii = 0, c = 0;
  for (size_t i = 0; i < n; ++i) {
    const bool finite = first_bad == n || (c < m && (size_t)orig[c] == i);
    if (out_mean_distances) out_mean_distances[i] = finite ? dist[c] : 0.f;
    if (finite && keep[c]) out_indices[oii++] = (int32_t)i;
    else { if (out_removed_indices) out_removed_indices[rii] = (int32_t)i; ++rii; }
    if (finite) ++c;
  }
  *out_count = oii;
  if (out_removed_count) *out_removed_count = rii;
  return B2_OK;
}

extern "C" int b2_mesh_squared_distance(const float* points, size_t n, const float* vertices, size_t num_vertices, const uint32_t* faces,
                                        size_t num_faces, float* out_squared_distance) {
  if ((n && (!points || !out_squared_distance)) || !vertices || !faces) return set_error(B2_ERR_ARG, "null argument");
  if (num_faces == 0 || num_vertices == 0) return set_error(B2_ERR_ARG, "empty mesh");
  if (num_faces >= (1ull << 27) || n >= (1ull << 31)) return set_error(B2_ERR_ARG, "mesh / point set too large");
  if (n == 0) return B2_OK;
  int dev = 0, sms = 0;
  B2_TRY(select_device(-1, &dev, &sms));
  cudaStream_t st; B2_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  TriBvh bvh; DevBuf d_pts, d_out;
  auto body = [&]() -> int {
    B2_TRY(bvh.build(vertices, num_vertices, faces, num_faces, sms, st));
    B2_TRY(d_pts.ensure(n * 12)); B2_TRY(d_out.ensure(n * 4));
    B2_CUDA(cudaMemcpyAsync(d_pts.p, points, n * 12, cudaMemcpyHostToDevice, st));
    kc_mesh_distance<<<bvh_div_up(n, 128), 128, 0, st>>>(bvh.view(abs_max(points, n, 3)), d_pts.as<float>(), n, d_out.as<float>());
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaMemcpyAsync(out_squared_distance, d_out.p, n * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    return B2_OK;
  };
  const int rc = body();
  bvh.release(); d_pts.release(); d_out.release();
  cudaStreamDestroy(st);
  return rc;
}

extern "C" int b2_splat_create(const float* xyz, const float* normals, size_t n, size_t stride_bytes, const float* vertices, size_t num_vertices,
                               const uint32_t* faces, size_t num_faces, float max_splat_size, float squared_distance_threshold,
                               float* out_corners, uint8_t* out_added, float* out_radius, size_t* out_splat_count) {
  if ((n && (!xyz || !normals || !out_corners || !out_added)) || !vertices || !faces) return set_error(B2_ERR_ARG, "null argument");
  if (stride_bytes < 12 || stride_bytes % 4) return set_error(B2_ERR_ARG, "stride_bytes must be a multiple of 4, >= 12");
  if (num_faces == 0 || num_vertices == 0) return set_error(B2_ERR_ARG, "empty mesh");
  if (num_faces >= (1ull << 27)) return set_error(B2_ERR_ARG, "mesh too large");
  if (out_splat_count) *out_splat_count = 0;
  if (n == 0) return B2_OK;
  constexpr int kNearestNeighborCount = 4;                       // splat_creator.cc:124
  if (n < (size_t)kNearestNeighborCount + 1) return set_error(B2_ERR_STATE, "cloud has %zu points, fewer than 5 (reference: CHECK_EQ, splat_creator.cc:164)", n);
  int dev = 0, sms = 0;
  B2_TRY(select_device(-1, &dev, &sms));
  const size_t sf = stride_bytes / 4;
  TriBvh bvh; DevBuf d_nrm, d_corners, d_added, d_radius;
  int rc = B2_OK;
  // magnitude bound of every query of the mesh test (the points and their splat corners: a neighbour distance never exceeds the cloud's
  // extent <= 2 sqrt(3) amax, a corner lies sqrt(2) radii from its point) -> the pruning margin of the triangle walk
  const float amax = abs_max(xyz, n, sf);
  const float corner_bound = amax + 1.5f * std::min(max_splat_size, 3.5f * amax) + 1e-30f;
  KnnHook hook;
  hook.stat_mode = kKnnStatLastD2;
  hook.need_idx = false;
  hook.run = [&](cudaStream_t st, const float* xyz_dev, const int*, const float* stat_dev, size_t cnt, int) -> int {
    B2_TRY(bvh.build(vertices, num_vertices, faces, num_faces, sms, st));
    B2_TRY(d_nrm.ensure(cnt * 12)); B2_TRY(d_corners.ensure(cnt * 48)); B2_TRY(d_added.ensure(cnt));
    if (out_radius) B2_TRY(d_radius.ensure(cnt * 4));
    if (stride_bytes == 12) B2_CUDA(cudaMemcpyAsync(d_nrm.p, normals, cnt * 12, cudaMemcpyHostToDevice, st));
    else B2_CUDA(cudaMemcpy2DAsync(d_nrm.p, 12, normals, stride_bytes, 12, cnt, cudaMemcpyHostToDevice, st));
    kc_splats<<<bvh_div_up(cnt, 128), 128, 0, st>>>(bvh.view(corner_bound), xyz_dev, d_nrm.as<float>(), cnt, stat_dev, max_splat_size,
                                                    squared_distance_threshold, d_corners.as<float>(), d_added.as<unsigned char>(),
                                                    out_radius ? d_radius.as<float>() : nullptr);
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaMemcpyAsync(out_corners, d_corners.p, cnt * 48, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(out_added, d_added.p, cnt, cudaMemcpyDeviceToHost, st));
    if (out_radius) B2_CUDA(cudaMemcpyAsync(out_radius, d_radius.p, cnt * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    return B2_OK;
  };
  rc = knn_with_hook(xyz, n, stride_bytes, kNearestNeighborCount + 1, hook);
  bvh.release(); for (DevBuf* b : {&d_nrm, &d_corners, &d_added, &d_radius}) b->release();
  if (rc != B2_OK) return rc;
  if (out_splat_count) { size_t c = 0; for (size_t i = 0; i < n; ++i) c += out_added[i] ? 1 : 0; *out_splat_count = c; }
  return B2_OK;
}
