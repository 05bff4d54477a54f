// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the point-cloud tools next to the hot paths:
//   * pcl::LocalStatisticalOutlierRemoval<PointT>::applyFilterIndices
//     (/root/reference/src/geometry/local_statistical_outlier_removal.hpp:72-176) as driven by PointCloudCleaner
//     (/root/reference/src/exe/point_cloud_cleaner.cc:80-96: setMeanK(knn), setDistanceFactorThresh(factor), filter);
//   * the splat geometry of SplatCreator (/root/reference/src/exe/splat_creator.cc:118-215): radius from the 4th nearest
//     neighbour, right / up from Eigen's unitOrthogonal(), four corners, and the "not represented by the mesh" test.
// Parity status: UNPINNED — the reference has no test for either tool. Third-party arithmetic restated from memory:
//   PCL 1.10 KdTreeFLANN::nearestKSearch (exact kNN, sorted; tie order unspecified in FLANN — here (d2, index) ascending; non-finite
//   points are not part of the tree), `sqrt (nn_dists[k])` on a float resolves to the float overload (libstdc++'s <math.h> wrapper is
//   in scope in PCL translation units), Eigen 3.3 MatrixBase::unitOrthogonal() (src/Geometry/OrthoMethods.h, 3-vector branch),
//   libigl is vendored (thirdparty/igl): point_simplex_squared_distance is followed operation by operation; AABB::squared_distance is
//   restated as the minimum over all triangles (libigl's tree prunes on `box distance < current minimum` without a rounding margin, so
//   its own answer can differ from that minimum by rounding in near-tie configurations).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "orc_api.h"
#include "orc_kdtree.h"

namespace {

inline bool finite3(const float* p) { return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); }

}  // namespace

// applyFilterIndices with indices_ = the whole cloud (what Filter::filter sets up), extract_removed_indices_ = true.
// out_indices / out_removed: capacity n each; returns 0, or 1 when the cloud has fewer than mean_k + 1 finite points (the reference
// reads nn_dists beyond what the search filled: undefined).
extern "C" int orc_lsor_filter(const float* xyz, size_t n, int mean_k, double distance_factor_threshold, int negative, int32_t* out_indices,
                               uint64_t* out_count, int32_t* out_removed, uint64_t* out_removed_count, float* out_distances /* n, nullable */) {
  // KdTreeFLANN indexes only the finite points (convertCloudToArray) and maps results back to cloud indices
  std::vector<float> fin; std::vector<int> orig;
  fin.reserve(3 * n); orig.reserve(n);
  for (size_t i = 0; i < n; ++i) if (finite3(xyz + 3 * i)) { fin.insert(fin.end(), xyz + 3 * i, xyz + 3 * i + 3); orig.push_back((int)i); }
  const size_t m = orig.size();
  const int k = mean_k + 1;
  if (mean_k < 1 || (m > 0 && m < (size_t)k)) return 1;
  orc::KdTree tree;
  tree.build(fin.data(), m, 3, 15);
  std::vector<float> distances(n, 0.f);
  std::vector<int> nn((size_t)m * k);
  // First pass (:92-120): mean distance to the mean_k nearest neighbours
#pragma omp parallel
  {
    std::vector<int> idx(k);
    std::vector<float> d2(k);
#pragma omp for schedule(dynamic, 1024)
    for (long long c = 0; c < (long long)m; ++c) {
      tree.knn(fin.data() + 3 * c, k, idx.data(), d2.data());
      double dist_sum = 0.0;
      for (int j = 1; j < k; ++j) dist_sum += std::sqrt(d2[j]);   // k = 0 is the query point; float sqrt, double sum
      distances[orig[c]] = static_cast<float>(dist_sum / mean_k);
      for (int j = 0; j < k; ++j) nn[(size_t)c * k + j] = orig[idx[j]];
    }
  }
  // Second pass (:122-170)
  uint64_t oii = 0, rii = 0;
  size_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    const bool problematic = !finite3(xyz + 3 * i);
    if (problematic) {
      // (!negative_: removed. negative_: the reference falls through with the previous point's neighbour list — not restated; such
      // points are reported as removed here as well.)
      out_removed[rii++] = (int32_t)i;
      continue;
    }
    int valid = 0;
    double sum = 0;
    for (int j = 1; j < k; ++j) {
      const double distance = distances[nn[c * k + j]];
      if (distance > 0) { ++valid; sum += distance; }
    }
    ++c;
    const double mean = sum / static_cast<double>(valid);
    const double distance_threshold = mean * distance_factor_threshold;
    if ((!negative && distances[i] > distance_threshold) || (negative && distances[i] <= distance_threshold)) {
      out_removed[rii++] = (int32_t)i;
      continue;
    }
    out_indices[oii++] = (int32_t)i;
  }
  *out_count = oii; *out_removed_count = rii;
  if (out_distances) std::copy(distances.begin(), distances.end(), out_distances);
  return 0;
}

namespace {

inline float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// igl::point_simplex_squared_distance<3> for a triangle (/root/reference/thirdparty/igl/point_simplex_squared_distance.cpp:44-135),
// Scalar = float (SplatCreator converts the mesh to Eigen::MatrixXf, splat_creator.cc:48-73).
float point_triangle_sqr(const float* p, const float* a, const float* b, const float* c) {
  float ab[3], ac[3], ap[3], q[3];
  for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
  const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  auto sqr = [&](const float* x) { float d[3] = {p[0] - x[0], p[1] - x[1], p[2] - x[2]}; return dot3(d, d); };
  if (d1 <= 0.0 && d2 <= 0.0) return sqr(a);
  float bp[3]; for (int k = 0; k < 3; ++k) bp[k] = p[k] - b[k];
  const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0.0 && d4 <= d3) return sqr(b);
  const float vc = d1 * d4 - d3 * d2;
  if (a[0] != b[0] || a[1] != b[1] || a[2] != b[2]) {
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
      const float v = d1 / (d1 - d3);
      for (int k = 0; k < 3; ++k) q[k] = a[k] + v * ab[k];
      return sqr(q);
    }
  }
  float cp[3]; for (int k = 0; k < 3; ++k) cp[k] = p[k] - c[k];
  const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0.0 && d5 <= d6) return sqr(c);
  const float vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
    const float w = d2 / (d2 - d6);
    for (int k = 0; k < 3; ++k) q[k] = a[k] + w * ac[k];
    return sqr(q);
  }
  const float va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
    const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    for (int k = 0; k < 3; ++k) q[k] = b[k] + w * (c[k] - b[k]);
    return sqr(q);
  }
  const float denom = (float)(1.0 / (double)((va + vb) + vc));
  const float v = vb * denom, w = vc * denom;
  for (int k = 0; k < 3; ++k) q[k] = (a[k] + ab[k] * v) + ac[k] * w;
  return sqr(q);
}

float mesh_sqr(const float* p, const float* V, const uint32_t* F, size_t nf) {
  float best = std::numeric_limits<float>::infinity();
  for (size_t t = 0; t < nf; ++t) best = std::min(best, point_triangle_sqr(p, V + 3 * (size_t)F[3 * t], V + 3 * (size_t)F[3 * t + 1], V + 3 * (size_t)F[3 * t + 2]));
  return best;
}

}  // namespace

// igl::AABB::squared_distance as the minimum over ALL triangles (the tree only prunes; AABB.cpp:357-430).
extern "C" void orc_mesh_squared_distance(const float* points, size_t n, const float* vertices, const uint32_t* faces, size_t nf, float* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (long long i = 0; i < (long long)n; ++i) out[i] = mesh_sqr(points + 3 * i, vertices, faces, nf);
}

// SplatCreator (splat_creator.cc:118-215) per point, in index order (the reference's `omp parallel for` appends in arrival order).
// corners: n x 4 x 3 (top right, bottom right, bottom left, top left; written for every point with a valid normal), added: n flags,
// radius: n (nullable). Returns the number of splats.
extern "C" uint64_t orc_splat_create(const float* xyz, const float* normals, size_t n, const float* vertices, const uint32_t* faces, size_t nf,
                                     float max_splat_size, float squared_distance_threshold, float* corners, uint8_t* added, float* radius) {
  constexpr int kNearestNeighborCount = 4;
  orc::KdTree tree;
  tree.build(xyz, n, 3, 15);
  uint64_t count = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : count)
  for (long long i = 0; i < (long long)n; ++i) {
    const float* p = xyz + 3 * i; const float* nr = normals + 3 * i;
    added[i] = 0;
    if (radius) radius[i] = 0.f;
    for (int k = 0; k < 12; ++k) corners[(size_t)i * 12 + k] = 0.f;
    if (std::isnan(nr[0]) || std::isnan(nr[1]) || std::isnan(nr[2])) continue;
    int idx[kNearestNeighborCount + 1]; float d2[kNearestNeighborCount + 1];
    tree.knn(p, kNearestNeighborCount + 1, idx, d2);
    const float splat_radius = std::min(sqrtf(d2[kNearestNeighborCount]), max_splat_size);
    if (radius) radius[i] = splat_radius;
    // Eigen::MatrixBase::unitOrthogonal(), 3-vector branch (Eigen 3.3 src/Geometry/OrthoMethods.h): dummy_precision<float>() = 1e-5f
    float right[3];
    if (!(std::fabs(nr[0]) <= std::fabs(nr[2]) * 1e-5f) || !(std::fabs(nr[1]) <= std::fabs(nr[2]) * 1e-5f)) {
      const float invnm = 1.f / std::sqrt(nr[0] * nr[0] + nr[1] * nr[1]);
      right[0] = -nr[1] * invnm; right[1] = nr[0] * invnm; right[2] = 0.f;
    } else {
      const float invnm = 1.f / std::sqrt(nr[1] * nr[1] + nr[2] * nr[2]);
      right[0] = 0.f; right[1] = -nr[2] * invnm; right[2] = nr[1] * invnm;
    }
    const float up[3] = {nr[1] * right[2] - nr[2] * right[1], nr[2] * right[0] - nr[0] * right[2], nr[0] * right[1] - nr[1] * right[0]};
    float c[4][3];
    for (int k = 0; k < 3; ++k) {
      c[0][k] = p[k] + splat_radius * (right[k] + up[k]);
      c[1][k] = p[k] + splat_radius * (right[k] - up[k]);
      c[2][k] = p[k] + splat_radius * (-right[k] - up[k]);
      c[3][k] = p[k] + splat_radius * (-right[k] + up[k]);
    }
    for (int v = 0; v < 4; ++v) for (int k = 0; k < 3; ++k) corners[((size_t)i * 4 + v) * 3 + k] = c[v][k];
    bool add = mesh_sqr(p, vertices, faces, nf) > squared_distance_threshold;
    for (int v = 0; v < 4 && !add; ++v) add = mesh_sqr(c[v], vertices, faces, nf) > squared_distance_threshold;
    added[i] = add ? 1 : 0;
    count += add ? 1 : 0;
  }
  return count;
}
