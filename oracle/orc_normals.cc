// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see orc_math.h header).
//
// CPU restatement of the reference's kNN two-pass normal estimation (Path A, rows A12/A13):
//   computeMeanAndCovarianceMatrixTwoPass  /root/reference/src/geometry/two_pass_centroid.hpp:155-259
//   computePointNormalTwoPass              /root/reference/src/geometry/two_pass_normal_3d.h:92-109
//   NormalEstimationTwoPassOMP::computeFeature /root/reference/src/geometry/two_pass_normal_3d_omp.hpp:47-119
// plus PCL 1.10 pieces NOT under /root/reference, restated from the published algorithm (from memory of
// pcl/common/impl/eigen.hpp `eigen33`/`computeRoots`, pcl/features/normal_3d.h `solvePlaneParameters`,
// `flipNormalTowardsViewpoint`): PARITY UNPINNED — the reference has no test for normals at all.
// kNN = exact k nearest including the query point itself, sorted by (d2, index) (kd-tree in orc_kdtree.h).
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "orc_api.h"
#include "orc_kdtree.h"

namespace orc {

static void roots2(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  float d = b * b - 4.f * c;
  if (d < 0.f) d = 0.f;
  const float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

// pcl::computeRoots for a symmetric 3x3 (row-major m[9]).
static void compute_roots(const float m[9], float roots[3]) {
  const float c0 = m[0] * m[4] * m[8] + 2.f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
  const float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  const float c2 = m[0] + m[4] + m[8];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
    roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = 1.0f / 3.0f;
  const float s_sqrt3 = std::sqrt(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f) q = 0.f;
  const float rho = std::sqrt(-a_over_3);
  const float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  const float cos_theta = std::cos(theta);
  const float sin_theta = std::sin(theta);
  roots[0] = c2_over_3 + 2.f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0.f) roots2(c2, c1, roots);
}

static inline void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline float sqn3(const float* a) { return a[0] * a[0] + (a[1] * a[1] + a[2] * a[2]); }

// pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenvalue + its eigenvector.
static void eigen33_smallest(const float cov[9], float* eigenvalue, float evec[3]) {
  float scale = 0.f;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(cov[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.f;
  float s[9];
  for (int i = 0; i < 9; ++i) s[i] = cov[i] / scale;
  float roots[3];
  compute_roots(s, roots);
  *eigenvalue = roots[0] * scale;
  s[0] -= roots[0]; s[4] -= roots[0]; s[8] -= roots[0];
  float v1[3], v2[3], v3[3];
  cross3(s + 0, s + 3, v1); cross3(s + 0, s + 6, v2); cross3(s + 3, s + 6, v3);
  const float l1 = sqn3(v1), l2 = sqn3(v2), l3 = sqn3(v3);
  const float* v; float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  else { v = v3; l = l3; }
  const float inv = std::sqrt(l);
  evec[0] = v[0] / inv; evec[1] = v[1] / inv; evec[2] = v[2] / inv;
}

// two_pass_centroid.hpp:164-258 (dense branch): pass 1 mean, pass 2 centred products; sequential fp32 sums in list order.
static void mean_cov_two_pass(const float* xyz, const int* idx, int count, float cov[9], float centroid[3]) {
  float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < count; ++i) { const float* p = xyz + 3 * (size_t)idx[i]; a[6] += p[0]; a[7] += p[1]; a[8] += p[2]; }
  for (int i = 0; i < 9; ++i) a[i] = a[i] / (float)count;
  for (int i = 0; i < count; ++i) {
    const float* p = xyz + 3 * (size_t)idx[i];
    a[0] += (p[0] - a[6]) * (p[0] - a[6]);
    a[1] += (p[0] - a[6]) * (p[1] - a[7]);
    a[2] += (p[0] - a[6]) * (p[2] - a[8]);
    a[3] += (p[1] - a[7]) * (p[1] - a[7]);
    a[4] += (p[1] - a[7]) * (p[2] - a[8]);
    a[5] += (p[2] - a[8]) * (p[2] - a[8]);
  }
  centroid[0] = a[6]; centroid[1] = a[7]; centroid[2] = a[8];
  const float n = (float)count;
  cov[0] = a[0] / n; cov[1] = a[1] / n; cov[2] = a[2] / n; cov[4] = a[3] / n; cov[5] = a[4] / n; cov[8] = a[5] / n;
  cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
}

void point_normal(const float* xyz, const int* idx, int count, const float* query, const float vp[3], float out[4]) {
  const float nan = std::numeric_limits<float>::quiet_NaN();
  if (count < 3) { out[0] = out[1] = out[2] = out[3] = nan; return; }   // two_pass_normal_3d.h:100-105
  float cov[9], cen[3];
  mean_cov_two_pass(xyz, idx, count, cov, cen);
  float ev, n[3];
  eigen33_smallest(cov, &ev, n);
  const float eig_sum = cov[0] + cov[4] + cov[8];
  out[3] = (eig_sum != 0.f) ? std::fabs(ev / eig_sum) : 0.f;
  // flipNormalTowardsViewpoint
  const float vx = vp[0] - query[0], vy = vp[1] - query[1], vz = vp[2] - query[2];
  const float cos_theta = (vx * n[0] + vy * n[1] + vz * n[2]);
  if (cos_theta < 0.f) { n[0] *= -1.f; n[1] *= -1.f; n[2] *= -1.f; }
  out[0] = n[0]; out[1] = n[1]; out[2] = n[2];
}

}  // namespace orc

extern "C" int orc_normals_knn(const float* xyz, size_t n, int k, const float viewpoint[3], float* out, int* out_idx) {
  orc::KdTree tree;
  tree.build(xyz, n, 3, 15);
#pragma omp parallel
  {
    std::vector<int> idx(k);
    std::vector<float> d2(k);
#pragma omp for schedule(dynamic, 1024)
    for (long long i = 0; i < (long long)n; ++i) {
      const int found = tree.knn(xyz + 3 * i, k, idx.data(), d2.data());
      if (out_idx) for (int j = 0; j < k; ++j) out_idx[(size_t)i * k + j] = j < found ? idx[j] : -1;
      orc::point_normal(xyz, idx.data(), found, xyz + 3 * i, viewpoint, out + 4 * i);
    }
  }
  return 0;
}

// setRadiusSearch mode (two_pass_normal_3d_omp.hpp:66 with search_parameter_ = radius; normal_estimator.cc:181-182): ALL points with
// d2 < (float)((double)r*r) (self included), in the order radiusSearch returns them (sorted by distance; ties to the lower index).
extern "C" int orc_normals_radius(const float* xyz, size_t n, float radius, const float viewpoint[3], float* out, int* out_count) {
  orc::KdTree tree;
  tree.build(xyz, n, 3, 15);
  const float r2 = (float)((double)radius * (double)radius);
#pragma omp parallel
  {
    std::vector<std::pair<float, int>> found;
    std::vector<int> idx;
#pragma omp for schedule(dynamic, 1024)
    for (long long i = 0; i < (long long)n; ++i) {
      tree.radius(xyz + 3 * i, r2, &found);
      std::sort(found.begin(), found.end());
      idx.resize(found.size());
      for (size_t j = 0; j < found.size(); ++j) idx[j] = found[j].second;
      if (out_count) out_count[i] = (int)found.size();
      orc::point_normal(xyz, idx.data(), (int)idx.size(), xyz + 3 * i, viewpoint, out + 4 * i);
    }
  }
  return 0;
}
