// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see orc_math.h header).
//
// CPU restatement of the multi-resolution point-cloud construction that produces Path B's inputs:
//   MergeClosePoints             /root/reference/src/opt/multi_scale_point_cloud.cc:44-124
//   CreateMultiScalePointCloud   /root/reference/src/opt/multi_scale_point_cloud.cc:263-368 (the scale loop, given the per-point
//                                 min / max radii of ComputeMinMaxPointRadius, :126-184)
//   Problem::DeterminePointNeighbors  /root/reference/src/opt/problem.cc:706-786
//
// Third-party behaviour NOT under /root/reference and restated here (PARITY UNPINNED unless noted):
//   * pcl::search::KdTree::radiusSearch(point, r, idx, d2, 0) (PCL 1.10 -> FLANN 1.9.1): every point with
//     d2 = ((dx*dx)+(dy*dy))+(dz*dz) < (float)((double)r*r), sorted by distance; ties are FLANN-internal, defined here as
//     ascending index. nearestKSearch: the k nearest sorted by (d2, index). The k-NN index SETS are pinned by the reference's
//     test_problem.cc:35-109 (tests/test_oracle_multiscale.py).
//   * std::shuffle + std::uniform_int_distribution with std::mt19937(0) (problem.cc:712,754,777): the permutation depends on the
//     libstdc++ version. The reference's Dockerfile pins Ubuntu 20.04 = GCC 9.3; its algorithm (bits/stl_algo.h `shuffle` with the
//     two-swaps-per-draw path, bits/uniform_int_dist.h down-scaling by rejection) is restated explicitly in gcc9_shuffle() below,
//     from memory of those headers — test_problem.cc does not observe the permutation (candidate count == neighbour count there).
//   * Eigen 3.3: Vector3f += is component-wise fp32 addition; `v /= int` is a component-wise division by (float)int.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <vector>

#include "orc_api.h"
#include "orc_kdtree.h"

namespace orc {

// multi_scale_point_cloud.cc:44-124. Outputs are appended (the reference clears points / colors / scan indices, :54-56, but NOT
// out_max_radius; CreateMultiScalePointCloud always passes an empty vector, so the difference is unobservable there).
static void merge_close_points(float merge_distance, int num_scans, const std::vector<float>& xyz, const std::vector<float>& colors,
                               const std::vector<uint8_t>& scan, const std::vector<float>& max_radius_in, std::vector<float>* oxyz,
                               std::vector<float>* ocol, std::vector<uint8_t>* oscan, std::vector<float>* omaxr) {
  const size_t n = colors.size();
  KdTree tree;
  tree.build(xyz.data(), n);
  const float r2 = (float)((double)merge_distance * (double)merge_distance);
  std::vector<bool> done(n, false);
  std::vector<int> merged(num_scans);
  std::vector<float> color_sum(num_scans);
  std::vector<std::pair<float, int>> nb;
  for (size_t i = 0; i < n; ++i) {
    if (done[i]) continue;
    int total = 0;
    float ax = 0.f, ay = 0.f, az = 0.f;
    int max_scan = -1, max_per_scan = 0;
    for (int s = 0; s < num_scans; ++s) { merged[s] = 0; color_sum[s] = 0; }
    float max_radius = -1;
    tree.radius(&xyz[3 * i], r2, &nb);
    for (const auto& e : nb) {
      const int idx = e.second;
      const int s = scan[idx];
      ax += xyz[3 * (size_t)idx]; ay += xyz[3 * (size_t)idx + 1]; az += xyz[3 * (size_t)idx + 2];
      color_sum[s] += colors[idx];
      if (max_radius_in[idx] > max_radius) max_radius = max_radius_in[idx];
      merged[s] += 1;
      if (merged[s] > max_per_scan) { max_per_scan = merged[s]; max_scan = s; }
      total += 1;
      done[idx] = true;
    }
    // CHECK_GT(total, 0): the centre itself is always found (d2 = 0 < r2 for r > 0)
    const float ft = (float)total;
    oxyz->push_back(ax / ft); oxyz->push_back(ay / ft); oxyz->push_back(az / ft);
    ocol->push_back(color_sum[max_scan] / merged[max_scan]);
    oscan->push_back((uint8_t)max_scan);
    omaxr->push_back(max_radius);
  }
}

// libstdc++ 9 `std::uniform_int_distribution<unsigned long>{0, range - 1}(g)` for a 32-bit engine (down-scaling branch).
static inline uint64_t gcc9_uniform(std::mt19937& g, uint64_t range) {
  const uint64_t urngrange = 0xFFFFFFFFull;     // g.max() - g.min()
  const uint64_t urange = range - 1;
  if (urngrange > urange) {
    const uint64_t uerange = urange + 1;
    const uint64_t scaling = urngrange / uerange;
    const uint64_t past = uerange * scaling;
    uint64_t ret;
    do ret = (uint64_t)g(); while (ret >= past);
    return ret / scaling;
  }
  // ranges >= 2^32 do not occur here (at most a few hundred candidates)
  return (uint64_t)g() % range;
}

// libstdc++ 9 `std::shuffle(first, last, g)` on ints.
static void gcc9_shuffle(int* first, int* last, std::mt19937& g) {
  if (first == last) return;
  const uint64_t urngrange = 0xFFFFFFFFull;
  const uint64_t urange = (uint64_t)(last - first);
  if (urngrange / urange >= urange) {
    int* i = first + 1;
    if ((urange % 2) == 0) { std::swap(*i, *(first + gcc9_uniform(g, 2))); ++i; }
    while (i != last) {
      const uint64_t swap_range = (uint64_t)(i - first) + 1;
      const uint64_t b1 = swap_range + 1;
      const uint64_t x = gcc9_uniform(g, swap_range * b1);
      std::swap(*i, *(first + x / b1)); ++i;
      std::swap(*i, *(first + x % b1)); ++i;
    }
    return;
  }
  for (int* i = first + 1; i != last; ++i) std::swap(*i, *(first + gcc9_uniform(g, (uint64_t)(i - first) + 1)));
}

}  // namespace orc

extern "C" {

uint64_t orc_ms_merge_close_points(const float* xyz, size_t n, const float* colors, const uint8_t* scan, const float* max_radius, int num_scans,
                                   float merge_distance, float* oxyz, float* ocol, uint8_t* oscan, float* omaxr) {
  std::vector<float> vx(xyz, xyz + 3 * n), vc(colors, colors + n), vm(max_radius, max_radius + n);
  std::vector<uint8_t> vs(scan, scan + n);
  std::vector<float> ox, oc, om; std::vector<uint8_t> os;
  orc::merge_close_points(merge_distance, num_scans, vx, vc, vs, vm, &ox, &oc, &os, &om);
  std::copy(ox.begin(), ox.end(), oxyz); std::copy(oc.begin(), oc.end(), ocol); std::copy(os.begin(), os.end(), oscan);
  std::copy(om.begin(), om.end(), omaxr);
  return oc.size();
}

// multi_scale_point_cloud.cc:263-368. min_radius / max_radius per input point as ComputeMinMaxPointRadius left them (+inf / -inf for
// points no image observes). Results per scale are concatenated: out_scale_count scales, scale s has out_counts[s] points and radius
// out_radius[s]. Output buffers must hold (max_scales) * n entries (every scale is a subset-merge of the input, so <= n points).
int orc_ms_create(const float* xyz, size_t n, const float* colors, const uint8_t* scan, const float* min_radius, const float* max_radius,
                  int num_scans, float min_radius_bias, float merge_distance_factor, int max_scales, float* out_radius, uint64_t* out_counts,
                  float* oxyz, float* ocol, uint8_t* oscan) {
  float min_radius_value = std::numeric_limits<float>::infinity();
  float max_radius_value = -1 * std::numeric_limits<float>::infinity();
  for (size_t i = 0; i < n; ++i) {
    if (min_radius[i] < min_radius_value) min_radius_value = min_radius[i];
    if (max_radius[i] > max_radius_value) max_radius_value = max_radius[i];
  }
  const float min_point_radius = min_radius_value * min_radius_bias;
  double radius = min_point_radius;
  std::vector<float> lx, lc, lm; std::vector<uint8_t> ls;
  for (size_t i = 0; i < n; ++i)
    if (radius >= min_radius[i]) {
      lx.insert(lx.end(), xyz + 3 * i, xyz + 3 * i + 3); lc.push_back(colors[i]); ls.push_back(scan[i]); lm.push_back(max_radius[i]);
    }
  float last_radius = -1;
  int scales = 0;
  size_t off = 0;
  while (true) {
    std::vector<float> nx, nc, nm; std::vector<uint8_t> ns;
    if (last_radius > 0) {
      for (size_t i = 0; i < lc.size(); ++i)
        if (radius <= lm[i]) { nx.insert(nx.end(), lx.begin() + 3 * i, lx.begin() + 3 * i + 3); nc.push_back(lc[i]); ns.push_back(ls[i]); nm.push_back(lm[i]); }
      for (size_t i = 0; i < n; ++i)
        if (last_radius < min_radius[i] && radius >= min_radius[i]) {
          nx.insert(nx.end(), xyz + 3 * i, xyz + 3 * i + 3); nc.push_back(colors[i]); ns.push_back(scan[i]); nm.push_back(max_radius[i]);
        }
      lx.swap(nx); lc.swap(nc); ls.swap(ns); lm.swap(nm);
      nx.clear(); nc.clear(); ns.clear(); nm.clear();
    }
    orc::merge_close_points((float)(merge_distance_factor * radius), num_scans, lx, lc, ls, lm, &nx, &nc, &ns, &nm);
    if (scales >= max_scales) return -1;
    out_radius[scales] = (float)radius; out_counts[scales] = nc.size();
    std::copy(nx.begin(), nx.end(), oxyz + 3 * off); std::copy(nc.begin(), nc.end(), ocol + off); std::copy(ns.begin(), ns.end(), oscan + off);
    off += nc.size();
    ++scales;
    last_radius = (float)radius;
    radius *= 2;
    const float kTolerance = 0.99f;
    if (radius >= max_radius_value * kTolerance) break;
    lx.swap(nx); lc.swap(nc); ls.swap(ns); lm.swap(nm);
  }
  return scales;
}

// problem.cc:706-786. out: neighbor_count x n indices (size_t in the reference), layout [point * neighbor_count + k].
int orc_ms_point_neighbors(const float* xyz, size_t n, const uint8_t* scan, int scan_count, int limit_to_same_scan, int candidate_count,
                           int neighbor_count, uint64_t* out) {
  std::mt19937 generator(0);
  const int k1 = candidate_count + 1;
  std::vector<int> indices(k1); std::vector<float> d2(k1);
  if (limit_to_same_scan) {
    std::vector<std::vector<float>> clouds(scan_count); std::vector<std::vector<size_t>> orig(scan_count);
    for (size_t i = 0; i < n; ++i) { const int s = scan[i]; clouds[s].insert(clouds[s].end(), xyz + 3 * i, xyz + 3 * i + 3); orig[s].push_back(i); }
    for (int s = 0; s < scan_count; ++s) if ((int)orig[s].size() < k1) return -1;      // CHECK_GE (:738)
    for (int s = 0; s < scan_count; ++s) {
      orc::KdTree tree; tree.build(clouds[s].data(), orig[s].size());
      for (size_t i = 0; i < orig[s].size(); ++i) {
        if (tree.knn(&clouds[s][3 * i], k1, indices.data(), d2.data()) != k1) return -1;
        orc::gcc9_shuffle(indices.data() + 1, indices.data() + k1, generator);
        for (int k = 0; k < neighbor_count; ++k) out[orig[s][i] * neighbor_count + k] = orig[s][indices[k + 1]];
      }
    }
  } else {
    if ((int)n < k1) return -1;
    orc::KdTree tree; tree.build(xyz, n);
    for (size_t i = 0; i < n; ++i) {
      if (tree.knn(&xyz[3 * i], k1, indices.data(), d2.data()) != k1) return -1;
      if (indices[0] != (int)i) return -2;                                              // CHECK_EQ(indices[0], i) (:773): self-match
      orc::gcc9_shuffle(indices.data() + 1, indices.data() + k1, generator);
      for (int k = 0; k < neighbor_count; ++k) out[i * neighbor_count + k] = indices[k + 1];
    }
  }
  return 0;
}

}  // extern "C"
