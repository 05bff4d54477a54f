// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Exact single-index kd-tree standing in for PCL 1.10 `search::KdTree` -> `KdTreeFLANN` ->
// FLANN 1.9.1 `KDTreeSingleIndex<L2_Simple<float>>` (leaf size 15, eps 0, sorted results), the
// structure the reference rebuilds per pair-direction (/root/reference/src/icp/icp_point_to_plane.cc:46-51)
// and queries with radiusSearch(p, d, idx, dist, max_nn=1) (:65-66), and with nearestKSearch for normals
// (/root/reference/src/geometry/two_pass_normal_3d_omp.hpp:66).
// FLANN is NOT under /root/reference (apt dependency, Dockerfile:6) -> parity UNPINNED at index level.
// The oracle DEFINES: squared distance d2 = ((dx*dx)+(dy*dy))+(dz*dz) in fp32 (L2_Simple accumulation order),
// radius test d2 < (float)((double)r*(double)r) strict, ties broken towards the LOWEST target index.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

namespace orc {

static inline float dist2_f32(const float* a, const float* b) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return ((dx * dx) + (dy * dy)) + (dz * dz);
}

class KdTree {
 public:
  // pts: n x 3 floats, contiguous with `stride` floats between points. Not copied.
  void build(const float* pts, size_t n, size_t stride = 3, int leaf_size = 15) {
    pts_ = pts; n_ = n; stride_ = stride; leaf_ = leaf_size;
    idx_.resize(n);
    for (size_t i = 0; i < n; ++i) idx_[i] = (int)i;
    nodes_.clear();
    nodes_.reserve(n / (leaf_size / 2 + 1) * 2 + 16);
    if (n == 0) return;
    for (int d = 0; d < 3; ++d) { lo_[d] = std::numeric_limits<float>::infinity(); hi_[d] = -lo_[d]; }
    for (size_t i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) { const float v = p(i)[d]; lo_[d] = std::min(lo_[d], v); hi_[d] = std::max(hi_[d], v); }
    float lo[3] = {lo_[0], lo_[1], lo_[2]}, hi[3] = {hi_[0], hi_[1], hi_[2]};
    build_rec(0, (int)n, lo, hi);
  }

  // Nearest target with d2 < r2 (strict); lowest index wins ties. Returns -1 if none.
  int nearest_within(const float* q, float r2, float* out_d2) const {
    if (n_ == 0) return -1;
    Best b{r2, -1};
    float off[3] = {0, 0, 0};
    float mind = 0.f;
    for (int d = 0; d < 3; ++d) {
      if (q[d] < lo_[d]) { off[d] = lo_[d] - q[d]; mind += off[d] * off[d]; }
      else if (q[d] > hi_[d]) { off[d] = q[d] - hi_[d]; mind += off[d] * off[d]; }
    }
    search1(0, q, mind, off, &b);
    if (b.idx >= 0) *out_d2 = b.d2;
    return b.idx;
  }

  // k nearest (including the query itself if it is in the set), sorted by (d2, index).
  // Returns number found (min(k, n)).
  int knn(const float* q, int k, int* out_idx, float* out_d2) const {
    if (n_ == 0 || k <= 0) return 0;
    Heap h{out_idx, out_d2, k, 0};
    float off[3] = {0, 0, 0};
    float mind = 0.f;
    for (int d = 0; d < 3; ++d) {
      if (q[d] < lo_[d]) { off[d] = lo_[d] - q[d]; mind += off[d] * off[d]; }
      else if (q[d] > hi_[d]) { off[d] = q[d] - hi_[d]; mind += off[d] * off[d]; }
    }
    searchk(0, q, mind, off, &h);
    // insertion-sorted already (kept ordered)
    return h.count;
  }

  // All targets with d2 < r2 (strict), sorted by (d2, index).
  void radius(const float* q, float r2, std::vector<std::pair<float, int>>* out) const {
    out->clear();
    if (n_ == 0) return;
    float off[3] = {0, 0, 0};
    float mind = 0.f;
    for (int d = 0; d < 3; ++d) {
      if (q[d] < lo_[d]) { off[d] = lo_[d] - q[d]; mind += off[d] * off[d]; }
      else if (q[d] > hi_[d]) { off[d] = q[d] - hi_[d]; mind += off[d] * off[d]; }
    }
    searchr(0, q, mind, off, r2, out);
    std::sort(out->begin(), out->end());
  }

 private:
  struct Node { int left, right, begin, end; int dim; float lo_split, hi_split; };
  struct Best { float d2; int idx; };
  struct Heap {  // ordered array of the k best (d2, idx) ascending
    int* idx; float* d2; int k; int count;
    float worst() const { return count < k ? std::numeric_limits<float>::infinity() : d2[k - 1]; }
    int worst_idx() const { return count < k ? std::numeric_limits<int>::max() : idx[k - 1]; }
    void push(float d, int i) {
      if (count == k && !(d < d2[k - 1] || (d == d2[k - 1] && i < idx[k - 1]))) return;
      int pos = count < k ? count : k - 1;
      while (pos > 0 && (d2[pos - 1] > d || (d2[pos - 1] == d && idx[pos - 1] > i))) {
        d2[pos] = d2[pos - 1]; idx[pos] = idx[pos - 1]; --pos;
      }
      d2[pos] = d; idx[pos] = i;
      if (count < k) ++count;
    }
  };

  const float* p(size_t i) const { return pts_ + i * stride_; }

  int build_rec(int begin, int end, float* lo, float* hi) {
    const int id = (int)nodes_.size();
    nodes_.push_back(Node{-1, -1, begin, end, -1, 0.f, 0.f});
    if (end - begin <= leaf_) return id;
    // tighten box to the data, split the widest dimension at the box midpoint (sliding to keep both sides non-empty)
    float dlo[3], dhi[3];
    for (int d = 0; d < 3; ++d) { dlo[d] = std::numeric_limits<float>::infinity(); dhi[d] = -dlo[d]; }
    for (int i = begin; i < end; ++i)
      for (int d = 0; d < 3; ++d) { const float v = p(idx_[i])[d]; dlo[d] = std::min(dlo[d], v); dhi[d] = std::max(dhi[d], v); }
    int dim = 0; float span = dhi[0] - dlo[0];
    for (int d = 1; d < 3; ++d) if (dhi[d] - dlo[d] > span) { span = dhi[d] - dlo[d]; dim = d; }
    if (!(span > 0.f)) return id;  // all points identical: keep as (large) leaf
    const float split = 0.5f * (dlo[dim] + dhi[dim]);
    int mid = (int)(std::partition(idx_.begin() + begin, idx_.begin() + end,
                                   [&](int i) { return p(i)[dim] < split; }) - idx_.begin());
    if (mid == begin || mid == end) {  // degenerate (rounding): fall back to median
      mid = (begin + end) / 2;
      std::nth_element(idx_.begin() + begin, idx_.begin() + mid, idx_.begin() + end,
                       [&](int a, int b) { return p(a)[dim] < p(b)[dim]; });
    }
    float lmax = -std::numeric_limits<float>::infinity(), rmin = std::numeric_limits<float>::infinity();
    for (int i = begin; i < mid; ++i) lmax = std::max(lmax, p(idx_[i])[dim]);
    for (int i = mid; i < end; ++i) rmin = std::min(rmin, p(idx_[i])[dim]);
    float sv;
    sv = hi[dim]; hi[dim] = lmax; const int l = build_rec(begin, mid, lo, hi); hi[dim] = sv;
    sv = lo[dim]; lo[dim] = rmin; const int r = build_rec(mid, end, lo, hi); lo[dim] = sv;
    Node& nd = nodes_[id];
    nd.left = l; nd.right = r; nd.dim = dim; nd.lo_split = lmax; nd.hi_split = rmin;
    return id;
  }

  // Branch-and-bound; bounds use <= so that equal-distance candidates in the far branch are still visited
  // (required for the lowest-index tie-break to be exact).
  void search1(int id, const float* q, float mind, float* off, Best* b) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; ++i) {
        const int t = idx_[i];
        const float d2 = dist2_f32(q, p(t));
        if (d2 < b->d2 || (d2 == b->d2 && b->idx >= 0 && t < b->idx)) { b->d2 = d2; b->idx = t; }
      }
      return;
    }
    const float v = q[nd.dim];
    const float d_lo = v - nd.lo_split, d_hi = nd.hi_split - v;
    int first, second; float cut;
    if (d_lo < d_hi) { first = nd.left; second = nd.right; cut = d_hi; }
    else { first = nd.right; second = nd.left; cut = d_lo; }
    search1(first, q, mind, off, b);
    const float old = off[nd.dim];
    // lower bound on the far side; computed conservatively (never above the true fp32 distance of any far point)
    const float c = cut > 0.f ? cut : 0.f;
    const float far_mind = lower_bound_sq(mind, old, c);
    if (far_mind <= b->d2) {
      off[nd.dim] = c > old ? c : old;
      search1(second, q, far_mind, off, b);
      off[nd.dim] = old;
    }
  }

  void searchk(int id, const float* q, float mind, float* off, Heap* h) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; ++i) { const int t = idx_[i]; h->push(dist2_f32(q, p(t)), t); }
      return;
    }
    const float v = q[nd.dim];
    const float d_lo = v - nd.lo_split, d_hi = nd.hi_split - v;
    int first, second; float cut;
    if (d_lo < d_hi) { first = nd.left; second = nd.right; cut = d_hi; }
    else { first = nd.right; second = nd.left; cut = d_lo; }
    searchk(first, q, mind, off, h);
    const float old = off[nd.dim];
    const float c = cut > 0.f ? cut : 0.f;
    const float far_mind = lower_bound_sq(mind, old, c);
    if (far_mind <= h->worst()) {
      off[nd.dim] = c > old ? c : old;
      searchk(second, q, far_mind, off, h);
      off[nd.dim] = old;
    }
  }

  void searchr(int id, const float* q, float mind, float* off, float r2, std::vector<std::pair<float, int>>* out) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; ++i) {
        const int t = idx_[i]; const float d2 = dist2_f32(q, p(t));
        if (d2 < r2) out->emplace_back(d2, t);
      }
      return;
    }
    const float v = q[nd.dim];
    const float d_lo = v - nd.lo_split, d_hi = nd.hi_split - v;
    int first, second; float cut;
    if (d_lo < d_hi) { first = nd.left; second = nd.right; cut = d_hi; }
    else { first = nd.right; second = nd.left; cut = d_lo; }
    searchr(first, q, mind, off, r2, out);
    const float old = off[nd.dim];
    const float c = cut > 0.f ? cut : 0.f;
    const float far_mind = lower_bound_sq(mind, old, c);
    if (far_mind <= r2) {
      off[nd.dim] = c > old ? c : old;
      searchr(second, q, far_mind, off, r2, out);
      off[nd.dim] = old;
    }
  }

  // Conservative lower bound of the squared distance to the far half-space: shrink by a few ulps so fp32
  // rounding in dist2_f32 can never make a real candidate look farther than the bound.
  static float lower_bound_sq(float mind, float old_off, float new_off) {
    if (new_off <= old_off) return mind * 0.999999f;
    const float lb = (mind - old_off * old_off) + new_off * new_off;
    return (lb > 0.f ? lb : 0.f) * 0.99999f;
  }

  const float* pts_ = nullptr;
  size_t n_ = 0, stride_ = 3;
  int leaf_ = 15;
  float lo_[3], hi_[3];
  std::vector<int> idx_;
  std::vector<Node> nodes_;
};

// Brute-force reference for small cases (validates the kd-tree itself in tests).
static inline int brute_nearest_within(const float* pts, size_t n, size_t stride, const float* q, float r2, float* out_d2) {
  int best = -1; float bd = r2;
  for (size_t i = 0; i < n; ++i) {
    const float d2 = dist2_f32(q, pts + i * stride);
    if (d2 < bd) { bd = d2; best = (int)i; }   // ascending i: first hit at a given d2 keeps the lowest index
  }
  if (best >= 0) *out_d2 = bd;
  return best;
}

}  // namespace orc
