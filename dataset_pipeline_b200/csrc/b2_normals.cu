// Path A — kNN two-pass normal estimation on sm_100a (kernel K7), behind b2_normals_estimate().
//
// Replaces pcl::NormalEstimationTwoPassOMP::computeFeature (/root/reference/src/geometry/two_pass_normal_3d_omp.hpp:47-119):
//   per point: k nearest neighbours incl. itself (kd-tree nearestKSearch, sorted) ->
//   computeMeanAndCovarianceMatrixTwoPass (/root/reference/src/geometry/two_pass_centroid.hpp:155-259, fp32 sequential sums in
//   neighbour order) -> smallest eigenvector of the 3x3 covariance (pcl::eigen33 closed form) -> curvature ->
//   flipNormalTowardsViewpoint; NaN when fewer than 3 neighbours (two_pass_normal_3d.h:100-105).
//
// Spatial index: the points are sorted by a 63-bit Morton code; consecutive runs of 8 sorted points are the leaves of an
// IMPLICIT binary BVH (node i of level l covers leaves [i*2^l, (i+1)*2^l); no pointers, 24 B AABB per node). One thread per
// sorted query walks it near-child-first with a k-best list in shared memory. The kNN is exact with the tie-break
// (squared distance, original index) ascending; the AABB lower bound is evaluated with the same fp32 operations as the
// point distance, so pruning can never drop a neighbour (rounding is monotone).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b2_bvh.cuh"
#include "b2_common.cuh"
#include "b2_knn.h"

namespace b2 {

static constexpr int kKnnThreads = 128;

// k-best list of one thread in shared memory, element e of thread t at [e * kKnnThreads + t] (conflict-free).
struct KBest {
  float* d2; unsigned int* pos; int k; int count; int worst; float worst_d2;
  __device__ __forceinline__ float& D(int e) { return d2[e * kKnnThreads]; }
  __device__ __forceinline__ unsigned int& P(int e) { return pos[e * kKnnThreads]; }
};

// (d2, original index) strict "less" with the index fetched lazily.
__device__ __forceinline__ bool less_than(float da, unsigned int ia, float db, unsigned int posb, const float4* __restrict__ s) {
  if (da != db) return da < db;
  return ia < __float_as_uint(__ldg(&s[posb].w));
}

__device__ __forceinline__ void kbest_rescan(KBest& h, const float4* __restrict__ s) {
  int w = 0; float wd = h.D(0);
  for (int e = 1; e < h.count; ++e) {
    const float d = h.D(e);
    if (d > wd || (d == wd && __float_as_uint(__ldg(&s[h.P(e)].w)) > __float_as_uint(__ldg(&s[h.P(w)].w)))) { w = e; wd = d; }
  }
  h.worst = w; h.worst_d2 = wd;
}

__device__ __forceinline__ void kbest_offer(KBest& h, float d, unsigned int p, unsigned int idx, const float4* __restrict__ s) {
  if (h.count < h.k) {
    h.D(h.count) = d; h.P(h.count) = p; ++h.count;
    if (h.count == h.k) kbest_rescan(h, s);
    return;
  }
  if (!less_than(d, idx, h.worst_d2, h.P(h.worst), s)) return;
  h.D(h.worst) = d; h.P(h.worst) = p;
  kbest_rescan(h, s);
}

__device__ __forceinline__ void scan_leaf(KBest& h, const float4& q, unsigned int leaf, size_t n, const float4* __restrict__ s) {
  const size_t b = (size_t)leaf * kLeaf, e = min(n, b + kLeaf);
  for (size_t p = b; p < e; ++p) {
    const float4 t = __ldg(&s[p]);
    const float d = dist2_pt(q, t);
    if (h.count < h.k || d <= h.worst_d2) kbest_offer(h, d, (unsigned int)p, __float_as_uint(t.w), s);
  }
}

// ---- closed-form smallest eigenpair of a symmetric 3x3 (pcl::eigen33 / computeRoots, fp32) ----
__device__ __forceinline__ void roots2(float b, float c, float r[3]) {
  r[0] = 0.f;
  float d = b * b - 4.f * c;
  if (d < 0.f) d = 0.f;
  const float sd = sqrtf(d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}
__device__ __forceinline__ void swapf(float& a, float& b) { const float t = a; a = b; b = t; }
__device__ void compute_roots(const float m[9], float r[3]) {
  const float c0 = m[0] * m[4] * m[8] + 2.f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
  const float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  const float c2 = m[0] + m[4] + m[8];
  if (fabsf(c0) < 1.1920929e-07f) { roots2(c2, c1, r); return; }
  const float s_inv3 = 1.0f / 3.0f;
  const float s_sqrt3 = sqrtf(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f) q = 0.f;
  const float rho = sqrtf(-a_over_3);
  const float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
  const float ct = cosf(theta), st = sinf(theta);
  r[0] = c2_over_3 + 2.f * rho * ct;
  r[1] = c2_over_3 - rho * (ct + s_sqrt3 * st);
  r[2] = c2_over_3 - rho * (ct - s_sqrt3 * st);
  if (r[0] >= r[1]) swapf(r[0], r[1]);
  if (r[1] >= r[2]) { swapf(r[1], r[2]); if (r[0] >= r[1]) swapf(r[0], r[1]); }
  if (r[0] <= 0.f) roots2(c2, c1, r);
}
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

// Normal + curvature from a neighbour list given in the reference's order (pos(a) = sorted position of neighbour a, cnt >= 3):
// two-pass mean / covariance in fp32 (two_pass_centroid.hpp:164-258), pcl::eigen33 smallest eigenpair, viewpoint flip.
template <typename PosFn>
__device__ __forceinline__ float4 normal_from_list(const float4* __restrict__ s_xyz, PosFn pos, int cnt, const float4& q, float vpx, float vpy, float vpz) {
  // two-pass mean / covariance, fp32, sequential in neighbour order, no contraction (two_pass_centroid.hpp:164-258)
  float a6 = 0.f, a7 = 0.f, a8 = 0.f;
  for (int a = 0; a < cnt; ++a) { const float4 p = __ldg(&s_xyz[pos(a)]); a6 = fadd(a6, p.x); a7 = fadd(a7, p.y); a8 = fadd(a8, p.z); }
  const float fn = (float)cnt;
  a6 = a6 / fn; a7 = a7 / fn; a8 = a8 / fn;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
  for (int a = 0; a < cnt; ++a) {
    const float4 p = __ldg(&s_xyz[pos(a)]);
    const float dx = fsub(p.x, a6), dy = fsub(p.y, a7), dz = fsub(p.z, a8);
    a0 = fadd(a0, fmul(dx, dx)); a1 = fadd(a1, fmul(dx, dy)); a2 = fadd(a2, fmul(dx, dz));
    a3 = fadd(a3, fmul(dy, dy)); a4 = fadd(a4, fmul(dy, dz)); a5 = fadd(a5, fmul(dz, dz));
  }
  float cov[9];
  cov[0] = a0 / fn; cov[1] = a1 / fn; cov[2] = a2 / fn; cov[4] = a3 / fn; cov[5] = a4 / fn; cov[8] = a5 / fn;
  cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];

  // pcl::eigen33: scale, closed-form roots, eigenvector from the largest cross product of rows of (A - l0 I)
  float scale = 0.f;
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(cov[i]));
  if (scale <= 1.17549435e-38f) scale = 1.f;
  float sc[9];
  for (int i = 0; i < 9; ++i) sc[i] = cov[i] / scale;
  float r[3];
  compute_roots(sc, r);
  const float ev = r[0] * scale;
  sc[0] -= r[0]; sc[4] -= r[0]; sc[8] -= r[0];
  float v1[3], v2[3], v3[3];
  cross3(sc, sc + 3, v1); cross3(sc, sc + 6, v2); cross3(sc + 3, sc + 6, v3);
  const float n1 = v1[0] * v1[0] + (v1[1] * v1[1] + v1[2] * v1[2]);
  const float n2 = v2[0] * v2[0] + (v2[1] * v2[1] + v2[2] * v2[2]);
  const float n3 = v3[0] * v3[0] + (v3[1] * v3[1] + v3[2] * v3[2]);
  const float* v; float len;
  if (n1 >= n2 && n1 >= n3) { v = v1; len = n1; } else if (n2 >= n1 && n2 >= n3) { v = v2; len = n2; } else { v = v3; len = n3; }
  const float inv = sqrtf(len);
  float nx = v[0] / inv, ny = v[1] / inv, nz = v[2] / inv;
  const float eig_sum = cov[0] + cov[4] + cov[8];
  const float curv = eig_sum != 0.f ? fabsf(ev / eig_sum) : 0.f;
  // flipNormalTowardsViewpoint
  const float wx = vpx - q.x, wy = vpy - q.y, wz = vpz - q.z;
  if ((wx * nx + wy * ny + wz * nz) < 0.f) { nx *= -1.f; ny *= -1.f; nz *= -1.f; }
  return make_float4(nx, ny, nz, curv);
}

__global__ void __launch_bounds__(kKnnThreads)
kn_knn_normals(const float4* __restrict__ s_xyz, size_t n, const Aabb* __restrict__ nodes, BvhLevels lv, int k, float vpx, float vpy, float vpz,
               float4* __restrict__ out, int* __restrict__ out_idx, unsigned int* __restrict__ nan_count, size_t q_begin, size_t q_end,
               int stat_mode, float* __restrict__ out_stat) {
  extern __shared__ unsigned char smem_raw[];
  float* sm_d2 = reinterpret_cast<float*>(smem_raw);
  unsigned int* sm_pos = reinterpret_cast<unsigned int*>(smem_raw + sizeof(float) * (size_t)k * kKnnThreads);
  const size_t j = q_begin + (size_t)blockIdx.x * kKnnThreads + threadIdx.x;   // this rank's slice of the Morton-sorted queries
  if (j >= q_end) return;
  const float4 q = s_xyz[j];
  KBest h{sm_d2 + threadIdx.x, sm_pos + threadIdx.x, k, 0, 0, INFINITY};

  // seed with the query's own leaf and its neighbours in Morton order: three leaves, or as many as hold k points (the list is then
  // full, with a near-final bound, before the walk starts)
  const unsigned int nleaf = lv.count[0];
  const unsigned int own = (unsigned int)(j / kLeaf);
  const unsigned int nseed = max(3u, (unsigned int)((k + kLeaf - 1) / kLeaf) + 1u);
  unsigned int l0 = own > nseed / 2 ? own - nseed / 2 : 0;
  if (l0 + nseed > nleaf) l0 = nleaf > nseed ? nleaf - nseed : 0;
  const unsigned int l1 = min(nleaf, l0 + nseed) - 1;
  for (unsigned int l = l0; l <= l1; ++l) scan_leaf(h, q, l, n, s_xyz);

  // near-first depth-first walk from the root; entries are (level << 27 | index)
  unsigned int stack[kBvhMaxLevels + 2];
  int sp = 0;
  stack[sp++] = ((unsigned int)(lv.nlevels - 1) << 27);
  while (sp > 0) {
    const unsigned int e = stack[--sp];
    const int level = (int)(e >> 27);
    const unsigned int i = e & 0x7FFFFFFu;
    if (h.count == h.k && dist2_box(q, nodes[lv.offset[level] + i]) > h.worst_d2) continue;
    if (level == 0) {
      if (i < l0 || i > l1) scan_leaf(h, q, i, n, s_xyz);
      continue;
    }
    const unsigned int c0 = 2 * i, c1 = 2 * i + 1;
    const unsigned int nchild = lv.count[level - 1];
    if (c1 >= nchild) { stack[sp++] = ((unsigned int)(level - 1) << 27) | c0; continue; }
    const float d0 = dist2_box(q, nodes[lv.offset[level - 1] + c0]);
    const float d1 = dist2_box(q, nodes[lv.offset[level - 1] + c1]);
    const bool full = h.count == h.k;
    const bool v0 = !full || d0 <= h.worst_d2, v1 = !full || d1 <= h.worst_d2;
    if (d0 <= d1) {
      if (v1) stack[sp++] = ((unsigned int)(level - 1) << 27) | c1;
      if (v0) stack[sp++] = ((unsigned int)(level - 1) << 27) | c0;
    } else {
      if (v0) stack[sp++] = ((unsigned int)(level - 1) << 27) | c0;
      if (v1) stack[sp++] = ((unsigned int)(level - 1) << 27) | c1;
    }
  }

  // sort the list by (d2, original index): neighbour order defines the fp32 summation order
  const int cnt = h.count;
  for (int a = 1; a < cnt; ++a) {
    const float d = h.D(a); const unsigned int p = h.P(a);
    const unsigned int pi = __float_as_uint(__ldg(&s_xyz[p].w));
    int b = a - 1;
    while (b >= 0 && less_than(d, pi, h.D(b), h.P(b), s_xyz)) { h.D(b + 1) = h.D(b); h.P(b + 1) = h.P(b); --b; }
    h.D(b + 1) = d; h.P(b + 1) = p;
  }
  const unsigned int qi = __float_as_uint(q.w);
  if (out_idx) {
    for (int a = 0; a < k; ++a) out_idx[(size_t)qi * k + a] = a < cnt ? (int)__float_as_uint(__ldg(&s_xyz[h.P(a)].w)) : -1;
  }
  if (stat_mode == kKnnStatMeanDistance) {
    // LocalStatisticalOutlierRemoval, first pass (local_statistical_outlier_removal.hpp:113-117): neighbour 0 is the query point
    double dist_sum = 0.0;
    for (int a = 1; a < cnt; ++a) dist_sum += (double)sqrtf(h.D(a));
    out_stat[qi] = (float)(dist_sum / (double)(k - 1));
  } else if (stat_mode == kKnnStatLastD2) {
    out_stat[qi] = cnt == k ? h.D(k - 1) : INFINITY;      // squared distance of the (k-1)-th neighbour (splat_creator.cc:166)
  }
  if (!out) return;
  const float nanv = __int_as_float(0x7fc00000);
  if (cnt < 3) { out[qi] = make_float4(nanv, nanv, nanv, nanv); atomicAdd(nan_count, 1u); return; }

  out[qi] = normal_from_list(s_xyz, [&](int a) { return h.P(a); }, cnt, q, vpx, vpy, vpz);
}

// ---- long neighbour lists (k > 64; PointCloudCleaner's recipe is kNN = 270): one WARP per query -------------------------------------------
// A per-thread list of hundreds of entries leaves room for two warps per SM. Here the list lives in shared memory per warp (24 B x k), the
// BVH walk is warp-uniform (no divergence), the eight points of a leaf are tested by eight lanes at once, and an insertion replaces the worst
// entry and re-finds the maximum with a strided scan + shuffle reduction. Keys are (d2 bits << 32 | original index): for non-negative floats
// the integer order is the (d2, index) order of the per-thread kernel, so both produce identical lists.
static constexpr int kWideWarps = 4;
__global__ void __launch_bounds__(32 * kWideWarps)
kn_knn_wide(const float4* __restrict__ s_xyz, size_t n, const Aabb* __restrict__ nodes, BvhLevels lv, int k, float vpx, float vpy, float vpz,
            float4* __restrict__ out, int* __restrict__ out_idx, unsigned int* __restrict__ nan_count, size_t q_begin, size_t q_end,
            int stat_mode, float* __restrict__ out_stat, unsigned long long* __restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned int c_nodes = 0, c_leaves = 0, c_inserts = 0;
  const size_t j = q_begin + (size_t)blockIdx.x * kWideWarps + w;
  if (j >= q_end) return;                                            // whole warp
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)w * 2 * k;   // [k] list, [k] sorted
  unsigned long long* skeys = keys + k;
  unsigned int* pos = reinterpret_cast<unsigned int*>(smem_raw + (size_t)kWideWarps * 16 * k) + (size_t)w * 2 * k;
  unsigned int* spos = pos + k;
  const float4 q = __ldg(&s_xyz[j]);
  int count = 0, worst_slot = 0;
  unsigned long long worst_key = ~0ull;

  auto refind_worst = [&]() {
    unsigned long long m = 0ull; int slot = 0;
    for (int e = lane; e < k; e += 32) { const unsigned long long v = keys[e]; if (v >= m) { m = v; slot = e; } }
    unsigned long long g = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, g, o); g = t > g ? t : g; }
    const int owner = __ffs(__ballot_sync(0xffffffffu, m == g && lane < k)) - 1;    // keys are unique (the index is part of the key)
    worst_key = g; worst_slot = __shfl_sync(0xffffffffu, slot, owner);
  };
  auto offer_points = [&](size_t first, int npts) {          // candidates first .. first + npts - 1 (npts <= 32), one per lane
    ++c_leaves;
    const size_t p = first + lane;
    const bool valid = lane < npts && p < n;
    unsigned long long key = ~0ull;
    if (valid) {
      const float4 t = __ldg(&s_xyz[p]);
      key = ((unsigned long long)__float_as_uint(dist2_pt(q, t)) << 32) | (unsigned long long)__float_as_uint(t.w);
    }
    unsigned int m = __ballot_sync(0xffffffffu, valid && (count < k || key < worst_key));
    while (m) {
      const int src = __ffs(m) - 1; m &= m - 1;
      const unsigned long long ck = __shfl_sync(0xffffffffu, key, src);
      const unsigned int cp = (unsigned int)(first + src);
      if (count < k) {
        if (lane == 0) { keys[count] = ck; pos[count] = cp; }
        ++count;
        if (count == k) { __syncwarp(); refind_worst(); }
      } else if (ck < worst_key) {
        ++c_inserts;
        if (lane == 0) { keys[worst_slot] = ck; pos[worst_slot] = cp; }
        __syncwarp();
        refind_worst();
      }
    }
  };

  auto offer_leaf = [&](unsigned int leaf) { offer_points((size_t)leaf * kLeaf, kLeaf); };

  // Seed: the list is FILLED, with coalesced loads and no insertion logic, from the ceil(k / 8) leaves around the query in Morton order
  // (spatially close points), so the walk below starts with a near-final bound instead of discovering it replacement by replacement.
  const unsigned int nleaf = lv.count[0];
  const unsigned int own = (unsigned int)(j / kLeaf);
  const unsigned int nfill_leaves = (unsigned int)((k + kLeaf - 1) / kLeaf);
  unsigned int l0 = own > nfill_leaves / 2 ? own - nfill_leaves / 2 : 0;
  if (l0 + nfill_leaves > nleaf) l0 = nleaf > nfill_leaves ? nleaf - nfill_leaves : 0;
  const unsigned int l1 = min(nleaf, l0 + nfill_leaves) - 1;
  {
    const size_t pb = (size_t)l0 * kLeaf, pe = min(n, (size_t)(l1 + 1) * kLeaf);
    const int nfill = (int)min((size_t)k, pe - pb);
    for (int e = lane; e < nfill; e += 32) {
      const float4 t = __ldg(&s_xyz[pb + e]);
      keys[e] = ((unsigned long long)__float_as_uint(dist2_pt(q, t)) << 32) | (unsigned long long)__float_as_uint(t.w);
      pos[e] = (unsigned int)(pb + e);
    }
    count = nfill;
    __syncwarp();
    if (count == k) refind_worst();
    if (pb + nfill < pe) offer_points(pb + nfill, (int)(pe - pb - nfill));     // the (< 8) points of the last seed leaf beyond k
  }
  unsigned int stack[kBvhMaxLevels + 2];
  int sp = 0;
  stack[sp++] = ((unsigned int)(lv.nlevels - 1) << 27);
  while (sp > 0) {
    const unsigned int e = stack[--sp];
    const int level = (int)(e >> 27);
    const unsigned int i = e & 0x7FFFFFFu;
    const bool full = count == k;
    const float worst_d2 = __uint_as_float((unsigned int)(worst_key >> 32));
    ++c_nodes;
    if (full && dist2_box(q, nodes[lv.offset[level] + i]) > worst_d2) continue;
    if (level == 0) {
      if (i < l0 || i > l1) offer_leaf(i);
      continue;
    }
    const unsigned int c0 = 2 * i, c1 = 2 * i + 1;
    const unsigned int nchild = lv.count[level - 1];
    if (c1 >= nchild) { stack[sp++] = ((unsigned int)(level - 1) << 27) | c0; continue; }
    const float d0 = dist2_box(q, nodes[lv.offset[level - 1] + c0]);
    const float d1 = dist2_box(q, nodes[lv.offset[level - 1] + c1]);
    const bool v0 = !full || d0 <= worst_d2, v1 = !full || d1 <= worst_d2;
    if (d0 <= d1) {
      if (v1) stack[sp++] = ((unsigned int)(level - 1) << 27) | c1;
      if (v0) stack[sp++] = ((unsigned int)(level - 1) << 27) | c0;
    } else {
      if (v0) stack[sp++] = ((unsigned int)(level - 1) << 27) | c0;
      if (v1) stack[sp++] = ((unsigned int)(level - 1) << 27) | c1;
    }
  }

  if (dbg && lane == 0) { atomicAdd(&dbg[0], (unsigned long long)c_nodes); atomicAdd(&dbg[1], (unsigned long long)c_leaves); atomicAdd(&dbg[2], (unsigned long long)c_inserts); }
  // rank every entry (keys are unique): sorted position = number of smaller keys
  const int cnt = count;
  __syncwarp();
  for (int e = lane; e < cnt; e += 32) {
    const unsigned long long mine = keys[e];
    int rank = 0;
    for (int f = 0; f < cnt; ++f) rank += keys[f] < mine ? 1 : 0;
    skeys[rank] = mine; spos[rank] = pos[e];
  }
  __syncwarp();
  const unsigned int qi = __float_as_uint(q.w);
  if (out_idx)
    for (int a = lane; a < k; a += 32) out_idx[(size_t)qi * k + a] = a < cnt ? (int)(unsigned int)(skeys[a] & 0xFFFFFFFFull) : -1;
  if (stat_mode == kKnnStatMeanDistance) {
    float* sq = reinterpret_cast<float*>(pos);                        // the unsorted positions are no longer needed
    for (int a = lane; a < cnt; a += 32) sq[a] = sqrtf(__uint_as_float((unsigned int)(skeys[a] >> 32)));
    __syncwarp();
    if (lane == 0) {
      double dist_sum = 0.0;
      for (int a = 1; a < cnt; ++a) dist_sum += (double)sq[a];        // neighbour order, as the reference (:113-117)
      out_stat[qi] = (float)(dist_sum / (double)(k - 1));
    }
  } else if (stat_mode == kKnnStatLastD2) {
    if (lane == 0) out_stat[qi] = cnt == k ? __uint_as_float((unsigned int)(skeys[k - 1] >> 32)) : INFINITY;
  }
  if (!out || lane != 0) return;
  const float nanv = __int_as_float(0x7fc00000);
  if (cnt < 3) { out[qi] = make_float4(nanv, nanv, nanv, nanv); atomicAdd(nan_count, 1u); return; }
  out[qi] = normal_from_list(s_xyz, [&](int a) { return spos[a]; }, cnt, q, vpx, vpy, vpz);
}
static int knn_wide_from() {     // list length from which the warp-per-query kernel takes over (B2_KNN_WIDE_FROM for A/B runs)
  static const int v = [] { const char* e = std::getenv("B2_KNN_WIDE_FROM"); const int x = e ? atoi(e) : 0; return x >= 8 && x <= 129 ? x : 56; }();
  return v;
}

// ---- radius mode (setRadiusSearch): every point with d2 < r2, in (d2, index) order --------------------------------------------
// Three kernels per batch of Morton-sorted queries: count -> (scan) -> fill keys ((d2 bits << 32) | original index, value = sorted
// position) -> (segmented radix sort: for non-negative floats the bit pattern orders like the value) -> normals.
__global__ void __launch_bounds__(128) kn_radius_count(const float4* __restrict__ s_xyz, size_t n, const Aabb* __restrict__ nodes, BvhLevels lv, float r2,
                                                       size_t q_begin, size_t q_end, unsigned int* __restrict__ counts) {
  const size_t j = q_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= q_end) return;
  unsigned int c = 0;
  radius_visit(s_xyz[j], r2, s_xyz, n, nodes, lv, [&](float, unsigned int, unsigned int) { ++c; });
  counts[j - q_begin] = c;
}
__global__ void __launch_bounds__(128) kn_radius_fill(const float4* __restrict__ s_xyz, size_t n, const Aabb* __restrict__ nodes, BvhLevels lv, float r2,
                                                      size_t q_begin, size_t q_end, const unsigned long long* __restrict__ offs,
                                                      unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals) {
  const size_t j = q_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= q_end) return;
  unsigned long long o = offs[j - q_begin];
  radius_visit(s_xyz[j], r2, s_xyz, n, nodes, lv, [&](float d, unsigned int p, unsigned int idx) {
    keys[o] = ((unsigned long long)__float_as_uint(d) << 32) | idx; vals[o] = p; ++o;
  });
}
__global__ void __launch_bounds__(128) kn_radius_normals(const float4* __restrict__ s_xyz, size_t q_begin, size_t q_end, const unsigned long long* __restrict__ offs,
                                                         const unsigned int* __restrict__ vals, float vpx, float vpy, float vpz, float4* __restrict__ out,
                                                         int* __restrict__ out_count, unsigned int* __restrict__ nan_count) {
  const size_t j = q_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= q_end) return;
  const float4 q = s_xyz[j];
  const unsigned long long b = offs[j - q_begin];
  const int cnt = (int)(offs[j - q_begin + 1] - b);
  const unsigned int qi = __float_as_uint(q.w);
  if (out_count) out_count[qi] = cnt;
  const float nanv = __int_as_float(0x7fc00000);
  if (cnt < 3) { out[qi] = make_float4(nanv, nanv, nanv, nanv); atomicAdd(nan_count, 1u); return; }
  out[qi] = normal_from_list(s_xyz, [&](int a) { return __ldg(&vals[b + a]); }, cnt, q, vpx, vpy, vpz);
}
__global__ void __launch_bounds__(256) kn_widen(const unsigned int* __restrict__ in, size_t n, unsigned long long* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

static inline unsigned int div_up_u(size_t a, size_t b) { return (unsigned int)((a + b - 1) / b); }

}  // namespace b2

using namespace b2;

// Multi-GPU: every rank holds the whole cloud (search set) and answers a contiguous slice of the Morton-sorted
// queries; the zero-initialised outputs are merged by one sum-allreduce (NaN normals survive the sum), so every rank returns
// the full result. No exchange during the search itself.
static int normals_impl(const float* xyz, size_t n, size_t stride_bytes, int k, const float viewpoint[3], float* out_nxyz_curv,
                        int32_t* out_knn_idx, int* is_dense, b2_comm* comm, int device, float radius = 0.f, int32_t* out_count = nullptr,
                        const KnnHook* hook = nullptr) {
  if ((n && (!xyz || (!out_nxyz_curv && !out_knn_idx && !hook))) || !viewpoint) return set_error(B2_ERR_ARG, "null argument");
  if (stride_bytes < 12) return set_error(B2_ERR_ARG, "stride_bytes must be >= 12");
  const bool radius_mode = radius > 0.f;
  if (!radius_mode && (k < 1 || k > 2048)) return set_error(B2_ERR_ARG, "k must be in [1,2048]");
  if (radius_mode && !(radius < INFINITY)) return set_error(B2_ERR_ARG, "radius must be finite");
  if (n >= (1ull << 30)) return set_error(B2_ERR_ARG, "clouds above 2^30 points are not supported");
  if (is_dense) *is_dense = 1;
  if (n == 0) return B2_OK;
  int dev = 0, sms = 0;
  B2_TRY(select_device(device, &dev, &sms));
  int rank = 0, world = 1;
  if (comm) B2_TRY(b2_comm_info(comm, &rank, &world));
  if (world > 1 && out_knn_idx) return set_error(B2_ERR_ARG, "neighbour index output is single-GPU only");
  const size_t q_begin = n * (size_t)rank / (size_t)world, q_end = n * (size_t)(rank + 1) / (size_t)world;
  cudaStream_t st; B2_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  DevBuf d_xyz, d_part, d_keys, d_keys2, d_idx, d_perm, d_sxyz, d_nodes, d_tmp, d_out, d_oidx, d_nan, d_stat, d_dbg;
  DevBuf r_cnt, r_cnt64, r_offs, r_keys, r_keys2, r_vals, r_vals2, r_ocount;
  PinnedBuf p_part;
  int rc = B2_OK;
  auto body = [&]() -> int {
    B2_TRY(d_xyz.ensure(n * 12));
    if (stride_bytes == 12) B2_CUDA(cudaMemcpyAsync(d_xyz.p, xyz, n * 12, cudaMemcpyHostToDevice, st));
    else B2_CUDA(cudaMemcpy2DAsync(d_xyz.p, 12, xyz, stride_bytes, 12, n, cudaMemcpyHostToDevice, st));
    const int bb = sms * 2;
    B2_TRY(d_part.ensure(sizeof(float) * 6 * bb)); B2_TRY(p_part.ensure(sizeof(float) * 6 * bb));
    kn_bbox<<<bb, 256, 0, st>>>(d_xyz.as<float>(), n, d_part.as<float>());
    B2_CUDA(cudaMemcpyAsync(p_part.p, d_part.p, sizeof(float) * 6 * bb, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int b = 0; b < bb; ++b) for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], p_part.as<float>()[6 * b + d]); mx[d] = std::max(mx[d], p_part.as<float>()[6 * b + 3 + d]);
    }
    for (int d = 0; d < 3; ++d) if (!std::isfinite(mn[d]) || !std::isfinite(mx[d])) return set_error(B2_ERR_ARG, "non-finite coordinates (dense clouds only)");
    const float ext = std::max({mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2], 1e-30f});
    const float scale = 2097151.f / ext;
    B2_TRY(d_keys.ensure(n * 8)); B2_TRY(d_keys2.ensure(n * 8)); B2_TRY(d_idx.ensure(n * 4)); B2_TRY(d_perm.ensure(n * 4));
    B2_TRY(d_sxyz.ensure(n * 16));
    kn_morton<<<div_up_u(n, 256), 256, 0, st>>>(d_xyz.as<float>(), n, mn[0], mn[1], mn[2], scale, d_keys.as<unsigned long long>(), d_idx.as<unsigned int>());
    size_t tmp = 0;
    B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_keys.as<unsigned long long>(), d_keys2.as<unsigned long long>(), d_idx.as<unsigned int>(),
                                            d_perm.as<unsigned int>(), (long long)n, 0, 63, st));
    B2_TRY(d_tmp.ensure(tmp));
    B2_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp, d_keys.as<unsigned long long>(), d_keys2.as<unsigned long long>(), d_idx.as<unsigned int>(),
                                            d_perm.as<unsigned int>(), (long long)n, 0, 63, st));
    kn_gather<<<div_up_u(n, 256), 256, 0, st>>>(d_xyz.as<float>(), n, d_perm.as<unsigned int>(), d_sxyz.as<float4>());
    // implicit BVH levels
    BvhLevels lv; std::memset(&lv, 0, sizeof(lv));
    unsigned int cnt = div_up_u(n, kLeaf), off = 0; int L = 0;
    while (true) { lv.offset[L] = off; lv.count[L] = cnt; off += cnt; ++L; if (cnt == 1) break; cnt = (cnt + 1) / 2; }
    lv.nlevels = L;
    B2_TRY(d_nodes.ensure(sizeof(Aabb) * (size_t)off));
    kn_leaf_aabb<<<div_up_u(lv.count[0], 256), 256, 0, st>>>(d_sxyz.as<float4>(), n, lv.count[0], d_nodes.as<Aabb>());
    for (int l = 1; l < L; ++l)
      kn_merge_level<<<div_up_u(lv.count[l], 256), 256, 0, st>>>(d_nodes.as<Aabb>() + lv.offset[l - 1], lv.count[l - 1],
                                                                 d_nodes.as<Aabb>() + lv.offset[l], lv.count[l]);
    B2_TRY(d_out.ensure(n * 16));
    const bool want_idx = out_knn_idx || (hook && hook->need_idx);
    const int stat_mode = hook ? hook->stat_mode : kKnnStatNone;
    if (want_idx) B2_TRY(d_oidx.ensure(n * (size_t)k * 4));
    if (stat_mode != kKnnStatNone) B2_TRY(d_stat.ensure(n * 4));
    B2_TRY(d_nan.ensure(4));
    B2_CUDA(cudaMemsetAsync(d_nan.p, 0, 4, st));
    if (world > 1) B2_CUDA(cudaMemsetAsync(d_out.p, 0, n * 16, st));
    if (radius_mode) {
      // setRadiusSearch: batches of Morton-sorted queries so that the neighbour lists (12 B per entry, double-buffered) stay bounded
      const float r2 = (float)((double)radius * (double)radius);      // PCL: radiusSearch(point, radius) -> FLANN with (float)(radius*radius)
      const size_t kBatch = 1u << 20;
      B2_TRY(r_cnt.ensure(kBatch * 4)); B2_TRY(r_cnt64.ensure((kBatch + 1) * 8)); B2_TRY(r_offs.ensure((kBatch + 1) * 8));
      if (out_count) { B2_TRY(r_ocount.ensure(n * 4)); B2_CUDA(cudaMemsetAsync(r_ocount.p, 0, n * 4, st)); }
      for (size_t qb = q_begin; qb < q_end; qb += kBatch) {
        const size_t qe = std::min(q_end, qb + kBatch), nb = qe - qb;
        kn_radius_count<<<div_up_u(nb, 128), 128, 0, st>>>(d_sxyz.as<float4>(), n, d_nodes.as<Aabb>(), lv, r2, qb, qe, r_cnt.as<unsigned int>());
        B2_CUDA(cudaMemsetAsync(r_cnt64.as<unsigned long long>() + nb, 0, 8, st));
        kn_widen<<<div_up_u(nb, 256), 256, 0, st>>>(r_cnt.as<unsigned int>(), nb, r_cnt64.as<unsigned long long>());
        size_t t1 = 0;
        B2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t1, r_cnt64.as<unsigned long long>(), r_offs.as<unsigned long long>(), (int)(nb + 1), st));
        B2_TRY(d_tmp.ensure(t1));
        B2_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, t1, r_cnt64.as<unsigned long long>(), r_offs.as<unsigned long long>(), (int)(nb + 1), st));
        unsigned long long total = 0;
        B2_CUDA(cudaMemcpyAsync(&total, r_offs.as<unsigned long long>() + nb, 8, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        if (total > 0x7FFFFFFFull) return set_error(B2_ERR_ARG, "radius %g gathers more than 2^31 neighbours per 2^20 points", (double)radius);
        const size_t tot = std::max<size_t>((size_t)total, 1);
        B2_TRY(r_keys.ensure(tot * 8)); B2_TRY(r_keys2.ensure(tot * 8)); B2_TRY(r_vals.ensure(tot * 4)); B2_TRY(r_vals2.ensure(tot * 4));
        kn_radius_fill<<<div_up_u(nb, 128), 128, 0, st>>>(d_sxyz.as<float4>(), n, d_nodes.as<Aabb>(), lv, r2, qb, qe, r_offs.as<unsigned long long>(),
                                                          r_keys.as<unsigned long long>(), r_vals.as<unsigned int>());
        size_t t2 = 0;
        B2_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, t2, r_keys.as<unsigned long long>(), r_keys2.as<unsigned long long>(), r_vals.as<unsigned int>(),
                                                    r_vals2.as<unsigned int>(), (int)total, (int)nb, r_offs.as<unsigned long long>(),
                                                    r_offs.as<unsigned long long>() + 1, st));
        B2_TRY(d_tmp.ensure(t2));
        B2_CUDA(cub::DeviceSegmentedSort::SortPairs(d_tmp.p, t2, r_keys.as<unsigned long long>(), r_keys2.as<unsigned long long>(), r_vals.as<unsigned int>(),
                                                    r_vals2.as<unsigned int>(), (int)total, (int)nb, r_offs.as<unsigned long long>(),
                                                    r_offs.as<unsigned long long>() + 1, st));
        kn_radius_normals<<<div_up_u(nb, 128), 128, 0, st>>>(d_sxyz.as<float4>(), qb, qe, r_offs.as<unsigned long long>(), r_vals2.as<unsigned int>(),
                                                             viewpoint[0], viewpoint[1], viewpoint[2], d_out.as<float4>(),
                                                             out_count ? r_ocount.as<int>() : nullptr, d_nan.as<unsigned int>());
      }
    }
    const size_t smem = (size_t)std::max(k, 1) * kKnnThreads * 8;
    const bool trace = std::getenv("B2_KNN_TRACE") != nullptr;
    cudaEvent_t te0 = nullptr, te1 = nullptr;
    if (trace) {
      B2_TRY(d_dbg.ensure(32)); B2_CUDA(cudaMemsetAsync(d_dbg.p, 0, 32, st));
      B2_CUDA(cudaEventCreate(&te0)); B2_CUDA(cudaEventCreate(&te1)); B2_CUDA(cudaEventRecord(te0, st));
    }
    if (!radius_mode && q_end > q_begin && k >= knn_wide_from()) {
      const size_t wsmem = (size_t)kWideWarps * 24 * k;
      B2_CUDA(cudaFuncSetAttribute(kn_knn_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
      kn_knn_wide<<<div_up_u(q_end - q_begin, kWideWarps), 32 * kWideWarps, wsmem, st>>>(d_sxyz.as<float4>(), n, d_nodes.as<Aabb>(), lv, k, viewpoint[0],
                                                                                        viewpoint[1], viewpoint[2],
                                                                                        (out_nxyz_curv || !hook) ? d_out.as<float4>() : nullptr,
                                                                                        want_idx ? d_oidx.as<int>() : nullptr,
                                                                                        d_nan.as<unsigned int>(), q_begin, q_end, stat_mode,
                                                                                        d_stat.as<float>(), trace ? d_dbg.as<unsigned long long>() : nullptr);
    } else if (!radius_mode) B2_CUDA(cudaFuncSetAttribute(kn_knn_normals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!radius_mode && q_end > q_begin && k < knn_wide_from())
      kn_knn_normals<<<div_up_u(q_end - q_begin, kKnnThreads), kKnnThreads, smem, st>>>(d_sxyz.as<float4>(), n, d_nodes.as<Aabb>(), lv, k, viewpoint[0],
                                                                                        viewpoint[1], viewpoint[2],
                                                                                        (out_nxyz_curv || !hook) ? d_out.as<float4>() : nullptr,
                                                                                        want_idx ? d_oidx.as<int>() : nullptr,
                                                                                        d_nan.as<unsigned int>(), q_begin, q_end, stat_mode,
                                                                                        d_stat.as<float>());
    B2_CUDA(cudaGetLastError());
    if (trace) {
      B2_CUDA(cudaEventRecord(te1, st)); B2_CUDA(cudaEventSynchronize(te1));
      float ms = 0.f; cudaEventElapsedTime(&ms, te0, te1);
      unsigned long long c[4] = {0, 0, 0, 0};
      B2_CUDA(cudaMemcpy(c, d_dbg.p, 32, cudaMemcpyDeviceToHost));
      fprintf(stderr, "[b2 knn] %zu points, k = %d: search kernel %.2f ms (%s); per query: %.1f nodes, %.1f leaves, %.1f replacements\n", n, k, ms,
              radius_mode ? "radius" : k >= knn_wide_from() ? "warp per query" : "thread per query", (double)c[0] / n, (double)c[1] / n, (double)c[2] / n);
      cudaEventDestroy(te0); cudaEventDestroy(te1);
    }
    if (hook && hook->run) B2_TRY(hook->run(st, d_xyz.as<float>(), want_idx ? d_oidx.as<int>() : nullptr, d_stat.as<float>(), n, k));
    if (world > 1) {
      B2_TRY(b2_comm_allreduce(comm, d_out.p, n * 4, B2_F32, (void*)st));
      B2_TRY(b2_comm_allreduce(comm, d_nan.p, 1, B2_I32, (void*)st));
    }
    unsigned int nans = 0;
    if (out_nxyz_curv) B2_CUDA(cudaMemcpyAsync(out_nxyz_curv, d_out.p, n * 16, cudaMemcpyDeviceToHost, st));
    if (out_knn_idx) B2_CUDA(cudaMemcpyAsync(out_knn_idx, d_oidx.p, n * (size_t)k * 4, cudaMemcpyDeviceToHost, st));
    if (radius_mode && out_count) {
      if (world > 1) B2_TRY(b2_comm_allreduce(comm, r_ocount.p, n, B2_I32, (void*)st));
      B2_CUDA(cudaMemcpyAsync(out_count, r_ocount.p, n * 4, cudaMemcpyDeviceToHost, st));
    }
    B2_CUDA(cudaMemcpyAsync(&nans, d_nan.p, 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (is_dense) *is_dense = nans == 0;
    return B2_OK;
  };
  rc = body();
  for (DevBuf* b : {&d_xyz, &d_part, &d_keys, &d_keys2, &d_idx, &d_perm, &d_sxyz, &d_nodes, &d_tmp, &d_out, &d_oidx, &d_nan, &d_stat, &d_dbg}) b->release();
  for (DevBuf* b : {&r_cnt, &r_cnt64, &r_offs, &r_keys, &r_keys2, &r_vals, &r_vals2, &r_ocount}) b->release();
  p_part.release();
  cudaStreamDestroy(st);
  return rc;
}

// Exact kNN lists / per-point statistics left on the device for a follow-up stage (b2_cleaner.cu).
int b2::knn_with_hook(const float* xyz, size_t n, size_t stride_bytes, int k, const KnnHook& hook) {
  const float vp[3] = {0.f, 0.f, 0.f};
  if (hook.stat_mode == kKnnStatNone && !hook.need_idx) return set_error(B2_ERR_ARG, "knn_with_hook: nothing requested");
  return normals_impl(xyz, n, stride_bytes, k, vp, nullptr, nullptr, nullptr, nullptr, -1, 0.f, nullptr, &hook);
}

extern "C" int b2_normals_estimate(const float* xyz, size_t n, size_t stride_bytes, int k, const float viewpoint[3], float* out_nxyz_curv,
                                   int32_t* out_knn_idx, int* is_dense) {
  return normals_impl(xyz, n, stride_bytes, k, viewpoint, out_nxyz_curv, out_knn_idx, is_dense, nullptr, -1);
}

extern "C" int b2_normals_estimate_radius(const float* xyz, size_t n, size_t stride_bytes, float radius, const float viewpoint[3], b2_comm* comm,
                                          int device, float* out_nxyz_curv, int32_t* out_neighbor_count, int* is_dense) {
  if (!(radius > 0.f)) return set_error(B2_ERR_ARG, "radius must be > 0");
  return normals_impl(xyz, n, stride_bytes, 0, viewpoint, out_nxyz_curv, nullptr, is_dense, comm, device, radius, out_neighbor_count);
}

extern "C" int b2_normals_estimate_dist(const float* xyz, size_t n, size_t stride_bytes, int k, const float viewpoint[3], b2_comm* comm, int device,
                                        float* out_nxyz_curv, int* is_dense) {
  if (!comm) return set_error(B2_ERR_ARG, "null communicator");
  return normals_impl(xyz, n, stride_bytes, k, viewpoint, out_nxyz_curv, nullptr, is_dense, comm, device);
}
