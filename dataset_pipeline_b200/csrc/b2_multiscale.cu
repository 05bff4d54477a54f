// Multi-resolution point-cloud construction on sm_100a — the producer of Path B's inputs.
//
//   b2_ms_merge_close_points   MergeClosePoints             /root/reference/src/opt/multi_scale_point_cloud.cc:44-124
//   b2_ms_create               CreateMultiScalePointCloud   /root/reference/src/opt/multi_scale_point_cloud.cc:263-368 (scale loop, given
//                                                            the per-point radii of ComputeMinMaxPointRadius)
//   b2_ms_point_neighbors      Problem::DeterminePointNeighbors  /root/reference/src/opt/problem.cc:706-786
//
// MergeClosePoints is a sequential greedy sweep: point i becomes the centre of a merged point unless an EARLIER centre lies within
// the merge distance; a centre averages ALL points within the merge distance (in radiusSearch order) and marks them done. The set of
// centres is therefore the lexicographically-first maximal independent set (LFMIS) of the "closer than r" graph in index order, and it
// is unique — so it can be computed in any schedule that respects the dependencies:
//   1. implicit BVH over the points (b2_bvh.cuh); per point the LATER neighbours (index > own) as a CSR list, and the number of
//      EARLIER neighbours as a counter `pending`;
//   2. worklist rounds: a point whose earlier neighbours are all decided and none of them is a centre becomes a centre
//      (km_release); a new centre marks its undecided later neighbours covered (km_cover). Work is O(edges), the number of rounds
//      is the depth of the dependency DAG (thousands for raster-ordered scans): rounds are launched 32 at a time without host syncs;
//   3. the centres in ascending index = output order; each centre's full neighbour list is gathered, sorted by (d2, index) with a
//      segmented radix sort (the reference's summation order) and reduced by one thread: fp32 position sum / count, per-scan colour
//      sums and counts, max of max_radius, first scan to reach the maximal count (multi_scale_point_cloud.cc:85-112).
// Bit-exact against the oracle (tests/test_gpu_multiscale.py). HBM / gather bound integer + fp32 work: no tensor cores.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include "b2_bvh.cuh"
#include "b2_common.cuh"

namespace b2 {

static constexpr int kMsMaxScans = 32;        // per-thread per-scan counters of the reduce kernel
static constexpr int kRoundsPerBatch = 32;

// ---- adjacency ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) km_count(const float4* __restrict__ s_xyz, size_t n, const Aabb* __restrict__ nodes, BvhLevels lv, float r2,
                                                unsigned long long* __restrict__ later_cnt, unsigned int* __restrict__ earlier_cnt,
                                                unsigned int* __restrict__ inv_perm) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float4 q = s_xyz[j];
  const unsigned int qi = __float_as_uint(q.w);
  unsigned int later = 0, earlier = 0;
  radius_visit(q, r2, s_xyz, n, nodes, lv, [&](float, unsigned int, unsigned int idx) { later += idx > qi; earlier += idx < qi; });
  later_cnt[qi] = later; earlier_cnt[qi] = earlier; inv_perm[qi] = (unsigned int)j;
}
__global__ void __launch_bounds__(128) km_fill(const float4* __restrict__ s_xyz, size_t n, const Aabb* __restrict__ nodes, BvhLevels lv, float r2,
                                               const unsigned long long* __restrict__ off, unsigned int* __restrict__ adj) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float4 q = s_xyz[j];
  const unsigned int qi = __float_as_uint(q.w);
  unsigned long long w = off[qi];
  radius_visit(q, r2, s_xyz, n, nodes, lv, [&](float, unsigned int, unsigned int idx) { if (idx > qi) adj[w++] = idx; });
}

// ---- LFMIS worklist -----------------------------------------------------------------------------------------------------------------
// state: 0 undecided, 1 centre, 2 covered.
__global__ void __launch_bounds__(256) km_seed(size_t n, const unsigned int* __restrict__ pending, unsigned int* __restrict__ state,
                                               unsigned int* __restrict__ qc, unsigned int* __restrict__ nc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pending[i] == 0) { state[i] = 1u; qc[atomicAdd(nc, 1u)] = (unsigned int)i; } else state[i] = 0u;
}
// new centres cover their undecided later neighbours
__global__ void __launch_bounds__(256) km_cover(const unsigned int* __restrict__ qc, const unsigned int* __restrict__ nc,
                                                const unsigned long long* __restrict__ off, const unsigned int* __restrict__ adj,
                                                unsigned int* __restrict__ state, unsigned int* __restrict__ qv, unsigned int* __restrict__ nv) {
  const unsigned int count = *nc;
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
  for (unsigned int t = warp; t < count; t += nwarps) {
    const unsigned int c = qc[t];
    const unsigned long long b = off[c], e = off[c + 1];
    for (unsigned long long p = b + lane; p < e; p += 32) {
      const unsigned int j = adj[p];
      if (atomicCAS(&state[j], 0u, 2u) == 0u) qv[atomicAdd(nv, 1u)] = j;
    }
  }
}
// newly covered points release their later neighbours; a point with no undecided earlier neighbour left becomes a centre
__global__ void __launch_bounds__(256) km_release(const unsigned int* __restrict__ qv, const unsigned int* __restrict__ nv,
                                                  const unsigned long long* __restrict__ off, const unsigned int* __restrict__ adj,
                                                  unsigned int* __restrict__ state, unsigned int* __restrict__ pending,
                                                  unsigned int* __restrict__ qc_next, unsigned int* __restrict__ nc_next) {
  const unsigned int count = *nv;
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
  for (unsigned int t = warp; t < count; t += nwarps) {
    const unsigned int v = qv[t];
    const unsigned long long b = off[v], e = off[v + 1];
    for (unsigned long long p = b + lane; p < e; p += 32) {
      const unsigned int w = adj[p];
      if (state[w] == 0u && atomicSub(&pending[w], 1u) == 1u) { state[w] = 1u; qc_next[atomicAdd(nc_next, 1u)] = w; }
    }
  }
}

// ---- centres -> merged points ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) km_center_flags(size_t n, const unsigned int* __restrict__ state, unsigned int* __restrict__ flags) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = state[i] == 1u ? 1u : 0u;
}
__global__ void __launch_bounds__(256) km_center_list(size_t n, const unsigned int* __restrict__ flags, const unsigned int* __restrict__ rank,
                                                      const unsigned long long* __restrict__ later_cnt, const unsigned int* __restrict__ earlier_cnt,
                                                      unsigned int* __restrict__ centers, unsigned long long* __restrict__ list_len) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  centers[rank[i]] = (unsigned int)i;
  list_len[rank[i]] = later_cnt[i] + earlier_cnt[i] + 1ull;     // + the centre itself
}
// keys of the centres [cb, ce): (d2 bits << 32) | original index — for non-negative floats the bit pattern orders like the value
__global__ void __launch_bounds__(128) km_center_keys(const float4* __restrict__ s_xyz, size_t n, const Aabb* __restrict__ nodes, BvhLevels lv, float r2,
                                                      const unsigned int* __restrict__ centers, const unsigned int* __restrict__ inv_perm,
                                                      size_t cb, size_t ce, const unsigned long long* __restrict__ list_off, unsigned long long base,
                                                      unsigned long long* __restrict__ keys) {
  const size_t c = cb + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ce) return;
  const float4 q = s_xyz[inv_perm[centers[c]]];
  unsigned long long w = list_off[c] - base;
  radius_visit(q, r2, s_xyz, n, nodes, lv, [&](float d, unsigned int, unsigned int idx) { keys[w++] = ((unsigned long long)__float_as_uint(d) << 32) | idx; });
}
// one thread per centre walks its sorted list (multi_scale_point_cloud.cc:85-122)
__global__ void __launch_bounds__(128) km_reduce(const float* __restrict__ xyz, const float* __restrict__ colors, const unsigned char* __restrict__ scan,
                                                 const float* __restrict__ max_radius, int num_scans, size_t cb, size_t ce,
                                                 const unsigned long long* __restrict__ list_off, unsigned long long base,
                                                 const unsigned long long* __restrict__ keys, float* __restrict__ oxyz, float* __restrict__ ocol,
                                                 unsigned char* __restrict__ oscan, float* __restrict__ omaxr) {
  const size_t c = cb + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ce) return;
  int merged[kMsMaxScans]; float color_sum[kMsMaxScans];
  for (int s = 0; s < num_scans; ++s) { merged[s] = 0; color_sum[s] = 0.f; }
  float ax = 0.f, ay = 0.f, az = 0.f, maxr = -1.f;
  int total = 0, max_scan = -1, max_per_scan = 0;
  const unsigned long long b = list_off[c] - base, e = list_off[c + 1] - base;
  for (unsigned long long p = b; p < e; ++p) {
    const size_t idx = (size_t)(keys[p] & 0xFFFFFFFFull);
    const int s = scan[idx];
    ax = fadd(ax, xyz[3 * idx]); ay = fadd(ay, xyz[3 * idx + 1]); az = fadd(az, xyz[3 * idx + 2]);
    color_sum[s] = fadd(color_sum[s], colors[idx]);
    const float mr = max_radius[idx];
    if (mr > maxr) maxr = mr;
    merged[s] += 1;
    if (merged[s] > max_per_scan) { max_per_scan = merged[s]; max_scan = s; }
    total += 1;
  }
  const float ft = (float)total;
  oxyz[3 * c] = ax / ft; oxyz[3 * c + 1] = ay / ft; oxyz[3 * c + 2] = az / ft;
  ocol[c] = color_sum[max_scan] / (float)merged[max_scan];
  oscan[c] = (unsigned char)max_scan;
  omaxr[c] = maxr;
}

// ---- order-preserving filters of the scale loop (multi_scale_point_cloud.cc:263-340) ------------------------------------------------
// mode 0: radius >= min_radius[i]; mode 1: last_radius < min_radius[i] && radius >= min_radius[i]; mode 2: radius <= max_radius[i]
__global__ void __launch_bounds__(256) km_select_flags(size_t n, const float* __restrict__ min_radius, const float* __restrict__ max_radius, int mode,
                                                       double radius, float last_radius, unsigned int* __restrict__ flags) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool f;
  if (mode == 0) f = radius >= (double)min_radius[i];
  else if (mode == 1) f = last_radius < min_radius[i] && radius >= (double)min_radius[i];
  else f = radius <= (double)max_radius[i];
  flags[i] = f ? 1u : 0u;
}
__global__ void __launch_bounds__(256) km_scatter(size_t n, const unsigned int* __restrict__ flags, const unsigned int* __restrict__ rank, size_t dst0,
                                                  const float* __restrict__ xyz, const float* __restrict__ col, const unsigned char* __restrict__ scan,
                                                  const float* __restrict__ maxr, float* __restrict__ oxyz, float* __restrict__ ocol,
                                                  unsigned char* __restrict__ oscan, float* __restrict__ omaxr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  const size_t d = dst0 + rank[i];
  oxyz[3 * d] = xyz[3 * i]; oxyz[3 * d + 1] = xyz[3 * i + 1]; oxyz[3 * d + 2] = xyz[3 * i + 2];
  ocol[d] = col[i]; oscan[d] = scan[i]; omaxr[d] = maxr[i];
}
__global__ void __launch_bounds__(256) km_minmax(size_t n, const float* __restrict__ lo, const float* __restrict__ hi, float* __restrict__ partial) {
  float mn = INFINITY, mx = -INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { mn = fminf(mn, lo[i]); mx = fmaxf(mx, hi[i]); }
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  __shared__ float s[8][2];
  if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5][0] = mn; s[threadIdx.x >> 5][1] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) { for (int i = 1; i < 8; ++i) { mn = fminf(mn, s[i][0]); mx = fmaxf(mx, s[i][1]); } partial[2 * blockIdx.x] = mn; partial[2 * blockIdx.x + 1] = mx; }
}

// A cloud of the scale loop, resident on the device.
struct MsCloud {
  DevBuf xyz, col, scan, maxr;
  size_t n = 0;
  int reserve(size_t cap) {
    cap = std::max<size_t>(cap, 1);
    B2_TRY(xyz.ensure(cap * 12)); B2_TRY(col.ensure(cap * 4)); B2_TRY(scan.ensure(cap)); B2_TRY(maxr.ensure(cap * 4));
    return B2_OK;
  }
  void release() { xyz.release(); col.release(); scan.release(); maxr.release(); n = 0; }
};

struct MsScratch {
  BvhIndex bvh;
  DevBuf later, earlier, pending, inv, off, adj, state, qc0, qc1, qv, cnt, flags, rank, centers, list_len, list_off, keys, keys2, tmp;
  PinnedBuf pin;
  void release() {
    bvh.release();
    for (DevBuf* b : {&later, &earlier, &pending, &inv, &off, &adj, &state, &qc0, &qc1, &qv, &cnt, &flags, &rank, &centers, &list_len, &list_off, &keys, &keys2, &tmp})
      b->release();
    pin.release();
  }
};

static int exclusive_sum_u64(DevBuf& tmp, const unsigned long long* in, unsigned long long* out, size_t count, cudaStream_t st) {
  size_t t = 0;
  B2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t, in, out, (long long)count, st));
  B2_TRY(tmp.ensure(t));
  B2_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, t, in, out, (long long)count, st));
  return B2_OK;
}
static int exclusive_sum_u32(DevBuf& tmp, const unsigned int* in, unsigned int* out, size_t count, cudaStream_t st) {
  size_t t = 0;
  B2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t, in, out, (long long)count, st));
  B2_TRY(tmp.ensure(t));
  B2_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, t, in, out, (long long)count, st));
  return B2_OK;
}

struct MsStats { uint64_t edges = 0; int rounds = 0; };
struct RebaseOp { unsigned long long base; __host__ __device__ unsigned long long operator()(unsigned long long v) const { return v - base; } };

// MergeClosePoints on device-resident clouds. out must not alias in.
static int merge_device(MsScratch& S, const MsCloud& in, float merge_distance, int num_scans, int sms, cudaStream_t st, MsCloud* out, MsStats* stats) {
  const size_t n = in.n;
  out->n = 0;
  if (n == 0) return B2_OK;
  if (n >= (1ull << 30)) return set_error(B2_ERR_ARG, "clouds above 2^30 points are not supported");
  const float r2 = (float)((double)merge_distance * (double)merge_distance);   // PCL radiusSearch -> FLANN: (float)(r*r), strict <
  B2_TRY(S.bvh.build(in.xyz.as<float>(), n, sms, st));
  const float4* sx = S.bvh.sxyz.as<float4>(); const Aabb* nodes = S.bvh.nodes.as<Aabb>(); const BvhLevels lv = S.bvh.lv;
  B2_TRY(S.later.ensure((n + 1) * 8)); B2_TRY(S.earlier.ensure(n * 4)); B2_TRY(S.pending.ensure(n * 4)); B2_TRY(S.inv.ensure(n * 4));
  B2_TRY(S.off.ensure((n + 1) * 8)); B2_TRY(S.pin.ensure(4096));
  B2_CUDA(cudaMemsetAsync(S.later.as<unsigned long long>() + n, 0, 8, st));
  km_count<<<bvh_div_up(n, 128), 128, 0, st>>>(sx, n, nodes, lv, r2, S.later.as<unsigned long long>(), S.earlier.as<unsigned int>(), S.inv.as<unsigned int>());
  B2_TRY(exclusive_sum_u64(S.tmp, S.later.as<unsigned long long>(), S.off.as<unsigned long long>(), n + 1, st));
  unsigned long long* h64 = S.pin.as<unsigned long long>();
  B2_CUDA(cudaMemcpyAsync(h64, S.off.as<unsigned long long>() + n, 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const unsigned long long edges = h64[0];
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if (edges * 4ull > (unsigned long long)free_b + S.adj.cap)
    return set_error(B2_ERR_ALLOC, "merge distance %g makes %llu neighbour pairs (%.1f GB): more than the free device memory", (double)merge_distance, edges, edges * 4e-9);
  B2_TRY(S.adj.ensure(std::max<unsigned long long>(edges, 1) * 4));
  km_fill<<<bvh_div_up(n, 128), 128, 0, st>>>(sx, n, nodes, lv, r2, S.off.as<unsigned long long>(), S.adj.as<unsigned int>());
  B2_CUDA(cudaMemcpyAsync(S.pending.p, S.earlier.p, n * 4, cudaMemcpyDeviceToDevice, st));
  // worklist rounds
  B2_TRY(S.state.ensure(n * 4)); B2_TRY(S.qc0.ensure(n * 4)); B2_TRY(S.qc1.ensure(n * 4)); B2_TRY(S.qv.ensure(n * 4));
  const int ncnt = 2 * kRoundsPerBatch + 2;
  B2_TRY(S.cnt.ensure(sizeof(unsigned int) * ncnt));
  unsigned int* cnt = S.cnt.as<unsigned int>();      // [2r] = centres entering round r, [2r+1] = points covered in round r
  B2_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * ncnt, st));
  km_seed<<<bvh_div_up(n, 256), 256, 0, st>>>(n, S.pending.as<unsigned int>(), S.state.as<unsigned int>(), S.qc0.as<unsigned int>(), cnt);
  unsigned int* qc[2] = {S.qc0.as<unsigned int>(), S.qc1.as<unsigned int>()};
  unsigned int* hcnt = S.pin.as<unsigned int>() + 16;
  const int grid = sms * 4;
  int rounds = 0, cur = 0;
  while (true) {
    for (int r = 0; r < kRoundsPerBatch; ++r) {
      km_cover<<<grid, 256, 0, st>>>(qc[cur], cnt + 2 * r, S.off.as<unsigned long long>(), S.adj.as<unsigned int>(), S.state.as<unsigned int>(),
                                     S.qv.as<unsigned int>(), cnt + 2 * r + 1);
      km_release<<<grid, 256, 0, st>>>(S.qv.as<unsigned int>(), cnt + 2 * r + 1, S.off.as<unsigned long long>(), S.adj.as<unsigned int>(),
                                       S.state.as<unsigned int>(), S.pending.as<unsigned int>(), qc[cur ^ 1], cnt + 2 * r + 2);
      cur ^= 1;
    }
    rounds += kRoundsPerBatch;
    B2_CUDA(cudaMemcpyAsync(hcnt, cnt + 2 * kRoundsPerBatch, 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    const unsigned int carry = hcnt[0];
    if (carry == 0) break;
    B2_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * ncnt, st));
    B2_CUDA(cudaMemcpyAsync(cnt, hcnt, 4, cudaMemcpyHostToDevice, st));   // centres entering the next batch's first round
  }
  // centres in ascending index
  B2_TRY(S.flags.ensure((n + 1) * 4)); B2_TRY(S.rank.ensure((n + 1) * 4));
  B2_CUDA(cudaMemsetAsync(S.flags.as<unsigned int>() + n, 0, 4, st));
  km_center_flags<<<bvh_div_up(n, 256), 256, 0, st>>>(n, S.state.as<unsigned int>(), S.flags.as<unsigned int>());
  B2_TRY(exclusive_sum_u32(S.tmp, S.flags.as<unsigned int>(), S.rank.as<unsigned int>(), n + 1, st));
  B2_CUDA(cudaMemcpyAsync(hcnt, S.rank.as<unsigned int>() + n, 4, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const size_t nc = hcnt[0];
  if (nc == 0) return set_error(B2_ERR_STATE, "internal: no centre found among %zu points", n);
  B2_TRY(out->reserve(nc));
  B2_TRY(S.centers.ensure(nc * 4)); B2_TRY(S.list_len.ensure((nc + 1) * 8)); B2_TRY(S.list_off.ensure((nc + 1) * 8));
  B2_CUDA(cudaMemsetAsync(S.list_len.as<unsigned long long>() + nc, 0, 8, st));
  km_center_list<<<bvh_div_up(n, 256), 256, 0, st>>>(n, S.flags.as<unsigned int>(), S.rank.as<unsigned int>(), S.later.as<unsigned long long>(),
                                                     S.earlier.as<unsigned int>(), S.centers.as<unsigned int>(), S.list_len.as<unsigned long long>());
  B2_TRY(exclusive_sum_u64(S.tmp, S.list_len.as<unsigned long long>(), S.list_off.as<unsigned long long>(), nc + 1, st));
  // batches of centres whose lists together stay below 2^30 entries (segmented sort counts items with int)
  std::vector<unsigned long long> hoff(nc + 1);
  B2_CUDA(cudaMemcpyAsync(hoff.data(), S.list_off.p, (nc + 1) * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const unsigned long long kMaxItems = 1ull << 30;
  size_t cb = 0;
  while (cb < nc) {
    size_t ce = cb + 1;
    if (hoff[ce] - hoff[cb] > kMaxItems) return set_error(B2_ERR_ARG, "a merged point gathers more than 2^30 neighbours");
    while (ce < nc && hoff[ce + 1] - hoff[cb] <= kMaxItems) ++ce;
    const unsigned long long base = hoff[cb], items = hoff[ce] - base;
    B2_TRY(S.keys.ensure(items * 8)); B2_TRY(S.keys2.ensure(items * 8));
    km_center_keys<<<bvh_div_up(ce - cb, 128), 128, 0, st>>>(sx, n, nodes, lv, r2, S.centers.as<unsigned int>(), S.inv.as<unsigned int>(), cb, ce,
                                                             S.list_off.as<unsigned long long>(), base, S.keys.as<unsigned long long>());
    // segment offsets relative to this batch's first entry
    size_t t = 0;
    auto begin_it = thrust::make_transform_iterator((const unsigned long long*)S.list_off.as<unsigned long long>() + cb, RebaseOp{base});
    auto end_it = thrust::make_transform_iterator((const unsigned long long*)S.list_off.as<unsigned long long>() + cb + 1, RebaseOp{base});
    B2_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, t, S.keys.as<unsigned long long>(), S.keys2.as<unsigned long long>(), (int)items, (int)(ce - cb), begin_it, end_it, st));
    B2_TRY(S.tmp.ensure(t));
    B2_CUDA(cub::DeviceSegmentedSort::SortKeys(S.tmp.p, t, S.keys.as<unsigned long long>(), S.keys2.as<unsigned long long>(), (int)items, (int)(ce - cb), begin_it, end_it, st));
    km_reduce<<<bvh_div_up(ce - cb, 128), 128, 0, st>>>(in.xyz.as<float>(), in.col.as<float>(), in.scan.as<unsigned char>(), in.maxr.as<float>(), num_scans, cb, ce,
                                                        S.list_off.as<unsigned long long>(), base, S.keys2.as<unsigned long long>(), out->xyz.as<float>(),
                                                        out->col.as<float>(), out->scan.as<unsigned char>(), out->maxr.as<float>());
    cb = ce;
  }
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaStreamSynchronize(st));
  out->n = nc;
  if (stats) { stats->edges += edges; stats->rounds += rounds; }
  return B2_OK;
}

// std::mt19937 + libstdc++ 9 (the reference's toolchain, Dockerfile: Ubuntu 20.04) std::shuffle for a 32-bit engine: one draw yields two
// swap positions while range^2 fits the engine's range; uniform_int_distribution scales the engine output down (scaling = 2^32-1 / range)
// and rejects the tail (v >= range * scaling). The engine is one sequential stream over ALL points of a call, so the work is split:
//   * sequential: draw the accepted engine outputs of every point (a compare per output; rejections are ~1e-7 per draw but shift the
//     stream, so they must be replayed exactly);
//   * parallel (host threads): turn the accepted outputs into swap positions and apply them to the neighbour lists.
struct Mt19937 {
  uint32_t mt[624]; uint32_t out[624]; int idx;
  explicit Mt19937(uint32_t seed) { mt[0] = seed; for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i; idx = 624; }
  static inline uint32_t twist(uint32_t u, uint32_t v, uint32_t m) { const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu); return m ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu); }
  // the whole state at once, in three branch-free loops the host compiler vectorises (no element depends on one written in the same loop
  // within 227 positions), then the tempering of all 624 outputs
#if defined(__GNUC__) && !defined(__clang__)
  __attribute__((optimize("O3")))
#endif
  void refill() {
    for (int i = 0; i < 227; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i + 397]);
    for (int i = 227; i < 454; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i - 227]);
    for (int i = 454; i < 623; ++i) mt[i] = twist(mt[i], mt[i + 1], mt[i - 227]);
    mt[623] = twist(mt[623], mt[0], mt[396]);
    for (int i = 0; i < 624; ++i) {
      uint32_t y = mt[i];
      y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
      out[i] = y;
    }
    idx = 0;
  }
  inline uint32_t next() {
    if (idx >= 624) refill();
    return out[idx++];
  }
};

// The draws std::shuffle makes for a list of `len` elements: draw d has `range` outcomes and moves one (b == 0) or two elements.
struct ShuffleDraw { uint64_t range, scaling, past, b; int pos; };
static std::vector<ShuffleDraw> shuffle_plan(uint64_t len) {
  std::vector<ShuffleDraw> plan;
  auto add = [&](uint64_t range, uint64_t b, int pos) {
    ShuffleDraw d; d.range = range; d.b = b; d.pos = pos;
    d.scaling = 0xFFFFFFFFull / range; d.past = range * d.scaling;       // urngrange = 2^32 - 1 > range - 1 always holds here
    plan.push_back(d);
  };
  if (len < 2) return plan;
  if (0xFFFFFFFFull / len >= len) {
    uint64_t i = 1;
    if ((len % 2) == 0) { add(2, 0, (int)i); ++i; }
    while (i != len) { add((i + 1) * (i + 2), i + 2, (int)i); i += 2; }
  } else {
    for (uint64_t i = 1; i != len; ++i) add(i + 1, 0, (int)i);
  }
  return plan;
}
static inline void shuffle_apply(int32_t* first, const std::vector<ShuffleDraw>& plan, const uint32_t* accepted) {
  for (size_t d = 0; d < plan.size(); ++d) {
    const ShuffleDraw& D = plan[d];
    const uint32_t x = accepted[d] / (uint32_t)D.scaling;          // scaling <= 2^32 - 1 and x < range: 32-bit divisions suffice
    if (D.b == 0) { std::swap(first[D.pos], first[x]); }
    else { const uint32_t b = (uint32_t)D.b, q = x / b; std::swap(first[D.pos], first[q]); std::swap(first[D.pos + 1], first[x - q * b]); }
  }
}
// Shuffles rows[i][1..k1) for i in [0, m) with the shared engine, in the reference's point order.
static void shuffle_rows(Mt19937& gen, int32_t* rows, size_t m, int k1) {
  const std::vector<ShuffleDraw> plan = shuffle_plan((uint64_t)(k1 - 1));
  const size_t D = plan.size();
  if (D == 0 || m == 0) return;
  const size_t kBlock = 1u << 20;
  std::vector<uint32_t> acc(std::min(m, kBlock) * D);
  const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  for (size_t b0 = 0; b0 < m; b0 += kBlock) {
    const size_t cnt = std::min(kBlock, m - b0);
    for (size_t i = 0; i < cnt; ++i)
      for (size_t d = 0; d < D; ++d) { uint32_t v; do v = gen.next(); while ((uint64_t)v >= plan[d].past); acc[i * D + d] = v; }
    auto work = [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) shuffle_apply(rows + (b0 + i) * (size_t)k1 + 1, plan, &acc[i * D]); };
    const unsigned nt = (unsigned)std::min<size_t>(hw, std::max<size_t>(1, cnt / 4096));
    if (nt <= 1) { work(0, cnt); continue; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, cnt * t / nt, cnt * (t + 1) / nt);
    for (auto& t : th) t.join();
  }
}

}  // namespace b2

using namespace b2;

namespace {
struct Guard {   // stream + scratch released on every exit path
  cudaStream_t st = nullptr; MsScratch S; MsCloud a, b, src;
  ~Guard() { S.release(); a.release(); b.release(); src.release(); if (st) cudaStreamDestroy(st); }
};
int upload(MsCloud* c, const float* xyz, const float* col, const uint8_t* scan, const float* maxr, size_t n, cudaStream_t st) {
  B2_TRY(c->reserve(n));
  c->n = n;
  if (!n) return B2_OK;
  B2_CUDA(cudaMemcpyAsync(c->xyz.p, xyz, n * 12, cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaMemcpyAsync(c->col.p, col, n * 4, cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaMemcpyAsync(c->scan.p, scan, n, cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaMemcpyAsync(c->maxr.p, maxr, n * 4, cudaMemcpyHostToDevice, st));
  return B2_OK;
}
int check_scans(const uint8_t* scan, size_t n, int num_scans) {
  if (num_scans < 1 || num_scans > kMsMaxScans) return set_error(B2_ERR_ARG, "num_scans must be in [1,%d]", kMsMaxScans);
  for (size_t i = 0; i < n; ++i) if (scan[i] >= num_scans) return set_error(B2_ERR_ARG, "scan index %d of point %zu is outside [0,%d)", (int)scan[i], i, num_scans);
  return B2_OK;
}
}  // namespace

extern "C" int b2_ms_merge_close_points(const float* xyz, size_t n, const float* colors, const uint8_t* scan_indices, const float* max_radius, int num_scans,
                                        float merge_distance, float* out_xyz, float* out_colors, uint8_t* out_scan_indices, float* out_max_radius,
                                        size_t* out_n, b2_ms_stats* stats) {
  if (!out_n || (n && (!xyz || !colors || !scan_indices || !max_radius || !out_xyz || !out_colors || !out_scan_indices || !out_max_radius)))
    return set_error(B2_ERR_ARG, "null argument");
  if (!(merge_distance > 0.f) || !(merge_distance < INFINITY)) return set_error(B2_ERR_ARG, "merge_distance must be positive and finite");
  *out_n = 0;
  B2_TRY(check_scans(scan_indices, n, num_scans));
  int dev = 0, sms = 0;
  B2_TRY(select_device(-1, &dev, &sms));
  if (n == 0) return B2_OK;
  Guard g;
  B2_CUDA(cudaStreamCreateWithFlags(&g.st, cudaStreamNonBlocking));
  B2_TRY(upload(&g.a, xyz, colors, scan_indices, max_radius, n, g.st));
  MsStats ms;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, g.st);
  int rc = merge_device(g.S, g.a, merge_distance, num_scans, sms, g.st, &g.b, &ms);
  cudaEventRecord(e1, g.st); cudaEventSynchronize(e1);
  float msec = 0; cudaEventElapsedTime(&msec, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc != B2_OK) return rc;
  const size_t m = g.b.n;
  B2_CUDA(cudaMemcpyAsync(out_xyz, g.b.xyz.p, m * 12, cudaMemcpyDeviceToHost, g.st));
  B2_CUDA(cudaMemcpyAsync(out_colors, g.b.col.p, m * 4, cudaMemcpyDeviceToHost, g.st));
  B2_CUDA(cudaMemcpyAsync(out_scan_indices, g.b.scan.p, m, cudaMemcpyDeviceToHost, g.st));
  B2_CUDA(cudaMemcpyAsync(out_max_radius, g.b.maxr.p, m * 4, cudaMemcpyDeviceToHost, g.st));
  B2_CUDA(cudaStreamSynchronize(g.st));
  *out_n = m;
  if (stats) { stats->neighbor_pairs = ms.edges; stats->rounds = ms.rounds; stats->ms_device = msec; stats->scales = 1; }
  return B2_OK;
}

extern "C" int b2_ms_create(const float* xyz, size_t n, const float* colors, const uint8_t* scan_indices, const float* min_radius, const float* max_radius,
                            int num_scans, float min_radius_bias, float merge_distance_factor, int max_scales, size_t out_capacity, int* out_scale_count,
                            float* out_radius, uint64_t* out_counts, float* out_xyz, float* out_colors, uint8_t* out_scan_indices, b2_ms_stats* stats) {
  if (!out_scale_count || !out_radius || !out_counts || (n && (!xyz || !colors || !scan_indices || !min_radius || !max_radius || !out_xyz || !out_colors || !out_scan_indices)))
    return set_error(B2_ERR_ARG, "null argument");
  *out_scale_count = 0;
  B2_TRY(check_scans(scan_indices, n, num_scans));
  int dev = 0, sms = 0;
  B2_TRY(select_device(-1, &dev, &sms));
  if (n == 0) return set_error(B2_ERR_ARG, "empty cloud");
  Guard g;
  B2_CUDA(cudaStreamCreateWithFlags(&g.st, cudaStreamNonBlocking));
  cudaStream_t st = g.st;
  B2_TRY(upload(&g.src, xyz, colors, scan_indices, max_radius, n, st));
  DevBuf d_min, d_part, d_flags, d_rank, d_tmp;
  struct Rel { DevBuf* b[5]; ~Rel() { for (DevBuf* x : b) x->release(); } } rel{{&d_min, &d_part, &d_flags, &d_rank, &d_tmp}};
  B2_TRY(d_min.ensure(n * 4)); B2_TRY(d_flags.ensure((n + 1) * 4)); B2_TRY(d_rank.ensure((n + 1) * 4));
  B2_CUDA(cudaMemcpyAsync(d_min.p, min_radius, n * 4, cudaMemcpyHostToDevice, st));
  const int pb = sms * 2;
  B2_TRY(d_part.ensure(sizeof(float) * 2 * pb)); B2_TRY(g.S.pin.ensure(std::max<size_t>(4096, sizeof(float) * 2 * pb + 64)));
  km_minmax<<<pb, 256, 0, st>>>(n, d_min.as<float>(), g.src.maxr.as<float>(), d_part.as<float>());
  float* hp = g.S.pin.as<float>() + 16;
  B2_CUDA(cudaMemcpyAsync(hp, d_part.p, sizeof(float) * 2 * pb, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  float min_radius_value = INFINITY, max_radius_value = -INFINITY;
  for (int b = 0; b < pb; ++b) { min_radius_value = std::min(min_radius_value, hp[2 * b]); max_radius_value = std::max(max_radius_value, hp[2 * b + 1]); }
  if (!(min_radius_value < INFINITY)) return set_error(B2_ERR_ARG, "no point has a finite minimum radius (no image observes the cloud)");
  const float min_point_radius = min_radius_value * min_radius_bias;     // multi_scale_point_cloud.cc:279-280
  double radius = min_point_radius;
  // order-preserving select of `from` (n_from points) by mode into dst at dst0; returns the number selected
  auto select_into = [&](const MsCloud& from, const float* lo, const float* hi, int mode, float last_radius, MsCloud* dst, size_t dst0, size_t* selected) -> int {
    const size_t m = from.n;
    *selected = 0;
    if (m == 0) return B2_OK;
    B2_TRY(d_flags.ensure((m + 1) * 4)); B2_TRY(d_rank.ensure((m + 1) * 4));
    B2_CUDA(cudaMemsetAsync(d_flags.as<unsigned int>() + m, 0, 4, st));
    km_select_flags<<<bvh_div_up(m, 256), 256, 0, st>>>(m, lo, hi, mode, radius, last_radius, d_flags.as<unsigned int>());
    B2_TRY(exclusive_sum_u32(d_tmp, d_flags.as<unsigned int>(), d_rank.as<unsigned int>(), m + 1, st));
    unsigned int* hc = g.S.pin.as<unsigned int>() + 8;
    B2_CUDA(cudaMemcpyAsync(hc, d_rank.as<unsigned int>() + m, 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    *selected = hc[0];
    B2_TRY(dst->reserve(dst0 + *selected));
    km_scatter<<<bvh_div_up(m, 256), 256, 0, st>>>(m, d_flags.as<unsigned int>(), d_rank.as<unsigned int>(), dst0, from.xyz.as<float>(), from.col.as<float>(),
                                                   from.scan.as<unsigned char>(), from.maxr.as<float>(), dst->xyz.as<float>(), dst->col.as<float>(),
                                                   dst->scan.as<unsigned char>(), dst->maxr.as<float>());
    return B2_OK;
  };
  MsCloud* last = &g.a; MsCloud* merged = &g.b;
  MsCloud staging;
  struct RelC { MsCloud* c; ~RelC() { c->release(); } } relc{&staging};
  size_t sel = 0;
  B2_TRY(last->reserve(n));
  B2_TRY(select_into(g.src, d_min.as<float>(), g.src.maxr.as<float>(), 0, -1.f, last, 0, &sel));
  last->n = sel;
  float last_radius = -1;
  int scales = 0; size_t off = 0;
  MsStats ms;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms_merge = 0.f;
  int rc = B2_OK;
  while (true) {
    if (last_radius > 0) {
      // survivors of the previous scale's merged cloud, then the points whose minimum radius was just passed (:297-330)
      size_t kept = 0, fresh = 0;
      B2_TRY(staging.reserve(merged->n + 1));
      B2_TRY(select_into(*merged, merged->maxr.as<float>(), merged->maxr.as<float>(), 2, last_radius, &staging, 0, &kept));
      staging.n = kept;
      // `staging` may be re-allocated by the second select: reserve for the worst case first so the first part stays in place
      MsCloud grown;
      B2_TRY(grown.reserve(kept + n));
      if (kept) {
        B2_CUDA(cudaMemcpyAsync(grown.xyz.p, staging.xyz.p, kept * 12, cudaMemcpyDeviceToDevice, st));
        B2_CUDA(cudaMemcpyAsync(grown.col.p, staging.col.p, kept * 4, cudaMemcpyDeviceToDevice, st));
        B2_CUDA(cudaMemcpyAsync(grown.scan.p, staging.scan.p, kept, cudaMemcpyDeviceToDevice, st));
        B2_CUDA(cudaMemcpyAsync(grown.maxr.p, staging.maxr.p, kept * 4, cudaMemcpyDeviceToDevice, st));
      }
      rc = select_into(g.src, d_min.as<float>(), g.src.maxr.as<float>(), 1, last_radius, &grown, kept, &fresh);
      if (rc != B2_OK) { grown.release(); break; }
      B2_CUDA(cudaStreamSynchronize(st));
      last->release();
      *last = grown;              // shallow hand-over of the buffers
      last->n = kept + fresh;
    }
    if (scales >= max_scales) { rc = set_error(B2_ERR_ARG, "more than %d point scales", max_scales); break; }
    cudaEventRecord(e0, st);
    rc = merge_device(g.S, *last, (float)(merge_distance_factor * radius), num_scans, sms, st, merged, &ms);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float t = 0; cudaEventElapsedTime(&t, e0, e1); ms_merge += t;
    if (rc != B2_OK) break;
    const size_t m = merged->n;
    if (off + m > out_capacity) { rc = set_error(B2_ERR_ARG, "output capacity %zu too small (scale %d needs %zu more)", out_capacity, scales, m); break; }
    out_radius[scales] = (float)radius; out_counts[scales] = m;
    if (m) {
      B2_CUDA(cudaMemcpyAsync(out_xyz + 3 * off, merged->xyz.p, m * 12, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(out_colors + off, merged->col.p, m * 4, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(out_scan_indices + off, merged->scan.p, m, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
    }
    off += m;
    ++scales;
    last_radius = (float)radius;
    radius *= 2;
    const float kTolerance = 0.99f;
    if (radius >= max_radius_value * kTolerance) break;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc != B2_OK) return rc;
  *out_scale_count = scales;
  if (stats) { stats->neighbor_pairs = ms.edges; stats->rounds = ms.rounds; stats->ms_device = ms_merge; stats->scales = scales; }
  return B2_OK;
}

extern "C" int b2_ms_point_neighbors(const float* xyz, size_t n, const uint8_t* scan_indices, int scan_count, int limit_neighbors_to_same_scan_index,
                                     int candidate_count, int neighbor_count, uint64_t* out_neighbor_indices) {
  if (n && (!xyz || !out_neighbor_indices)) return set_error(B2_ERR_ARG, "null argument");
  if (candidate_count < 1 || candidate_count + 1 > 128 || neighbor_count < 1 || neighbor_count > candidate_count)
    return set_error(B2_ERR_ARG, "need 1 <= neighbor_count <= candidate_count <= 127");
  const int k1 = candidate_count + 1;
  Mt19937 gen(0);                                        // std::mt19937 generator(/*seed*/ 0)  (problem.cc:712)
  const bool trace = std::getenv("B2_MS_TRACE") != nullptr;
  const float vp[3] = {0.f, 0.f, 0.f};
  // neighbour lists land in a pinned, grow-only scratch of the library (one D2H at PCIe rate; pageable destinations cost 3-10x)
  static PinnedBuf scratch;
  static std::mutex scratch_mutex;
  std::lock_guard<std::mutex> lock(scratch_mutex);
  struct Rows { int32_t* p = nullptr; int32_t* data() const { return p; } int32_t& operator[](size_t i) const { return p[i]; } } idx;
  auto knn = [&](const float* pts, size_t m, Rows* rows) -> int {
    B2_TRY(scratch.ensure(std::max<size_t>(m, 1) * (size_t)k1 * 4));
    rows->p = scratch.as<int32_t>();
    int dense = 0;
    return b2_normals_estimate(pts, m, 12, k1, vp, nullptr, rows->p, &dense);            // K7's exact kNN lists, (d2, index) order
  };
  if (limit_neighbors_to_same_scan_index) {
    if (!scan_indices || scan_count < 1 || scan_count > 256) return set_error(B2_ERR_ARG, "bad scan arguments");
    std::vector<std::vector<float>> clouds(scan_count); std::vector<std::vector<size_t>> orig(scan_count);
    {
      std::vector<size_t> cnt(scan_count, 0);
      for (size_t i = 0; i < n; ++i) if (scan_indices[i] < scan_count) ++cnt[scan_indices[i]];
      for (int s = 0; s < scan_count; ++s) { clouds[s].reserve(3 * cnt[s]); orig[s].reserve(cnt[s]); }
    }
    for (size_t i = 0; i < n; ++i) {
      const int s = scan_indices[i];
      if (s >= scan_count) return set_error(B2_ERR_ARG, "scan index %d of point %zu is outside [0,%d)", s, i, scan_count);
      clouds[s].insert(clouds[s].end(), xyz + 3 * i, xyz + 3 * i + 3); orig[s].push_back(i);
    }
    for (int s = 0; s < scan_count; ++s)
      if ((int)orig[s].size() < k1) return set_error(B2_ERR_STATE, "scan %d has %zu points, fewer than point_neighbor_candidate_count + 1 (reference: CHECK_GE, problem.cc:738)", s, orig[s].size());
    // two pinned buffers: the host shuffles scan s (one worker thread, so the engine stream stays in scan order) while the GPU searches scan s + 1
    static PinnedBuf scratch2;
    std::thread worker;
    int rc = B2_OK;
    for (int s = 0; s < scan_count && rc == B2_OK; ++s) {
      const auto t0 = std::chrono::steady_clock::now();
      PinnedBuf& buf = (s & 1) ? scratch2 : scratch;
      const size_t m = orig[s].size();
      rc = buf.ensure(std::max<size_t>(m, 1) * (size_t)k1 * 4);
      int dense = 0;
      if (rc == B2_OK) rc = b2_normals_estimate(clouds[s].data(), m, 12, k1, vp, nullptr, buf.as<int32_t>(), &dense);
      const auto t1 = std::chrono::steady_clock::now();
      if (worker.joinable()) worker.join();
      if (rc != B2_OK) break;
      int32_t* rows = buf.as<int32_t>();
      worker = std::thread([&, s, m, rows, t0, t1]() {
        const auto t2 = std::chrono::steady_clock::now();
        shuffle_rows(gen, rows, m, k1);
        for (size_t i = 0; i < m; ++i) {
          const int32_t* row = rows + i * (size_t)k1;
          for (int k = 0; k < neighbor_count; ++k) out_neighbor_indices[orig[s][i] * (size_t)neighbor_count + k] = orig[s][(size_t)row[k + 1]];
        }
        if (trace) fprintf(stderr, "[b2_ms_point_neighbors] scan %d: %zu points, kNN %.1f ms, shuffle + scatter %.1f ms (overlapped with the next search)\n", s, m,
                           std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count());
      });
    }
    if (worker.joinable()) worker.join();
    return rc;
    return B2_OK;
  }
  if ((int)std::min<size_t>(n, 1u << 20) < k1) return set_error(B2_ERR_STATE, "cloud has %zu points, fewer than point_neighbor_candidate_count + 1", n);
  const auto t0 = std::chrono::steady_clock::now();
  B2_TRY(knn(xyz, n, &idx));
  const auto t1 = std::chrono::steady_clock::now();
  for (size_t i = 0; i < n; ++i)
    if (idx[i * (size_t)k1] != (int32_t)i) return set_error(B2_ERR_STATE, "point %zu is not its own nearest neighbour (duplicate points; reference: CHECK_EQ, problem.cc:773)", i);
  shuffle_rows(gen, idx.data(), n, k1);
  for (size_t i = 0; i < n; ++i) {
    const int32_t* row = idx.data() + i * (size_t)k1;
    for (int k = 0; k < neighbor_count; ++k) out_neighbor_indices[i * (size_t)neighbor_count + k] = (uint64_t)row[k + 1];
  }
  if (trace) fprintf(stderr, "[b2_ms_point_neighbors] %zu points, kNN %.1f ms, shuffle + scatter %.1f ms\n", n,
                     std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
  return B2_OK;
}
