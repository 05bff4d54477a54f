// Device kernels of Path A (multi-scan point-to-plane ICP) for sm_100a.
//
//   K1  k_bbox / k_keys / k_apply   transform to the global frame + AABB + cell keys + sorted SoA copies
//                                   (replaces pcl::transformPointCloudWithNormals + bbox loops, icp_point_to_plane.cc:189-205)
//   K2  k_count_cells / k_hash_*    occupied-cell hash over the cell-sorted target (replaces the per-pair kd-tree build, :46-51)
//   K3  k_nn_radius1                nearest target within radius per source point (replaces radiusSearch loop, :63-102)
//   K4  k_pack                      48 B packed correspondence records (p_s,n_s,p_t,n_t), three float4 planes
//   K5  k_accumulate                one streaming pass: cost + 6x6 S + 6-vector g per correspondence set, fp64
//                                   (replaces compute() loops, icp_point_to_plane_impl.h:129-211 and :240-266)
//   K6  k_finalize                  fixed-order reduction of the per-CTA partials + assembly of the normal equations
//                                   with the reference's upper-triangle quirk (impl.h:82-113 + :226)
// All of these are HBM / gather bound integer+fp32+fp64 streaming work: no tensor cores.
#pragma once
#include "b2_common.cuh"

namespace b2 {

static constexpr int kAccThreads = 256;
static constexpr int kAccVals = 28;   // 21 (upper S) + 6 (g) + 1 (cost)

// ------------------------------------------------------------------------------------------------------------------
// K1a: AABB of the transformed cloud. One partial (6 floats) per block; the host finishes the reduction.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bbox(const float* __restrict__ xyz, size_t n, Mat4 T, float* __restrict__ partial) {
  float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float3 p = xform_point(T, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    mnx = fminf(mnx, p.x); mny = fminf(mny, p.y); mnz = fminf(mnz, p.z);
    mxx = fmaxf(mxx, p.x); mxy = fmaxf(mxy, p.y); mxz = fmaxf(mxz, p.z);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
  }
  __shared__ float s[8][6];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s[w][0] = mnx; s[w][1] = mny; s[w][2] = mnz; s[w][3] = mxx; s[w][4] = mxy; s[w][5] = mxz; }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = s[0][threadIdx.x];
    for (int i = 1; i < 8; ++i) v = threadIdx.x < 3 ? fminf(v, s[i][threadIdx.x]) : fmaxf(v, s[i][threadIdx.x]);
    partial[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

// K1b: cell key of every transformed point (+ identity permutation).
__global__ void __launch_bounds__(256) k_keys(const float* __restrict__ xyz, size_t n, Mat4 T, GridParams g,
                                              unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float3 p = xform_point(T, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  const double fx = ((double)p.x - g.ox) * g.inv, fy = ((double)p.y - g.oy) * g.inv, fz = ((double)p.z - g.oz) * g.inv;
  const int cx = (int)floor(fx), cy = (int)floor(fy), cz = (int)floor(fz);
  keys[i] = (cell_key(g, cx, cy, cz) << (3 * g.fbits)) | fine_code(fx - cx, fy - cy, fz - cz, g.fbits);
  idx[i] = (unsigned int)i;
}

// K1c: cell-sorted global-frame copies: s_xyz[j] = (p, bits(original index)), s_nrm[j] = (n, 0).
__global__ void __launch_bounds__(256) k_apply(const float* __restrict__ xyz, const float* __restrict__ nrm, size_t n, Mat4 T,
                                               const unsigned int* __restrict__ perm, float4* __restrict__ s_xyz,
                                               float4* __restrict__ s_nrm) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned int i = perm[j];
  const float3 p = xform_point(T, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]);
  s_xyz[j] = make_float4(p.x, p.y, p.z, __uint_as_float(i));
  if (nrm) {
    const float3 q = xform_normal(T, nrm[3 * (size_t)i], nrm[3 * (size_t)i + 1], nrm[3 * (size_t)i + 2]);
    s_nrm[j] = make_float4(q.x, q.y, q.z, 0.f);
  }
}

// Plain transform into packed float3 arrays (fixed-cloud concatenation, icp_point_to_plane.cc:118-126).
__global__ void __launch_bounds__(256) k_transform(const float* __restrict__ xyz, const float* __restrict__ nrm, size_t n, Mat4 T,
                                                   float* __restrict__ oxyz, float* __restrict__ onrm) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float3 p = xform_point(T, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  oxyz[3 * i] = p.x; oxyz[3 * i + 1] = p.y; oxyz[3 * i + 2] = p.z;
  const float3 q = xform_normal(T, nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
  onrm[3 * i] = q.x; onrm[3 * i + 1] = q.y; onrm[3 * i + 2] = q.z;
}

// ------------------------------------------------------------------------------------------------------------------
// K2: occupied-cell hash table over the sorted keys. Entry = {key, begin, end} (16 B, one LDG.128 per probe).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_cells(const unsigned long long* __restrict__ keys, size_t n, unsigned int* __restrict__ count,
                                                     int kFineBits) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool head = j < n && (j == 0 || (keys[j] >> kFineBits) != (keys[j - 1] >> kFineBits));
  const unsigned int m = __ballot_sync(0xffffffffu, head);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned int)__popc(m));
}

__global__ void __launch_bounds__(256) k_hash_insert(const unsigned long long* __restrict__ keys, size_t n, HashEntry* __restrict__ table,
                                                     int log2size, int kFineBits) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long key = keys[j] >> kFineBits;
  if (j != 0 && (keys[j - 1] >> kFineBits) == key) return;
  const unsigned int mask = (1u << log2size) - 1u;
  unsigned int s = hash_slot(key, log2size);
  while (true) {
    const unsigned long long prev = atomicCAS(&table[s].key, kEmptyKey, key);
    if (prev == kEmptyKey) { table[s].begin = (unsigned int)j; return; }
    s = (s + 1) & mask;
  }
}

__device__ __forceinline__ bool hash_find(const HashEntry* __restrict__ table, int log2size, unsigned long long key,
                                          unsigned int* begin, unsigned int* end) {
  const unsigned int mask = (1u << log2size) - 1u;
  unsigned int s = hash_slot(key, log2size);
  while (true) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(table + s));
    const unsigned long long k = ((unsigned long long)e.y << 32) | e.x;
    if (k == key) { *begin = e.z; *end = e.w; return true; }
    if (k == kEmptyKey) return false;
    s = (s + 1) & mask;
  }
}

__global__ void __launch_bounds__(256) k_hash_ends(const unsigned long long* __restrict__ keys, size_t n, HashEntry* __restrict__ table,
                                                   int log2size, int kFineBits) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long key = keys[j] >> kFineBits;
  if (j + 1 != n && (keys[j + 1] >> kFineBits) == key) return;
  const unsigned int mask = (1u << log2size) - 1u;
  unsigned int s = hash_slot(key, log2size);
  while (table[s].key != key) s = (s + 1) & mask;
  table[s].end = (unsigned int)(j + 1);
}

// ------------------------------------------------------------------------------------------------------------------
// K3: nearest target within radius (strict d2 < r2), lowest ORIGINAL target index on exact ties.
// d2 = ((dx*dx)+(dy*dy))+(dz*dz) in fp32 without contraction (FLANN L2_Simple order).
// Grid cells are >= 2d wide, so every target within d of a query lies in the 2x2x2 block of cells on the query's side of
// its own cell (per axis: the neighbour across the NEARER face; the farther face is >= cell/2 >= d away). One thread per
// cell-sorted source point: the query's own cell is probed and scanned first, the other seven only when their nearest face
// is not already farther than the best match (most matched queries touch 1-2 cells).
// Neighbouring threads share cells, so probes and candidate rows are served from L1/L2. Output at the sorted source position.
// ------------------------------------------------------------------------------------------------------------------
// Chunk boxes over the cell-sorted points: level 1 = 32 consecutive points (one warp each), level 2 = 32 level-1 boxes.
__global__ void __launch_bounds__(256) k_chunk_boxes1(const float4* __restrict__ s_xyz, size_t n, Aabb* __restrict__ box1, unsigned int nbox1) {
  const unsigned int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= nbox1) return;
  const size_t p = (size_t)c * kChunk1 + lane;
  float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
  if (p < n) { const float4 v = s_xyz[p]; lx = hx = v.x; ly = hy = v.y; lz = hz = v.z; }
  for (int o = 16; o > 0; o >>= 1) {
    lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
    lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
    hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
  }
  if (lane == 0) { Aabb b; b.lo[0] = lx; b.lo[1] = ly; b.lo[2] = lz; b.hi[0] = hx; b.hi[1] = hy; b.hi[2] = hz; box1[c] = b; }
}
__global__ void __launch_bounds__(256) k_chunk_boxes2(const Aabb* __restrict__ box1, unsigned int nbox1, Aabb* __restrict__ box2, unsigned int nbox2) {
  const unsigned int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= nbox2) return;
  const unsigned int i = c * 32 + lane;
  float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
  if (i < nbox1) { const Aabb v = box1[i]; lx = v.lo[0]; ly = v.lo[1]; lz = v.lo[2]; hx = v.hi[0]; hy = v.hi[1]; hz = v.hi[2]; }
  for (int o = 16; o > 0; o >>= 1) {
    lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
    lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
    hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
  }
  if (lane == 0) { Aabb b; b.lo[0] = lx; b.lo[1] = ly; b.lo[2] = lz; b.hi[0] = hx; b.hi[1] = hy; b.hi[2] = hz; box2[c] = b; }
}

struct SearchWork { unsigned int points, box1, box2, cells; };   // per-query work counters of the diagnostic K3 variant (B2_K3_WORK)

// candidate test; ties on d2 go to the lower original target index so the result does not depend on the visiting order.
// (d2, index) is ONE 64-bit key, (d2 bits << 32) | index: for non-negative floats the integer order is the float order, so a single
// unsigned compare implements "d2 < best, or d2 == best and lower index"; the initial key (r2 bits << 32) | 0 rejects d2 == r2 for every
// index (the radius test is strict). The pruning tests read the best d2 back from the key's high word (key_d2): no second accumulator.
#define B2_NN_TEST(T, P)                                                                                              \
  {                                                                                                                   \
    const float ax_ = fsub(q.x, (T).x), ay_ = fsub(q.y, (T).y), az_ = fsub(q.z, (T).z);                                \
    const float d_ = fadd(fadd(fmul(ax_, ax_), fmul(ay_, ay_)), fmul(az_, az_));                                      \
    const unsigned long long k_ = ((unsigned long long)__float_as_uint(d_) << 32) | (unsigned long long)__float_as_uint((T).w); \
    if (k_ < best_key) { best_key = k_; best_pos = (int)(P); }                                                        \
  }

__device__ __forceinline__ float key_d2(unsigned long long key) { return __uint_as_float((unsigned int)(key >> 32)); }

__device__ __forceinline__ void scan_range(const float4* __restrict__ tgt, unsigned int b, unsigned int e, const float4& q,
                                           int& best_pos, unsigned long long& best_key, SearchWork& wk) {
  wk.points += e - b;
  unsigned int p = b;
  for (; p + 3 < e; p += 4) {   // four candidate loads in flight: the heavy lanes of this kernel are latency bound
    const float4 t0 = __ldg(tgt + p), t1 = __ldg(tgt + p + 1), t2 = __ldg(tgt + p + 2), t3 = __ldg(tgt + p + 3);
    B2_NN_TEST(t0, p) B2_NN_TEST(t1, p + 1) B2_NN_TEST(t2, p + 2) B2_NN_TEST(t3, p + 3)
  }
  for (; p < e; ++p) {
    const float4 t0 = __ldg(tgt + p);
    B2_NN_TEST(t0, p)
  }
}
// candidates p, p+stride, p+2 stride, p+3 stride (those below e)
__device__ __forceinline__ void scan_strided4(const float4* __restrict__ tgt, unsigned int p, unsigned int stride, unsigned int e, const float4& q,
                                              int& best_pos, unsigned long long& best_key, SearchWork& wk) {
  const unsigned int p1 = p + stride, p2 = p + 2u * stride, p3 = p + 3u * stride;
  const float4 t0 = __ldg(tgt + p), t1 = __ldg(tgt + min(p1, e - 1)), t2 = __ldg(tgt + min(p2, e - 1)), t3 = __ldg(tgt + min(p3, e - 1));
  wk.points += 1u + (p1 < e) + (p2 < e) + (p3 < e);
  B2_NN_TEST(t0, p)
  if (p1 < e) B2_NN_TEST(t1, p1)
  if (p2 < e) B2_NN_TEST(t2, p2)
  if (p3 < e) B2_NN_TEST(t3, p3)
}

// One cell's candidates [b,e). Small cells are scanned directly; dense cells (scanner-zenith clusters reach 10^4..10^5 points in
// one cell) go through the chunk boxes so the work per query stays bounded. What remains expensive is inherent to exact search:
// a query a few millimetres off a dense slab (range-noise outliers) has to visit every chunk whose box is nearer than its true
// neighbour — ~10^3 candidates against a mean of ~20 (measured with B2_K3_WORK / tools/k3_work.py). Those lanes are latency
// bound, so scan_range keeps four candidate loads in flight, and the launch order of the CTAs is longest-first (k_cta_cost).
__device__ __forceinline__ void scan_cell(const float4* __restrict__ tgt, const Aabb* __restrict__ box1, const Aabb* __restrict__ box2,
                                          unsigned int b, unsigned int e, const float4& q, int& best_pos, unsigned long long& best_key,
                                          SearchWork& wk) {
  ++wk.cells;
  if (e - b <= 48u) { scan_range(tgt, b, e, q, best_pos, best_key, wk); return; }
  const unsigned int last = e - 1;
  if (e - b > 2u * kChunk2) {
    // Very dense cell: a strided sample of 64 candidates first (independent loads), so that `best` is already tight when the
    // chunk boxes are tested. Sampled points are real candidates; re-visiting them later changes nothing.
    const unsigned int stride = (e - b) / 64u;
    for (unsigned int p = b; p < e; p += 4u * stride) scan_strided4(tgt, p, stride, e, q, best_pos, best_key, wk);
  }
  for (unsigned int c2 = b / kChunk2; c2 <= last / kChunk2; ++c2) {
    ++wk.box2;
    if (e - b > 2u * kChunk2 && dist2_box(q.x, q.y, q.z, box2[c2]) > key_d2(best_key)) continue;
    const unsigned int c1b = max(b / kChunk1, c2 * 32u), c1e = min(last / kChunk1, c2 * 32u + 31u);
    for (unsigned int c1 = c1b; c1 <= c1e; ++c1) {
      ++wk.box1;
      if (dist2_box(q.x, q.y, q.z, box1[c1]) > key_d2(best_key)) continue;
      scan_range(tgt, max(b, c1 * kChunk1), min(e, (c1 + 1u) * kChunk1), q, best_pos, best_key, wk);
    }
  }
}

// Launch-order heuristic for K3: estimated cost of each 128-query CTA = target population of the cells of four of its queries.
// The CTAs are then issued longest-first (the ids sorted by descending cost), so the expensive ones (queries inside scanner-zenith
// clusters; with the z-major cell key they would otherwise sit at the very end of the grid and run alone) overlap with the rest.
__global__ void __launch_bounds__(256) k_cta_cost(const float4* __restrict__ src, size_t ns, const HashEntry* __restrict__ table, int log2size,
                                                  GridParams g, unsigned int ncta, unsigned int* __restrict__ cost, unsigned int* __restrict__ ids) {
  const unsigned int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncta) return;
  const unsigned int mask = (1u << log2size) - 1u;
  unsigned int total = 0;
  for (int k = 0; k < 4; ++k) {
    const size_t j = (size_t)c * 128 + 32 * k;
    if (j >= ns) break;
    const float4 q = src[j];
    const int cx = cell_of(q.x, g.ox, g.inv), cy = cell_of(q.y, g.oy, g.inv), cz = cell_of(q.z, g.oz, g.inv);
    if (cx < 0 || cx >= g.nx || cy < 0 || cy >= g.ny || cz < 0 || cz >= g.nz) continue;
    const unsigned long long key = cell_key(g, cx, cy, cz);
    unsigned int s = hash_slot(key, log2size);
    uint4 e = __ldg(reinterpret_cast<const uint4*>(table + s));
    unsigned long long kk = ((unsigned long long)e.y << 32) | e.x;
    while (kk != key && kk != kEmptyKey) { s = (s + 1) & mask; e = __ldg(reinterpret_cast<const uint4*>(table + s)); kk = ((unsigned long long)e.y << 32) | e.x; }
    if (kk == key) total += e.w - e.z;
  }
  cost[c] = total; ids[c] = c;
}

template <bool STATS>
__global__ void __launch_bounds__(128) k_nn_radius1(const float4* __restrict__ src, size_t ns, const float4* __restrict__ tgt,
                                                    const Aabb* __restrict__ box1, const Aabb* __restrict__ box2,
                                                    const HashEntry* __restrict__ table, int log2size, GridParams g, float r2,
                                                    int* __restrict__ match_pos, float* __restrict__ match_d2,
                                                    unsigned int* __restrict__ flags, uint4* __restrict__ work,
                                                    const unsigned int* __restrict__ order) {
  const size_t j = (size_t)(order ? order[blockIdx.x] : blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const float4 q = src[j];
  SearchWork wk = {0u, 0u, 0u, 0u};
  const double fx = ((double)q.x - g.ox) * g.inv, fy = ((double)q.y - g.oy) * g.inv, fz = ((double)q.z - g.oz) * g.inv;
  const int cx = (int)floor(fx), cy = (int)floor(fy), cz = (int)floor(fz);
  const double rx = fx - cx, ry = fy - cy, rz = fz - cz;               // position inside the cell, [0,1)
  const int sx = rx < 0.5 ? -1 : 1, sy = ry < 0.5 ? -1 : 1, sz = rz < 0.5 ? -1 : 1;
  // distance to the nearer face per axis, shrunk so that fp32 rounding of d2 can never beat the bound
  const double cell = g.cell;
  const float ex = (float)((rx < 0.5 ? rx : 1.0 - rx) * cell * 0.9999);
  const float ey = (float)((ry < 0.5 ? ry : 1.0 - ry) * cell * 0.9999);
  const float ez = (float)((rz < 0.5 ? rz : 1.0 - rz) * cell * 0.9999);
  const float ex2 = ex * ex * 0.9999f, ey2 = ey * ey * 0.9999f, ez2 = ez * ez * 0.9999f;

  // Own cell first; a neighbour (bit0 = x, bit1 = y, bit2 = z) is probed only if its nearest face is not already farther than the
  // best match, which removes most of the 8 probes for matched queries. The neighbours a lane still needs are kept as a bit mask and
  // popped in a per-lane loop: in round r every lane probes ITS r-th remaining neighbour, whichever that is, so the ~0.35 neighbour
  // probes per query of a warp run side by side instead of one mostly idle pass per neighbour slot (measured: search 36.8 -> 33.9 ms
  // per outer iteration at config 2).
  const unsigned int mask = (1u << log2size) - 1u;
  int best_pos = -1;
  unsigned long long best_key = (unsigned long long)__float_as_uint(r2) << 32;
  auto probe = [&](int c) {
    const int x = cx + ((c & 1) ? sx : 0), y = cy + ((c & 2) ? sy : 0), z = cz + ((c & 4) ? sz : 0);
    if (x < 0 || x >= g.nx || y < 0 || y >= g.ny || z < 0 || z >= g.nz) return;
    const unsigned long long key = cell_key(g, x, y, z);
    unsigned int s = hash_slot(key, log2size);
    uint4 e = __ldg(reinterpret_cast<const uint4*>(table + s));
    unsigned long long k = ((unsigned long long)e.y << 32) | e.x;
    while (k != key && k != kEmptyKey) {      // linear probing (rare at load factor <= 0.5)
      s = (s + 1) & mask;
      e = __ldg(reinterpret_cast<const uint4*>(table + s));
      k = ((unsigned long long)e.y << 32) | e.x;
    }
    if (k == key) scan_cell(tgt, box1, box2, e.z, e.w, q, best_pos, best_key, wk);
  };
  probe(0);
  unsigned int todo = 0;
#pragma unroll
  for (int c = 1; c < 8; ++c) {
    const float lb = ((c & 1) ? ex2 : 0.f) + ((c & 2) ? ey2 : 0.f) + ((c & 4) ? ez2 : 0.f);
    if (!(lb > key_d2(best_key))) todo |= 1u << c;
  }
  while (todo) {
    const int c = __ffs(todo) - 1;
    todo &= todo - 1u;
    const float lb = ((c & 1) ? ex2 : 0.f) + ((c & 2) ? ey2 : 0.f) + ((c & 4) ? ez2 : 0.f);
    if (lb > key_d2(best_key)) continue;      // an earlier neighbour tightened the bound
    probe(c);
  }
  if (STATS) work[j] = make_uint4(wk.points, wk.box1, wk.box2, wk.cells);
  match_pos[j] = best_pos;
  match_d2[j] = key_d2(best_key);
  flags[j] = best_pos >= 0 ? 1u : 0u;
}

// ------------------------------------------------------------------------------------------------------------------
// K4: packed correspondence records, three float4 planes (48 B / correspondence, fully coalesced in K5):
//   A = (ps.x, ps.y, ps.z, ns.x)  B = (ns.y, ns.z, pt.x, pt.y)  C = (pt.z, nt.x, nt.y, nt.z)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const float4* __restrict__ s_xyz_src, const float4* __restrict__ s_nrm_src, size_t ns,
                                              const float4* __restrict__ s_xyz_tgt, const float4* __restrict__ s_nrm_tgt,
                                              const int* __restrict__ match_pos, const unsigned int* __restrict__ offs,
                                              unsigned long long base, float4* __restrict__ ra, float4* __restrict__ rb,
                                              float4* __restrict__ rc) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const int p = match_pos[j];
  if (p < 0) return;
  const float4 ps = s_xyz_src[j], nsr = s_nrm_src[j];
  const float4 pt = __ldg(s_xyz_tgt + p), nt = __ldg(s_nrm_tgt + p);
  const unsigned long long o = base + offs[j];
  ra[o] = make_float4(ps.x, ps.y, ps.z, nsr.x);
  rb[o] = make_float4(nsr.y, nsr.z, pt.x, pt.y);
  rc[o] = make_float4(pt.z, nt.x, nt.y, nt.z);
}

// Correspondence list in the caller's (original) indexing, scattered to the original query index.
__global__ void __launch_bounds__(256) k_scatter_matches(const float4* __restrict__ s_xyz_src, size_t ns, const float4* __restrict__ s_xyz_tgt,
                                                         const int* __restrict__ match_pos, const float* __restrict__ match_d2,
                                                         int* __restrict__ out_match, float* __restrict__ out_d2) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const unsigned int qi = __float_as_uint(s_xyz_src[j].w);
  const int p = match_pos[j];
  out_match[qi] = p >= 0 ? (int)__float_as_uint(__ldg(s_xyz_tgt + p).w) : -1;
  out_d2[qi] = match_d2[j];
}

// ------------------------------------------------------------------------------------------------------------------
// K5: one streaming pass over the packed records.
// The record array is the concatenation of the correspondence sets ("segments"); CTA b owns the contiguous range
// [b*per_cta, (b+1)*per_cta) and flushes a 28-double partial per segment it touches: a fixed partition and a fixed
// reduction tree, so the result (and therefore every LM accept/reject decision) is reproducible run to run.
// Per record (fp32, evaluation order of icp_point_to_plane_impl.h:146-204, no FMA contraction):
//   ps = Rs*ps0 + ts, ns = Rs*ns0, pt = Rt*pt0 + tt, nt = Rt*nt0
//   r1 = ns.(pt-ps)   j1 = [ns ; pt x ns]        (d r1 / d target pose; d/d source pose = -j1)
//   r2 = nt.(ps-pt)   j2 = [nt ; ps x nt]        (d r2 / d source pose; d/d target pose = -j2)
// accumulated in fp64 (products of fp32-valued doubles are exact):  S += j1 j1^T + j2 j2^T,  g += r1 j1 - r2 j2,
// cost += r1*r1 + r2*r2 (fp32 squares, as the reference).
// ------------------------------------------------------------------------------------------------------------------
struct CloudPose { float R[9]; float t[3]; };      // increment of one impl cloud (row-major R)
struct Segment { unsigned long long begin, end; int src, tgt; };   // records [begin,end), impl cloud indices

__device__ __forceinline__ void load_pose(const CloudPose* __restrict__ P, int i, float R[9], float t[3]) {
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = __ldg(&P[i].R[k]);
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = __ldg(&P[i].t[k]);
}

__device__ __forceinline__ void accumulate_record(const float4 a, const float4 b, const float4 c, const float Rs[9], const float ts[3],
                                                  const float Rt[9], const float tt[3], double acc[kAccVals]) {
  const float psx = fadd(sum3(fmul(Rs[0], a.x), fmul(Rs[1], a.y), fmul(Rs[2], a.z)), ts[0]);
  const float psy = fadd(sum3(fmul(Rs[3], a.x), fmul(Rs[4], a.y), fmul(Rs[5], a.z)), ts[1]);
  const float psz = fadd(sum3(fmul(Rs[6], a.x), fmul(Rs[7], a.y), fmul(Rs[8], a.z)), ts[2]);
  const float nsx = sum3(fmul(Rs[0], a.w), fmul(Rs[1], b.x), fmul(Rs[2], b.y));
  const float nsy = sum3(fmul(Rs[3], a.w), fmul(Rs[4], b.x), fmul(Rs[5], b.y));
  const float nsz = sum3(fmul(Rs[6], a.w), fmul(Rs[7], b.x), fmul(Rs[8], b.y));
  const float ptx = fadd(sum3(fmul(Rt[0], b.z), fmul(Rt[1], b.w), fmul(Rt[2], c.x)), tt[0]);
  const float pty = fadd(sum3(fmul(Rt[3], b.z), fmul(Rt[4], b.w), fmul(Rt[5], c.x)), tt[1]);
  const float ptz = fadd(sum3(fmul(Rt[6], b.z), fmul(Rt[7], b.w), fmul(Rt[8], c.x)), tt[2]);
  const float ntx = sum3(fmul(Rt[0], c.y), fmul(Rt[1], c.z), fmul(Rt[2], c.w));
  const float nty = sum3(fmul(Rt[3], c.y), fmul(Rt[4], c.z), fmul(Rt[5], c.w));
  const float ntz = sum3(fmul(Rt[6], c.y), fmul(Rt[7], c.z), fmul(Rt[8], c.w));

  const float r1 = dot3(nsx, nsy, nsz, fsub(ptx, psx), fsub(pty, psy), fsub(ptz, psz));
  const float r2 = dot3(ntx, nty, ntz, fsub(psx, ptx), fsub(psy, pty), fsub(psz, ptz));
  float j1[6], j2[6];
  j1[0] = nsx; j1[1] = nsy; j1[2] = nsz;
  j1[3] = fadd(fmul(-nsy, ptz), fmul(nsz, pty));
  j1[4] = fsub(fmul(nsx, ptz), fmul(nsz, ptx));
  j1[5] = fadd(fmul(-nsx, pty), fmul(nsy, ptx));
  j2[0] = ntx; j2[1] = nty; j2[2] = ntz;
  j2[3] = fadd(fmul(-nty, psz), fmul(ntz, psy));
  j2[4] = fsub(fmul(ntx, psz), fmul(ntz, psx));
  j2[5] = fadd(fmul(-ntx, psy), fmul(nty, psx));

  double d1[6], d2[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) { d1[i] = (double)j1[i]; d2[i] = (double)j2[i]; }
  const double dr1 = (double)r1, dr2 = -(double)r2;
  int k = 0;
#pragma unroll
  for (int cidx = 0; cidx < 6; ++cidx)
#pragma unroll
    for (int r = 0; r <= cidx; ++r) { acc[k] = fma(d1[r], d1[cidx], acc[k]); acc[k] = fma(d2[r], d2[cidx], acc[k]); ++k; }
#pragma unroll
  for (int i = 0; i < 6; ++i) { acc[21 + i] = fma(dr1, d1[i], acc[21 + i]); acc[21 + i] = fma(dr2, d2[i], acc[21 + i]); }
  acc[27] += (double)fmul(r1, r1);
  acc[27] += (double)fmul(r2, r2);
}

__device__ __forceinline__ void cost_record(const float4 a, const float4 b, const float4 c, const float Rs[9], const float ts[3],
                                            const float Rt[9], const float tt[3], double* cost) {
  const float psx = fadd(sum3(fmul(Rs[0], a.x), fmul(Rs[1], a.y), fmul(Rs[2], a.z)), ts[0]);
  const float psy = fadd(sum3(fmul(Rs[3], a.x), fmul(Rs[4], a.y), fmul(Rs[5], a.z)), ts[1]);
  const float psz = fadd(sum3(fmul(Rs[6], a.x), fmul(Rs[7], a.y), fmul(Rs[8], a.z)), ts[2]);
  const float nsx = sum3(fmul(Rs[0], a.w), fmul(Rs[1], b.x), fmul(Rs[2], b.y));
  const float nsy = sum3(fmul(Rs[3], a.w), fmul(Rs[4], b.x), fmul(Rs[5], b.y));
  const float nsz = sum3(fmul(Rs[6], a.w), fmul(Rs[7], b.x), fmul(Rs[8], b.y));
  const float ptx = fadd(sum3(fmul(Rt[0], b.z), fmul(Rt[1], b.w), fmul(Rt[2], c.x)), tt[0]);
  const float pty = fadd(sum3(fmul(Rt[3], b.z), fmul(Rt[4], b.w), fmul(Rt[5], c.x)), tt[1]);
  const float ptz = fadd(sum3(fmul(Rt[6], b.z), fmul(Rt[7], b.w), fmul(Rt[8], c.x)), tt[2]);
  const float ntx = sum3(fmul(Rt[0], c.y), fmul(Rt[1], c.z), fmul(Rt[2], c.w));
  const float nty = sum3(fmul(Rt[3], c.y), fmul(Rt[4], c.z), fmul(Rt[5], c.w));
  const float ntz = sum3(fmul(Rt[6], c.y), fmul(Rt[7], c.z), fmul(Rt[8], c.w));
  const float r1 = dot3(nsx, nsy, nsz, fsub(ptx, psx), fsub(pty, psy), fsub(ptz, psz));
  const float r2 = dot3(ntx, nty, ntz, fsub(psx, ptx), fsub(psy, pty), fsub(psz, ptz));
  *cost += (double)fmul(r1, r1);
  *cost += (double)fmul(r2, r2);
}

// Block-wide fixed-tree reduction of NV doubles per thread; result valid in threads [0,NV) of the block.
template <int NV>
__device__ __forceinline__ void block_reduce(double acc[NV], double (*smem)[NV], double* out_first_nv_threads) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[k] = v;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();   // smem may still be read from a previous flush
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[w][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double v = smem[0][threadIdx.x];
    for (int i = 1; i < kAccThreads / 32; ++i) v += smem[i][threadIdx.x];
    *out_first_nv_threads = v;
  }
}

template <bool WITH_H>
__global__ void __launch_bounds__(kAccThreads, 2)
k_accumulate(const float4* __restrict__ ra, const float4* __restrict__ rb, const float4* __restrict__ rc,
             const Segment* __restrict__ segs, int nseg, const CloudPose* __restrict__ poses, unsigned long long total,
             unsigned long long per_cta, double* __restrict__ partials /* [nseg][gridDim.x][kAccVals] */) {
  constexpr int NV = WITH_H ? kAccVals : 1;
  __shared__ double smem[kAccThreads / 32][NV];
  unsigned long long r0 = (unsigned long long)blockIdx.x * per_cta;
  const unsigned long long r1 = min(total, r0 + per_cta);
  if (r0 >= r1) return;
  // first segment whose end is beyond r0 (segments are sorted, non-overlapping, possibly empty)
  int lo = 0, hi = nseg - 1;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (segs[mid].end > r0) hi = mid; else lo = mid + 1; }
  int seg = lo;
  while (r0 < r1) {
    const Segment sg = segs[seg];
    const unsigned long long e = min(r1, sg.end);
    if (e <= r0) { ++seg; continue; }
    float Rs[9], ts[3], Rt[9], tt[3];
    load_pose(poses, sg.src, Rs, ts);
    load_pose(poses, sg.tgt, Rt, tt);
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    unsigned long long r = r0 + threadIdx.x;
    // two records in flight per thread: 6 independent LDG.128 before the math
    for (; r + kAccThreads < e; r += 2 * kAccThreads) {
      const float4 a0 = __ldcs(ra + r), b0 = __ldcs(rb + r), c0 = __ldcs(rc + r);
      const float4 a1 = __ldcs(ra + r + kAccThreads), b1 = __ldcs(rb + r + kAccThreads), c1 = __ldcs(rc + r + kAccThreads);
      if (WITH_H) { accumulate_record(a0, b0, c0, Rs, ts, Rt, tt, acc); accumulate_record(a1, b1, c1, Rs, ts, Rt, tt, acc); }
      else { cost_record(a0, b0, c0, Rs, ts, Rt, tt, &acc[0]); cost_record(a1, b1, c1, Rs, ts, Rt, tt, &acc[0]); }
    }
    if (r < e) {
      const float4 a0 = __ldcs(ra + r), b0 = __ldcs(rb + r), c0 = __ldcs(rc + r);
      if (WITH_H) accumulate_record(a0, b0, c0, Rs, ts, Rt, tt, acc);
      else cost_record(a0, b0, c0, Rs, ts, Rt, tt, &acc[0]);
    }
    double out;
    block_reduce<NV>(acc, smem, &out);
    if (threadIdx.x < NV) {
      const int slot = WITH_H ? threadIdx.x : (kAccVals - 1);
      partials[((size_t)seg * gridDim.x + blockIdx.x) * kAccVals + slot] = out;
    }
    r0 = e;
    ++seg;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K5 (Blackwell path): the same pass with the record stream staged through shared memory by the bulk-copy engine.
// One producer lane issues cp.async.bulk (TMA 1-D) copies of 256-record tiles (3 planes x 4 KB) into a ring of stages guarded by
// full/empty mbarriers; 8 consumer warps read their record with three conflict-free LDS.128 and run the identical arithmetic.
// The bytes in flight per SM (stages x 12 KB x 2 CTAs) are decoupled from the consumers' register budget, which is what the
// register-staged variant above runs out of (28 fp64 accumulators per thread). Same record partition, same reduction tree:
// results are bit-identical to k_accumulate.
// ------------------------------------------------------------------------------------------------------------------
// Speculative LM trials (NX > 0): the poses of NX further trial states (the next damping factors of the LM loop,
// icp_point_to_plane_impl.h:217-285: lambda doubles after every rejected try) ride along with the pass. Their costs are evaluated on
// the records while these are in registers, with the arithmetic, record partition and reduction tree of the trial-0 cost, so every
// value is bit-identical to what a pass of its own would return and the accept / reject sequence is unchanged; a rejected chain of
// 10 tries then costs 3-4 passes over the 16 GB of records instead of 10. Extra poses are staged in shared memory per segment.
static constexpr int kMaxExtraTrials = 3;
static constexpr int kTmaStages = 4;
static constexpr int kTmaTile = 2 * kAccThreads;                   // records per tile: two per consumer thread
static constexpr size_t kTmaSmemBytes = (size_t)kTmaStages * 3 * kTmaTile * 16;

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
  unsigned int ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned int bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Walks the tile sequence of one CTA: tiles never straddle a segment ("correspondence set") boundary.
struct TileCursor {
  unsigned long long r, seg_end, r_end; int seg;
  __device__ __forceinline__ bool valid() const { return r < r_end; }
  __device__ __forceinline__ void settle(const Segment* __restrict__ segs) {      // move to the segment that contains r
    while (r < r_end) { seg_end = min(r_end, segs[seg].end); if (seg_end > r) break; ++seg; }
  }
  __device__ __forceinline__ unsigned int count() const { return (unsigned int)min((unsigned long long)kTmaTile, seg_end - r); }
  __device__ __forceinline__ void advance(const Segment* __restrict__ segs) { r += count(); if (r >= seg_end) { ++seg; settle(segs); } }
};

template <bool WITH_H, int NX>
__global__ void __launch_bounds__(kAccThreads, 2)
k_accumulate_tma(const float4* __restrict__ ra, const float4* __restrict__ rb, const float4* __restrict__ rc,
                 const Segment* __restrict__ segs, int nseg, const CloudPose* __restrict__ poses /* [1 + NX][nclouds] */, int nclouds,
                 unsigned long long total, unsigned long long per_cta, double* __restrict__ partials /* [nseg][gridDim.x][kAccVals] */,
                 double* __restrict__ xpartials /* [nseg][gridDim.x][kMaxExtraTrials] */) {
  constexpr int NV = WITH_H ? kAccVals : 1;
  constexpr int NXS = NX > 0 ? NX : 1;
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ double red[kAccThreads / 32][NV];
  __shared__ double xred[kAccThreads / 32][NXS];
  __shared__ __align__(16) float xpose[NXS][24];                 // per extra trial: source pose (R, t), target pose (R, t)
  __shared__ __align__(8) unsigned long long full_bar[kTmaStages], empty_bar[kTmaStages];
  const unsigned long long r_begin = (unsigned long long)blockIdx.x * per_cta;
  const unsigned long long r_end = min(total, r_begin + per_cta);
  if (r_begin >= r_end) return;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTmaStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kAccThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int lo = 0, hi = nseg - 1;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (segs[mid].end > r_begin) hi = mid; else lo = mid + 1; }
  float4* const sa = reinterpret_cast<float4*>(tile_smem);       // stage s: planes at sa + (3*s + p) * kTmaTile

  // Thread 0 doubles as the producer: it keeps kTmaStages-1 tiles in flight ahead of the tile being consumed.
  TileCursor prod{r_begin, 0, r_end, lo}; unsigned int pit = 0;
  auto produce = [&]() {
    const unsigned int s = pit % kTmaStages, cnt = prod.count();
    if (pit >= (unsigned int)kTmaStages) mbar_wait(&empty_bar[s], ((pit / kTmaStages) - 1) & 1);
    mbar_expect_tx(&full_bar[s], 3u * cnt * 16u);
    bulk_g2s(sa + (3 * s + 0) * kTmaTile, ra + prod.r, cnt * 16u, &full_bar[s]);
    bulk_g2s(sa + (3 * s + 1) * kTmaTile, rb + prod.r, cnt * 16u, &full_bar[s]);
    bulk_g2s(sa + (3 * s + 2) * kTmaTile, rc + prod.r, cnt * 16u, &full_bar[s]);
    prod.advance(segs); ++pit;
  };
  if (threadIdx.x == 0) {
    prod.settle(segs);
    for (int k = 0; k < kTmaStages - 1 && prod.valid(); ++k) produce();
  }

  unsigned long long r0 = r_begin; int seg = lo; unsigned int it = 0;
  while (r0 < r_end) {
    const Segment sg = segs[seg];
    const unsigned long long e = min(r_end, sg.end);
    if (e <= r0) { ++seg; continue; }
    float Rs[9], ts[3], Rt[9], tt[3];
    load_pose(poses, sg.src, Rs, ts);
    load_pose(poses, sg.tgt, Rt, tt);
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    double xacc[NXS];
#pragma unroll
    for (int j = 0; j < NXS; ++j) xacc[j] = 0.0;
    if (NX > 0) {   // (the reductions at the end of the previous segment separate its readers from this write)
      if ((int)threadIdx.x < NX * 24) {
        const int j = threadIdx.x / 24, w = threadIdx.x % 24;
        const CloudPose* P = poses + (size_t)(1 + j) * nclouds + (w < 12 ? sg.src : sg.tgt);
        const int q = w % 12;
        xpose[j][w] = q < 9 ? __ldg(&P->R[q]) : __ldg(&P->t[q - 9]);
      }
      __syncthreads();
    }
    for (unsigned long long r = r0; r < e; r += kTmaTile, ++it) {
      if (threadIdx.x == 0 && prod.valid()) produce();            // refill the stage released one tile ago
      const unsigned int s = it % kTmaStages, cnt = (unsigned int)min((unsigned long long)kTmaTile, e - r);
      mbar_wait(&full_bar[s], (it / kTmaStages) & 1);
      float4 a0, b0, c0, a1, b1, c1;
      const bool m0 = threadIdx.x < cnt, m1 = threadIdx.x + kAccThreads < cnt;
      const float4* st = sa + (3 * s) * kTmaTile;
      if (m0) { a0 = st[threadIdx.x]; b0 = st[kTmaTile + threadIdx.x]; c0 = st[2 * kTmaTile + threadIdx.x]; }
      if (m1) { a1 = st[kAccThreads + threadIdx.x]; b1 = st[kTmaTile + kAccThreads + threadIdx.x]; c1 = st[2 * kTmaTile + kAccThreads + threadIdx.x]; }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[s]);     // this warp's values are in registers: one arrival per warp
      if (m0) { if (WITH_H) accumulate_record(a0, b0, c0, Rs, ts, Rt, tt, acc); else cost_record(a0, b0, c0, Rs, ts, Rt, tt, &acc[0]); }
      if (m1) { if (WITH_H) accumulate_record(a1, b1, c1, Rs, ts, Rt, tt, acc); else cost_record(a1, b1, c1, Rs, ts, Rt, tt, &acc[0]); }
      if (NX > 0) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
          float P[24];
          const float4* xp = reinterpret_cast<const float4*>(xpose[j]);
#pragma unroll
          for (int q = 0; q < 6; ++q) { const float4 v = xp[q]; P[4 * q] = v.x; P[4 * q + 1] = v.y; P[4 * q + 2] = v.z; P[4 * q + 3] = v.w; }
          if (m0) cost_record(a0, b0, c0, P, P + 9, P + 12, P + 21, &xacc[j]);
          if (m1) cost_record(a1, b1, c1, P, P + 9, P + 12, P + 21, &xacc[j]);
        }
      }
    }
    double out;
    block_reduce<NV>(acc, red, &out);
    if (threadIdx.x < NV) {
      const int slot = WITH_H ? threadIdx.x : (kAccVals - 1);
      partials[((size_t)seg * gridDim.x + blockIdx.x) * kAccVals + slot] = out;
    }
    if (NX > 0) {
      double xout;
      block_reduce<NXS>(xacc, xred, &xout);
      if ((int)threadIdx.x < NX) xpartials[((size_t)seg * gridDim.x + blockIdx.x) * kMaxExtraTrials + threadIdx.x] = xout;
    }
    r0 = e;
    ++seg;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K6: per-segment sums in ascending CTA order, then the normal equations [H (nv*nv, col-major, symmetric) | b | cost].
// Assembly follows Accumulate (impl.h:82-113) with w = 1 and the solver's Upper view (impl.h:226):
//   H(src,src) += S, H(tgt,tgt) += S, H(src,tgt) += -S only when that block lies in the upper triangle (src var < tgt var),
//   b(src) -= g, b(tgt) += g; impl cloud 0 has no variables.
// Single block; every sum runs in a fixed order.
// ------------------------------------------------------------------------------------------------------------------
template <bool WITH_H>
__global__ void __launch_bounds__(1024) k_finalize(const double* __restrict__ partials, const Segment* __restrict__ segs, int nseg,
                                                   int grid_acc, unsigned long long per_cta, int nv, double* __restrict__ segsum,
                                                   double* __restrict__ eq, double extra0, double extra1,
                                                   const double* __restrict__ xpartials, int nx, double* __restrict__ xsegsum) {
  // costs of the speculative trials: the same two-stage order as the trial-0 cost (CTAs ascending, then segments ascending)
  for (int w = threadIdx.x; w < nseg * nx; w += blockDim.x) {
    const int s = w / nx, j = w % nx;
    double v = 0.0;
    if (segs[s].end > segs[s].begin) {
      const int b0 = (int)(segs[s].begin / per_cta), b1 = (int)((segs[s].end - 1) / per_cta);
      for (int b = b0; b <= b1; ++b) v += xpartials[((size_t)s * grid_acc + b) * kMaxExtraTrials + j];
    }
    xsegsum[w] = v;
  }
  for (int w = threadIdx.x; w < nseg * kAccVals; w += blockDim.x) {
    const int s = w / kAccVals, k = w % kAccVals;
    double v = 0.0;
    if ((WITH_H || k == kAccVals - 1) && segs[s].end > segs[s].begin) {
      const int b0 = (int)(segs[s].begin / per_cta), b1 = (int)((segs[s].end - 1) / per_cta);
      for (int b = b0; b <= b1; ++b) v += partials[((size_t)s * grid_acc + b) * kAccVals + k];
    }
    segsum[w] = v;
  }
  __syncthreads();
  const int nh = nv * nv;
  if (WITH_H) {
    for (int w = threadIdx.x; w < nh; w += blockDim.x) {
      const int r = w % nv, c = w / nv;          // column-major
      const int rr = r <= c ? r : c, cc = r <= c ? c : r;   // mirror the upper triangle
      const int vr = rr / 6, vc = cc / 6, ir = rr % 6, ic = cc % 6;
      const int lo6 = ir <= ic ? ir : ic, hi6 = ir <= ic ? ic : ir;
      const int sidx = hi6 * (hi6 + 1) / 2 + lo6;           // packed upper index of S
      double v = 0.0;
      for (int s = 0; s < nseg; ++s) {
        const int sv = segs[s].src - 1, tv = segs[s].tgt - 1;   // variable block index, -1 = fixed
        const double S = segsum[s * kAccVals + sidx];
        if (vr == vc) { if (sv == vr) v += S; if (tv == vr) v += S; }
        else if (sv == vr && tv == vc) v -= S;
      }
      eq[w] = v;
    }
    for (int w = threadIdx.x; w < nv; w += blockDim.x) {
      const int vb = w / 6, i = w % 6;
      double v = 0.0;
      for (int s = 0; s < nseg; ++s) {
        const double g = segsum[s * kAccVals + 21 + i];
        if (segs[s].tgt - 1 == vb) v += g;
        if (segs[s].src - 1 == vb) v -= g;
      }
      eq[nh + w] = v;
    }
  }
  if (threadIdx.x == 0) {
    double c = 0.0;
    for (int s = 0; s < nseg; ++s) c += segsum[s * kAccVals + kAccVals - 1];
    eq[nh + nv] = c;
    eq[nh + nv + 1] = extra0;
    eq[nh + nv + 2] = extra1;
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kMaxExtraTrials) {
    const int j = threadIdx.x - 32;
    double c = 0.0;
    if (j < nx) for (int s = 0; s < nseg; ++s) c += xsegsum[s * nx + j];
    eq[nh + nv + 3 + j] = c;
  }
}

}  // namespace b2
