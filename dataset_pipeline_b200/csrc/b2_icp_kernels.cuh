// Device kernels of Path A (multi-scan point-to-plane ICP) for sm_100a.
//
// Static index, once per cloud and search radius, in the CLOUD's own frame (a rigid transform maps a uniform grid to a uniform
// grid, so the index survives every pose update; replaces the per-pair kd-tree build, icp_point_to_plane.cc:46-51):
//   K1  k_bbox / k_keys / k_gather_sorted   cell keys of the cloud-frame points, cell-sorted copies, inverse permutation
//   K2  k_mark_cells / k_word_counts / k_word_prefix / k_cell_starts / k_corner_occupancy   (grids up to 2^31 cells) rank bitmap,
//       first point of every occupied cell, one occupancy byte per lattice corner;  k_hash_cells: occupied-cell hash beyond
// Per outer iteration:
//   K1x k_xform_sorted     global-frame copies of the sorted points + chunk boxes + AABB in ONE streaming pass
//                          (replaces pcl::transformPointCloudWithNormals + bbox loops, icp_point_to_plane.cc:189-205)
//   K3  k_nn_tiles         nearest target within radius per source point (replaces the radiusSearch loop, :63-102): persistent CTAs
//                          over the longest-first tile order, warp-autonomous tiles of 32 queries staged (and prefetched) in shared
//                          memory by the bulk-copy engine, own cell per lane, occupied neighbour cells (corner map) through a per-warp
//                          work queue, dense cells by the whole warp
//   K4  k_scan_tiles / k_pack_tiles   48 B packed correspondence records, three float4 planes with source / target values interleaved
//   K5  k_accumulate_tma   one streaming pass over bulk-copied record tiles: cost + 6x6 S + 6-vector g per correspondence set in fp64,
//                          plus the costs of up to three further LM tries in packed fp32 (FFMA2); k_accumulate = register-staged variant
//                          (replaces compute() loops, icp_point_to_plane_impl.h:129-211 and :240-266)
//   K6  k_finalize         fixed-order reduction of the per-CTA partials + assembly of the normal equations
//                          with the reference's upper-triangle quirk (impl.h:82-113 + :226)
// All of these are HBM / gather bound integer+fp32+fp64 streaming work: no tensor cores.
#pragma once
#include <type_traits>

#include "b2_common.cuh"

namespace b2 {

static constexpr int kAccThreads = 256;
static constexpr int kAccVals = 28;   // 21 (upper S) + 6 (g) + 1 (cost)

// ---- mbarrier / bulk-copy (TMA 1-D) helpers, used by K3 (query tiles) and K5 (record tiles) ------------------------
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

// K1b: cell key of every point (+ identity permutation). The grid lives in the cloud's INDEX frame x' = F l (F = the pose the cloud
// had when it was indexed, frozen from then on: all clouds of a handle then share one lattice, so the cell-sorted queries of one
// cloud walk the cells of another in order). Positions and cell coordinates in double, so the key of a point is a function of
// its fp32 coordinates, F and the grid alone.
struct IndexFrame { double f[12]; };   // row-major 3x4
__global__ void __launch_bounds__(256) k_keys(const float* __restrict__ xyz, size_t n, IndexFrame F, GridParams g,
                                              unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  const double fx = (F.f[0] * x + F.f[1] * y + F.f[2] * z + F.f[3] - g.ox) * g.inv;
  const double fy = (F.f[4] * x + F.f[5] * y + F.f[6] * z + F.f[7] - g.oy) * g.inv;
  const double fz = (F.f[8] * x + F.f[9] * y + F.f[10] * z + F.f[11] - g.oz) * g.inv;
  const int cx = (int)floor(fx), cy = (int)floor(fy), cz = (int)floor(fz);
  keys[i] = (cell_key(g, cx, cy, cz) << (3 * g.fbits)) | fine_code(fx - cx, fy - cy, fz - cz, g.fbits);
  idx[i] = (unsigned int)i;
}

// K1c: cell-sorted cloud-frame copies: l_xyz[j] = (p, bits(original index)), l_nrm[j] = (n, 0), and the inverse permutation
// (sorted position of every original index; K4 looks the matched target up through it).
__global__ void __launch_bounds__(256) k_gather_sorted(const float* __restrict__ xyz, const float* __restrict__ nrm, size_t n,
                                                       const unsigned int* __restrict__ perm, float4* __restrict__ l_xyz,
                                                       float4* __restrict__ l_nrm, unsigned int* __restrict__ perm_inv) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned int i = perm[j];
  l_xyz[j] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __uint_as_float(i));
  l_nrm[j] = make_float4(nrm[3 * (size_t)i], nrm[3 * (size_t)i + 1], nrm[3 * (size_t)i + 2], 0.f);
  perm_inv[i] = (unsigned int)j;
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
// K1x (every outer iteration): s_xyz[j] = (T * l_xyz[j], index), s_nrm[j] = R * l_nrm[j] in the reference's fp32 operation order,
// the level-1 chunk box of every 32 consecutive sorted points (one warp = one chunk) and one AABB partial per block — a single
// 64 B/point stream; nothing is sorted or hashed after the first iteration. Persistent grid (a multiple of the SM count), every
// warp walks whole chunks.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_xform_sorted(const float4* __restrict__ l_xyz, const float4* __restrict__ l_nrm, size_t n, Mat4 T,
                                                      float4* __restrict__ s_xyz, float4* __restrict__ s_nrm, Aabb* __restrict__ box1,
                                                      float* __restrict__ partial) {
  float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
  const int lane = threadIdx.x & 31;
  const size_t nround = (n + 31) & ~(size_t)31;      // whole warps stay in the loop (the shuffles need all 32 lanes)
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += (size_t)gridDim.x * blockDim.x) {
    float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
    if (j < n) {
      const float4 l = __ldcs(l_xyz + j), ln = __ldcs(l_nrm + j);
      const float3 p = xform_point(T, l.x, l.y, l.z);
      const float3 q = xform_normal(T, ln.x, ln.y, ln.z);
      s_xyz[j] = make_float4(p.x, p.y, p.z, l.w);
      s_nrm[j] = make_float4(q.x, q.y, q.z, 0.f);
      lx = hx = p.x; ly = hy = p.y; lz = hz = p.z;
      mnx = fminf(mnx, p.x); mny = fminf(mny, p.y); mnz = fminf(mnz, p.z);
      mxx = fmaxf(mxx, p.x); mxy = fmaxf(mxy, p.y); mxz = fmaxf(mxz, p.z);
    }
    for (int o = 16; o > 0; o >>= 1) {
      lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
      lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
      hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
    }
    if (lane == 0) { Aabb b; b.lo[0] = lx; b.lo[1] = ly; b.lo[2] = lz; b.hi[0] = hx; b.hi[1] = hy; b.hi[2] = hz; box1[j >> 5] = b; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
  }
  __shared__ float s[8][6];
  const int w = threadIdx.x >> 5;
  if (lane == 0) { s[w][0] = mnx; s[w][1] = mny; s[w][2] = mnz; s[w][3] = mxx; s[w][4] = mxy; s[w][5] = mxz; }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = s[0][threadIdx.x];
    for (int i = 1; i < 8; ++i) v = threadIdx.x < 3 ? fminf(v, s[i][threadIdx.x]) : fmaxf(v, s[i][threadIdx.x]);
    partial[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

// Level-2 chunk boxes: one per 32 level-1 boxes (1024 points).
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

// ------------------------------------------------------------------------------------------------------------------
// K2: occupied-cell hash table over the sorted keys. Entry = {key, begin, end} (16 B, one LDG.128 per probe).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_cells(const unsigned long long* __restrict__ keys, size_t n, unsigned int* __restrict__ count,
                                                     int kFineBits) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool head = j < n && (j == 0 || (keys[j] >> kFineBits) != (keys[j - 1] >> kFineBits));
  const int c = __syncthreads_count(head);
  if (threadIdx.x == 0 && c) atomicAdd(count, (unsigned int)c);
}

// One read of the sorted keys: the first point of a cell stores `begin`, the last one stores `end`; whichever of the two arrives
// first claims the slot (both run the same claim-or-find probe), so no second pass is needed.
__device__ __forceinline__ unsigned int hash_claim(HashEntry* __restrict__ table, int log2size, unsigned long long key) {
  const unsigned int mask = (1u << log2size) - 1u;
  unsigned int s = hash_slot(key, log2size);
  while (true) {
    const unsigned long long prev = atomicCAS(&table[s].key, kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) return s;
    s = (s + 1) & mask;
  }
}
__global__ void __launch_bounds__(256) k_hash_cells(const unsigned long long* __restrict__ keys, size_t n, HashEntry* __restrict__ table,
                                                    int log2size, int kFineBits) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long key = keys[j] >> kFineBits;
  const bool head = j == 0 || (keys[j - 1] >> kFineBits) != key;
  const bool tail = j + 1 == n || (keys[j + 1] >> kFineBits) != key;
  if (!head && !tail) return;
  const unsigned int s = hash_claim(table, log2size, key);
  if (head) table[s].begin = (unsigned int)j;
  if (tail) table[s].end = (unsigned int)(j + 1);
}

__device__ __forceinline__ bool hash_find(const HashEntry* __restrict__ table, int log2size, unsigned long long key,
                                          unsigned int* begin, unsigned int* end) {
  const unsigned int mask = (1u << log2size) - 1u;
  unsigned int s = hash_slot(key, log2size);
  while (true) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(table + s));
    const unsigned long long k = ((unsigned long long)e.y << 32) | e.x;
    if (k == key) { *begin = e.z; *end = e.w; return true; }
    if (k == kEmptyKey) return false;      // linear probing, load factor <= 0.5: rare second probe
    s = (s + 1) & mask;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K3: nearest target within radius (strict d2 < r2), lowest ORIGINAL target index on exact ties.
// d2 = ((dx*dx)+(dy*dy))+(dz*dz) in fp32 on the GLOBAL-frame coordinates, without contraction (FLANN L2_Simple order) — the
// arithmetic of the reference. Only the candidate LOOKUP runs in the target cloud's frame: the query is mapped there by the
// inverse pose (fp32 FMA), its cell and its half of the cell are read off, and the 2x2x2 block of cells on that side is
// searched. The grid cell is >= 2 (d sigma + margin), where sigma bounds the stretch of the inverse pose and margin every fp32
// rounding on the way (host: grid_for_cloud / search_grid), so every target whose fp32 d2 is below r2 lies in that block.
//
// Work decomposition (one CTA = one tile of kTile consecutive cell-sorted queries):
//   stage  the tile's query rows are copied into shared memory by the bulk-copy engine (cp.async.bulk + mbarrier);
//   A      one thread per query: cell lookup, scan of the query's OWN target cell (neighbouring threads share cells, so
//          candidate rows are broadcast loads), then the neighbour cells whose nearest face is not already farther than the
//          best match are appended to a work queue in shared memory — ~0.4 items per query instead of up to 7 mostly idle
//          per-thread rounds;
//   B      full warps drain the queue, one (query, neighbour cell) item per lane; results meet in a shared 64-bit key per query
//          (atomicMin);
//   C      key out (8 B/query) + matched count of the tile (for K4's offsets).
// (d2, index) is ONE 64-bit key, (d2 bits << 32) | original index: for non-negative floats the integer order is the float order,
// so min() over keys implements "d2 < best, or d2 == best and lower index" independently of the visiting order; the initial key
// (r2 bits << 32) | 0 rejects d2 == r2 for every index (the radius test is strict) and doubles as the "no match" value.
// ------------------------------------------------------------------------------------------------------------------
// (the B2_K3_* macros exist for A/B builds: tools/k3_variants.sh)
#ifndef B2_K3_TILE
#define B2_K3_TILE 256
#endif
#ifndef B2_K3_INLINE
#define B2_K3_INLINE 64
#endif
#ifndef B2_K3_COOP
#define B2_K3_COOP 256
#endif
#ifndef B2_K3_MINB
#define B2_K3_MINB 7
#endif
static constexpr int kTile = B2_K3_TILE;   // queries per CTA
static constexpr unsigned int kInlineCell = B2_K3_INLINE;   // cells up to this many points are scanned without chunk boxes
static constexpr unsigned int kCoopCell = B2_K3_COOP;       // cells above this many points are scanned by a whole warp (scan_cell_warp)
static constexpr unsigned int kDenseCap = 48u;     // dense-cell queue entries per warp (overflow: the thread scans the cell itself)

struct SearchGrid {              // a target cloud's static grid + this iteration's global -> cloud-frame map
  float m[12];                   // row-major 3x4: position relative to the grid origin = m[4r..4r+2] . q + m[4r+3]
  float inv, cell;
  float inv_sigma, margin;       // global distance >= cloud-frame distance * inv_sigma - margin
  int nx, ny, nz;
  long long sy, sz;              // cell key = cz*sz + cy*sy + cx
  int log2size;
  float one;                     // 1.0f at run time (see B2_NN_KEY)
  const unsigned int* occ;       // sparse layout: unused (nullptr: every neighbour is probed in the hash table)
  const unsigned char* corner;   // DUAL: per lattice corner, the occupancy of its 8 adjacent cells (bit dx + 2 dy + 4 dz, d = 1: the cell on the upper side)
  const uint2* rb;               // DENSE layout: rank bitmap, {occupancy bits, occupied cells before this word} per 32 cells
  const unsigned int* starts;    // DENSE layout: first sorted point of every occupied cell, + the total
};
__device__ __forceinline__ bool cell_occupied(const SearchGrid& g, long long key) {
  return !g.occ || ((__ldg(g.occ + (key >> 5)) >> (unsigned int)(key & 31ll)) & 1u);
}

struct SearchWork { unsigned int points, box1, box2, cells; };   // work counters of the diagnostic K3 variant (B2_K3_WORK)

// The five additions of a candidate test are issued as FMAs with a unit factor, fma(b, -1, a) = fl(a - b) and fma(a, 1, b) =
// fl(a + b) bit for bit: FADD shares the ALU pipe with the key compare / select (ISETP, SEL), which was the binding pipe of this
// kernel (ncu r02b: ALU 51 %, FMA 19 % of peak); as FFMA they run beside them. `one` is a run-time 1.0f so that the multiply
// survives ptxas (a literal 1.0 is folded back into FADD).
#define B2_NN_KEY(T)                                                                                                  \
  {                                                                                                                   \
    const float ax_ = __fmaf_rn((T).x, -one, q.x), ay_ = __fmaf_rn((T).y, -one, q.y), az_ = __fmaf_rn((T).z, -one, q.z);  \
    const float d_ = __fmaf_rn(__fmaf_rn(fmul(ax_, ax_), one, fmul(ay_, ay_)), one, fmul(az_, az_));                  \
    B2_NN_MIN(d_, (T).w)                                                                                              \
  }
// The 64-bit key minimum. Default: unsigned compare + selects (ISETP, ISETP.EX, 2 SEL on the ALU pipe). B2_K3_DMIN: the same minimum
// taken as an fp64 one (DSETP.MIN + 2 selects): for a finite non-negative d2 the key's bit pattern is a finite non-negative double
// whose order is the order of the bit patterns, and a NaN d2 loses either way.
#ifdef B2_K3_DMIN
#define B2_NN_MIN(D, W)                                                                                               \
  best = (unsigned long long)__double_as_longlong(fmin(__longlong_as_double((long long)best), __hiloint2double(__float_as_int(D), __float_as_int(W))));
#else
#define B2_NN_MIN(D, W)                                                                                               \
  {                                                                                                                   \
    const unsigned long long k_ = ((unsigned long long)__float_as_uint(D) << 32) | (unsigned long long)__float_as_uint(W); \
    best = k_ < best ? k_ : best;                                                                                     \
  }
#endif

__device__ __forceinline__ float key_d2(unsigned long long key) { return __uint_as_float((unsigned int)(key >> 32)); }

// Candidates [b,e), four loads in flight: whole groups of four first, then ONE clamped group for the remainder — it re-reads
// candidate e-1 (testing a candidate twice changes nothing: the key minimum is idempotent), so there is no per-candidate tail loop
// and the lanes of a warp run the same loop shape.
__device__ __forceinline__ void scan_range(const float4* __restrict__ tgt, unsigned int b, unsigned int e, const float4& q, float one,
                                           unsigned long long& best, SearchWork& wk) {
  wk.points += e - b;
  unsigned int p = b;
  for (; p + 4u <= e; p += 4u) {
    const float4 t0 = __ldg(tgt + p), t1 = __ldg(tgt + p + 1u), t2 = __ldg(tgt + p + 2u), t3 = __ldg(tgt + p + 3u);
    B2_NN_KEY(t0) B2_NN_KEY(t1) B2_NN_KEY(t2) B2_NN_KEY(t3)
  }
  if (p < e) {
    const unsigned int last = e - 1u;
    const float4 t0 = __ldg(tgt + p), t1 = __ldg(tgt + min(p + 1u, last)), t2 = __ldg(tgt + min(p + 2u, last));
    B2_NN_KEY(t0) B2_NN_KEY(t1) B2_NN_KEY(t2)
  }
}

// One cell's candidates [b,e). Small cells are scanned directly; dense cells (near the target's scanner a 2 cm cell holds
// 10^2..10^3 points, at its zenith 10^4..10^5) go through the chunk boxes — 32 / 1024 consecutive points of the in-cell Morton
// order; the box bound is evaluated with the same fp32 operations as the point distance, so pruning is exact. Cells above 2048
// points first take a strided sample of 64 candidates so that `best` is tight before the boxes are tested.
__device__ __forceinline__ void scan_cell(const float4* __restrict__ tgt, const Aabb* __restrict__ box1, const Aabb* __restrict__ box2,
                                          unsigned int b, unsigned int e, const float4& q, float one, unsigned long long& best,
                                          SearchWork& wk) {
  ++wk.cells;
  if (e - b <= kInlineCell) { scan_range(tgt, b, e, q, one, best, wk); return; }
  const unsigned int last = e - 1u;
  const bool huge = e - b > 2u * kChunk2;
  if (huge) {
    const unsigned int stride = (e - b) / 64u;
    wk.points += 64u;
    for (unsigned int p = b; p < b + 64u * stride; p += 4u * stride) {
      const float4 t0 = __ldg(tgt + p), t1 = __ldg(tgt + p + stride), t2 = __ldg(tgt + p + 2u * stride), t3 = __ldg(tgt + p + 3u * stride);
      B2_NN_KEY(t0) B2_NN_KEY(t1) B2_NN_KEY(t2) B2_NN_KEY(t3)
    }
  }
  for (unsigned int c2 = b / kChunk2; c2 <= last / kChunk2; ++c2) {
    if (huge) { ++wk.box2; if (dist2_box(q.x, q.y, q.z, box2[c2]) > key_d2(best)) continue; }
    const unsigned int c1b = max(b / kChunk1, c2 * 32u), c1e = min(last / kChunk1, c2 * 32u + 31u);
    for (unsigned int c1 = c1b; c1 <= c1e; ++c1) {
      ++wk.box1;
      if (dist2_box(q.x, q.y, q.z, box1[c1]) > key_d2(best)) continue;
      scan_range(tgt, max(b, c1 * kChunk1), min(e, (c1 + 1u) * kChunk1), q, one, best, wk);
    }
  }
}

// One DENSE cell scanned by a whole warp (all 32 lanes hold the same query): per step the lanes test 32 chunk boxes or 32 candidates
// (one coalesced 512 B row load); every lane keeps the minimum of the keys it has seen, the pruning bound is the warp minimum of
// their d2 (one REDUX per step), the result the warp minimum of the keys. A thread walking such a cell alone is a chain of hundreds
// of dependent loads — with 10^4..10^5 points per cell at a scanner's zenith those few threads WERE the kernel's duration (ncu r02d:
// one SM active for the whole launch, the average SM for half of it). Same candidates, same keys, same minimum: the result is
// identical to scan_cell's.
__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long k) {
  const unsigned int hi = __reduce_min_sync(0xffffffffu, (unsigned int)(k >> 32));
  const unsigned int lo = __reduce_min_sync(0xffffffffu, (unsigned int)(k >> 32) == hi ? (unsigned int)k : 0xffffffffu);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long scan_cell_warp(const float4* __restrict__ tgt, const Aabb* __restrict__ box1,
                                                             const Aabb* __restrict__ box2, unsigned int b, unsigned int e, const float4& q,
                                                             float one, unsigned long long start, unsigned int lane, SearchWork& wk) {
  unsigned long long best = start;                       // this lane's minimum
  float bound = key_d2(start);                           // warp-uniform: min over lanes of key_d2(best)
  const unsigned int last = e - 1u;
  const bool huge = e - b > 2u * kChunk2;
  if (huge) {                                            // 32 strided candidates first: a tight bound before any box is tested
    const float4 t = __ldg(tgt + b + lane * ((e - b) / 32u));
    B2_NN_KEY(t)
    bound = __uint_as_float(__reduce_min_sync(0xffffffffu, (unsigned int)(best >> 32)));
    if (lane == 0) wk.points += 32u;
  }
  for (unsigned int c2 = b / kChunk2; c2 <= last / kChunk2; ++c2) {
    if (huge) { if (lane == 0) ++wk.box2; if (dist2_box(q.x, q.y, q.z, box2[c2]) > bound) continue; }
    // the (up to) 32 level-1 boxes of this level-2 chunk, one per lane
    const unsigned int c1 = c2 * 32u + lane;
    const bool in = c1 >= b / kChunk1 && c1 <= last / kChunk1;
    const float lb = in ? dist2_box(q.x, q.y, q.z, box1[c1]) : INFINITY;
    const unsigned int inside = __ballot_sync(0xffffffffu, in);
    if (lane == 0) wk.box1 += __popc(inside);
    // nearest box first (its candidates tighten the bound most), then the others in order, each re-checked against the bound
    const unsigned int lbmin = __reduce_min_sync(0xffffffffu, __float_as_uint(lb));
    unsigned int todo = __ballot_sync(0xffffffffu, in && !(lb > bound));
    const unsigned int first = __ballot_sync(0xffffffffu, __float_as_uint(lb) == lbmin) & todo;
    bool pick_first = first != 0u;
    while (todo) {
      const unsigned int i = pick_first ? __ffs(first) - 1u : __ffs(todo) - 1u;
      pick_first = false;
      todo &= ~(1u << i);
      if (__shfl_sync(0xffffffffu, lb, i) > bound) continue;
      const unsigned int p = (c2 * 32u + i) * kChunk1 + lane;
      if (p >= b && p < e) { const float4 t = __ldg(tgt + p); B2_NN_KEY(t) }
      if (lane == 0) wk.points += kChunk1;
      bound = __uint_as_float(__reduce_min_sync(0xffffffffu, (unsigned int)(best >> 32)));
    }
  }
  return warp_min_key(best);
}

// Cell of a global-frame point in the target's grid: cell coordinates, the half of the cell per axis and the conservative
// distances to the nearer faces.
struct CellLookup { int cx, cy, cz; unsigned int upper; float ex2, ey2, ez2; };
__device__ __forceinline__ CellLookup lookup_cell(const SearchGrid& g, const float4& q) {
  const float lx = __fmaf_rn(g.m[0], q.x, __fmaf_rn(g.m[1], q.y, __fmaf_rn(g.m[2], q.z, g.m[3])));
  const float ly = __fmaf_rn(g.m[4], q.x, __fmaf_rn(g.m[5], q.y, __fmaf_rn(g.m[6], q.z, g.m[7])));
  const float lz = __fmaf_rn(g.m[8], q.x, __fmaf_rn(g.m[9], q.y, __fmaf_rn(g.m[10], q.z, g.m[11])));
  // clamped so that the int conversion is defined; a query more than a cell outside the grid has no candidate at all
  const float ux = fminf(fmaxf(fmul(lx, g.inv), -2.f), (float)g.nx + 1.f);
  const float uy = fminf(fmaxf(fmul(ly, g.inv), -2.f), (float)g.ny + 1.f);
  const float uz = fminf(fmaxf(fmul(lz, g.inv), -2.f), (float)g.nz + 1.f);
  const float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
  const float rx = fsub(ux, fx), ry = fsub(uy, fy), rz = fsub(uz, fz);
  CellLookup c;
  c.cx = (int)fx; c.cy = (int)fy; c.cz = (int)fz;
  const bool hx = rx >= 0.5f, hy = ry >= 0.5f, hz = rz >= 0.5f;
  c.upper = (hx ? 1u : 0u) | (hy ? 2u : 0u) | (hz ? 4u : 0u);
  // distance to the nearer face: cloud-frame -> global (inv_sigma), minus every rounding on the way (margin), squared and shrunk so
  // that the fp32 rounding of d2 itself can never beat the bound
  const float ex = fmaxf(fsub(fmul(fmul(hx ? fsub(1.f, rx) : rx, g.cell), g.inv_sigma), g.margin), 0.f);
  const float ey = fmaxf(fsub(fmul(fmul(hy ? fsub(1.f, ry) : ry, g.cell), g.inv_sigma), g.margin), 0.f);
  const float ez = fmaxf(fsub(fmul(fmul(hz ? fsub(1.f, rz) : rz, g.cell), g.inv_sigma), g.margin), 0.f);
  c.ex2 = fmul(fmul(ex, ex), 0.9999f); c.ey2 = fmul(fmul(ey, ey), 0.9999f); c.ez2 = fmul(fmul(ez, ez), 0.9999f);
  return c;
}

// Cell lookup. Two index layouts, chosen per cloud when it is indexed:
//   DENSE  (grids up to 2^31 cells — a 10 x 8 x 3 m room at 2 cm has 3 * 10^7): a rank bitmap over the grid, one {bits, prefix}
//          pair per 32 cells; an occupied cell's rank = prefix + popc(bits below it) indexes `starts` (first point of every occupied
//          cell in sorted order, + the total). One 8 B load answers "occupied?" and "where?"; no probing, 32-bit cell keys; ~10 MB
//          for a 10 M-point room scan, so the whole index stays in L2.
//   sparse (anything larger): open-addressing hash of the occupied cells, 64-bit keys.
__device__ __forceinline__ bool dense_rank(const SearchGrid& g, int key, unsigned int* rank) {
  const uint2 w = __ldg(g.rb + (key >> 5));
  const unsigned int bit = 1u << (key & 31);
  *rank = w.y + __popc(w.x & (bit - 1u));
  return (w.x & bit) != 0u;
}
template <bool DENSE>
__device__ __forceinline__ bool find_cell(const SearchGrid& g, const HashEntry* __restrict__ table, long long key, unsigned int* b, unsigned int* e) {
  if (DENSE) {
    unsigned int rank;
    if (!dense_rank(g, (int)key, &rank)) return false;
    *b = __ldg(g.starts + rank); *e = __ldg(g.starts + rank + 1);
    return true;
  }
  return hash_find(table, g.log2size, (unsigned long long)key, b, e);
}

// Launch-order heuristic for K3: estimated cost of each tile = target population of the cells of eight of its queries. The
// tiles are then issued longest-first (ids sorted by descending cost), so the expensive ones (queries inside the target's
// scanner-zenith clusters) overlap with the rest instead of forming the tail of the launch.
template <bool DENSE>
__global__ void __launch_bounds__(256) k_tile_cost(const float4* __restrict__ src, size_t ns, const HashEntry* __restrict__ table, SearchGrid g,
                                                   unsigned int ntiles, unsigned int* __restrict__ cost, unsigned int* __restrict__ ids) {
  const unsigned int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ntiles) return;
  unsigned int total = 0;
  for (int k = 0; k < 8; ++k) {
    const size_t j = (size_t)c * kTile + 32 * k + 16;
    if (j >= ns) break;
    const float4 q = src[j];
    const CellLookup L = lookup_cell(g, q);
    if ((unsigned int)L.cx >= (unsigned int)g.nx || (unsigned int)L.cy >= (unsigned int)g.ny || (unsigned int)L.cz >= (unsigned int)g.nz) continue;
    unsigned int b, e;
    if (find_cell<DENSE>(g, table, (long long)L.cz * g.sz + (long long)L.cy * g.sy + L.cx, &b, &e)) total += e - b;
  }
  cost[c] = total; ids[c] = c;
}

// Per-warp working set in shared memory (a CTA is kTile / 32 independent warps; nothing in K3 synchronises the CTA).
template <typename key_t>
struct __align__(128) WarpTile {
  float4 q[2][32];                         // the warp's 32 query rows (its own bulk copy), double-buffered across the tiles of a persistent CTA
  unsigned long long best[32];             // per query: best (d2, index) key
  key_t item_cell[32 * 7];                 // work queue of (query, neighbour cell): DENSE: the cell's rank; sparse: the cell's key
  uint2 dense_range[kDenseCap];            // dense-cell queue: candidate range [x, y) ...
  unsigned char item_q[32 * 7];
  unsigned char dense_q[kDenseCap];        // ... and the query it belongs to
  unsigned long long bar[2];
};

template <bool STATS, bool DENSE, bool DUAL>
__global__ void __launch_bounds__(kTile, B2_K3_MINB) k_nn_tiles(const float4* __restrict__ src, size_t ns, const float4* __restrict__ tgt,
                                                    const Aabb* __restrict__ box1, const Aabb* __restrict__ box2,
                                                    const HashEntry* __restrict__ table, SearchGrid g, float r2,
                                                    unsigned long long* __restrict__ out_key, unsigned int* __restrict__ tile_count,
                                                    unsigned long long* __restrict__ work, const unsigned int* __restrict__ order,
                                                    unsigned int ntiles) {
  typedef typename std::conditional<DENSE, int, long long>::type key_t;   // cell key: 32 bits suffice for a dense grid
  __shared__ WarpTile<key_t> s_warp[kTile / 32];
  const unsigned int lane = threadIdx.x & 31u;
  WarpTile<key_t>& W = s_warp[threadIdx.x >> 5];
  // Persistent CTAs: CTA b works through tiles order[b], order[b + grid], ... (with the longest-first order every CTA receives an equal
  // share of long and short tiles, so all of them finish together instead of leaving the launch's tail to a few long tiles), and each
  // warp has the query rows of its NEXT tile copied into the other half of its buffer while it works on the current one. With
  // grid == ntiles this is the one-tile-per-CTA kernel.
  const unsigned int wofs = threadIdx.x & ~31u;
  if (lane == 0) {
    mbar_init(&W.bar[0], 1); mbar_init(&W.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto stage = [&](unsigned int t, unsigned int buf) {                    // lane 0: start the copy of this warp's rows of tile number t
    const size_t j = (size_t)(order ? order[t] : t) * kTile + wofs;
    if (j < ns) {
      const unsigned int c = (unsigned int)min((size_t)32, ns - j);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the rows read from this half are overwritten by the copy engine
      mbar_expect_tx(&W.bar[buf], c * 16u);
      bulk_g2s(W.q[buf], src + j, c * 16u, &W.bar[buf]);
    }
  };
  if (lane == 0 && blockIdx.x < ntiles) stage(blockIdx.x, 0u);
  const unsigned long long init = (unsigned long long)__float_as_uint(r2) << 32;
  SearchWork wk = {0u, 0u, 0u, 0u};
  unsigned int total_items = 0u, phases = 0u, it = 0u;
  for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
  const unsigned int buf = it & 1u;
  const unsigned int tile = order ? order[t] : t;
  const size_t j0 = (size_t)tile * kTile + wofs;                          // this warp's first query
  if (lane == 0 && t + gridDim.x < ntiles) stage(t + gridDim.x, buf ^ 1u);
  if (j0 >= ns) continue;                                                 // (whole warp; nothing was staged for it)
  const unsigned int cnt = (unsigned int)min((size_t)32, ns - j0);
  mbar_wait(&W.bar[buf], (phases >> buf) & 1u);
  phases ^= 1u << buf;
  const float4* __restrict__ Wq = W.q[buf];

  // ---- A: own cell per thread ----
  unsigned long long best = init;
  key_t base = 0, dxk = 0, dyk = 0, dzk = 0;
  unsigned int todo = 0u;
  bool own_dense = false;
  uint2 own_range = make_uint2(0u, 0u);
  if (lane < cnt) {
    const float4 q = Wq[lane];
    const CellLookup L = lookup_cell(g, q);
    const unsigned int upper = L.upper;
    base = (key_t)((key_t)L.cz * (key_t)g.sz + (key_t)L.cy * (key_t)g.sy + (key_t)L.cx);
    dxk = (upper & 1u) ? (key_t)1 : (key_t)-1;
    dyk = (upper & 2u) ? (key_t)g.sy : (key_t)-g.sy;
    dzk = (upper & 4u) ? (key_t)g.sz : (key_t)-g.sz;
    const bool x0 = (unsigned int)L.cx < (unsigned int)g.nx, y0 = (unsigned int)L.cy < (unsigned int)g.ny, z0 = (unsigned int)L.cz < (unsigned int)g.nz;
    const bool x1 = (unsigned int)(L.cx + ((upper & 1u) ? 1 : -1)) < (unsigned int)g.nx;
    const bool y1 = (unsigned int)(L.cy + ((upper & 2u) ? 1 : -1)) < (unsigned int)g.ny;
    const bool z1 = (unsigned int)(L.cz + ((upper & 4u) ? 1 : -1)) < (unsigned int)g.nz;
    // DUAL: the 2x2x2 block of cells a query can have matches in is the set of cells around ONE lattice corner — the corner of its cell
    // on the side of the octant it lies in. One byte of the corner map says which of the eight are inside the grid and occupied; it
    // replaces seven bitmap lookups (key, word load, bit test) and all range checks. Bit j of occ8, after the permutation: the cell
    // reached from the query's own by stepping along the axes in j (bit 0 = its own cell).
    unsigned int occ8 = 0u;
    if (DUAL) {
      const int kx = L.cx + (int)(upper & 1u), ky = L.cy + (int)((upper >> 1) & 1u), kz = L.cz + (int)((upper >> 2) & 1u);
      if ((unsigned int)kx <= (unsigned int)g.nx && (unsigned int)ky <= (unsigned int)g.ny && (unsigned int)kz <= (unsigned int)g.nz) {
        unsigned int m = __ldg(g.corner + ((size_t)kz * (size_t)(g.ny + 1) + (size_t)ky) * (size_t)(g.nx + 1) + (size_t)kx);
        // the query's own cell is the corner's cell number (~upper & 7); bit j of the result = bit (j ^ own) of m
        if (!(upper & 1u)) m = ((m & 0x55u) << 1) | ((m >> 1) & 0x55u);
        if (!(upper & 2u)) m = ((m & 0x33u) << 2) | ((m >> 2) & 0x33u);
        if (!(upper & 4u)) m = ((m & 0x0fu) << 4) | (m >> 4);
        occ8 = m;
      }
    }
    if (DUAL ? (occ8 & 1u) != 0u : (x0 && y0 && z0)) {
      unsigned int b, e;
      if (find_cell<DENSE>(g, table, base, &b, &e)) {
        if (e - b > kCoopCell) {
          // dense own cell: the whole warp scans it in phase B2; 16 strided candidates now, so that the neighbours are still pruned
          own_dense = true; own_range = make_uint2(b, e);
          const unsigned int stride = (e - b) / 16u;
          for (unsigned int p = b; p < b + 16u * stride; p += 4u * stride) {
            const float4 t0 = __ldg(tgt + p), t1 = __ldg(tgt + p + stride), t2 = __ldg(tgt + p + 2u * stride), t3 = __ldg(tgt + p + 3u * stride);
            const float one = g.one;
            B2_NN_KEY(t0) B2_NN_KEY(t1) B2_NN_KEY(t2) B2_NN_KEY(t3)
          }
        } else {
          scan_cell(tgt, box1, box2, b, e, q, g.one, best, wk);
        }
      }
    }
    // neighbours worth a visit: inside the grid, nearest face not farther than the best match, and OCCUPIED (most neighbour cells
    // of a surface scan are empty; the test is one bitmap word)
    const float bd = key_d2(best);
#pragma unroll
    for (int c = 1; c < 8; ++c) {
      const float lb = fadd(fadd((c & 1) ? L.ex2 : 0.f, (c & 2) ? L.ey2 : 0.f), (c & 4) ? L.ez2 : 0.f);
      if (DUAL) { if (((occ8 >> c) & 1u) && !(lb > bd)) todo |= 1u << c; continue; }
      const bool valid = ((c & 1) ? x1 : x0) && ((c & 2) ? y1 : y0) && ((c & 4) ? z1 : z0);
      if (valid && !(lb > bd)) {
        const key_t key = base + ((c & 1) ? dxk : (key_t)0) + ((c & 2) ? dyk : (key_t)0) + ((c & 4) ? dzk : (key_t)0);
        bool occ = true;
        if (DENSE) { const uint2 w = __ldg(g.rb + ((int)key >> 5)); occ = (w.x >> ((int)key & 31)) & 1u; }
        if (occ) todo |= 1u << c;
      }
    }
  }
  // queue slots by prefix sums over the warp: no atomics, the queues are private to the warp
  unsigned int ndense;
  {
    const unsigned int dm = __ballot_sync(0xffffffffu, own_dense);
    ndense = __popc(dm);                                                // <= 32 < kDenseCap
    if (own_dense) { const unsigned int slot = __popc(dm & ((1u << lane) - 1u)); W.dense_range[slot] = own_range; W.dense_q[slot] = (unsigned char)lane; }
  }
  unsigned int nitems;
  {
    const unsigned int mine = __popc(todo);
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned int)o) incl += v; }
    nitems = __shfl_sync(0xffffffffu, incl, 31);
    unsigned int at = incl - mine;
    while (todo) {
      const unsigned int c = __ffs(todo) - 1u;
      todo &= todo - 1u;
      const key_t key = base + ((c & 1u) ? dxk : (key_t)0) + ((c & 2u) ? dyk : (key_t)0) + ((c & 4u) ? dzk : (key_t)0);
      if (DENSE) { unsigned int rank; dense_rank(g, (int)key, &rank); W.item_cell[at] = (key_t)rank; }   // (the word is in L1 from the test above)
      else W.item_cell[at] = key;
      W.item_q[at] = (unsigned char)lane;
      ++at;
    }
  }
  W.best[lane] = best;
  __syncwarp();

  // ---- B: neighbour cells, one queue item per lane ----
  for (unsigned int i0 = 0; i0 < nitems; i0 += 32u) {
    const unsigned int i = i0 + lane;
    bool dense_item = false;
    uint2 range = make_uint2(0u, 0u);
    unsigned int ql = 0;
    if (i < nitems) {
      ql = W.item_q[i];
      const key_t cell = W.item_cell[i];
      unsigned int b = 0, e = 0;
      bool found = true;
      if (DENSE) { b = __ldg(g.starts + (unsigned int)cell); e = __ldg(g.starts + (unsigned int)cell + 1u); }
      else found = hash_find(table, g.log2size, (unsigned long long)cell, &b, &e);
      if (found) {
        if (e - b > kCoopCell) { dense_item = true; range = make_uint2(b, e); }
        else {
          const float4 q = Wq[ql];
          const unsigned long long seen = W.best[ql];     // possibly lowered by another item of this query already: a tighter start
          unsigned long long bk = seen;
          scan_cell(tgt, box1, box2, b, e, q, g.one, bk, wk);
          if (bk < seen) atomicMin(&W.best[ql], bk);
        }
      }
    }
    const unsigned int dm = __ballot_sync(0xffffffffu, dense_item);
    if (dense_item) {
      const unsigned int slot = ndense + __popc(dm & ((1u << lane) - 1u));
      if (slot < kDenseCap) { W.dense_range[slot] = range; W.dense_q[slot] = (unsigned char)ql; }
      else {                                             // queue full: this lane walks the cell itself
        const float4 q = Wq[ql];
        const unsigned long long seen = W.best[ql];
        unsigned long long bk = seen;
        scan_cell(tgt, box1, box2, range.x, range.y, q, g.one, bk, wk);
        if (bk < seen) atomicMin(&W.best[ql], bk);
      }
    }
    ndense = min(ndense + __popc(dm), kDenseCap);
  }
  __syncwarp();

  // ---- B2: dense cells, the whole warp on one queue entry at a time ----
  for (unsigned int i = 0; i < ndense; ++i) {
    const uint2 r = W.dense_range[i];
    const unsigned int ql = W.dense_q[i];
    const float4 q = Wq[ql];
    const unsigned long long seen = W.best[ql];
    const unsigned long long bk = scan_cell_warp(tgt, box1, box2, r.x, r.y, q, g.one, seen, lane, wk);
    if (lane == 0u) { if (bk < seen) W.best[ql] = bk; ++wk.cells; }
    __syncwarp();
  }

  // ---- C: results ----
  bool matched = false;
  if (lane < cnt) {
    best = W.best[lane];
    matched = best < init;
    out_key[j0 + lane] = best;
  }
  const unsigned int mm = __ballot_sync(0xffffffffu, matched);
  if (lane == 0u && mm) atomicAdd(tile_count + tile, (unsigned int)__popc(mm));   // (zeroed by the host before the launch)
  if (STATS) total_items += nitems;
  __syncwarp();                                                           // the queues and W.best are reused by the next tile
  }
  if (STATS) {
    // per-launch totals: candidates tested, level-1 / level-2 box tests, cells scanned, queue items
    unsigned int v[4] = {wk.points, wk.box1, wk.box2, wk.cells};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned int x = v[k];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0u && x) atomicAdd(work + k, (unsigned long long)x);
    }
    if (lane == 0u && total_items) atomicAdd(work + 4, (unsigned long long)total_items);
  }
}

// ---- one-time construction of the rank bitmap (DENSE layout) ----
__global__ void __launch_bounds__(256) k_mark_cells(const unsigned long long* __restrict__ keys, size_t n, int kFineBits, uint2* __restrict__ rb) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long key = keys[j] >> kFineBits;
  if (j != 0 && (keys[j - 1] >> kFineBits) == key) return;
  atomicOr(&rb[key >> 5].x, 1u << (unsigned int)(key & 31ull));
}
// The corner map of the DUAL lookup: thread per lattice corner (kx, ky, kz) in [0, nx] x [0, ny] x [0, nz]; cell (kx-1+dx, ky-1+dy, kz-1+dz).
__global__ void __launch_bounds__(256) k_corner_occupancy(const uint2* __restrict__ rb, int nx, int ny, int nz, unsigned char* __restrict__ corner) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)(nx + 1) * (size_t)(ny + 1) * (size_t)(nz + 1);
  if (i >= total) return;
  const int kx = (int)(i % (size_t)(nx + 1));
  const size_t t = i / (size_t)(nx + 1);
  const int ky = (int)(t % (size_t)(ny + 1)), kz = (int)(t / (size_t)(ny + 1));
  unsigned int m = 0u;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const int x = kx - 1 + (d & 1), y = ky - 1 + ((d >> 1) & 1), z = kz - 1 + (d >> 2);
    if ((unsigned int)x < (unsigned int)nx && (unsigned int)y < (unsigned int)ny && (unsigned int)z < (unsigned int)nz) {
      const long long key = ((long long)z * ny + y) * nx + x;
      m |= ((__ldg(&rb[key >> 5].x) >> (unsigned int)(key & 31ll)) & 1u) << d;
    }
  }
  corner[i] = (unsigned char)m;
}
static constexpr unsigned int kWordsPerBlock = 2048;   // bitmap words per block of the two-level prefix
__global__ void __launch_bounds__(256) k_word_counts(const uint2* __restrict__ rb, size_t nwords, unsigned int* __restrict__ block_sum) {
  __shared__ unsigned int s[8];
  const size_t w0 = (size_t)blockIdx.x * kWordsPerBlock;
  unsigned int c = 0;
  for (unsigned int i = threadIdx.x; i < kWordsPerBlock; i += 256) if (w0 + i < nwords) c += __popc(rb[w0 + i].x);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned int v = 0; for (int i = 0; i < 8; ++i) v += s[i]; block_sum[blockIdx.x] = v; }
}
__global__ void __launch_bounds__(256) k_word_prefix(uint2* __restrict__ rb, size_t nwords, const unsigned int* __restrict__ block_off) {
  // block-wide exclusive scan of the 2048 word counts of this block (8 consecutive words per thread), + the block's offset
  __shared__ unsigned int s[256];
  const size_t w0 = (size_t)blockIdx.x * kWordsPerBlock + (size_t)threadIdx.x * 8;
  unsigned int c[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i] = w0 + i < nwords ? __popc(rb[w0 + i].x) : 0u; sum += c[i]; }
  s[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const unsigned int v = threadIdx.x >= (unsigned int)o ? s[threadIdx.x - o] : 0u;
    __syncthreads();
    s[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned int run = block_off[blockIdx.x] + s[threadIdx.x] - sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) { if (w0 + i < nwords) rb[w0 + i].y = run; run += c[i]; }
}
__global__ void __launch_bounds__(256) k_cell_starts(const unsigned long long* __restrict__ keys, size_t n, int kFineBits, const uint2* __restrict__ rb,
                                                     unsigned int* __restrict__ starts, unsigned int ncells) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (j == 0) starts[ncells] = (unsigned int)n;
  const unsigned long long key = keys[j] >> kFineBits;
  if (j != 0 && (keys[j - 1] >> kFineBits) == key) return;
  const uint2 w = rb[key >> 5];
  starts[w.y + __popc(w.x & ((1u << (unsigned int)(key & 31ull)) - 1u))] = (unsigned int)j;
}

// ------------------------------------------------------------------------------------------------------------------
// K4: packed correspondence records, three float4 planes (48 B / correspondence, fully coalesced in K5):
//   A = (ps.x, ps.y, ps.z, ns.x)  B = (ns.y, ns.z, pt.x, pt.y)  C = (pt.z, nt.x, nt.y, nt.z)
// Record order = ascending cell-sorted source position: the tile offsets are an exclusive scan of K3's per-tile match counts
// (k_scan_tiles, one block), the position inside a tile comes from ballots — no per-query flag / offset arrays.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_tiles(const unsigned int* __restrict__ tile_count, unsigned int ntiles,
                                                     unsigned int* __restrict__ tile_off, unsigned int* __restrict__ total) {
  __shared__ unsigned int s[1024];
  const unsigned int per = (ntiles + 1023u) / 1024u, b = threadIdx.x * per, e = min(ntiles, b + per);
  unsigned int sum = 0;
  for (unsigned int i = b; i < e; ++i) sum += tile_count[i];
  s[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned int v = threadIdx.x >= (unsigned int)o ? s[threadIdx.x - o] : 0u;
    __syncthreads();
    s[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned int run = s[threadIdx.x] - sum;
  for (unsigned int i = b; i < e; ++i) { tile_off[i] = run; run += tile_count[i]; }
  if (threadIdx.x == 1023) *total = s[1023];
}

// Where a set's records start: either the host's value (`base`, chain == nullptr) or — when the sets are packed while later sets are
// still being searched — a device-side running offset: set number `pos` of the iteration starts at chain->base[pos] and publishes
// chain->base[pos + 1] = its start + its match count (the packs of an iteration run in set order on one stream). A set that would not
// fit into the record arrays raises chain->overflow and writes nothing; the host then packs the classic way.
struct PackChain { unsigned long long* base; const unsigned int* total; unsigned int* overflow; unsigned long long cap; int pos; };

__global__ void __launch_bounds__(kTile) k_pack_tiles(const float4* __restrict__ s_xyz_src, const float4* __restrict__ s_nrm_src, size_t ns,
                                                      const float4* __restrict__ s_xyz_tgt, const float4* __restrict__ s_nrm_tgt,
                                                      const unsigned int* __restrict__ perm_inv_tgt, const unsigned long long* __restrict__ key,
                                                      const unsigned int* __restrict__ tile_off, unsigned long long init_key,
                                                      unsigned long long base, PackChain chain, float4* __restrict__ ra, float4* __restrict__ rb,
                                                      float4* __restrict__ rc) {
  __shared__ unsigned int warp_sum[kTile / 32];
  if (chain.base) {                                   // (block-uniform)
    base = chain.base[chain.pos];
    const unsigned long long end = base + (unsigned long long)*chain.total;
    if (blockIdx.x == 0 && threadIdx.x == 0) { chain.base[chain.pos + 1] = end; if (end > chain.cap) *chain.overflow = 1u; }
    if (end > chain.cap) return;
  }
  const size_t j = (size_t)blockIdx.x * kTile + threadIdx.x;
  const unsigned int lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const unsigned long long k = j < ns ? key[j] : init_key;
  const bool matched = k < init_key;
  const unsigned int m = __ballot_sync(0xffffffffu, matched);
  if (lane == 0u) warp_sum[w] = __popc(m);
  __syncthreads();
  if (!matched) return;
  unsigned int before = __popc(m & ((1u << lane) - 1u));
  for (unsigned int i = 0; i < w; ++i) before += warp_sum[i];
  const unsigned int p = perm_inv_tgt[(unsigned int)k];
  const float4 ps = s_xyz_src[j], nsr = s_nrm_src[j];
  const float4 pt = __ldg(s_xyz_tgt + p), nt = __ldg(s_nrm_tgt + p);
  const unsigned long long o = base + tile_off[blockIdx.x] + before;
  // source and target components interleaved (see "Record layout" at K5): the packed fp32 path reads its operand pairs as they lie
  ra[o] = make_float4(ps.x, pt.x, ps.y, pt.y);
  rb[o] = make_float4(ps.z, pt.z, nsr.x, nt.x);
  rc[o] = make_float4(nsr.y, nt.y, nsr.z, nt.z);
}

// Correspondence list in the caller's (original) indexing, scattered to the original query index.
__global__ void __launch_bounds__(256) k_scatter_matches(const float4* __restrict__ s_xyz_src, size_t ns, const unsigned long long* __restrict__ key,
                                                         unsigned long long init_key, int* __restrict__ out_match, float* __restrict__ out_d2) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const unsigned int qi = __float_as_uint(s_xyz_src[j].w);
  const unsigned long long k = key[j];
  out_match[qi] = k < init_key ? (int)(unsigned int)k : -1;
  out_d2[qi] = key_d2(k);
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

// Record layout (three float4 planes, 48 B): a = (ps.x, pt.x, ps.y, pt.y), b = (ps.z, pt.z, ns.x, nt.x), c = (ns.y, nt.y, ns.z, nt.z) —
// the source and target value of every component side by side, which is how the packed path below consumes them.
struct RecordFields { float psx, psy, psz, nsx, nsy, nsz, ptx, pty, ptz, ntx, nty, ntz; };
__device__ __forceinline__ RecordFields record_fields(const float4 a, const float4 b, const float4 c) {
  RecordFields f;
  f.psx = a.x; f.ptx = a.y; f.psy = a.z; f.pty = a.w; f.psz = b.x; f.ptz = b.y; f.nsx = b.z; f.ntx = b.w; f.nsy = c.x; f.nty = c.y; f.nsz = c.z; f.ntz = c.w;
  return f;
}

__device__ __forceinline__ void accumulate_record(const float4 a, const float4 b, const float4 c, const float Rs[9], const float ts[3],
                                                  const float Rt[9], const float tt[3], double acc[kAccVals]) {
  const RecordFields f = record_fields(a, b, c);
  const float psx = fadd(sum3(fmul(Rs[0], f.psx), fmul(Rs[1], f.psy), fmul(Rs[2], f.psz)), ts[0]);
  const float psy = fadd(sum3(fmul(Rs[3], f.psx), fmul(Rs[4], f.psy), fmul(Rs[5], f.psz)), ts[1]);
  const float psz = fadd(sum3(fmul(Rs[6], f.psx), fmul(Rs[7], f.psy), fmul(Rs[8], f.psz)), ts[2]);
  const float nsx = sum3(fmul(Rs[0], f.nsx), fmul(Rs[1], f.nsy), fmul(Rs[2], f.nsz));
  const float nsy = sum3(fmul(Rs[3], f.nsx), fmul(Rs[4], f.nsy), fmul(Rs[5], f.nsz));
  const float nsz = sum3(fmul(Rs[6], f.nsx), fmul(Rs[7], f.nsy), fmul(Rs[8], f.nsz));
  const float ptx = fadd(sum3(fmul(Rt[0], f.ptx), fmul(Rt[1], f.pty), fmul(Rt[2], f.ptz)), tt[0]);
  const float pty = fadd(sum3(fmul(Rt[3], f.ptx), fmul(Rt[4], f.pty), fmul(Rt[5], f.ptz)), tt[1]);
  const float ptz = fadd(sum3(fmul(Rt[6], f.ptx), fmul(Rt[7], f.pty), fmul(Rt[8], f.ptz)), tt[2]);
  const float ntx = sum3(fmul(Rt[0], f.ntx), fmul(Rt[1], f.nty), fmul(Rt[2], f.ntz));
  const float nty = sum3(fmul(Rt[3], f.ntx), fmul(Rt[4], f.nty), fmul(Rt[5], f.ntz));
  const float ntz = sum3(fmul(Rt[6], f.ntx), fmul(Rt[7], f.nty), fmul(Rt[8], f.ntz));

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
  const RecordFields f = record_fields(a, b, c);
  const float psx = fadd(sum3(fmul(Rs[0], f.psx), fmul(Rs[1], f.psy), fmul(Rs[2], f.psz)), ts[0]);
  const float psy = fadd(sum3(fmul(Rs[3], f.psx), fmul(Rs[4], f.psy), fmul(Rs[5], f.psz)), ts[1]);
  const float psz = fadd(sum3(fmul(Rs[6], f.psx), fmul(Rs[7], f.psy), fmul(Rs[8], f.psz)), ts[2]);
  const float nsx = sum3(fmul(Rs[0], f.nsx), fmul(Rs[1], f.nsy), fmul(Rs[2], f.nsz));
  const float nsy = sum3(fmul(Rs[3], f.nsx), fmul(Rs[4], f.nsy), fmul(Rs[5], f.nsz));
  const float nsz = sum3(fmul(Rs[6], f.nsx), fmul(Rs[7], f.nsy), fmul(Rs[8], f.nsz));
  const float ptx = fadd(sum3(fmul(Rt[0], f.ptx), fmul(Rt[1], f.pty), fmul(Rt[2], f.ptz)), tt[0]);
  const float pty = fadd(sum3(fmul(Rt[3], f.ptx), fmul(Rt[4], f.pty), fmul(Rt[5], f.ptz)), tt[1]);
  const float ptz = fadd(sum3(fmul(Rt[6], f.ptx), fmul(Rt[7], f.pty), fmul(Rt[8], f.ptz)), tt[2]);
  const float ntx = sum3(fmul(Rt[0], f.ntx), fmul(Rt[1], f.nty), fmul(Rt[2], f.ntz));
  const float nty = sum3(fmul(Rt[3], f.ntx), fmul(Rt[4], f.nty), fmul(Rt[5], f.ntz));
  const float ntz = sum3(fmul(Rt[6], f.ntx), fmul(Rt[7], f.nty), fmul(Rt[8], f.ntz));
  const float r1 = dot3(nsx, nsy, nsz, fsub(ptx, psx), fsub(pty, psy), fsub(ptz, psz));
  const float r2 = dot3(ntx, nty, ntz, fsub(psx, ptx), fsub(psy, pty), fsub(psz, ptz));
  *cost += (double)fmul(r1, r1);
  *cost += (double)fmul(r2, r2);
}

// ---- the cost of a record with packed fp32 (sm_100 FFMA2: one instruction, two IEEE-rn results) ----------------------------------
// Lane 0 carries the source side of a quantity, lane 1 the target side: (ps_k, pt_k) = (Rs, Rt)(row k) . (p_s0, p_t0) + (ts_k, tt_k),
// (ns_k, nt_k) likewise; then d = pt - ps (scalar), (r1, -r2) = (ns, nt) . d — the reference's r2 = nt . (ps - pt) is exactly the negative,
// products and round-to-nearest sums being sign-symmetric — and (r1^2, r2^2). Every lane performs the operations of cost_record in
// cost_record's order, each rounded once, so the two functions return the same bits; the instruction count per record and trial
// drops from ~96 to ~46, which is what bounds a pass that evaluates four LM tries (ncu r02D: 76 % issue-active, 63 % FMA pipe).
// ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even under -fmad=false, so products and sums are written as FMAs that cannot be
// contracted any further: a * b = fma(a, b, -0), a + b = fma(a, 1, b), with 1 and -0 passed at run time so that they are not folded back.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
struct PackedOps {
  f32x2 one, nzero;
  __device__ __forceinline__ f32x2 mul(f32x2 a, f32x2 b) const { return fma2(a, b, nzero); }
  __device__ __forceinline__ f32x2 add(f32x2 a, f32x2 b) const { return fma2(a, one, b); }
  __device__ __forceinline__ f32x2 sum3(f32x2 a, f32x2 b, f32x2 c) const { return add(a, add(b, c)); }
};
// P: 12 pairs of one trial — (Rs[k], Rt[k]) for k = 0..8, then (ts[k], tt[k]) for k = 0..2.
__device__ __forceinline__ void cost_record_packed(const float4 a, const float4 b, const float4 c, const f32x2 P[12], const PackedOps& K, double* cost) {
  const f32x2 x0 = pk2(a.x, a.y), x1 = pk2(a.z, a.w), x2 = pk2(b.x, b.y);      // (ps0, pt0) by component: adjacent in the record
  const f32x2 n0 = pk2(b.z, b.w), n1 = pk2(c.x, c.y), n2 = pk2(c.z, c.w);      // (ns0, nt0)
  f32x2 p[3], n[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    p[r] = K.add(K.sum3(K.mul(P[3 * r], x0), K.mul(P[3 * r + 1], x1), K.mul(P[3 * r + 2], x2)), P[9 + r]);
    n[r] = K.sum3(K.mul(P[3 * r], n0), K.mul(P[3 * r + 1], n1), K.mul(P[3 * r + 2], n2));
  }
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) { float ps, pt; upk2(p[r], ps, pt); d[r] = fsub(pt, ps); }
  const f32x2 rr = K.sum3(K.mul(n[0], pk2(d[0], d[0])), K.mul(n[1], pk2(d[1], d[1])), K.mul(n[2], pk2(d[2], d[2])));
  float q1, q2;
  upk2(K.mul(rr, rr), q1, q2);
  *cost += (double)q1;
  *cost += (double)q2;
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
// Passes without the normal equations need a third of the registers (no 28 fp64 accumulators): with three stages (72 KB) instead of four
// three of their CTAs fit an SM (24 warps instead of 16) — they are bound by fp32 latency / issue, not by bytes in flight.
__host__ __device__ constexpr int tma_stages(bool with_h) { return with_h ? kTmaStages : 3; }
__host__ __device__ constexpr size_t tma_smem_bytes(bool with_h) { return (size_t)tma_stages(with_h) * 3 * kTmaTile * 16; }

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
__global__ void __launch_bounds__(kAccThreads, WITH_H ? 2 : 3)
k_accumulate_tma(const float4* __restrict__ ra, const float4* __restrict__ rb, const float4* __restrict__ rc,
                 const Segment* __restrict__ segs, int nseg, const CloudPose* __restrict__ poses /* [1 + NX][nclouds] */, int nclouds,
                 unsigned long long total, unsigned long long per_cta, double* __restrict__ partials /* [nseg][gridDim.x][kAccVals] */,
                 double* __restrict__ xpartials /* [nseg][gridDim.x][kMaxExtraTrials] */, float one, float nzero /* 1.0f, -0.0f */) {
  constexpr int NV = WITH_H ? kAccVals : 1;
  constexpr int NXS = NX > 0 ? NX : 1;
  constexpr int kStages = tma_stages(WITH_H);
  constexpr int kFirstPacked = WITH_H ? 1 : 0;                   // trial 0 goes through the packed cost path unless it carries H
  constexpr bool kAnyPacked = NX > 0 || !WITH_H;
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ double red[kAccThreads / 32][NV];
  __shared__ double xred[kAccThreads / 32][NXS];
  // per trial (0 = the pass's own state, 1.. = the speculative ones): the 12 (source, target) pairs of cost_record_packed
  __shared__ __align__(16) float ppose[1 + NXS][24];
  const PackedOps K = {pk2(one, one), pk2(nzero, nzero)};
  __shared__ __align__(8) unsigned long long full_bar[kStages], empty_bar[kStages];
  const unsigned long long r_begin = (unsigned long long)blockIdx.x * per_cta;
  const unsigned long long r_end = min(total, r_begin + per_cta);
  if (r_begin >= r_end) return;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kAccThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int lo = 0, hi = nseg - 1;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (segs[mid].end > r_begin) hi = mid; else lo = mid + 1; }
  float4* const sa = reinterpret_cast<float4*>(tile_smem);       // stage s: planes at sa + (3*s + p) * kTmaTile

  // Thread 0 doubles as the producer: it keeps kStages-1 tiles in flight ahead of the tile being consumed.
  TileCursor prod{r_begin, 0, r_end, lo}; unsigned int pit = 0;
  auto produce = [&]() {
    const unsigned int s = pit % kStages, cnt = prod.count();
    if (pit >= (unsigned int)kStages) mbar_wait(&empty_bar[s], ((pit / kStages) - 1) & 1);
    mbar_expect_tx(&full_bar[s], 3u * cnt * 16u);
    bulk_g2s(sa + (3 * s + 0) * kTmaTile, ra + prod.r, cnt * 16u, &full_bar[s]);
    bulk_g2s(sa + (3 * s + 1) * kTmaTile, rb + prod.r, cnt * 16u, &full_bar[s]);
    bulk_g2s(sa + (3 * s + 2) * kTmaTile, rc + prod.r, cnt * 16u, &full_bar[s]);
    prod.advance(segs); ++pit;
  };
  if (threadIdx.x == 0) {
    prod.settle(segs);
    for (int k = 0; k < kStages - 1 && prod.valid(); ++k) produce();
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
    if (kAnyPacked) {   // (the reductions at the end of the previous segment separate its readers from this write)
      if ((int)threadIdx.x < (1 + NX) * 24 && (int)threadIdx.x >= kFirstPacked * 24) {
        const int j = threadIdx.x / 24, w = threadIdx.x % 24;
        const CloudPose* P = poses + (size_t)j * nclouds + ((w & 1) ? sg.tgt : sg.src);
        const int q = w >> 1;
        ppose[j][w] = q < 9 ? __ldg(&P->R[q]) : __ldg(&P->t[q - 9]);
      }
      __syncthreads();
    }
    for (unsigned long long r = r0; r < e; r += kTmaTile, ++it) {
      if (threadIdx.x == 0 && prod.valid()) produce();            // refill the stage released one tile ago
      const unsigned int s = it % kStages, cnt = (unsigned int)min((unsigned long long)kTmaTile, e - r);
      mbar_wait(&full_bar[s], (it / kStages) & 1);
      float4 a0, b0, c0, a1, b1, c1;
      const bool m0 = threadIdx.x < cnt, m1 = threadIdx.x + kAccThreads < cnt;
      const float4* st = sa + (3 * s) * kTmaTile;
      if (m0) { a0 = st[threadIdx.x]; b0 = st[kTmaTile + threadIdx.x]; c0 = st[2 * kTmaTile + threadIdx.x]; }
      if (m1) { a1 = st[kAccThreads + threadIdx.x]; b1 = st[kTmaTile + kAccThreads + threadIdx.x]; c1 = st[2 * kTmaTile + kAccThreads + threadIdx.x]; }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[s]);     // this warp's values are in registers: one arrival per warp
      if (WITH_H) {
        if (m0) accumulate_record(a0, b0, c0, Rs, ts, Rt, tt, acc);
        if (m1) accumulate_record(a1, b1, c1, Rs, ts, Rt, tt, acc);
      }
      if (kAnyPacked) {
#pragma unroll
        for (int j = kFirstPacked; j <= NX; ++j) {
          f32x2 P[12];
          const ulonglong2* pp = reinterpret_cast<const ulonglong2*>(ppose[j]);
#pragma unroll
          for (int q = 0; q < 6; ++q) { const ulonglong2 v = pp[q]; P[2 * q] = v.x; P[2 * q + 1] = v.y; }
          double* sum = j == 0 ? &acc[0] : &xacc[j > 0 ? j - 1 : 0];
          if (m0) cost_record_packed(a0, b0, c0, P, K, sum);
          if (m1) cost_record_packed(a1, b1, c1, P, K, sum);
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
