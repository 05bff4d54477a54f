// Implicit BVH over Morton-sorted points, shared by the kNN / radius normals (b2_normals.cu) and the multi-resolution point-cloud
// construction (b2_multiscale.cu).
//
// The points are sorted by a 63-bit Morton code; consecutive runs of 8 sorted points are the leaves of an IMPLICIT binary BVH (node i
// of level l covers leaves [i*2^l, (i+1)*2^l); no pointers, 24 B AABB per node). The AABB lower bound is evaluated with the same fp32
// operations as the point distance, so pruning can never drop a neighbour (rounding is monotone).
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "b2_common.cuh"

namespace b2 {

static constexpr int kLeaf = 8;


static __global__ void __launch_bounds__(256) kn_bbox(const float* __restrict__ xyz, size_t n, float* __restrict__ partial) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    for (int d = 0; d < 3; ++d) { const float v = xyz[3 * i + d]; mn[d] = fminf(mn[d], v); mx[d] = fmaxf(mx[d], v); }
  for (int o = 16; o > 0; o >>= 1)
    for (int d = 0; d < 3; ++d) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  __shared__ float s[8][6];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) for (int d = 0; d < 3; ++d) { s[w][d] = mn[d]; s[w][3 + d] = mx[d]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = s[0][threadIdx.x];
    for (int i = 1; i < 8; ++i) v = threadIdx.x < 3 ? fminf(v, s[i][threadIdx.x]) : fmaxf(v, s[i][threadIdx.x]);
    partial[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

static __global__ void __launch_bounds__(256) kn_morton(const float* __restrict__ xyz, size_t n, float ox, float oy, float oz, float scale,
                                                 unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float m = 2097151.f;
  const unsigned int x = (unsigned int)fminf(fmaxf((xyz[3 * i] - ox) * scale, 0.f), m);
  const unsigned int y = (unsigned int)fminf(fmaxf((xyz[3 * i + 1] - oy) * scale, 0.f), m);
  const unsigned int z = (unsigned int)fminf(fmaxf((xyz[3 * i + 2] - oz) * scale, 0.f), m);
  keys[i] = spread21(x) | (spread21(y) << 1) | (spread21(z) << 2);
  idx[i] = (unsigned int)i;
}

static __global__ void __launch_bounds__(256) kn_gather(const float* __restrict__ xyz, size_t n, const unsigned int* __restrict__ perm,
                                                 float4* __restrict__ s_xyz) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned int i = perm[j];
  s_xyz[j] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __uint_as_float(i));
}

static __global__ void __launch_bounds__(256) kn_leaf_aabb(const float4* __restrict__ s_xyz, size_t n, unsigned int nleaf, Aabb* __restrict__ nodes) {
  const unsigned int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nleaf) return;
  Aabb b; for (int d = 0; d < 3; ++d) { b.lo[d] = INFINITY; b.hi[d] = -INFINITY; }
  const size_t e = min(n, (size_t)(l + 1) * kLeaf);
  for (size_t p = (size_t)l * kLeaf; p < e; ++p) {
    const float4 v = s_xyz[p];
    b.lo[0] = fminf(b.lo[0], v.x); b.lo[1] = fminf(b.lo[1], v.y); b.lo[2] = fminf(b.lo[2], v.z);
    b.hi[0] = fmaxf(b.hi[0], v.x); b.hi[1] = fmaxf(b.hi[1], v.y); b.hi[2] = fmaxf(b.hi[2], v.z);
  }
  nodes[l] = b;
}

static __global__ void __launch_bounds__(256) kn_merge_level(const Aabb* __restrict__ child, unsigned int nchild, Aabb* __restrict__ parent,
                                                      unsigned int nparent) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nparent) return;
  Aabb b = child[2 * i];
  if (2 * i + 1 < nchild) {
    const Aabb c = child[2 * i + 1];
    for (int d = 0; d < 3; ++d) { b.lo[d] = fminf(b.lo[d], c.lo[d]); b.hi[d] = fmaxf(b.hi[d], c.hi[d]); }
  }
  parent[i] = b;
}

static constexpr int kBvhMaxLevels = 32;
struct BvhLevels { unsigned int offset[kBvhMaxLevels]; unsigned int count[kBvhMaxLevels]; int nlevels; };

__device__ __forceinline__ float dist2_pt(const float4& q, const float4& t) {
  const float dx = fsub(q.x, t.x), dy = fsub(q.y, t.y), dz = fsub(q.z, t.z);
  return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}
__device__ __forceinline__ float dist2_box(const float4& q, const Aabb& b) { return dist2_box(q.x, q.y, q.z, b); }

template <typename F>
__device__ __forceinline__ void radius_visit(const float4& q, float r2, const float4* __restrict__ s_xyz, size_t n, const Aabb* __restrict__ nodes,
                                             const BvhLevels& lv, F&& f) {
  unsigned int stack[2 * kBvhMaxLevels + 2];
  int sp = 0;
  stack[sp++] = ((unsigned int)(lv.nlevels - 1) << 27);
  while (sp > 0) {
    const unsigned int e = stack[--sp];
    const int level = (int)(e >> 27);
    const unsigned int i = e & 0x7FFFFFFu;
    if (dist2_box(q, nodes[lv.offset[level] + i]) >= r2) continue;    // bound <= every d2 inside (monotone fp32), and the test is strict
    if (level == 0) {
      const size_t b = (size_t)i * kLeaf, e2 = min(n, b + kLeaf);
      for (size_t p = b; p < e2; ++p) { const float4 t = __ldg(&s_xyz[p]); const float d = dist2_pt(q, t); if (d < r2) f(d, (unsigned int)p, __float_as_uint(t.w)); }
      continue;
    }
    const unsigned int c0 = 2 * i, c1 = 2 * i + 1;
    stack[sp++] = ((unsigned int)(level - 1) << 27) | c0;
    if (c1 < lv.count[level - 1]) stack[sp++] = ((unsigned int)(level - 1) << 27) | c1;
  }
}
static inline unsigned int bvh_div_up(size_t a, size_t b) { return (unsigned int)((a + b - 1) / b); }

// Host side: sort `n` packed float3 device points and build the node array. s_xyz[j] = (x, y, z, bits(original index)).
struct BvhIndex {
  DevBuf part, keys, keys2, idx, perm, sxyz, nodes, tmp;
  PinnedBuf ppart;
  BvhLevels lv;
  size_t n = 0;
  int build(const float* d_xyz, size_t count, int sms, cudaStream_t st) {
    n = count;
    const int bb = sms * 2;
    B2_TRY(part.ensure(sizeof(float) * 6 * bb)); B2_TRY(ppart.ensure(sizeof(float) * 6 * bb));
    kn_bbox<<<bb, 256, 0, st>>>(d_xyz, n, part.as<float>());
    B2_CUDA(cudaMemcpyAsync(ppart.p, part.p, sizeof(float) * 6 * bb, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int b = 0; b < bb; ++b) for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], ppart.as<float>()[6 * b + d]); mx[d] = std::max(mx[d], ppart.as<float>()[6 * b + 3 + d]);
    }
    for (int d = 0; d < 3; ++d) if (!std::isfinite(mn[d]) || !std::isfinite(mx[d])) return set_error(B2_ERR_ARG, "non-finite coordinates (dense clouds only)");
    const float ext = std::max({mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2], 1e-30f});
    const float scale = 2097151.f / ext;
    B2_TRY(keys.ensure(n * 8)); B2_TRY(keys2.ensure(n * 8)); B2_TRY(idx.ensure(n * 4)); B2_TRY(perm.ensure(n * 4)); B2_TRY(sxyz.ensure(n * 16));
    kn_morton<<<bvh_div_up(n, 256), 256, 0, st>>>(d_xyz, n, mn[0], mn[1], mn[2], scale, keys.as<unsigned long long>(), idx.as<unsigned int>());
    size_t t = 0;
    B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), idx.as<unsigned int>(),
                                            perm.as<unsigned int>(), (long long)n, 0, 63, st));
    B2_TRY(tmp.ensure(t));
    B2_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, t, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), idx.as<unsigned int>(),
                                            perm.as<unsigned int>(), (long long)n, 0, 63, st));
    kn_gather<<<bvh_div_up(n, 256), 256, 0, st>>>(d_xyz, n, perm.as<unsigned int>(), sxyz.as<float4>());
    std::memset(&lv, 0, sizeof(lv));
    unsigned int cnt = bvh_div_up(n, kLeaf), off = 0; int L = 0;
    while (true) { lv.offset[L] = off; lv.count[L] = cnt; off += cnt; ++L; if (cnt == 1) break; cnt = (cnt + 1) / 2; }
    lv.nlevels = L;
    B2_TRY(nodes.ensure(sizeof(Aabb) * (size_t)off));
    kn_leaf_aabb<<<bvh_div_up(lv.count[0], 256), 256, 0, st>>>(sxyz.as<float4>(), n, lv.count[0], nodes.as<Aabb>());
    for (int l = 1; l < L; ++l)
      kn_merge_level<<<bvh_div_up(lv.count[l], 256), 256, 0, st>>>(nodes.as<Aabb>() + lv.offset[l - 1], lv.count[l - 1], nodes.as<Aabb>() + lv.offset[l], lv.count[l]);
    B2_CUDA(cudaGetLastError());
    return B2_OK;
  }
  void release() { for (DevBuf* b : {&part, &keys, &keys2, &idx, &perm, &sxyz, &nodes, &tmp}) b->release(); ppart.release(); }
};

}  // namespace b2
