// Shared device/host plumbing for libeth3d_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>
#include <string>

#include "../../include/eth3d_b200.h"

namespace b2 {

// ---- error plumbing -------------------------------------------------------------------------------------------
inline std::string& last_error_ref() { static thread_local std::string e; return e; }
inline int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  last_error_ref() = buf;
  return code;
}
#define B2_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      return ::b2::set_error(B2_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,          \
                             cudaGetErrorString(_e));                                                   \
  } while (0)
#define B2_TRY(expr) do { int _rc = (expr); if (_rc != B2_OK) return _rc; } while (0)

// ---- grow-only device buffer, carved from a stream-ordered memory pool PRIVATE to this library -------------------------
// A handle owns GBs of index / direction / record buffers; with plain cudaMalloc / cudaFree a create -> add clouds -> run -> destroy
// cycle spent more time mapping and unmapping memory than computing. The library therefore allocates from its own cudaMemPool_t
// (one per device, release threshold = max, so every cycle after the first re-uses the same physical memory). The device's DEFAULT
// pool is left untouched: a host process that also uses cudaMallocAsync (PyTorch's allocator among them) keeps its own policy, and
// b2_trim() hands everything this library has cached back to the driver. Semantics stay those of cudaMalloc / cudaFree: memory
// returned by ensure() is usable on any stream at once, release() waits for the device first.
// B2_POOL=0 restores cudaMalloc / cudaFree (A/B runs, memory debugging tools).
inline bool pool_enabled() {
  static const bool on = [] { const char* e = std::getenv("B2_POOL"); return !(e && e[0] == '0'); }();
  return on;
}
struct PoolTable { cudaMemPool_t pool[64]; bool made[64]; std::mutex mu; };
inline PoolTable& pool_table() { static PoolTable t = {}; return t; }
inline cudaError_t private_pool(int dev, cudaMemPool_t* out) {
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  PoolTable& t = pool_table();
  std::lock_guard<std::mutex> lock(t.mu);
  if (!t.made[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaError_t e = cudaMemPoolCreate(&t.pool[dev], &props);
    if (e != cudaSuccess) return e;
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(t.pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    t.made[dev] = true;
  }
  *out = t.pool[dev];
  return cudaSuccess;
}
inline cudaError_t dev_alloc(void** p, size_t bytes) {
  if (!pool_enabled()) return cudaMalloc(p, bytes);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t pool;
  cudaError_t e = private_pool(dev, &pool);
  if (e != cudaSuccess) return e;
  e = cudaMallocFromPoolAsync(p, bytes, pool, cudaStreamPerThread);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
  return e;
}
inline void dev_free(void* p) {
  if (!p) return;
  if (!pool_enabled()) { cudaFree(p); return; }
  cudaDeviceSynchronize();                         // cudaFree's implicit guarantee: nothing in flight still uses the block
  cudaFreeAsync(p, cudaStreamPerThread);
}
// Returns the cached (unused) memory of every device's private pool to the driver.
inline void pinned_cache_trim();
inline void pool_trim_all() {
  pinned_cache_trim();
  PoolTable& t = pool_table();
  std::lock_guard<std::mutex> lock(t.mu);
  for (int d = 0; d < 64; ++d)
    if (t.made[d]) { cudaStreamSynchronize(cudaStreamPerThread); cudaMemPoolTrimTo(t.pool[d], 0); }
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return B2_OK;
    if (p) dev_free(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;         // slack so small growth does not re-allocate
    // large buffers in 32 MB steps: a create -> run -> destroy cycle whose counts differ a little from the previous one's (record
    // arrays follow the number of correspondences) then asks the pool for the sizes it has cached instead of for new physical memory
    constexpr size_t kStep = (size_t)32 << 20;
    if (want > kStep) want = (want + kStep - 1) / kStep * kStep;
    cudaError_t e = dev_alloc(&p, want);
    if (e != cudaSuccess) return set_error(B2_ERR_ALLOC, "device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return B2_OK;
  }
  void release() { if (p) dev_free(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Page-locked host buffers. cudaMallocHost / cudaFreeHost take 0.1 .. 10+ ms each (page locking, and they synchronise with the device);
// a handle needs eight small ones, so a create -> run -> destroy cycle returns them to a process-wide free list instead (power-of-two
// size classes from 4 KB; at most 64 cached blocks; pool_trim_all() / b2_trim() frees them).
struct PinnedCache {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_list;
};
inline PinnedCache& pinned_cache() { static PinnedCache c; return c; }
inline size_t pinned_class(size_t bytes) { size_t c = 4096; while (c < bytes) c <<= 1; return c; }
inline void pinned_cache_trim() {
  PinnedCache& c = pinned_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  for (auto& b : c.free_list) cudaFreeHost(b.first);
  c.free_list.clear();
}
struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return B2_OK;
    release();
    const size_t want = pinned_class(bytes);
    {
      PinnedCache& c = pinned_cache();
      std::lock_guard<std::mutex> lock(c.mu);
      for (size_t i = 0; i < c.free_list.size(); ++i)
        if (c.free_list[i].second == want) { p = c.free_list[i].first; cap = want; c.free_list.erase(c.free_list.begin() + i); return B2_OK; }
    }
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { p = nullptr; return set_error(B2_ERR_ALLOC, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); }
    cap = want;
    return B2_OK;
  }
  void release() {
    if (!p) return;
    PinnedCache& c = pinned_cache();
    bool kept = false;
    {
      std::lock_guard<std::mutex> lock(c.mu);
      if (c.free_list.size() < 64 && cap <= ((size_t)64 << 20)) { c.free_list.emplace_back(p, cap); kept = true; }
    }
    if (!kept) cudaFreeHost(p);
    p = nullptr; cap = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- device selection: fail loudly, no CPU fallback ------------------------------------------------------------
inline int select_device(int requested, int* out_dev, int* out_sms) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return set_error(B2_ERR_NO_DEVICE, "no CUDA device available (%s); libeth3d_b200 has no CPU fallback",
                     e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  int dev = requested;
  if (dev < 0) { B2_CUDA(cudaGetDevice(&dev)); }
  if (dev >= count) return set_error(B2_ERR_ARG, "device %d out of range (%d devices)", dev, count);
  // (single attributes: cudaGetDeviceProperties fills ~100 fields and takes milliseconds, on every handle creation)
  int major = 0, minor = 0, sms = 0;
  B2_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B2_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (major != 10)
    return set_error(B2_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
  B2_CUDA(cudaSetDevice(dev));
  *out_dev = dev;
  *out_sms = sms;
  return B2_OK;
}

// ---- fp32 arithmetic without FMA contraction (bit-parity with the CPU evaluation order) -------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
// Eigen 3-term reduction order a0 + (a1 + a2).
__device__ __forceinline__ float sum3(float a, float b, float c) { return fadd(a, fadd(b, c)); }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return sum3(fmul(ax, bx), fmul(ay, by), fmul(az, bz));
}

// Column-major 4x4 rigid transform applied the way pcl::transformPointCloudWithNormals does (PCL 1.10 SSE path):
//   point : x*c0 + (y*c1 + (z*c2 + c3))      normal: x*c0 + (y*c1 + z*c2)
struct Mat4 { float m[16]; };
__device__ __forceinline__ float3 xform_point(const Mat4& T, float x, float y, float z) {
  float3 o;
  o.x = fadd(fmul(x, T.m[0]), fadd(fmul(y, T.m[4]), fadd(fmul(z, T.m[8]), T.m[12])));
  o.y = fadd(fmul(x, T.m[1]), fadd(fmul(y, T.m[5]), fadd(fmul(z, T.m[9]), T.m[13])));
  o.z = fadd(fmul(x, T.m[2]), fadd(fmul(y, T.m[6]), fadd(fmul(z, T.m[10]), T.m[14])));
  return o;
}
__device__ __forceinline__ float3 xform_normal(const Mat4& T, float x, float y, float z) {
  float3 o;
  o.x = fadd(fmul(x, T.m[0]), fadd(fmul(y, T.m[4]), fmul(z, T.m[8])));
  o.y = fadd(fmul(x, T.m[1]), fadd(fmul(y, T.m[5]), fmul(z, T.m[9])));
  o.z = fadd(fmul(x, T.m[2]), fadd(fmul(y, T.m[6]), fmul(z, T.m[10])));
  return o;
}

// ---- uniform grid over the union bounding box ------------------------------------------------------------------
struct GridParams {
  double ox, oy, oz;   // origin (bbox min)
  double inv;          // 1 / cell size
  double cell;         // 1.0 / inv, precomputed (a double division per query otherwise)
  int nx, ny, nz;      // cells per axis (each < 2^21)
  int fbits;           // bits per axis of the in-cell Morton code appended to the cell key (5..8)
};
__host__ __device__ __forceinline__ int cell_of(float v, double o, double inv) {
  return (int)floor(((double)v - o) * inv);
}
__host__ __device__ __forceinline__ unsigned long long cell_key(const GridParams& g, int cx, int cy, int cz) {
  return ((unsigned long long)cz * (unsigned long long)g.ny + (unsigned long long)cy) * (unsigned long long)g.nx +
         (unsigned long long)cx;
}

// Sort key = (cell key << 3*fbits) | Morton code of the position inside the cell (fbits bits per axis): points of one cell
// stay contiguous (the hash table is keyed by key >> 3*fbits) and are ordered along a space-filling curve, so the per-32 /
// per-1024 point chunk boxes of dense cells are compact and prune well. fbits is chosen by compute_grid(): 5..8 bits per axis,
// as many as keep the whole key within 40 bits (= 5 radix-sort passes).
static constexpr int kMinFineBits = 5, kMaxFineBits = 8;
__host__ __device__ __forceinline__ unsigned int spread3(unsigned int v) {   // 10 bits -> every third bit
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__host__ __device__ __forceinline__ unsigned int fine_code(double rx, double ry, double rz, int fbits) {   // r* in [0,1)
  const int m = (1 << fbits) - 1;
  const double s = (double)(1 << fbits);
  const int x = min(m, max(0, (int)(rx * s))), y = min(m, max(0, (int)(ry * s))), z = min(m, max(0, (int)(rz * s)));
  return spread3((unsigned)x) | (spread3((unsigned)y) << 1) | (spread3((unsigned)z) << 2);
}

struct Aabb { float lo[3], hi[3]; };
static constexpr int kChunk1 = 32, kChunk2 = 1024;     // points per level-1 / level-2 chunk box
// Lower bound of the fp32 squared distance from q to any point in the box, evaluated with the SAME operations as the point
// distance (rounding is monotone), so pruning with `bound > best` can never drop a candidate.
__device__ __forceinline__ float dist2_box(float qx, float qy, float qz, const Aabb& b) {
  const float dx = fmaxf(fmaxf(fsub(b.lo[0], qx), fsub(qx, b.hi[0])), 0.f);
  const float dy = fmaxf(fmaxf(fsub(b.lo[1], qy), fsub(qy, b.hi[1])), 0.f);
  const float dz = fmaxf(fmaxf(fsub(b.lo[2], qz), fsub(qz, b.hi[2])), 0.f);
  return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

struct HashEntry { unsigned long long key; unsigned int begin, end; };   // 16 B: one LDG.128 per probe
static constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
__host__ __device__ __forceinline__ unsigned int hash_slot(unsigned long long key, int log2size) {
  return (unsigned int)((key * 0x9E3779B97F4A7C15ull) >> (64 - log2size));
}

}  // namespace b2
