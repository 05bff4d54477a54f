// Host side of Path A behind the C ABI (include/eth3d_b200.h): owns device memory, builds the static per-cloud index,
// schedules the pair-directions, runs the LM loop of PointToPlaneICPImpl::compute on top of the streaming kernels.
//
// Reference seams replaced (see the header for the signatures):
//   PointToPlaneICP::AddPointCloud  /root/reference/src/icp/icp_point_to_plane.cc:109-135
//   PointToPlaneICP::Run            :137-163        AlignMeshes :169-342
//   PointToPlaneICPImpl::compute    /root/reference/src/icp/icp_point_to_plane_impl.h:115-293
//
// Index design. The reference rebuilds a kd-tree over the transformed target for every pair-direction of every outer iteration
// (:46-51). Here a cloud is indexed ONCE: cell-sorted copies of its points and a hash table (+ bitmap) of the occupied cells of a
// uniform grid that is attached to the cloud. The grid is laid out in the cloud's "index frame" = the global frame as it was when
// the cloud was indexed (x' = F l, F = the pose at that time, frozen), with its origin on the lattice of multiples of the cell
// size — so the grids of all clouds of a handle coincide when they are built and drift apart only by what ICP moves the poses
// (millimetres), and the cell-sorted queries of one cloud walk the cells of another in order. A pose update moves the grid rigidly
// with its cloud, so nothing is re-sorted or re-hashed; per outer iteration a cloud costs one streaming pass (K1x: global-frame
// copies, chunk boxes, AABB). A query reaches a target's grid through F T^-1; candidate distances are evaluated on the global-frame
// fp32 coordinates exactly as the reference does, so the results do not depend on the frame the lookup ran in (see the error
// budget at grid_for_cloud).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "b2_common.cuh"
#include "b2_hostmath.h"
#include "b2_icp_kernels.cuh"

namespace b2 {

struct Cloud {
  size_t n = 0;
  bool is_fixed = false;
  DevBuf local_xyz, local_nrm;          // packed float3, caller's order (cloud frame; global frame for the fixed cloud)
  float T[16];                          // global_T_cloud, column-major
  float bmin[3], bmax[3];               // AABB of the global-frame cloud (this outer iteration)
  // ---- static index in the cloud frame (valid for index_d / index_sigma / index_mtot) ----
  bool have_lbox = false, indexed = false;
  float lmin[3], lmax[3];               // AABB in the cloud frame
  double F[12];                         // index frame (row-major 3x4): the pose at index time, frozen; grid coordinates = F l
  float index_d = 0.f;
  double index_sigma = 0, index_mtot = 0, margin = 0;
  GridParams g;
  DevBuf l_xyz, l_nrm, perm_inv;        // cell-sorted cloud-frame rows (float4), original index -> sorted position
  bool dense = false;                   // layout of the occupied-cell index (k_nn_tiles<., DENSE>)
  DevBuf rb, starts, table;             // DENSE: rank bitmap + cell starts; sparse: hash table of the occupied cells
  DevBuf corner;                        // DENSE, lattices up to 2^29 corners: the corner map of the DUAL lookup (1 B per corner)
  bool dual = false;
  int log2size = 0;
  unsigned int ncells = 0;
  // ---- per outer iteration ----
  DevBuf s_xyz, s_nrm, box1, box2;      // global-frame rows in the sorted order + chunk boxes
  unsigned int index_epoch = 0;         // bumped by every index build (the sorted order of the rows changes with it)
  bool rows_ahead = false;              // search_ahead: s_xyz / boxes hold the rows at pose rows_T of index build rows_epoch
  float rows_T[16];
  unsigned int rows_epoch = 0;
  cudaEvent_t ready_ev = nullptr;       // sharded upload: the cloud's bytes have arrived on this rank (recorded on the broadcast stream)
};

struct Direction {
  int src = 0, tgt = 0;                 // impl cloud indices
  bool local = false;                   // searched / accumulated on this rank
  DevBuf key, tile_count, tile_off, order;   // per sorted query: (d2, index) key; per tile: matches, exclusive offsets; launch order
  int order_src = -1, order_tgt = -1, order_age = 0;
  unsigned int order_tiles = 0;
  unsigned long long count = 0, rec_begin = 0;
  int ahead_slot = -1;                  // >= 0: this iteration's search was done ahead of b2_icp_run (slot of its match count)
};

// A pair-direction searched ahead of b2_icp_run (cfg.search_ahead): its result is adopted by the first outer iteration if, by then,
// neither pose nor either index has changed and the radius is the hinted one; otherwise it is dropped and the search runs as usual.
struct Ahead {
  Cloud* S = nullptr; Cloud* T = nullptr;
  float Ts[16], Tt[16];
  unsigned int s_epoch = 0, t_epoch = 0;
  float r2 = 0.f;
  int slot = 0;
  Direction d;
};

// Temporaries of a call, released on every exit path. With a stream: a stream-ORDERED free (every use of the buffer was queued on that
// stream), which does not wait for the device — a device-wide wait here would also wait for a peer's upload that is still being broadcast
// into another cloud (sharded uploads), serialising what the broadcast stream exists to overlap.
struct Scoped {
  DevBuf b;
  cudaStream_t stream = nullptr;
  Scoped() {}
  explicit Scoped(cudaStream_t s) : stream(s) {}
  ~Scoped() {
    if (b.p && stream && pool_enabled()) { cudaFreeAsync(b.p, stream); b.p = nullptr; b.cap = 0; }
    else b.release();
  }
};

static inline Mat4 mat4_of(const float T[16]) { Mat4 m; std::memcpy(m.m, T, sizeof(m.m)); return m; }
static inline unsigned int div_up(size_t a, size_t b) { return (unsigned int)((a + b - 1) / b); }

}  // namespace b2

using namespace b2;

struct b2_icp {
  b2_icp_config cfg;
  int device = 0, sms = 148;
  bool lpt_order = true;                // K3 tiles issued longest-first (B2_K3_ORDER=grid disables, for A/B runs)
  bool work_stats = false;              // B2_K3_WORK=1: the diagnostic K3 variant that counts its work
  cudaStream_t stream = nullptr, copy_stream = nullptr;   // copy_stream: uploads of b2_icp_add_cloud (so that an index build can run beside them)
  cudaStream_t bcast_stream = nullptr;  // sharded uploads: the NCCL broadcasts, ordered behind the owner's copy by an event
  bool own_stream = false;
  std::vector<b2::Cloud*> pending_index;   // index_distance_hint: clouds whose index is built behind the following uploads
  // search_ahead: pair-directions among the clouds indexed so far, searched on the auxiliary streams behind the following uploads
  static constexpr int kMaxAhead = 4096;
  std::vector<std::unique_ptr<b2::Ahead>> ahead;
  std::vector<b2::Cloud*> ahead_clouds;
  DevBuf ahead_totals_dev, ahead_bbox;
  PinnedBuf ahead_totals_pin;
  cudaEvent_t ahead_ev = nullptr;
  int ahead_rr = 0;
  // K4 overlapped with K3: a set is packed on pack_stream as soon as its search has finished, at a device-side running offset
  // Off by default: measured on config 2 (r02D) the step took 51.35 ms with it and 51.49 ms without — K3 needs its occupancy and K4
  // its threads in flight, they take SM slots from each other one for one — and sizing the record arrays for the worst case made
  // create / destroy of a handle slower. Kept as an option (B2_PACK=overlap | overlap_nosize, b2_icp_set_option "pack_overlap").
  bool pack_overlap = false;
  bool pack_presize = true;             // overlap_nosize: no worst-case sizing of the record arrays in the first iteration (tests)
  cudaStream_t pack_stream = nullptr;
  cudaEvent_t pack_join_ev = nullptr;
  std::vector<cudaEvent_t> done_ev;     // search of the i-th issued set finished
  DevBuf chain_dev;                     // [sets + 1] running offsets + overflow flag
  PinnedBuf pin_chain;
  // K3 runs the pair-directions of an iteration round-robin over `nsearch` streams (the handle's + auxiliaries) so that one
  // direction's tail overlaps the next direction's head; each stream has its own CUB scratch.
  static constexpr int kMaxSearchStreams = 8;
  int nsearch = 4;
  cudaStream_t aux[kMaxSearchStreams - 1] = {};
  cudaEvent_t fork_ev = nullptr, join_ev[kMaxSearchStreams - 1] = {};
  DevBuf search_tmp[kMaxSearchStreams];
  std::vector<std::unique_ptr<Cloud>> movable;
  std::unique_ptr<Cloud> fixed;         // concatenated global-frame fixed cloud (may be null)
  std::vector<std::unique_ptr<Direction>> dirs;   // pool, reused across outer iterations
  int ndirs = 0;
  // scratch
  DevBuf bbox_partial, cub_tmp, cell_counts, rec_a, rec_b, rec_c, segs_dev, poses_dev, partials, xpartials, segsum, xsegsum, eq_dev, scatter_m, scatter_d, work_dev, totals_dev;
  PinnedBuf pin_bbox, pin_counts, pin_eq, pin_poses, pin_segs, pin_misc;
  unsigned int tma_attr_mask = 0;       // K5 instantiations whose dynamic shared memory opt-in has been set on THIS handle's device
  // last outer iteration
  b2_icp_stats stats;
  std::vector<int32_t> tries;
  std::vector<double> H0, b0;
  double cost0 = 0;
  std::vector<Segment> segs_host;       // local non-empty segments
  std::vector<int> seg_dir;             // segs_host[k] -> dirs index
  unsigned long long total_records = 0;
  int grid_acc = 0;
  unsigned long long per_cta = 0;
  int launches = 0;
  float ms_index_build = 0.f;
  unsigned long long last_init_key = 0; // "no match" key of the last search (r2 bits << 32)
  int prev_inner_iterations = 0;        // LM iterations of the previous outer iteration (0 = none yet): gates the speculative second try
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> acc_events, nn_events;
};

namespace b2 {

static int num_impl_clouds(const b2_icp* h) { return (int)h->movable.size() + (h->fixed ? 1 : 0); }
static Cloud* impl_cloud(b2_icp* h, int idx) {
  if (h->fixed) return idx == 0 ? h->fixed.get() : h->movable[idx - 1].get();
  return h->movable[idx].get();
}
static int impl_index_of_movable(const b2_icp* h, int m) { return h->fixed ? m + 1 : m; }

static int build_pending_index(b2_icp* h);
// owner >= 0: sharded upload (cfg.shard_uploads) — only rank `owner` reads the caller's buffers, everyone else receives the cloud
// by ncclBroadcast on the copy stream.
static int upload_cloud(b2_icp* h, Cloud* c, const float* xyz, const float* nrm, size_t n, size_t stride_bytes, bool from_device, int owner = -1) {
  c->n = n;
  c->have_lbox = c->indexed = false;
  B2_TRY(c->local_xyz.ensure(std::max<size_t>(n, 1) * 12));
  B2_TRY(c->local_nrm.ensure(std::max<size_t>(n, 1) * 12));
  if (n) {
    cudaStream_t cs = h->copy_stream;
    if (owner >= 0 && owner != h->cfg.rank) {
      // nothing to read on this rank
    } else if (from_device) {
      B2_CUDA(cudaMemcpyAsync(c->local_xyz.p, xyz, n * 12, cudaMemcpyDeviceToDevice, cs));
      B2_CUDA(cudaMemcpyAsync(c->local_nrm.p, nrm, n * 12, cudaMemcpyDeviceToDevice, cs));
    } else if (stride_bytes == 12) {
      B2_CUDA(cudaMemcpyAsync(c->local_xyz.p, xyz, n * 12, cudaMemcpyHostToDevice, cs));
      B2_CUDA(cudaMemcpyAsync(c->local_nrm.p, nrm, n * 12, cudaMemcpyHostToDevice, cs));
    } else {
      B2_CUDA(cudaMemcpy2DAsync(c->local_xyz.p, 12, xyz, stride_bytes, 12, n, cudaMemcpyHostToDevice, cs));
      B2_CUDA(cudaMemcpy2DAsync(c->local_nrm.p, 12, nrm, stride_bytes, 12, n, cudaMemcpyHostToDevice, cs));
    }
    if (owner >= 0) {
      // The broadcasts run on their own stream: behind this cloud's copy on the owner, behind nothing on the other ranks — whose
      // call returns at once, so that THEIR next upload (a cloud they own) crosses PCIe at the same time as this one.
      if (!c->ready_ev) B2_CUDA(cudaEventCreateWithFlags(&c->ready_ev, cudaEventDisableTiming));
      if (owner == h->cfg.rank) { B2_CUDA(cudaEventRecord(c->ready_ev, cs)); B2_CUDA(cudaStreamWaitEvent(h->bcast_stream, c->ready_ev, 0)); }
      B2_TRY(b2_comm_broadcast(h->cfg.comm, c->local_xyz.p, n * 12, owner, (void*)h->bcast_stream));
      B2_TRY(b2_comm_broadcast(h->cfg.comm, c->local_nrm.p, n * 12, owner, (void*)h->bcast_stream));
      B2_CUDA(cudaEventRecord(c->ready_ev, h->bcast_stream));
    }
  }
  // while this cloud's bytes are on the wire: the search index of the previously added cloud (index_distance_hint)
  const int rc = build_pending_index(h);
  if (n && (owner < 0 || owner == h->cfg.rank)) B2_CUDA(cudaStreamSynchronize(h->copy_stream));   // the caller may free / reuse its buffers after return
  return rc;
}

// ---- the linear part of a pose -------------------------------------------------------------------------------------
// sigma >= || B ||_2 for the linear part B of the map from the global frame into a cloud's index frame: index-frame distances are at
// most sigma times the global ones (and global ones at least 1 / sigma times the index-frame ones). While poses stay rigid B is a
// rotation and sigma = 1 + O(1e-7); anything invertible is accepted (a pose that starts to scale or shear only enlarges the cells).
static bool invert3(const double A[9], double Ai[9]) {      // row-major
  const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
  if (!(std::fabs(det) > 1e-300) || !std::isfinite(det)) return false;
  const double id = 1.0 / det;
  Ai[0] = c0 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c1 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c2 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}
static void linear_part(const float T[16], double A[9]) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[3 * r + c] = T[r + 4 * c]; }
// B = F_lin A^-1 maps global offsets to index-frame offsets (F_lin = linear part of the frozen index frame, nullptr = A itself, i.e.
// the index is about to be built at this pose and B = I up to rounding).
static int pose_sigma(const float T[16], const double* F, double* sigma) {
  double A[9], Ai[9], B[9];
  linear_part(T, A);
  if (!invert3(A, Ai)) return set_error(B2_ERR_ARG, "pose is not invertible");
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      double v = 0;
      for (int m = 0; m < 3; ++m) v += (F ? F[4 * r + m] : A[3 * r + m]) * Ai[3 * m + k];
      B[3 * r + k] = v;
    }
  double eta = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double e = -(i == j ? 1.0 : 0.0);
      for (int k = 0; k < 3; ++k) e += B[3 * k + i] * B[3 * k + j];
      eta += e * e;
    }
  eta = std::sqrt(eta);
  if (eta < 0.25) *sigma = std::sqrt(1.0 + eta);                 // sigma_max(B)^2 <= 1 + ||B^T B - I||
  else { double f = 0; for (double v : B) f += v * v; *sigma = std::sqrt(f); }   // || . ||_2 <= || . ||_F
  return B2_OK;
}
// Largest coordinate magnitude anything on the lookup path can take: global coordinates of every cloud (bounded from its pose
// and its cloud-frame AABB) and the cloud-frame coordinates themselves.
static double magnitude_bound(b2_icp* h) {
  double m = 0;
  const int nc = num_impl_clouds(h);
  for (int i = 0; i < nc; ++i) {
    const Cloud* c = impl_cloud(h, i);
    if (c->n == 0) continue;
    double la[3];
    for (int k = 0; k < 3; ++k) { la[k] = std::max(std::fabs((double)c->lmin[k]), std::fabs((double)c->lmax[k])); m = std::max(m, la[k]); }
    for (int a = 0; a < 3; ++a) {
      double gsum = std::fabs((double)c->T[12 + a]);
      for (int k = 0; k < 3; ++k) gsum += std::fabs((double)c->T[a + 4 * k]) * la[k];
      m = std::max(m, gsum);
    }
  }
  return m;
}

// Grid of one cloud in its index frame (the global frame as it was when the cloud was indexed).
// Error budget of the lookup (why every target with fp32 d2 < r2 lies in the 2x2x2 block that k_nn_tiles searches):
//   * d2 < r2 in the reference's fp32 arithmetic  =>  true distance of the fp32 global coordinates D <= d (1 + 2^-21);
//   * the target's global row is fl(T l): |rounding| <= 3 * 2^-24 * M per coordinate (M = magnitude_bound);
//   * the query is mapped by fl32(F T^-1) with three FMAs: <= 7 * 2^-24 * M per coordinate, and the cell fraction is formed in
//     fp32: <= 4 * 2^-24 * M;
//   * index-frame distance <= sigma * global distance (pose_sigma).
// Hence |index-frame offset per axis| <= d sigma + 32 * 2^-24 * M. margin = 64 * 2^-24 * M, the index is built for 2 M and
// sigma (1 + 1e-4) and rebuilt should an iteration exceed either. cell = 2 (d sigma + margin) (1 + 1e-4): from a query's
// (computed) half of its cell, the far faces of the 2x2x2 block on that side are >= cell / 2 away.
static int grid_for_cloud(Cloud* c, const float fmin[3], const float fmax[3], float max_dist, double sigma, double mtot, int* key_bits) {
  for (int d = 0; d < 3; ++d)
    if (!std::isfinite(fmin[d]) || !std::isfinite(fmax[d]))
      return set_error(B2_ERR_ARG, "non-finite point coordinates (clouds must be dense, as the reference's is_dense path assumes)");
  c->index_sigma = sigma * (1.0 + 1e-4);
  // magnitude classes are powers of two, so that clouds indexed at different times (index_distance_hint) still get the same margin,
  // hence the same cell size and the same lattice
  c->index_mtot = std::ldexp(1.0, (int)std::ceil(std::log2(std::max(2.0 * mtot, 1e-30))));
  c->margin = 64.0 * std::ldexp(1.0, -24) * c->index_mtot;
  double cell = 2.0 * ((double)max_dist * c->index_sigma + c->margin) * 1.0001;
  if (!(cell > 0.0)) cell = 1e-30;
  GridParams& g = c->g;
  // the box was measured on fp32 positions, the keys use double ones: pad by the margin (it bounds that difference many times over)
  const double lo[3] = {fmin[0] - c->margin, fmin[1] - c->margin, fmin[2] - c->margin};
  const double hi[3] = {fmax[0] + c->margin, fmax[1] + c->margin, fmax[2] + c->margin};
  const double ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]});
  cell = std::max(cell, ext / 2096000.0);    // keep every axis below 2^21 cells
  auto dims = [&] {
    g.inv = 1.0 / cell;
    // origin on the lattice of multiples of the cell size: clouds indexed with the same radius and magnitude bound share one lattice
    g.ox = std::floor(lo[0] / cell) * cell; g.oy = std::floor(lo[1] / cell) * cell; g.oz = std::floor(lo[2] / cell) * cell;
    g.nx = (int)std::floor((hi[0] - g.ox) * g.inv) + 2; g.ny = (int)std::floor((hi[1] - g.oy) * g.inv) + 2; g.nz = (int)std::floor((hi[2] - g.oz) * g.inv) + 2;
  };
  dims();
  // keep the cell key below 2^47 so that (key << 3*kMinFineBits) fits 63 bits: enlarge the cells of enormous sparse clouds
  while ((double)g.nx * (double)g.ny * (double)g.nz >= 140737488355328.0) { cell *= 2.0; dims(); }
  g.cell = 1.0 / g.inv;
  const unsigned long long maxkey = cell_key(g, g.nx - 1, g.ny - 1, g.nz - 1);
  int bits = 1;
  while (bits < 64 && (maxkey >> bits) != 0ull) ++bits;
  // in-cell Morton resolution: as fine as fits 42 key bits, never below 5 bits per axis (dense cells are pruned through chunk boxes
  // over this order; the sort runs once per cloud, so a sixth radix pass is affordable)
  g.fbits = std::max(kMinFineBits, std::min(kMaxFineBits, (42 - bits) / 3));
  *key_bits = bits + 3 * g.fbits;
  return B2_OK;
}

// One-time index of a cloud (K1, K2): AABB in the cloud frame -> grid -> keys -> radix sort -> cell-sorted rows, inverse
// permutation, occupied-cell table. The sort is the only library call (cub::DeviceRadixSort) and it is off the per-iteration path.
static int cloud_local_box(b2_icp* h, Cloud* c) {
  if (c->ready_ev) {       // sharded upload: everything that touches the cloud comes after this point
    B2_CUDA(cudaEventSynchronize(c->ready_ev));
    cudaEventDestroy(c->ready_ev); c->ready_ev = nullptr;
  }
  if (c->have_lbox) return B2_OK;
  for (int d = 0; d < 3; ++d) { c->lmin[d] = INFINITY; c->lmax[d] = -INFINITY; }
  if (c->n) {
    const int blocks = h->sms * 2;
    B2_TRY(h->bbox_partial.ensure((size_t)blocks * 6 * sizeof(float)));
    B2_TRY(h->pin_bbox.ensure((size_t)blocks * 6 * sizeof(float)));
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    k_bbox<<<blocks, 256, 0, h->stream>>>(c->local_xyz.as<float>(), c->n, mat4_of(I), h->bbox_partial.as<float>());
    ++h->launches;
    B2_CUDA(cudaMemcpyAsync(h->pin_bbox.p, h->bbox_partial.p, (size_t)blocks * 6 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    const float* p = h->pin_bbox.as<float>();
    for (int b = 0; b < blocks; ++b)
      for (int d = 0; d < 3; ++d) { c->lmin[d] = std::min(c->lmin[d], p[b * 6 + d]); c->lmax[d] = std::max(c->lmax[d], p[b * 6 + 3 + d]); }
  } else {
    for (int d = 0; d < 3; ++d) c->lmin[d] = c->lmax[d] = 0.f;
  }
  c->have_lbox = true;
  return B2_OK;
}

static int build_index(b2_icp* h, Cloud* c, float max_dist, double sigma, double mtot) {
  const size_t n = c->n;
  c->indexed = false;
  ++c->index_epoch;
  if (c->rows_ahead) {     // searches issued ahead of b2_icp_run may still be reading the index this call rewrites
    for (int i = 0; i + 1 < h->nsearch; ++i) B2_CUDA(cudaStreamSynchronize(h->aux[i]));
    c->rows_ahead = false;
  }
  if (n == 0) { c->ncells = 0; c->index_d = max_dist; c->index_sigma = sigma * (1.0 + 1e-4); c->index_mtot = 4.0 * mtot; c->indexed = true; return B2_OK; }
  // freeze the index frame at the current pose and measure the cloud's box in it
  for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) c->F[4 * r + k] = c->T[r + 4 * k]; c->F[4 * r + 3] = c->T[12 + r]; }
  float fmin[3] = {INFINITY, INFINITY, INFINITY}, fmax[3] = {-INFINITY, -INFINITY, -INFINITY};
  {
    const int blocks = h->sms * 2;
    B2_TRY(h->bbox_partial.ensure((size_t)blocks * 6 * sizeof(float)));
    B2_TRY(h->pin_bbox.ensure((size_t)blocks * 6 * sizeof(float)));
    k_bbox<<<blocks, 256, 0, h->stream>>>(c->local_xyz.as<float>(), n, mat4_of(c->T), h->bbox_partial.as<float>());
    ++h->launches;
    B2_CUDA(cudaMemcpyAsync(h->pin_bbox.p, h->bbox_partial.p, (size_t)blocks * 6 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    const float* p = h->pin_bbox.as<float>();
    for (int b = 0; b < blocks; ++b)
      for (int d = 0; d < 3; ++d) { fmin[d] = std::min(fmin[d], p[b * 6 + d]); fmax[d] = std::max(fmax[d], p[b * 6 + 3 + d]); }
  }
  int key_bits = 0;
  B2_TRY(grid_for_cloud(c, fmin, fmax, max_dist, sigma, mtot, &key_bits));
  const GridParams& g = c->g;
  IndexFrame F; std::memcpy(F.f, c->F, sizeof(F.f));
  Scoped keys_in(h->stream), keys_out(h->stream), idx_in(h->stream), perm(h->stream);
  B2_TRY(keys_in.b.ensure(n * 8 + 16)); B2_TRY(keys_out.b.ensure(n * 8 + 16)); B2_TRY(idx_in.b.ensure(n * 4)); B2_TRY(perm.b.ensure(n * 4));
  B2_TRY(c->l_xyz.ensure(n * 16)); B2_TRY(c->l_nrm.ensure(n * 16)); B2_TRY(c->perm_inv.ensure(n * 4));
  B2_TRY(c->s_xyz.ensure(n * 16)); B2_TRY(c->s_nrm.ensure(n * 16));
  k_keys<<<div_up(n, 256), 256, 0, h->stream>>>(c->local_xyz.as<float>(), n, F, g, keys_in.b.as<unsigned long long>(), idx_in.b.as<unsigned int>());
  size_t tmp = 0;
  B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys_in.b.as<unsigned long long>(), keys_out.b.as<unsigned long long>(),
                                          idx_in.b.as<unsigned int>(), perm.b.as<unsigned int>(), (long long)n, 0, key_bits, h->stream));
  B2_TRY(h->cub_tmp.ensure(tmp));
  B2_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, tmp, keys_in.b.as<unsigned long long>(), keys_out.b.as<unsigned long long>(),
                                          idx_in.b.as<unsigned int>(), perm.b.as<unsigned int>(), (long long)n, 0, key_bits, h->stream));
  k_gather_sorted<<<div_up(n, 256), 256, 0, h->stream>>>(c->local_xyz.as<float>(), c->local_nrm.as<float>(), n, perm.b.as<unsigned int>(),
                                                         c->l_xyz.as<float4>(), c->l_nrm.as<float4>(), c->perm_inv.as<unsigned int>());
  B2_TRY(h->cell_counts.ensure(sizeof(unsigned int)));
  B2_TRY(h->pin_counts.ensure(sizeof(unsigned long long) * 8));
  B2_CUDA(cudaMemsetAsync(h->cell_counts.p, 0, sizeof(unsigned int), h->stream));
  k_count_cells<<<div_up(n, 256), 256, 0, h->stream>>>(keys_out.b.as<unsigned long long>(), n, h->cell_counts.as<unsigned int>(), 3 * g.fbits);
  B2_CUDA(cudaMemcpyAsync(h->pin_counts.p, h->cell_counts.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  c->ncells = h->pin_counts.as<unsigned int>()[0];
  // layout of the occupied-cell index: rank bitmap while the grid has at most 2^31 cells (8 B per 32 cells: 512 MB at the limit,
  // 7.5 MB for a 10 x 8 x 3 m room at 2 cm), hash table of the occupied cells beyond
  const double cells_total = (double)g.nx * (double)g.ny * (double)g.nz;
  c->dense = cells_total <= 2147483648.0;
  c->dual = false;
  if (c->dense) {
    const size_t nwords = (size_t)(cells_total / 32.0) + 2;
    const unsigned int nblocks = div_up(nwords, kWordsPerBlock);
    Scoped block_sum(h->stream), block_off(h->stream);
    B2_TRY(c->rb.ensure(nwords * 8)); B2_TRY(c->starts.ensure(((size_t)c->ncells + 2) * 4));
    B2_TRY(block_sum.b.ensure((size_t)nblocks * 4)); B2_TRY(block_off.b.ensure((size_t)nblocks * 4 + 4));
    B2_CUDA(cudaMemsetAsync(c->rb.p, 0, nwords * 8, h->stream));
    k_mark_cells<<<div_up(n, 256), 256, 0, h->stream>>>(keys_out.b.as<unsigned long long>(), n, 3 * g.fbits, c->rb.as<uint2>());
    k_word_counts<<<nblocks, 256, 0, h->stream>>>(c->rb.as<uint2>(), nwords, block_sum.b.as<unsigned int>());
    k_scan_tiles<<<1, 1024, 0, h->stream>>>(block_sum.b.as<unsigned int>(), nblocks, block_off.b.as<unsigned int>(), block_off.b.as<unsigned int>() + nblocks);
    k_word_prefix<<<nblocks, 256, 0, h->stream>>>(c->rb.as<uint2>(), nwords, block_off.b.as<unsigned int>());
    k_cell_starts<<<div_up(n, 256), 256, 0, h->stream>>>(keys_out.b.as<unsigned long long>(), n, 3 * g.fbits, c->rb.as<uint2>(),
                                                         c->starts.as<unsigned int>(), c->ncells);
    h->launches += 5;
    // corner map (30 MB for a 10 x 8 x 3 m room at 2 cm; skipped for lattices above 2^29 corners, the search then tests the bitmap)
    const double corners = ((double)g.nx + 1.0) * ((double)g.ny + 1.0) * ((double)g.nz + 1.0);
    static const bool dual_off = [] { const char* e = getenv("B2_K3_DUAL"); return e && e[0] == '0'; }();
    c->dual = !dual_off && corners <= 536870912.0;
    if (c->dual) {
      B2_TRY(c->corner.ensure((size_t)corners));
      k_corner_occupancy<<<div_up((size_t)corners, 256), 256, 0, h->stream>>>(c->rb.as<uint2>(), g.nx, g.ny, g.nz, c->corner.as<unsigned char>());
      ++h->launches;
    }
  } else {
    int lg = 4;
    while ((1ull << lg) < 2ull * c->ncells) ++lg;
    c->log2size = lg;
    B2_TRY(c->table.ensure(sizeof(HashEntry) << lg));
    B2_CUDA(cudaMemsetAsync(c->table.p, 0xFF, sizeof(HashEntry) << lg, h->stream));
    k_hash_cells<<<div_up(n, 256), 256, 0, h->stream>>>(keys_out.b.as<unsigned long long>(), n, c->table.as<HashEntry>(), lg, 3 * g.fbits);
    ++h->launches;
  }
  h->launches += 4;
  c->index_d = max_dist;
  c->indexed = true;
  return B2_OK;
}

// This iteration's map into a target's grid: position relative to the grid origin = F T^-1 (q - t) - o, as one fp32 3x4.
static int search_grid(const Cloud* c, SearchGrid* sg) {
  double A[9], Ai[9];
  linear_part(c->T, A);
  if (!invert3(A, Ai)) return set_error(B2_ERR_ARG, "pose is not invertible");
  const double t[3] = {c->T[12], c->T[13], c->T[14]}, o[3] = {c->g.ox, c->g.oy, c->g.oz};
  for (int r = 0; r < 3; ++r) {
    double B[3];
    for (int k = 0; k < 3; ++k) { B[k] = c->F[4 * r] * Ai[k] + c->F[4 * r + 1] * Ai[3 + k] + c->F[4 * r + 2] * Ai[6 + k]; sg->m[4 * r + k] = (float)B[k]; }
    sg->m[4 * r + 3] = (float)(c->F[4 * r + 3] - (B[0] * t[0] + B[1] * t[1] + B[2] * t[2]) - o[r]);
  }
  sg->inv = (float)c->g.inv; sg->cell = (float)c->g.cell;
  sg->inv_sigma = (float)((1.0 - 1e-6) / c->index_sigma);
  sg->margin = (float)c->margin;
  sg->nx = c->g.nx; sg->ny = c->g.ny; sg->nz = c->g.nz;
  sg->sy = c->g.nx; sg->sz = (long long)c->g.nx * (long long)c->g.ny;
  sg->log2size = c->log2size;
  sg->one = 1.0f;
  sg->occ = nullptr;
  sg->corner = (c->dense && c->dual) ? c->corner.as<unsigned char>() : nullptr;
  sg->rb = c->dense ? c->rb.as<uint2>() : nullptr;
  sg->starts = c->dense ? c->starts.as<unsigned int>() : nullptr;
  return B2_OK;
}

static bool boxes_intersect(const Cloud* a, const Cloud* b) {
  // Eigen::AlignedBox::intersection(...).isEmpty(): empty iff (min > max) on any axis (icp_point_to_plane.cc:214-215).
  for (int d = 0; d < 3; ++d) if (std::max(a->bmin[d], b->bmin[d]) > std::min(a->bmax[d], b->bmax[d])) return false;
  return true;
}

static Direction* next_direction(b2_icp* h) {
  if (h->ndirs == (int)h->dirs.size()) h->dirs.emplace_back(new Direction());
  return h->dirs[h->ndirs++].get();
}

// One streaming pass (K5 + K6) at the given increments; returns [H|b|cost|extras|costs of the speculative trials] in h->pin_eq (host)
// after the optional cross-rank reduction. trials[0] is the state the normal equations (with_h) and the cost are evaluated at;
// trials[1..] (at most kMaxExtraTrials) are further LM trial states whose costs ride along on the same read of the records.
template <bool WITH_H, int NX>
static int launch_accumulate_tma(b2_icp* h, int nseg, int nc) {
  // the > 48 KB dynamic shared memory opt-in is per device: tracked per handle (a handle is bound to one device)
  const unsigned int bit = 1u << ((WITH_H ? 4 : 0) + NX);
  if (!(h->tma_attr_mask & bit)) {
    B2_CUDA(cudaFuncSetAttribute(k_accumulate_tma<WITH_H, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem_bytes(WITH_H)));
    h->tma_attr_mask |= bit;
  }
  k_accumulate_tma<WITH_H, NX><<<h->grid_acc, kAccThreads, tma_smem_bytes(WITH_H), h->stream>>>(
      h->rec_a.as<float4>(), h->rec_b.as<float4>(), h->rec_c.as<float4>(), h->segs_dev.as<Segment>(), nseg, h->poses_dev.as<CloudPose>(), nc,
      h->total_records, h->per_cta, h->partials.as<double>(), h->xpartials.as<double>(), 1.0f, -0.0f);
  return B2_OK;
}

static int run_pass(b2_icp* h, const std::vector<std::vector<Pose>>& trials, bool with_h, int nv) {
  const int nc = (int)trials[0].size();
  const int nx = (int)trials.size() - 1;
  if (nx < 0 || nx > kMaxExtraTrials) return set_error(B2_ERR_ARG, "run_pass: bad trial count");
  CloudPose* pp = h->pin_poses.as<CloudPose>();
  for (int j = 0; j <= nx; ++j)
    for (int i = 0; i < nc; ++i) {
      CloudPose& o = pp[(size_t)j * nc + i];
      quat_matrix(trials[j][i].q, o.R); for (int k = 0; k < 3; ++k) o.t[k] = trials[j][i].t[k];
    }
  B2_CUDA(cudaMemcpyAsync(h->poses_dev.p, pp, sizeof(CloudPose) * nc * (1 + nx), cudaMemcpyHostToDevice, h->stream));
  const int nseg = (int)h->segs_host.size();
  const size_t eq_count = (size_t)nv * nv + nv + 3 + kMaxExtraTrials;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->total_records > 0) {
    B2_CUDA(cudaEventCreate(&e0)); B2_CUDA(cudaEventCreate(&e1));
    B2_CUDA(cudaEventRecord(e0, h->stream));
    static const bool use_tma = [] { const char* e = getenv("B2_K5"); return !(e && std::string(e) == "ldg"); }();
    if (use_tma) {
      // Blackwell path: record tiles staged by cp.async.bulk + mbarrier (bit-identical results, see k_accumulate_tma)
      switch ((with_h ? 4 : 0) + nx) {
        case 0: B2_TRY((launch_accumulate_tma<false, 0>(h, nseg, nc))); break;
        case 1: B2_TRY((launch_accumulate_tma<false, 1>(h, nseg, nc))); break;
        case 2: B2_TRY((launch_accumulate_tma<false, 2>(h, nseg, nc))); break;
        case 3: B2_TRY((launch_accumulate_tma<false, 3>(h, nseg, nc))); break;
        case 4: B2_TRY((launch_accumulate_tma<true, 0>(h, nseg, nc))); break;
        case 5: B2_TRY((launch_accumulate_tma<true, 1>(h, nseg, nc))); break;
        case 6: B2_TRY((launch_accumulate_tma<true, 2>(h, nseg, nc))); break;
        default: B2_TRY((launch_accumulate_tma<true, 3>(h, nseg, nc))); break;
      }
    } else {
      if (nx != 0) return set_error(B2_ERR_STATE, "run_pass: the register-staged K5 has no speculative trials");
      if (with_h)
        k_accumulate<true><<<h->grid_acc, kAccThreads, 0, h->stream>>>(h->rec_a.as<float4>(), h->rec_b.as<float4>(), h->rec_c.as<float4>(),
                                                                      h->segs_dev.as<Segment>(), nseg, h->poses_dev.as<CloudPose>(),
                                                                      h->total_records, h->per_cta, h->partials.as<double>());
      else
        k_accumulate<false><<<h->grid_acc, kAccThreads, 0, h->stream>>>(h->rec_a.as<float4>(), h->rec_b.as<float4>(), h->rec_c.as<float4>(),
                                                                       h->segs_dev.as<Segment>(), nseg, h->poses_dev.as<CloudPose>(),
                                                                       h->total_records, h->per_cta, h->partials.as<double>());
    }
    B2_CUDA(cudaEventRecord(e1, h->stream));
    h->acc_events.emplace_back(e0, e1);
    ++h->launches;
  }
  double local_pairs = (double)nseg, local_corr = (double)h->total_records;
  const int nxf = h->total_records > 0 ? nx : 0;
  if (with_h)
    k_finalize<true><<<1, 1024, 0, h->stream>>>(h->partials.as<double>(), h->segs_dev.as<Segment>(), nseg, h->grid_acc, h->per_cta, nv,
                                                h->segsum.as<double>(), h->eq_dev.as<double>(), local_pairs, local_corr,
                                                h->xpartials.as<double>(), nxf, h->xsegsum.as<double>());
  else
    k_finalize<false><<<1, 1024, 0, h->stream>>>(h->partials.as<double>(), h->segs_dev.as<Segment>(), nseg, h->grid_acc, h->per_cta, nv,
                                                 h->segsum.as<double>(), h->eq_dev.as<double>(), local_pairs, local_corr,
                                                 h->xpartials.as<double>(), nxf, h->xsegsum.as<double>());
  ++h->launches;
  B2_CUDA(cudaGetLastError());
  if (h->cfg.world_size > 1) {
    // Data-parallel exchange: ONE sum-allreduce of the packed normal equations per pass (NCCL over NVLink on the host side).
    double* buf = h->eq_dev.as<double>();
    size_t off = 0, cnt = eq_count;
    if (!with_h) { off = (size_t)nv * nv + nv; cnt = 3 + kMaxExtraTrials; }
    if (h->cfg.comm) {
      B2_TRY(b2_comm_allreduce_f64(h->cfg.comm, buf + off, cnt, (void*)h->stream));   // NCCL on the handle's stream: no host sync
    } else {
      // Host-language hook: the buffer is complete before it runs (it may reduce on a stream of its own), and it must have ordered
      // its result on `stream` (or completed it) before it returns.
      B2_CUDA(cudaStreamSynchronize(h->stream));
      if (h->cfg.allreduce(h->cfg.allreduce_user, buf + off, cnt, (void*)h->stream) != 0)
        return set_error(B2_ERR_COMM, "allreduce hook failed");
    }
  }
  B2_CUDA(cudaMemcpyAsync(h->pin_eq.p, h->eq_dev.p, eq_count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  ++h->stats.passes;
  return B2_OK;
}

static float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms = 0.f; cudaEventElapsedTime(&ms, a, b); return ms; }

// K3 + tile offsets of one pair-direction on stream st; the match count lands in *total_host (pinned) once st has drained.
static int launch_search(b2_icp* h, Direction* d, Cloud* S, Cloud* T, float r2, cudaStream_t st, DevBuf& cub_tmp, unsigned int* total_dev,
                         unsigned int* total_host, bool timed) {
  static const bool grid_order = [] { const char* e = getenv("B2_K3_ORDER"); return e && std::string(e) == "grid"; }();
  const size_t ns = S->n;
  const unsigned int ntiles = div_up(ns, kTile);
  B2_TRY(d->key.ensure(ns * 8)); B2_TRY(d->tile_count.ensure((size_t)ntiles * 4)); B2_TRY(d->tile_off.ensure((size_t)ntiles * 4));
  SearchGrid sg;
  B2_TRY(search_grid(T, &sg));
  cudaEvent_t n0 = nullptr, n1 = nullptr;
  if (timed) { B2_CUDA(cudaEventCreate(&n0)); B2_CUDA(cudaEventCreate(&n1)); B2_CUDA(cudaEventRecord(n0, st)); }
  // longest-first launch order (k_tile_cost + a 10^4..10^5-element radix sort); the geometry of a direction changes by millimetres
  // between outer iterations, so the order is kept for eight of them
  const unsigned int* order = nullptr;
  if (h->lpt_order && !grid_order && ntiles > 4u * (unsigned int)h->sms) {
    B2_TRY(d->order.ensure((size_t)ntiles * 16));
    unsigned int* cc = d->order.as<unsigned int>();
    if (d->order_src != d->src || d->order_tgt != d->tgt || d->order_tiles != ntiles || d->order_age >= 8) {
      if (T->dense) k_tile_cost<true><<<div_up(ntiles, 256), 256, 0, st>>>(S->s_xyz.as<float4>(), ns, T->table.as<HashEntry>(), sg, ntiles, cc, cc + ntiles);
      else k_tile_cost<false><<<div_up(ntiles, 256), 256, 0, st>>>(S->s_xyz.as<float4>(), ns, T->table.as<HashEntry>(), sg, ntiles, cc, cc + ntiles);
      size_t tmp2 = 0;
      B2_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp2, cc, cc + 2 * (size_t)ntiles, cc + ntiles, cc + 3 * (size_t)ntiles, (int)ntiles, 0, 32, st));
      B2_TRY(cub_tmp.ensure(tmp2));
      B2_CUDA(cub::DeviceRadixSort::SortPairsDescending(cub_tmp.p, tmp2, cc, cc + 2 * (size_t)ntiles, cc + ntiles, cc + 3 * (size_t)ntiles, (int)ntiles, 0, 32, st));
      d->order_src = d->src; d->order_tgt = d->tgt; d->order_tiles = ntiles; d->order_age = 0;
      h->launches += 2;
    }
    ++d->order_age;
    order = cc + 3 * (size_t)ntiles;
  }
  B2_CUDA(cudaMemsetAsync(d->tile_count.p, 0, (size_t)ntiles * 4, st));
  // persistent CTAs (as many as are resident at once) unless B2_K3_PERSIST=0; the longest-first order is what balances them
  // (B2_K3_PERSIST=n: 1/n of the resident CTA slots per launch, so that n launches of different streams share the GPU)
  static const int persist = [] { const char* e = getenv("B2_K3_PERSIST"); return e ? std::max(0, atoi(e)) : 1; }();
  const unsigned int nn_grid = persist > 0 && order ? std::min(ntiles, (unsigned int)std::max(1, h->sms * B2_K3_MINB / persist)) : ntiles;
#define B2_LAUNCH_NN(STATS, DENSE, DUAL)                                                                                            \
  k_nn_tiles<STATS, DENSE, DUAL><<<nn_grid, kTile, 0, st>>>(S->s_xyz.as<float4>(), ns, T->s_xyz.as<float4>(), T->box1.as<Aabb>(), T->box2.as<Aabb>(),  \
                                                     T->table.as<HashEntry>(), sg, r2, d->key.as<unsigned long long>(),                 \
                                                     d->tile_count.as<unsigned int>(), h->work_stats ? h->work_dev.as<unsigned long long>() : nullptr, order, ntiles)
  const bool dual = T->dense && T->dual;
  if (h->work_stats) { if (dual) B2_LAUNCH_NN(true, true, true); else if (T->dense) B2_LAUNCH_NN(true, true, false); else B2_LAUNCH_NN(true, false, false); }
  else { if (dual) B2_LAUNCH_NN(false, true, true); else if (T->dense) B2_LAUNCH_NN(false, true, false); else B2_LAUNCH_NN(false, false, false); }
#undef B2_LAUNCH_NN
  if (timed) {
    B2_CUDA(cudaEventRecord(n1, st));
    h->nn_events.emplace_back(n0, n1);
    h->stats.search_algorithmic_bytes += 12ull * ns + 12ull * T->n;
  }
  k_scan_tiles<<<1, 1024, 0, st>>>(d->tile_count.as<unsigned int>(), ntiles, d->tile_off.as<unsigned int>(), total_dev);
  h->launches += 2;
  B2_CUDA(cudaMemcpyAsync(total_host, total_dev, 4, cudaMemcpyDeviceToHost, st));
  return B2_OK;
}

// The global-frame rows, chunk boxes and per-block bounding boxes of one cloud at its current pose (K1x).
static int launch_rows(b2_icp* h, Cloud* c, float* bbox_partial, int xf_blocks) {
  const unsigned int nb1 = div_up(c->n, kChunk1), nb2 = div_up(nb1, 32);
  B2_TRY(c->box1.ensure(sizeof(Aabb) * nb1)); B2_TRY(c->box2.ensure(sizeof(Aabb) * nb2));
  k_xform_sorted<<<xf_blocks, 256, 0, h->stream>>>(c->l_xyz.as<float4>(), c->l_nrm.as<float4>(), c->n, mat4_of(c->T), c->s_xyz.as<float4>(),
                                                   c->s_nrm.as<float4>(), c->box1.as<Aabb>(), bbox_partial);
  k_chunk_boxes2<<<div_up((size_t)nb2 * 32, 256), 256, 0, h->stream>>>(c->box1.as<Aabb>(), nb1, c->box2.as<Aabb>(), nb2);
  h->launches += 2;
  return B2_OK;
}

static void release_direction(Direction* d) { for (DevBuf* b : {&d->key, &d->tile_count, &d->tile_off, &d->order}) b->release(); }

// Drops every search done ahead of b2_icp_run (their streams are drained first: the buffers go back to the pool).
static void drop_ahead(b2_icp* h) {
  if (h->ahead.empty() && h->ahead_clouds.empty()) return;
  for (int i = 0; i + 1 < h->nsearch; ++i) cudaStreamSynchronize(h->aux[i]);
  for (auto& a : h->ahead) release_direction(&a->d);
  h->ahead.clear();
  for (Cloud* c : h->ahead_clouds) c->rows_ahead = false;
  h->ahead_clouds.clear();
}

static bool rows_current(const Cloud* c) {
  return c->rows_ahead && c->indexed && c->rows_epoch == c->index_epoch && std::memcmp(c->rows_T, c->T, sizeof(c->T)) == 0;
}

static Ahead* find_ahead(b2_icp* h, const Cloud* S, const Cloud* T, float r2) {
  if (h->work_stats) return nullptr;
  for (auto& a : h->ahead)
    if (a->S == S && a->T == T && a->d.key.p && a->r2 == r2 && a->s_epoch == S->index_epoch && a->t_epoch == T->index_epoch && S->indexed && T->indexed &&
        std::memcmp(a->Ts, S->T, sizeof(a->Ts)) == 0 && std::memcmp(a->Tt, T->T, sizeof(a->Tt)) == 0)
      return a.get();
  return nullptr;
}

// cfg.search_ahead: cloud c has just been indexed (behind the upload of the cloud after it). Its global-frame rows are written at the
// pose it was added with, and the pair-directions between c and every cloud prepared the same way before it are searched on the
// auxiliary streams: while the remaining clouds cross PCIe the GPU is otherwise idle, and a first outer iteration that is called with
// the hinted radius and unchanged poses finds these searches done (align_once adopts them; bit-identical, it is the same kernel on
// the same rows). One rank only: with several, which rank owns a direction is not known before the last cloud has been added.
static int search_ahead_for(b2_icp* h, Cloud* c) {
  if (!h->cfg.search_ahead || h->cfg.world_size > 1 || h->work_stats || h->nsearch < 2 || c->n == 0 || !c->indexed) return B2_OK;
  const float hint = h->cfg.index_distance_hint;
  const float r2 = (float)((double)hint * (double)hint);
  const int xf_blocks = h->sms * 4;
  B2_TRY(h->ahead_bbox.ensure((size_t)xf_blocks * 6 * sizeof(float)));
  B2_TRY(launch_rows(h, c, h->ahead_bbox.as<float>(), xf_blocks));
  c->rows_ahead = true; c->rows_epoch = c->index_epoch; std::memcpy(c->rows_T, c->T, sizeof(c->T));
  if (!h->ahead_ev) B2_CUDA(cudaEventCreateWithFlags(&h->ahead_ev, cudaEventDisableTiming));
  B2_CUDA(cudaEventRecord(h->ahead_ev, h->stream));
  const int naux = h->nsearch - 1;
  for (int i = 0; i < naux; ++i) B2_CUDA(cudaStreamWaitEvent(h->aux[i], h->ahead_ev, 0));
  B2_TRY(h->ahead_totals_dev.ensure(sizeof(unsigned int) * b2_icp::kMaxAhead));
  B2_TRY(h->ahead_totals_pin.ensure(sizeof(unsigned int) * b2_icp::kMaxAhead));
  for (Cloud* o : h->ahead_clouds) {
    if (o == c || o->n == 0 || !rows_current(o)) continue;
    for (int dir = 0; dir < 2; ++dir) {
      if ((int)h->ahead.size() >= b2_icp::kMaxAhead) break;
      Cloud* S = dir == 0 ? o : c; Cloud* T = dir == 0 ? c : o;
      std::unique_ptr<Ahead> a(new Ahead());
      a->S = S; a->T = T; a->s_epoch = S->index_epoch; a->t_epoch = T->index_epoch; a->r2 = r2; a->slot = (int)h->ahead.size();
      std::memcpy(a->Ts, S->T, sizeof(a->Ts)); std::memcpy(a->Tt, T->T, sizeof(a->Tt));
      a->d.src = -2; a->d.tgt = -2;
      h->ahead_totals_pin.as<unsigned int>()[a->slot] = 0;
      const int si = h->ahead_rr++ % naux;
      const int rc = launch_search(h, &a->d, S, T, r2, h->aux[si], h->search_tmp[si + 1], h->ahead_totals_dev.as<unsigned int>() + a->slot,
                                   h->ahead_totals_pin.as<unsigned int>() + a->slot, false);
      if (rc != B2_OK) { cudaStreamSynchronize(h->aux[si]); release_direction(&a->d); return rc; }
      h->ahead.push_back(std::move(a));
    }
  }
  h->ahead_clouds.push_back(c);
  return B2_OK;
}

// index_distance_hint: the clouds added so far are indexed now — the caller is b2_icp_add_cloud with the NEXT cloud's copy in flight.
// Only a cloud's own magnitude is known yet; should a later cloud raise the handle's bound above the class chosen here, ensure_indexes
// rebuilds.
static int build_pending_index(b2_icp* h) {
  if (!(h->cfg.index_distance_hint > 0.f)) { h->pending_index.clear(); return B2_OK; }
  while (!h->pending_index.empty()) {
    Cloud* c = h->pending_index.front();
    // sharded uploads: a cloud another rank owns is indexed only once its bytes are here — waiting for them would delay THIS rank's
    // own upload (the next call), and with it every other rank (b2_icp_run builds whatever is still pending)
    if (c->ready_ev && cudaEventQuery(c->ready_ev) != cudaSuccess) break;
    h->pending_index.erase(h->pending_index.begin());
    if (c->indexed) continue;
    B2_TRY(cloud_local_box(h, c));
    double m = 0, sigma = 1.0;
    for (int k = 0; k < 3; ++k) m = std::max({m, std::fabs((double)c->lmin[k]), std::fabs((double)c->lmax[k])});
    for (int a = 0; a < 3; ++a) {
      double gsum = std::fabs((double)c->T[12 + a]);
      for (int k = 0; k < 3; ++k) gsum += std::fabs((double)c->T[a + 4 * k]) * std::max(std::fabs((double)c->lmin[k]), std::fabs((double)c->lmax[k]));
      m = std::max(m, gsum);
    }
    if (!std::isfinite(m)) continue;             // reported by b2_icp_run
    B2_TRY(pose_sigma(c->T, nullptr, &sigma));
    B2_TRY(build_index(h, c, h->cfg.index_distance_hint, sigma, m));
    B2_TRY(search_ahead_for(h, c));
  }
  return B2_OK;
}

// Index maintenance at the start of an outer iteration: (re)build the indexes the current radius / poses are not covered by.
static int ensure_indexes(b2_icp* h, float max_dist) {
  const int nc = num_impl_clouds(h);
  h->pending_index.clear();                      // whatever is still pending is built below, with the handle's full magnitude bound
  for (int i = 0; i < nc; ++i) B2_TRY(cloud_local_box(h, impl_cloud(h, i)));
  const double mtot = magnitude_bound(h);
  if (!std::isfinite(mtot)) return set_error(B2_ERR_ARG, "non-finite pose or point coordinates");
  bool built = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  for (int i = 0; i < nc; ++i) {
    Cloud* c = impl_cloud(h, i);
    double sigma = 1.0;
    if (c->indexed) B2_TRY(pose_sigma(c->T, c->F, &sigma));
    if (c->indexed && c->index_d == max_dist && sigma <= c->index_sigma && mtot <= c->index_mtot) continue;
    B2_TRY(pose_sigma(c->T, nullptr, &sigma));                    // (re)built at the current pose: B = I up to rounding
    if (!built) { B2_CUDA(cudaEventCreate(&e0)); B2_CUDA(cudaEventCreate(&e1)); B2_CUDA(cudaEventRecord(e0, h->stream)); built = true; }
    B2_TRY(build_index(h, c, max_dist, sigma, mtot));
  }
  if (built) {
    B2_CUDA(cudaEventRecord(e1, h->stream));
    B2_CUDA(cudaEventSynchronize(e1));
    h->ms_index_build = elapsed(e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  return B2_OK;
}

// AlignMeshes (icp_point_to_plane.cc:169-342).
static int align_once(b2_icp* h, float max_dist, float thr, int print, bool* converged, bool search_only = false, bool gate = true) {
  h->launches = 0;
  std::memset(&h->stats, 0, sizeof(h->stats));
  h->ms_index_build = 0.f;
  h->tries.clear();
  for (auto& pr : h->acc_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  h->acc_events.clear();
  for (auto& pr : h->nn_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  h->nn_events.clear();
  const int nc = num_impl_clouds(h);
  const int nmov = (int)h->movable.size();
  const int nv = 6 * (nc - 1);
  B2_TRY(ensure_indexes(h, max_dist));
  B2_CUDA(cudaEventRecord(h->ev[0], h->stream));

  // ---- K1x: global-frame rows, chunk boxes and bounding boxes of all clouds (one stream per cloud, one sync for all) ----
  if (!h->ahead.empty())     // the rows are rewritten below (with the same values where a search done ahead is still reading them)
    for (int i = 0; i + 1 < h->nsearch; ++i) { B2_CUDA(cudaEventRecord(h->join_ev[i], h->aux[i])); B2_CUDA(cudaStreamWaitEvent(h->stream, h->join_ev[i], 0)); }
  const int xf_blocks = h->sms * 4;
  B2_TRY(h->bbox_partial.ensure((size_t)nc * xf_blocks * 6 * sizeof(float)));
  B2_TRY(h->pin_bbox.ensure((size_t)nc * xf_blocks * 6 * sizeof(float)));
  for (int i = 0; i < nc; ++i) {
    Cloud* c = impl_cloud(h, i);
    if (c->n == 0) continue;
    B2_TRY(launch_rows(h, c, h->bbox_partial.as<float>() + (size_t)i * xf_blocks * 6, xf_blocks));
  }
  B2_CUDA(cudaMemcpyAsync(h->pin_bbox.p, h->bbox_partial.p, (size_t)nc * xf_blocks * 6 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < nc; ++i) {
    Cloud* c = impl_cloud(h, i);
    for (int d = 0; d < 3; ++d) { c->bmin[d] = INFINITY; c->bmax[d] = -INFINITY; }
    if (c->n == 0) continue;
    const float* p = h->pin_bbox.as<float>() + (size_t)i * xf_blocks * 6;
    for (int b = 0; b < xf_blocks; ++b)
      for (int d = 0; d < 3; ++d) { c->bmin[d] = std::min(c->bmin[d], p[b * 6 + d]); c->bmax[d] = std::max(c->bmax[d], p[b * 6 + 3 + d]); }
  }
  B2_CUDA(cudaEventRecord(h->ev[1], h->stream));

  // ---- pair scheduling in ik order (icp_point_to_plane.cc:208-309); ownership from the shared planner ----
  h->ndirs = 0;
  const int world = std::max(1, h->cfg.world_size), rank = h->cfg.rank;
  {
    const int cap = nmov * nmov + 2 * nmov + 1;
    std::vector<int32_t> ps(cap), pt(cap), po(cap);
    int cnt = 0;
    B2_TRY(b2_icp_plan_directions(nmov, h->fixed ? 1 : 0, world, ps.data(), pt.data(), po.data(), cap, &cnt));
    for (int k = 0; k < cnt; ++k) {
      if (!gate || boxes_intersect(impl_cloud(h, ps[k]), impl_cloud(h, pt[k]))) {
        Direction* d = next_direction(h); d->src = ps[k]; d->tgt = pt[k]; d->local = po[k] == rank;
      }
    }
  }
  const float r2 = (float)((double)max_dist * (double)max_dist);
  unsigned int r2_bits; std::memcpy(&r2_bits, &r2, 4);
  const unsigned long long init_key = (unsigned long long)r2_bits << 32;
  h->last_init_key = init_key;

  // ---- K3 + tile offsets: one pair-direction on stream st ----
  B2_TRY(h->pin_misc.ensure(sizeof(unsigned int) * (size_t)std::max(1, h->ndirs)));
  B2_TRY(h->totals_dev.ensure(sizeof(unsigned int) * (size_t)std::max(1, h->ndirs)));
  unsigned int* totals = h->pin_misc.as<unsigned int>();
  if (h->work_stats) { B2_TRY(h->work_dev.ensure(5 * sizeof(unsigned long long))); B2_CUDA(cudaMemsetAsync(h->work_dev.p, 0, 5 * sizeof(unsigned long long), h->stream)); }
  const int nstreams = h->nsearch;
  for (int k = 0; k < h->ndirs; ++k) { h->dirs[k]->count = 0; totals[k] = 0; }
  std::vector<std::pair<int, int>> adopted;      // (direction, slot): searched ahead of this call
  bool launched = false;
  int adopted_slot = -1;
  auto issue_search = [&](int k, cudaStream_t st, DevBuf& cub_tmp) -> int {
    Direction* d = h->dirs[k].get();
    Cloud* S = impl_cloud(h, d->src); Cloud* T = impl_cloud(h, d->tgt);
    d->ahead_slot = -1;
    if (S->n == 0 || T->n == 0) return B2_OK;
    if (Ahead* a = find_ahead(h, S, T, r2)) {
      std::swap(d->key, a->d.key); std::swap(d->tile_count, a->d.tile_count); std::swap(d->tile_off, a->d.tile_off); std::swap(d->order, a->d.order);
      d->order_src = d->src; d->order_tgt = d->tgt; d->order_tiles = a->d.order_tiles; d->order_age = a->d.order_age;
      d->ahead_slot = a->slot;
      adopted.emplace_back(k, a->slot);
      adopted_slot = a->slot;
      ++h->stats.searches_ahead;
      return B2_OK;
    }
    launched = true;
    return launch_search(h, d, S, T, r2, st, cub_tmp, h->totals_dev.as<unsigned int>() + k, &totals[k], true);
  };

  // ---- optional (pack_overlap): K4 behind K3 — a set is packed (on pack_stream) as soon as its search has finished, while later sets
  // are still being searched. A set's first record
  // is a running offset kept on the device (PackChain); the host learns the counts at the one synchronisation below, as before, and
  // derives the same offsets for the segment table. Needs record arrays that are already large enough: sized from the previous
  // iteration, or — first iteration — for the worst case (every query matched) when that is a small part of the free memory.
  int nlocal = 0;
  unsigned long long worst = 0;
  for (int k = 0; k < h->ndirs; ++k)
    if (h->dirs[k]->local) { ++nlocal; const Cloud* S = impl_cloud(h, h->dirs[k]->src); if (impl_cloud(h, h->dirs[k]->tgt)->n) worst += S->n; }
  bool chain = h->pack_overlap && !search_only && !h->work_stats && nlocal > 0;
  if (chain && h->pack_presize && h->rec_a.cap == 0 && worst > 0) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (double)worst * 48.0 * 1.2 < 0.25 * (double)free_b) {
      B2_TRY(h->rec_a.ensure(worst * 16)); B2_TRY(h->rec_b.ensure(worst * 16)); B2_TRY(h->rec_c.ensure(worst * 16));
    }
  }
  const unsigned long long rec_cap = std::min({h->rec_a.cap, h->rec_b.cap, h->rec_c.cap}) / 16;
  chain = chain && rec_cap > 0;
  unsigned long long* chain_base = nullptr;
  unsigned int* chain_overflow = nullptr;
  int chain_pos = 0;
  if (chain && !h->pack_stream) {
    int least = 0, greatest = 0;
    B2_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    B2_CUDA(cudaStreamCreateWithPriority(&h->pack_stream, cudaStreamNonBlocking, greatest));   // its short kernels slot in between K3's tiles
    B2_CUDA(cudaEventCreateWithFlags(&h->pack_join_ev, cudaEventDisableTiming));
  }
  if (chain) {
    const size_t bytes = ((size_t)nlocal + 2) * sizeof(unsigned long long);
    B2_TRY(h->chain_dev.ensure(bytes)); B2_TRY(h->pin_chain.ensure(sizeof(unsigned int)));
    B2_CUDA(cudaMemsetAsync(h->chain_dev.p, 0, bytes, h->stream));
    chain_base = h->chain_dev.as<unsigned long long>();
    chain_overflow = reinterpret_cast<unsigned int*>(chain_base + nlocal + 1);
    *h->pin_chain.as<unsigned int>() = 1u;         // pessimistic until the flag has been read back
    while ((int)h->done_ev.size() < nlocal) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->done_ev.push_back(e); }
  }

  int issued = 0;
  if (nstreams > 1 || chain) {
    B2_CUDA(cudaEventRecord(h->fork_ev, h->stream));
    for (int i = 0; i + 1 < nstreams; ++i) B2_CUDA(cudaStreamWaitEvent(h->aux[i], h->fork_ev, 0));
    if (chain) B2_CUDA(cudaStreamWaitEvent(h->pack_stream, h->fork_ev, 0));
  }
  for (int k = 0; k < h->ndirs; ++k) {
    if (!h->dirs[k]->local) continue;
    const int si = issued++ % nstreams;
    cudaStream_t st = si == 0 ? h->stream : h->aux[si - 1];
    launched = false; adopted_slot = -1;
    B2_TRY(issue_search(k, st, h->search_tmp[si]));
    if (chain && (launched || adopted_slot >= 0)) {
      Direction* d = h->dirs[k].get();
      Cloud* S = impl_cloud(h, d->src); Cloud* T = impl_cloud(h, d->tgt);
      if (launched) {
        B2_CUDA(cudaEventRecord(h->done_ev[chain_pos], st));
        B2_CUDA(cudaStreamWaitEvent(h->pack_stream, h->done_ev[chain_pos], 0));
      }   // (a set searched ahead of this call is complete: this call's work was queued behind the auxiliary streams above)
      const unsigned int* set_total = launched ? h->totals_dev.as<unsigned int>() + k : h->ahead_totals_dev.as<unsigned int>() + adopted_slot;
      const PackChain pc = {chain_base, set_total, chain_overflow, rec_cap, chain_pos};
      k_pack_tiles<<<div_up(S->n, kTile), kTile, 0, h->pack_stream>>>(S->s_xyz.as<float4>(), S->s_nrm.as<float4>(), S->n, T->s_xyz.as<float4>(),
                                                                      T->s_nrm.as<float4>(), T->perm_inv.as<unsigned int>(),
                                                                      d->key.as<unsigned long long>(), d->tile_off.as<unsigned int>(), init_key, 0ull, pc,
                                                                      h->rec_a.as<float4>(), h->rec_b.as<float4>(), h->rec_c.as<float4>());
      ++h->launches; ++chain_pos;
    }
  }
  if (nstreams > 1)
    for (int i = 0; i + 1 < nstreams; ++i) { B2_CUDA(cudaEventRecord(h->join_ev[i], h->aux[i])); B2_CUDA(cudaStreamWaitEvent(h->stream, h->join_ev[i], 0)); }
  if (chain) {
    B2_CUDA(cudaMemcpyAsync(h->pin_chain.p, chain_overflow, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->pack_stream));
    B2_CUDA(cudaEventRecord(h->pack_join_ev, h->pack_stream));
    B2_CUDA(cudaStreamWaitEvent(h->stream, h->pack_join_ev, 0));
  }
  B2_CUDA(cudaStreamSynchronize(h->stream));
  B2_CUDA(cudaEventRecord(h->ev[2], h->stream));
  if (!h->ahead.empty() || !h->ahead_clouds.empty()) {
    for (int i = 0; i + 1 < h->nsearch; ++i) B2_CUDA(cudaStreamSynchronize(h->aux[i]));     // (also drained when nothing was launched above)
    for (auto& ks : adopted) totals[ks.first] = h->ahead_totals_pin.as<unsigned int>()[ks.second];
    drop_ahead(h);           // whatever was not adopted is stale by now
  }

  // ---- K4: pack local non-empty sets into one record array ----
  h->segs_host.clear(); h->seg_dir.clear();
  unsigned long long total = 0;
  for (int k = 0; k < h->ndirs; ++k) {
    Direction* d = h->dirs[k].get();
    if (!d->local) continue;
    d->count = totals[k];
    if (d->ahead_slot < 0) h->stats.search_algorithmic_bytes += 8ull * d->count;
    d->rec_begin = total;
    if (d->count == 0) continue;   // empty sets are not registered (icp_point_to_plane.cc:240)
    h->segs_host.push_back(Segment{total, total + d->count, d->src, d->tgt});
    h->seg_dir.push_back(k);
    total += d->count;
  }
  h->total_records = total;
  const int nseg = (int)h->segs_host.size();
  if (search_only) { *converged = false; h->stats.kernel_launches = h->launches; return B2_OK; }
  B2_TRY(h->rec_a.ensure(std::max<unsigned long long>(total, 1) * 16)); B2_TRY(h->rec_b.ensure(std::max<unsigned long long>(total, 1) * 16));
  B2_TRY(h->rec_c.ensure(std::max<unsigned long long>(total, 1) * 16));
  // (the device-side offsets of an overlapped pack are the prefix sums computed above: same counts, same set order)
  const bool packed = chain && *h->pin_chain.as<unsigned int>() == 0u;
  h->stats.packs_overlapped = packed ? chain_pos : 0;
  for (int s = 0; s < nseg && !packed; ++s) {
    Direction* d = h->dirs[h->seg_dir[s]].get();
    Cloud* S = impl_cloud(h, d->src); Cloud* T = impl_cloud(h, d->tgt);
    const PackChain none = {nullptr, nullptr, nullptr, 0ull, 0};
    k_pack_tiles<<<div_up(S->n, kTile), kTile, 0, h->stream>>>(S->s_xyz.as<float4>(), S->s_nrm.as<float4>(), S->n, T->s_xyz.as<float4>(),
                                                               T->s_nrm.as<float4>(), T->perm_inv.as<unsigned int>(), d->key.as<unsigned long long>(),
                                                               d->tile_off.as<unsigned int>(), init_key, d->rec_begin, none, h->rec_a.as<float4>(),
                                                               h->rec_b.as<float4>(), h->rec_c.as<float4>());
    ++h->launches;
  }
  B2_TRY(h->segs_dev.ensure(sizeof(Segment) * std::max(1, nseg)));
  B2_TRY(h->pin_segs.ensure(sizeof(Segment) * std::max(1, nseg)));
  if (nseg) {
    std::memcpy(h->pin_segs.p, h->segs_host.data(), sizeof(Segment) * nseg);
    B2_CUDA(cudaMemcpyAsync(h->segs_dev.p, h->pin_segs.p, sizeof(Segment) * nseg, cudaMemcpyHostToDevice, h->stream));
  }
  B2_CUDA(cudaEventRecord(h->ev[3], h->stream));
  if (print) {
    for (int k = 0; k < h->ndirs; ++k) {
      Direction* d = h->dirs[k].get();
      if (d->local) printf("  found correspondences from %d to %d: %llu\n", d->src, d->tgt, d->count);
    }
  }

  // ---- streaming-pass geometry ----
  // record ranges (= CTAs of a pass; the partition fixes the summation order, so it is the same for every kind of pass): 6 per SM — three
  // waves of the passes that carry the normal equations (2 CTAs of 128 registers per SM), two waves of the cost-only passes (3 per SM)
  static const int ranges_per_sm = [] { const char* e = getenv("B2_K5_RANGES_PER_SM"); return e ? std::max(1, std::min(16, atoi(e))) : 6; }();
  h->grid_acc = h->sms * ranges_per_sm;
  unsigned long long per = (total + h->grid_acc - 1) / (unsigned long long)h->grid_acc;
  per = std::max<unsigned long long>(kAccThreads, (per + kAccThreads - 1) / kAccThreads * kAccThreads);
  h->per_cta = per;
  const size_t eq_count = (size_t)nv * nv + nv + 3 + kMaxExtraTrials;
  B2_TRY(h->partials.ensure(sizeof(double) * kAccVals * (size_t)std::max(1, nseg) * h->grid_acc));
  B2_TRY(h->xpartials.ensure(sizeof(double) * kMaxExtraTrials * (size_t)std::max(1, nseg) * h->grid_acc));
  B2_TRY(h->segsum.ensure(sizeof(double) * kAccVals * std::max(1, nseg)));
  B2_TRY(h->xsegsum.ensure(sizeof(double) * kMaxExtraTrials * std::max(1, nseg)));
  B2_TRY(h->eq_dev.ensure(sizeof(double) * eq_count));
  B2_TRY(h->pin_eq.ensure(sizeof(double) * eq_count));
  B2_TRY(h->poses_dev.ensure(sizeof(CloudPose) * nc * (1 + kMaxExtraTrials)));
  B2_TRY(h->pin_poses.ensure(sizeof(CloudPose) * nc * (1 + kMaxExtraTrials)));

  // ---- compute() (icp_point_to_plane_impl.h:115-293): LM over the fixed correspondence sets ----
  // The reference evaluates the tries of one LM iteration one after the other (lambda doubles after each rejection, :217-285); the
  // trial states do not depend on the outcome of earlier tries, so a pass evaluates the costs of several consecutive tries at once
  // (see k_accumulate_tma) and the decisions are then taken in the reference's order. B2_LM_SPEC="a,b": extra tries riding along with
  // the first pass of an iteration (which also carries the normal equations at its first try) / with the follow-up passes.
  static const std::pair<int, int> spec = [] {
    int a = 1, b = kMaxExtraTrials;
    if (const char* e = getenv("B2_LM_SPEC")) { if (sscanf(e, "%d,%d", &a, &b) < 2) b = a; }
    if (const char* e = getenv("B2_K5")) if (std::string(e) == "ldg") a = b = 0;
    return std::make_pair(std::max(0, std::min(a, kMaxExtraTrials)), std::max(0, std::min(b, kMaxExtraTrials)));
  }();
  std::vector<Pose> poses(nc);
  std::vector<std::vector<Pose>> batch;
  std::vector<double> H((size_t)nv * nv), b(nv), x(nv), HL;
  const double* eq = h->pin_eq.as<double>();
  const size_t cost_at = (size_t)nv * nv + nv;
  batch.assign(1, poses);
  B2_TRY(run_pass(h, batch, true, nv));
  std::copy(eq, eq + (size_t)nv * nv, H.begin()); std::copy(eq + (size_t)nv * nv, eq + cost_at, b.begin());
  double cost = eq[cost_at];
  h->stats.num_pairs = (int)std::llround(eq[cost_at + 1]);
  h->stats.num_correspondences = (uint64_t)std::llround(eq[cost_at + 2]);
  h->stats.local_correspondences = total;
  h->stats.num_variables = nv;
  h->stats.first_cost = cost;
  h->H0 = H; h->b0 = b; h->cost0 = cost;
  double lambda = 0.1;
  const int max_inner = h->cfg.inner_max_iterations > 0 ? h->cfg.inner_max_iterations : 150;
  // The first pass of an LM iteration carries a speculative second try only once the alignment has settled (the previous outer
  // iteration took at most three LM iterations, i.e. it was mostly its final chain of ten rejected tries): far from convergence every
  // iteration is accepted at its first try and the extra cost evaluation (+32 % on the pass) would be thrown away each time. The
  // decision depends on counts the reference produces identically, never on timing, so runs stay reproducible.
  const bool speculate_first = h->prev_inner_iterations > 0 && h->prev_inner_iterations <= 3;
  for (int it = 0; it < max_inner; ++it) {
    ++h->stats.inner_iterations;
    bool applied = false;
    int ntries = 0;
    for (int lm = 0; lm < 10 && !applied;) {
      // tries lm .. lm + nb - 1 of this iteration: damping lambda, 2 lambda, 4 lambda, ...
      const int nb = std::min(10 - lm, 1 + (lm == 0 ? (speculate_first ? spec.first : 0) : spec.second));
      const bool with_h = lm == 0;
      batch.assign(nb, poses);
      double lam_j = lambda;
      for (int j = 0; j < nb; ++j, lam_j = 2.f * lam_j) {
        HL = H;
        for (int i = 0; i < nv; ++i) HL[(size_t)i * nv + i] += lam_j;
        if (nv > 0) sym_solve(HL, nv, b.data(), x.data());
        for (int ci = 1; ci < nc; ++ci) {
          double neg[6];
          for (int k = 0; k < 6; ++k) neg[k] = -x[6 * (ci - 1) + k];
          batch[j][ci] = pose_mul(pose_exp(neg), poses[ci]);
        }
      }
      // Fused pass: the costs at the nb trial states and (first pass of an iteration) the normal equations at its first try; if
      // that try is accepted they are exactly what the next iteration of compute() would recompute (same poses, same arithmetic).
      B2_TRY(run_pass(h, batch, with_h, nv));
      for (int j = 0; j < nb; ++j) {
        ++ntries;
        ++h->stats.lm_tries_total;
        const double new_cost = j == 0 ? eq[cost_at] : eq[cost_at + 3 + (j - 1)];
        if (new_cost < cost) {
          poses = batch[j];
          if (j == 0 && with_h) {
            std::copy(eq, eq + (size_t)nv * nv, H.begin()); std::copy(eq + (size_t)nv * nv, eq + cost_at, b.begin());
          } else {
            // accepted a try whose normal equations were not speculated: one pass at the accepted state
            batch.assign(1, poses);
            B2_TRY(run_pass(h, batch, true, nv));
            std::copy(eq, eq + (size_t)nv * nv, H.begin()); std::copy(eq + (size_t)nv * nv, eq + cost_at, b.begin());
          }
          cost = new_cost;
          lambda = 0.5f * lambda;
          applied = true;
          break;
        } else {
          lambda = 2.f * lambda;
        }
      }
      lm += nb;
    }
    h->tries.push_back(ntries);
    if (!applied) break;
  }
  h->prev_inner_iterations = h->stats.inner_iterations;
  h->stats.last_cost = cost;
  h->stats.final_lambda = lambda;
  B2_CUDA(cudaEventRecord(h->ev[4], h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));

  // ---- pose composition + convergence (icp_point_to_plane.cc:320-341) ----
  *converged = true;
  for (int m = 0; m < nmov; ++m) {
    Cloud* c = h->movable[m].get();
    float U[16], N[16];
    affine_from_pose(poses[impl_index_of_movable(h, m)], U);
    affine_mul(U, c->T, N);
    const float dx = c->T[12] - N[12], dy = c->T[13] - N[13], dz = c->T[14] - N[14];
    const float movement = std::sqrt(dx * dx + (dy * dy + dz * dz));
    if (movement > thr) *converged = false;
    if (print) printf("  %d moved by %g\n", impl_index_of_movable(h, m), movement);
    std::memcpy(c->T, N, sizeof(N));
  }
  h->stats.kernel_launches = h->launches;
  h->stats.ms_index = elapsed(h->ev[0], h->ev[1]);
  h->stats.ms_search = elapsed(h->ev[1], h->ev[2]);
  h->stats.ms_pack = elapsed(h->ev[2], h->ev[3]);
  h->stats.ms_inner = elapsed(h->ev[3], h->ev[4]);
  h->stats.ms_total = elapsed(h->ev[0], h->ev[4]);
  double acc = 0; for (auto& pr : h->acc_events) acc += elapsed(pr.first, pr.second);
  h->stats.ms_accum_kernel_avg = h->acc_events.empty() ? 0.f : (float)(acc / h->acc_events.size());
  double nn = 0; for (auto& pr : h->nn_events) nn += elapsed(pr.first, pr.second);
  // with several search streams the per-launch spans overlap, so the per-launch figure is the phase time over the launches
  h->stats.ms_search_kernel_avg = h->nn_events.empty() ? 0.f : (float)((h->nsearch > 1 ? (double)h->stats.ms_search : nn) / h->nn_events.size());
  h->stats.search_launches = (int)h->nn_events.size();
  h->stats.ms_index_build = h->ms_index_build;
  for (int i = 0; i < nc; ++i) if (impl_cloud(h, i)->n && !impl_cloud(h, i)->dense) ++h->stats.sparse_grids;
  if (h->work_stats) {
    B2_CUDA(cudaMemcpy(h->stats.search_work, h->work_dev.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  return B2_OK;
}

}  // namespace b2

extern "C" {

int b2_icp_plan_directions(int n_movable, int has_fixed, int world_size, int32_t* src_impl, int32_t* tgt_impl, int32_t* owner, int cap,
                           int* count) {
  if (!count || n_movable < 0 || world_size < 1) return set_error(B2_ERR_ARG, "bad argument");
  int n = 0;
  auto emit = [&](int s, int t) {
    if (n < cap) { if (src_impl) src_impl[n] = s; if (tgt_impl) tgt_impl[n] = t; if (owner) owner[n] = n % world_size; }
    ++n;
  };
  const int off = has_fixed ? 1 : 0;   // impl index 0 is the fixed cloud when there is one
  for (int i = 0; i < n_movable; ++i)
    for (int k = 0; k < n_movable; ++k) {
      if (i != k) emit(i + off, k + off);
      else if (has_fixed) { emit(i + off, 0); emit(0, i + off); }
    }
  *count = n;
  return n <= cap ? B2_OK : set_error(B2_ERR_ARG, "capacity %d too small for %d directions", cap, n);
}

int b2_icp_upload_owner(int cloud_id, int world_size) { return world_size > 0 && cloud_id >= 0 ? cloud_id % world_size : -1; }

const char* b2_last_error(void) { return last_error_ref().c_str(); }
int b2_abi_version(void) { return B2_ABI_VERSION; }
int b2_trim(void) { pool_trim_all(); return B2_OK; }

int b2_device_info(int* device, char* name, size_t name_cap, int* sm_count, int* cc_major, int* cc_minor) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return set_error(B2_ERR_NO_DEVICE, "no CUDA device available; libeth3d_b200 has no CPU fallback");
  int dev = 0; B2_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p; B2_CUDA(cudaGetDeviceProperties(&p, dev));
  if (device) *device = dev;
  if (name && name_cap) { std::strncpy(name, p.name, name_cap - 1); name[name_cap - 1] = 0; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return B2_OK;
}

void b2_icp_default_config(b2_icp_config* cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->device = -1; cfg->inner_max_iterations = 150; cfg->world_size = 1; cfg->search_ahead = 1;
}

int b2_icp_create(const b2_icp_config* cfg, b2_icp** out) {
  if (!out) return set_error(B2_ERR_ARG, "out is null");
  *out = nullptr;
  b2_icp_config c; b2_icp_default_config(&c);
  if (cfg) c = *cfg;
  if (c.world_size < 1) c.world_size = 1;
  if (c.world_size > 1 && !c.allreduce && !c.comm) return set_error(B2_ERR_ARG, "world_size > 1 needs a b2_comm or an allreduce hook");
  if (c.shard_uploads && !c.comm) return set_error(B2_ERR_ARG, "shard_uploads needs a b2_comm");
  if (c.rank < 0 || c.rank >= c.world_size) return set_error(B2_ERR_ARG, "rank %d out of range", c.rank);
  std::unique_ptr<b2_icp> h(new b2_icp());
  h->cfg = c;
  B2_TRY(select_device(c.device, &h->device, &h->sms));
  if (c.stream) { h->stream = (cudaStream_t)c.stream; }
  else { B2_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  B2_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  B2_CUDA(cudaStreamCreateWithFlags(&h->bcast_stream, cudaStreamNonBlocking));
  if (const char* e = getenv("B2_K3_STREAMS")) h->nsearch = std::max(1, std::min((int)b2_icp::kMaxSearchStreams, atoi(e)));
  if (const char* e = getenv("B2_K3_WORK")) h->work_stats = e[0] && e[0] != '0';
  for (int i = 0; i + 1 < h->nsearch; ++i) {
    B2_CUDA(cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking));
    B2_CUDA(cudaEventCreateWithFlags(&h->join_ev[i], cudaEventDisableTiming));
  }
  B2_CUDA(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
  {   // (the pack stream itself is created on first use: a stream of another priority is a new channel for the driver, tens of ms)
    if (const char* e = getenv("B2_PACK")) { const std::string v(e); h->pack_overlap = v == "overlap" || v == "overlap_nosize"; h->pack_presize = v != "overlap_nosize"; }
  }
  for (auto& e : h->ev) B2_CUDA(cudaEventCreate(&e));
  std::memset(&h->stats, 0, sizeof(h->stats));
  *out = h.release();
  return B2_OK;
}

int b2_icp_destroy(b2_icp* h) {
  if (!h) return B2_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  auto free_cloud = [](Cloud* c) {
    if (!c) return;
    if (c->ready_ev) { cudaEventSynchronize(c->ready_ev); cudaEventDestroy(c->ready_ev); c->ready_ev = nullptr; }
    for (DevBuf* b : {&c->local_xyz, &c->local_nrm, &c->l_xyz, &c->l_nrm, &c->perm_inv, &c->rb, &c->starts, &c->corner, &c->s_xyz, &c->s_nrm, &c->table, &c->box1, &c->box2}) b->release();
  };
  for (auto& c : h->movable) free_cloud(c.get());
  free_cloud(h->fixed.get());
  drop_ahead(h);
  if (h->pack_stream) { cudaStreamSynchronize(h->pack_stream); cudaStreamDestroy(h->pack_stream); }
  if (h->pack_join_ev) cudaEventDestroy(h->pack_join_ev);
  for (cudaEvent_t e : h->done_ev) cudaEventDestroy(e);
  h->chain_dev.release(); h->pin_chain.release();
  for (DevBuf* b : {&h->ahead_totals_dev, &h->ahead_bbox}) b->release();
  h->ahead_totals_pin.release();
  if (h->ahead_ev) cudaEventDestroy(h->ahead_ev);
  for (auto& d : h->dirs) release_direction(d.get());
  for (DevBuf* b : {&h->bbox_partial, &h->cub_tmp, &h->cell_counts, &h->rec_a, &h->rec_b, &h->rec_c, &h->segs_dev, &h->poses_dev, &h->partials,
                    &h->segsum, &h->eq_dev, &h->scatter_m, &h->scatter_d, &h->xpartials, &h->xsegsum, &h->work_dev, &h->totals_dev}) b->release();
  for (PinnedBuf* b : {&h->pin_bbox, &h->pin_counts, &h->pin_eq, &h->pin_poses, &h->pin_segs, &h->pin_misc}) b->release();
  for (auto& pr : h->acc_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto& pr : h->nn_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (int i = 0; i < b2_icp::kMaxSearchStreams - 1; ++i) { if (h->aux[i]) cudaStreamDestroy(h->aux[i]); if (h->join_ev[i]) cudaEventDestroy(h->join_ev[i]); }
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  for (DevBuf& b : h->search_tmp) b.release();
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->bcast_stream) { cudaStreamSynchronize(h->bcast_stream); cudaStreamDestroy(h->bcast_stream); }
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return B2_OK;
}

static int add_cloud_impl(b2_icp* h, const float* xyz, const float* nrm, size_t n, size_t stride, const float T[16], int fixed, int* out_id,
                          bool from_device) {
  if (!h || !T) return set_error(B2_ERR_ARG, "null argument");
  const int owner = (!fixed && h->cfg.shard_uploads && h->cfg.comm && h->cfg.world_size > 1) ? b2_icp_upload_owner((int)h->movable.size(), h->cfg.world_size) : -1;
  if (n > 0 && (!xyz || !nrm) && (owner < 0 || owner == h->cfg.rank)) return set_error(B2_ERR_ARG, "null argument");
  if (!from_device && stride < 12) return set_error(B2_ERR_ARG, "stride_bytes must be >= 12");
  if (n >= (1ull << 31)) return set_error(B2_ERR_ARG, "clouds above 2^31 points are not supported (pcl::Correspondence indices are int)");
  B2_CUDA(cudaSetDevice(h->device));
  if (fixed) {
    // icp_point_to_plane.cc:112-127: transform to the global frame, concatenate, return -1.
    Cloud tmp;
    struct Guard { Cloud* c; ~Guard() { c->local_xyz.release(); c->local_nrm.release(); } } tmp_guard{&tmp};
    B2_TRY(upload_cloud(h, &tmp, xyz, nrm, n, stride, from_device));
    if (!h->fixed) { h->fixed.reset(new Cloud()); h->fixed->is_fixed = true; const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; std::memcpy(h->fixed->T, I, sizeof(I)); }
    Cloud* f = h->fixed.get();
    const size_t old = f->n, tot = old + n;
    Scoped sx, sn;      // become the fixed cloud's buffers on success (swapped below), released on every error path
    DevBuf &nx = sx.b, &nn = sn.b;
    B2_TRY(nx.ensure(std::max<size_t>(tot, 1) * 12)); B2_TRY(nn.ensure(std::max<size_t>(tot, 1) * 12));
    if (old) {
      B2_CUDA(cudaMemcpyAsync(nx.p, f->local_xyz.p, old * 12, cudaMemcpyDeviceToDevice, h->stream));
      B2_CUDA(cudaMemcpyAsync(nn.p, f->local_nrm.p, old * 12, cudaMemcpyDeviceToDevice, h->stream));
    }
    if (n) k_transform<<<div_up(n, 256), 256, 0, h->stream>>>(tmp.local_xyz.as<float>(), tmp.local_nrm.as<float>(), n, mat4_of(T),
                                                             nx.as<float>() + 3 * old, nn.as<float>() + 3 * old);
    B2_CUDA(cudaStreamSynchronize(h->stream));
    std::swap(f->local_xyz, nx); std::swap(f->local_nrm, nn);   // the old buffers leave with the Scoped temporaries
    f->n = tot; f->have_lbox = f->indexed = false;
    if (out_id) *out_id = -1;
    return B2_OK;
  }
  std::unique_ptr<Cloud> c(new Cloud());
  std::memcpy(c->T, T, sizeof(float) * 16);
  B2_TRY(upload_cloud(h, c.get(), xyz, nrm, n, stride, from_device, owner));
  h->pending_index.push_back(c.get());
  h->movable.push_back(std::move(c));
  if (out_id) *out_id = (int)h->movable.size() - 1;
  return B2_OK;
}

int b2_icp_add_cloud(b2_icp* h, const float* xyz, const float* normals, size_t n, size_t stride_bytes, const float T[16], int fixed, int* out_id) {
  return add_cloud_impl(h, xyz, normals, n, stride_bytes, T, fixed, out_id, false);
}
int b2_icp_add_cloud_dev(b2_icp* h, const float* xyz_dev, const float* normals_dev, size_t n, const float T[16], int fixed, int* out_id) {
  return add_cloud_impl(h, xyz_dev, normals_dev, n, 12, T, fixed, out_id, true);
}

int b2_icp_run(b2_icp* h, float max_dist, int initial_iteration, int max_iters, float thr, int print, int* converged) {
  if (!h || !converged) return set_error(B2_ERR_ARG, "null argument");
  *converged = 0;
  if (h->movable.empty()) return set_error(B2_ERR_STATE, "Run() without any movable cloud (reference: CHECK(!clouds_.empty()))");
  B2_CUDA(cudaSetDevice(h->device));
  for (int i = initial_iteration; i < initial_iteration + max_iters; ++i) {
    if (print) printf("-- Alignment iteration %d --\n", i);
    bool conv = false;
    B2_TRY(align_once(h, max_dist, thr, print, &conv));
    if (conv) {
      if (print) printf("Convergence is assumed as the maximum movement is less than the threshold.\n");
      *converged = 1;
      return B2_OK;
    }
  }
  return B2_OK;
}

int b2_icp_get_pose(b2_icp* h, int id, float T[16]) {
  if (!h || !T || id < 0 || id >= (int)h->movable.size()) return set_error(B2_ERR_ARG, "bad cloud id %d", id);
  std::memcpy(T, h->movable[id]->T, sizeof(float) * 16);
  return B2_OK;
}
int b2_icp_set_pose(b2_icp* h, int id, const float T[16]) {
  if (!h || !T || id < 0 || id >= (int)h->movable.size()) return set_error(B2_ERR_ARG, "bad cloud id %d", id);
  std::memcpy(h->movable[id]->T, T, sizeof(float) * 16);
  return B2_OK;
}

int b2_icp_set_option(b2_icp* h, const char* name, int value) {
  if (!h || !name) return set_error(B2_ERR_ARG, "null argument");
  const std::string n(name);
  if (n == "pack_overlap") h->pack_overlap = value != 0;
  else if (n == "lpt_order") h->lpt_order = value != 0;
  else return set_error(B2_ERR_ARG, "unknown option '%s'", name);
  return B2_OK;
}

int b2_icp_last_stats(b2_icp* h, b2_icp_stats* out) {
  if (!h || !out) return set_error(B2_ERR_ARG, "null argument");
  *out = h->stats;
  return B2_OK;
}
int b2_icp_get_lm_tries(b2_icp* h, int32_t* tries, int cap, int* count) {
  if (!h || !count) return set_error(B2_ERR_ARG, "null argument");
  *count = (int)h->tries.size();
  for (int i = 0; i < std::min(cap, *count); ++i) tries[i] = h->tries[i];
  return B2_OK;
}
int b2_icp_get_pair_info(b2_icp* h, int k, int* src, int* tgt, uint64_t* count) {
  if (!h || k < 0 || k >= (int)h->segs_host.size()) return set_error(B2_ERR_ARG, "bad pair index %d", k);
  if (src) *src = h->segs_host[k].src;
  if (tgt) *tgt = h->segs_host[k].tgt;
  if (count) *count = h->segs_host[k].end - h->segs_host[k].begin;
  return B2_OK;
}

int b2_icp_get_pair_correspondences(b2_icp* h, int k, int32_t* iq, int32_t* im, float* dist) {
  if (!h || k < 0 || k >= (int)h->segs_host.size() || !iq || !im || !dist) return set_error(B2_ERR_ARG, "bad argument");
  B2_CUDA(cudaSetDevice(h->device));
  Direction* d = h->dirs[h->seg_dir[k]].get();
  Cloud* S = impl_cloud(h, d->src);
  const size_t ns = S->n;
  B2_TRY(h->scatter_m.ensure(ns * 4)); B2_TRY(h->scatter_d.ensure(ns * 4));
  k_scatter_matches<<<div_up(ns, 256), 256, 0, h->stream>>>(S->s_xyz.as<float4>(), ns, d->key.as<unsigned long long>(), h->last_init_key,
                                                           h->scatter_m.as<int>(), h->scatter_d.as<float>());
  std::vector<int> m(ns); std::vector<float> dd(ns);
  B2_CUDA(cudaMemcpyAsync(m.data(), h->scatter_m.p, ns * 4, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaMemcpyAsync(dd.data(), h->scatter_d.p, ns * 4, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  size_t o = 0;
  for (size_t i = 0; i < ns; ++i) if (m[i] >= 0) { iq[o] = (int32_t)i; im[o] = m[i]; dist[o] = dd[i]; ++o; }
  return B2_OK;
}

int b2_icp_get_normal_equations(b2_icp* h, double* H, double* b, double* cost, int* nv) {
  if (!h) return set_error(B2_ERR_ARG, "null argument");
  const int n = (int)h->b0.size();
  if (nv) *nv = n;
  if (H) std::copy(h->H0.begin(), h->H0.end(), H);
  if (b) std::copy(h->b0.begin(), h->b0.end(), b);
  if (cost) *cost = h->cost0;
  return B2_OK;
}

int b2_find_correspondences(const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt, float max_dist, int32_t* iq, int32_t* im,
                            float* dist, uint64_t* count) {
  if (!count || (n_src && (!src_xyz || !iq || !im || !dist)) || (n_tgt && !tgt_xyz)) return set_error(B2_ERR_ARG, "null argument");
  *count = 0;
  b2_icp_config cfg; b2_icp_default_config(&cfg);
  b2_icp* h = nullptr;
  B2_TRY(b2_icp_create(&cfg, &h));
  int rc = B2_OK;
  do {
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    // the handle owns both clouds from the start, so every exit path releases them; normals are irrelevant for the search: xyz stands in
    h->movable.emplace_back(new Cloud()); h->movable.emplace_back(new Cloud());
    Cloud* S = h->movable[0].get(); Cloud* T = h->movable[1].get();
    std::memcpy(S->T, I, sizeof(I)); std::memcpy(T->T, I, sizeof(I));
    if ((rc = upload_cloud(h, S, src_xyz, src_xyz, n_src, 12, false)) != B2_OK) break;
    if ((rc = upload_cloud(h, T, tgt_xyz, tgt_xyz, n_tgt, 12, false)) != B2_OK) break;
    if (n_src == 0 || n_tgt == 0) break;
    // run only the index + search part of an outer iteration
    bool conv;
    if ((rc = align_once(h, max_dist, 0.f, 0, &conv, /*search_only=*/true, /*gate=*/false)) != B2_OK) break;
    for (int k = 0; k < (int)h->segs_host.size(); ++k) {
      if (h->segs_host[k].src == 0 && h->segs_host[k].tgt == 1) {
        *count = h->segs_host[k].end - h->segs_host[k].begin;
        rc = b2_icp_get_pair_correspondences(h, k, iq, im, dist);
      }
    }
  } while (false);
  b2_icp_destroy(h);
  return rc;
}

}  // extern "C"
