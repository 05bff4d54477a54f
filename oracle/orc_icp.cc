// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see orc_math.h header).
//
// CPU restatement of the reference's multi-scan point-to-plane ICP (Path A):
//   FindCorrespondencesFast        /root/reference/src/icp/icp_point_to_plane.cc:42-105
//   PointToPlaneICP::AddPointCloud /root/reference/src/icp/icp_point_to_plane.cc:109-135
//   PointToPlaneICP::Run           /root/reference/src/icp/icp_point_to_plane.cc:137-163
//   AlignMeshes                    /root/reference/src/icp/icp_point_to_plane.cc:169-342
//   PointToPlaneICPImpl::Accumulate /root/reference/src/icp/icp_point_to_plane_impl.h:82-113
//   PointToPlaneICPImpl::compute    /root/reference/src/icp/icp_point_to_plane_impl.h:115-293
// Third-party arithmetic not in /root/reference (PCL 1.10 transforms, FLANN 1.9.1 kd-tree, Eigen 3.3.7 LDLT)
// is restated from its published behaviour; index-level NN parity and the fp32 op order of
// pcl::transformPointCloudWithNormals are UNPINNED (no reference test pins them) — the oracle
// defines them and says so. What IS pinned: the reference's own ICP tests (src/opt/test/test_icp.cc:39-172),
// ported in tests/test_oracle_icp.py.
//
// Same parallel structure as the reference: OpenMP over the n*n pair loop only (:208); the accumulate and LM
// cost loops are serial (impl.h:129-285). Differences (documented): correspondence sets are stored in ik order
// instead of critical-section arrival order (the reference order is nondeterministic).
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "orc_api.h"
#include "orc_kdtree.h"
#include "orc_math.h"

namespace orc {

// pcl::transformPointCloudWithNormals, dense float path (PCL 1.10 common/impl/transforms.hpp, SSE2 Transformer):
//   se3: x*c0 + (y*c1 + (z*c2 + c3));   so3: x*c0 + (y*c1 + z*c2)       [from memory of PCL 1.10; UNPINNED]
static void transform_cloud(const float* xyz, const float* nrm, size_t n, const Affine3f& T, float* oxyz, float* onrm) {
  for (size_t i = 0; i < n; ++i) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    for (int r = 0; r < 3; ++r) oxyz[3 * i + r] = x * T.at(r, 0) + (y * T.at(r, 1) + (z * T.at(r, 2) + T.at(r, 3)));
    if (nrm) {
      const float a = nrm[3 * i], b = nrm[3 * i + 1], c = nrm[3 * i + 2];
      for (int r = 0; r < 3; ++r) onrm[3 * i + r] = a * T.at(r, 0) + (b * T.at(r, 1) + c * T.at(r, 2));
    }
  }
}

struct Corr { int q, m; float d2; };

// FindCorrespondencesFast (icp_point_to_plane.cc:42-105): tree on target, nearest within radius per source point,
// output in ascending source index.
// query_stride > 1 is a SAMPLING knob for the bounded CPU-baseline timing only (bench.py): every stride-th source point.
static void find_correspondences(const float* src, size_t ns, const float* tgt, size_t nt, float max_dist,
                                 bool use_kdtree, std::vector<Corr>* out, size_t query_stride = 1) {
  out->clear();
  const float r2 = (float)((double)max_dist * (double)max_dist);
  if (use_kdtree) {
    KdTree tree;
    tree.build(tgt, nt, 3, 15);
    for (size_t i = 0; i < ns; i += query_stride) {
      float d2;
      const int m = tree.nearest_within(src + 3 * i, r2, &d2);
      if (m >= 0) out->push_back(Corr{(int)i, m, d2});
    }
  } else {
    for (size_t i = 0; i < ns; ++i) {
      float d2;
      const int m = brute_nearest_within(tgt, nt, 3, src + 3 * i, r2, &d2);
      if (m >= 0) out->push_back(Corr{(int)i, m, d2});
    }
  }
}

struct ImplCloud {
  const float* xyz; const float* nrm; size_t n;  // global-frame cloud of this outer iteration
  SE3f T;                                        // increment, starts at identity (impl.h:59)
};
struct CorrSet { int src, tgt; std::vector<Corr> corr; };

struct InnerStats {
  int inner_iterations = 0;   // number of accumulate passes executed
  int lm_tries_total = 0;     // number of cost passes executed
  double first_cost = 0, last_cost = 0, final_lambda = 0;
  std::vector<int> tries_per_iteration;  // tries used in each inner iteration (10 & not applied => abort)
  std::vector<double> H0, b0;            // normal equations of the first inner iteration (full matrix as written)
  double t_acc = 0, t_cost = 0;          // wall seconds in accumulate passes / cost passes
};

static inline void rigid(const float R[9], const V3f& t, const float* p, V3f* out) {
  const M3f& M = *reinterpret_cast<const M3f*>(R);
  const V3f v = mul(M, V3f{p[0], p[1], p[2]});
  *out = {v.x + t.x, v.y + t.y, v.z + t.z};
}
static inline void rot(const float R[9], const float* p, V3f* out) {
  const M3f& M = *reinterpret_cast<const M3f*>(R);
  *out = mul(M, V3f{p[0], p[1], p[2]});
}

// The four 6-vectors of one correspondence, fp32, op order of impl.h:161-204 (left-to-right evaluation).
static inline void jacobians(const V3f& ps, const V3f& ns, const V3f& pt, const V3f& nt,
                             float* r_src, float j_src_t[6], float j_src_s[6],
                             float* r_tgt, float j_tgt_t[6], float j_tgt_s[6]) {
  *r_src = dot(ns, sub(pt, ps));
  j_src_t[0] = ns.x; j_src_t[1] = ns.y; j_src_t[2] = ns.z;
  j_src_t[3] = -ns.y * pt.z + ns.z * pt.y;
  j_src_t[4] = ns.x * pt.z - ns.z * pt.x;
  j_src_t[5] = -ns.x * pt.y + ns.y * pt.x;
  j_src_s[0] = -ns.x; j_src_s[1] = -ns.y; j_src_s[2] = -ns.z;
  j_src_s[3] = ns.y * ps.z - ns.y * (ps.z - pt.z) - ns.z * ps.y + ns.z * (ps.y - pt.y);
  j_src_s[4] = -ns.x * ps.z + ns.x * (ps.z - pt.z) + ns.z * ps.x - ns.z * (ps.x - pt.x);
  j_src_s[5] = ns.x * ps.y - ns.x * (ps.y - pt.y) - ns.y * ps.x + ns.y * (ps.x - pt.x);

  *r_tgt = dot(nt, sub(ps, pt));
  j_tgt_t[0] = -nt.x; j_tgt_t[1] = -nt.y; j_tgt_t[2] = -nt.z;
  j_tgt_t[3] = nt.y * pt.z - nt.y * (pt.z - ps.z) - nt.z * pt.y + nt.z * (pt.y - ps.y);
  j_tgt_t[4] = -nt.x * pt.z + nt.x * (pt.z - ps.z) + nt.z * pt.x - nt.z * (pt.x - ps.x);
  j_tgt_t[5] = nt.x * pt.y - nt.x * (pt.y - ps.y) - nt.y * pt.x + nt.y * (pt.x - ps.x);
  j_tgt_s[0] = nt.x; j_tgt_s[1] = nt.y; j_tgt_s[2] = nt.z;
  j_tgt_s[3] = -nt.y * ps.z + nt.z * ps.y;
  j_tgt_s[4] = nt.x * ps.z - nt.z * ps.x;
  j_tgt_s[5] = -nt.x * ps.y + nt.y * ps.x;
}

// Accumulate (impl.h:82-113); weight == 1. H is column-major nv x nv, written exactly where the reference writes.
static inline void accumulate(double residual, int sv, const float js[6], int tv, const float jt[6],
                              std::vector<double>* H, std::vector<double>* b, int nv) {
  double s[6], t[6];
  for (int i = 0; i < 6; ++i) { s[i] = (double)js[i]; t[i] = (double)jt[i]; }
  auto h = [&](int r, int c) -> double& { return (*H)[(size_t)c * nv + r]; };
  if (sv >= 0) {
    for (int c = 0; c < 6; ++c) for (int r = 0; r <= c; ++r) h(sv + r, sv + c) += s[r] * s[c];
    for (int i = 0; i < 6; ++i) (*b)[sv + i] += residual * s[i];
    if (tv >= 0) for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) h(sv + r, tv + c) += s[r] * t[c];
  }
  if (tv >= 0) {
    for (int c = 0; c < 6; ++c) for (int r = 0; r <= c; ++r) h(tv + r, tv + c) += t[r] * t[c];
    for (int i = 0; i < 6; ++i) (*b)[tv + i] += residual * t[i];
  }
}

static double cost_pass(const std::vector<ImplCloud>& clouds, const std::vector<CorrSet>& sets) {
  double cost = 0.0;
  for (const CorrSet& cs : sets) {
    const ImplCloud& S = clouds[cs.src]; const ImplCloud& T = clouds[cs.tgt];
    float Rs[9], Rt[9]; quat_to_matrix(S.T.q, Rs); quat_to_matrix(T.T.q, Rt);
    for (const Corr& c : cs.corr) {
      V3f ps, ns, pt, nt;
      rigid(Rs, S.T.t, S.xyz + 3 * (size_t)c.q, &ps); rot(Rs, S.nrm + 3 * (size_t)c.q, &ns);
      rigid(Rt, T.T.t, T.xyz + 3 * (size_t)c.m, &pt); rot(Rt, T.nrm + 3 * (size_t)c.m, &nt);
      const float r1 = dot(ns, sub(pt, ps));
      cost += r1 * r1;
      const float r2 = dot(nt, sub(ps, pt));
      cost += r2 * r2;
    }
  }
  return cost;
}

// compute() (impl.h:115-293).
static void inner_compute(std::vector<ImplCloud>* clouds_io, const std::vector<CorrSet>& sets, int max_iterations,
                          InnerStats* st) {
  std::vector<ImplCloud>& clouds = *clouds_io;
  const int nv = 6 * ((int)clouds.size() - 1);
  double lambda = 0.1;
  for (int iteration = 0; iteration < max_iterations; ++iteration) {
    std::vector<double> H((size_t)nv * nv, 0.0), b(nv, 0.0);
    double cost = 0.0;
    const double ta0 = omp_get_wtime();
    for (const CorrSet& cs : sets) {
      const int sv = 6 * (cs.src - 1), tv = 6 * (cs.tgt - 1);
      const ImplCloud& S = clouds[cs.src]; const ImplCloud& T = clouds[cs.tgt];
      float Rs[9], Rt[9]; quat_to_matrix(S.T.q, Rs); quat_to_matrix(T.T.q, Rt);
      for (const Corr& c : cs.corr) {
        V3f ps, ns, pt, nt;
        rigid(Rs, S.T.t, S.xyz + 3 * (size_t)c.q, &ps); rot(Rs, S.nrm + 3 * (size_t)c.q, &ns);
        rigid(Rt, T.T.t, T.xyz + 3 * (size_t)c.m, &pt); rot(Rt, T.nrm + 3 * (size_t)c.m, &nt);
        float r1, r2, j1t[6], j1s[6], j2t[6], j2s[6];
        jacobians(ps, ns, pt, nt, &r1, j1t, j1s, &r2, j2t, j2s);
        cost += r1 * r1;
        accumulate((double)r1, sv, j1s, tv, j1t, &H, &b, nv);
        cost += r2 * r2;
        accumulate((double)r2, sv, j2s, tv, j2t, &H, &b, nv);
      }
    }
    st->t_acc += omp_get_wtime() - ta0;
    st->inner_iterations++;
    if (iteration == 0) { st->first_cost = cost; st->H0 = H; st->b0 = b; }
    st->last_cost = cost;

    bool applied = false;
    int tries = 0;
    for (int lm = 0; lm < 10; ++lm) {
      ++tries;
      std::vector<double> HL = H;
      for (int i = 0; i < nv; ++i) HL[(size_t)i * nv + i] += lambda;
      std::vector<double> x;
      ldlt_solve_upper(HL, nv, b, &x);
      std::vector<ImplCloud> upd = clouds;
      for (size_t ci = 1; ci < clouds.size(); ++ci) {
        double neg[6];
        for (int k = 0; k < 6; ++k) neg[k] = -x[6 * (ci - 1) + k];
        upd[ci].T = se3_mul(se3d_exp_cast_float(neg), clouds[ci].T);
      }
      const double tc0 = omp_get_wtime();
      const double new_cost = cost_pass(upd, sets);
      st->t_cost += omp_get_wtime() - tc0;
      st->lm_tries_total++;
      if (new_cost < cost) {
        clouds = upd;
        lambda = 0.5f * lambda;
        applied = true;
        st->last_cost = new_cost;
        break;
      } else {
        lambda = 2.f * lambda;
      }
    }
    st->tries_per_iteration.push_back(tries);
    if (!applied) break;
  }
  st->final_lambda = lambda;
}

struct MovCloud {
  std::vector<float> xyz, nrm;          // local frame (caller's data, copied: the oracle keeps no caller pointers)
  Affine3f global_T_cloud;
  std::vector<float> gxyz, gnrm;        // global frame of the current outer iteration
  float bmin[3], bmax[3];
  int cloud_index = -1;
};

static inline bool boxes_intersect(const float* amin, const float* amax, const float* bmin, const float* bmax) {
  // Eigen AlignedBox::intersection(...).isEmpty(): empty iff any (max(min) > min(max)); empty boxes have min=+inf,max=-inf.
  for (int d = 0; d < 3; ++d) if (std::max(amin[d], bmin[d]) > std::min(amax[d], bmax[d])) return false;
  return true;
}
static void bbox_of(const std::vector<float>& xyz, float* mn, float* mx) {
  for (int d = 0; d < 3; ++d) { mn[d] = std::numeric_limits<float>::infinity(); mx[d] = -mn[d]; }
  for (size_t i = 0; i < xyz.size() / 3; ++i)
    for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], xyz[3 * i + d]); mx[d] = std::max(mx[d], xyz[3 * i + d]); }
}

}  // namespace orc

using namespace orc;

struct orc_icp {
  std::vector<MovCloud> clouds;
  bool has_fixed = false;
  std::vector<float> fixed_xyz, fixed_nrm;
  bool use_kdtree = true;
  int inner_max_iterations = 150;
  size_t query_stride = 1;
  // last AlignMeshes
  std::vector<CorrSet> last_sets;
  InnerStats last_stats;
  std::vector<float> last_movement;
  double t_transform = 0, t_search = 0, t_inner = 0;
};

static bool align_meshes(orc_icp* h, float max_dist, float thr) {
  std::vector<ImplCloud> impl;
  int fixed_vertex = -1;
  float fmin[3], fmax[3];
  const double t0 = omp_get_wtime();
  if (h->has_fixed) {
    fixed_vertex = (int)impl.size();
    impl.push_back(ImplCloud{h->fixed_xyz.data(), h->fixed_nrm.data(), h->fixed_xyz.size() / 3, SE3f()});
    bbox_of(h->fixed_xyz, fmin, fmax);
  }
  for (MovCloud& c : h->clouds) {
    const size_t n = c.xyz.size() / 3;
    c.gxyz.resize(3 * n); c.gnrm.resize(3 * n);
    transform_cloud(c.xyz.data(), c.nrm.data(), n, c.global_T_cloud, c.gxyz.data(), c.gnrm.data());
    c.cloud_index = (int)impl.size();
    impl.push_back(ImplCloud{c.gxyz.data(), c.gnrm.data(), n, SE3f()});
    bbox_of(c.gxyz, c.bmin, c.bmax);
  }
  const double t1 = omp_get_wtime();
  const int n = (int)h->clouds.size();
  // slots in ik order: [ik*3 + 0] = i->k (or i->fixed when i==k), [ik*3+1] = fixed->i
  std::vector<CorrSet> slots((size_t)n * n * 2);
  std::vector<char> used((size_t)n * n * 2, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int ik = 0; ik < n * n; ++ik) {
    const int i = ik / n, k = ik % n;
    MovCloud& A = h->clouds[i]; MovCloud& B = h->clouds[k];
    if (i != k && boxes_intersect(A.bmin, A.bmax, B.bmin, B.bmax)) {
      CorrSet cs; cs.src = A.cloud_index; cs.tgt = B.cloud_index;
      find_correspondences(A.gxyz.data(), A.gxyz.size() / 3, B.gxyz.data(), B.gxyz.size() / 3, max_dist, h->use_kdtree, &cs.corr, h->query_stride);
      if (!cs.corr.empty()) { slots[2 * (size_t)ik] = std::move(cs); used[2 * (size_t)ik] = 1; }
    }
    if (i == k && h->has_fixed && boxes_intersect(fmin, fmax, A.bmin, A.bmax)) {
      CorrSet cs; cs.src = A.cloud_index; cs.tgt = fixed_vertex;
      find_correspondences(A.gxyz.data(), A.gxyz.size() / 3, h->fixed_xyz.data(), h->fixed_xyz.size() / 3, max_dist, h->use_kdtree, &cs.corr, h->query_stride);
      if (!cs.corr.empty()) { slots[2 * (size_t)ik] = std::move(cs); used[2 * (size_t)ik] = 1; }
      CorrSet cs2; cs2.src = fixed_vertex; cs2.tgt = A.cloud_index;
      find_correspondences(h->fixed_xyz.data(), h->fixed_xyz.size() / 3, A.gxyz.data(), A.gxyz.size() / 3, max_dist, h->use_kdtree, &cs2.corr, h->query_stride);
      if (!cs2.corr.empty()) { slots[2 * (size_t)ik + 1] = std::move(cs2); used[2 * (size_t)ik + 1] = 1; }
    }
  }
  h->last_sets.clear();
  for (size_t s = 0; s < slots.size(); ++s) if (used[s]) h->last_sets.push_back(std::move(slots[s]));
  const double t2 = omp_get_wtime();

  h->last_stats = InnerStats();
  inner_compute(&impl, h->last_sets, h->inner_max_iterations, &h->last_stats);
  const double t3 = omp_get_wtime();
  h->t_transform += t1 - t0; h->t_search += t2 - t1; h->t_inner += t3 - t2;

  bool converged = true;
  h->last_movement.clear();
  for (MovCloud& c : h->clouds) {
    const Affine3f upd = se3_to_affine(impl[c.cloud_index].T);
    const Affine3f nw = affine_mul(upd, c.global_T_cloud);
    const V3f dt{c.global_T_cloud.at(0, 3) - nw.at(0, 3), c.global_T_cloud.at(1, 3) - nw.at(1, 3), c.global_T_cloud.at(2, 3) - nw.at(2, 3)};
    const float movement = std::sqrt(dot(dt, dt));
    if (movement > thr) converged = false;
    h->last_movement.push_back(movement);
    c.global_T_cloud = nw;
  }
  return converged;
}

extern "C" {

orc_icp* orc_icp_create(void) { return new orc_icp(); }
void orc_icp_destroy(orc_icp* h) { delete h; }
void orc_icp_set_options(orc_icp* h, int use_kdtree, int inner_max_iterations) {
  h->use_kdtree = use_kdtree != 0;
  if (inner_max_iterations > 0) h->inner_max_iterations = inner_max_iterations;
}
void orc_icp_set_query_stride(orc_icp* h, size_t stride) { h->query_stride = stride ? stride : 1; }

int orc_icp_add_cloud(orc_icp* h, const float* xyz, const float* nrm, size_t n, const float T_colmajor[16], int fixed) {
  Affine3f T; std::memcpy(T.m, T_colmajor, sizeof(T.m));
  if (fixed) {
    // icp_point_to_plane.cc:112-127: transform to global frame and concatenate.
    const size_t off = h->fixed_xyz.size();
    h->fixed_xyz.resize(off + 3 * n); h->fixed_nrm.resize(off + 3 * n);
    transform_cloud(xyz, nrm, n, T, h->fixed_xyz.data() + off, h->fixed_nrm.data() + off);
    h->has_fixed = true;
    return -1;
  }
  MovCloud c;
  c.xyz.assign(xyz, xyz + 3 * n); c.nrm.assign(nrm, nrm + 3 * n);
  c.global_T_cloud = T;
  h->clouds.push_back(std::move(c));
  return (int)h->clouds.size() - 1;
}

int orc_icp_run(orc_icp* h, float max_dist, int initial_iteration, int max_iters, float thr, int* converged) {
  if (h->clouds.empty()) return 1;  // reference: CHECK(!clouds_.empty()) aborts (icp_point_to_plane.cc:142)
  *converged = 0;
  for (int i = initial_iteration; i < initial_iteration + max_iters; ++i) {
    if (align_meshes(h, max_dist, thr)) { *converged = 1; return 0; }
  }
  return 0;
}

int orc_icp_get_pose(orc_icp* h, int id, float T[16]) {
  if (id < 0 || id >= (int)h->clouds.size()) return 1;
  std::memcpy(T, h->clouds[id].global_T_cloud.m, sizeof(float) * 16);
  return 0;
}
int orc_icp_set_pose(orc_icp* h, int id, const float T[16]) {
  if (id < 0 || id >= (int)h->clouds.size()) return 1;
  std::memcpy(h->clouds[id].global_T_cloud.m, T, sizeof(float) * 16);
  return 0;
}

void orc_icp_last_stats(orc_icp* h, orc_icp_stats* s) {
  s->inner_iterations = h->last_stats.inner_iterations;
  s->lm_tries_total = h->last_stats.lm_tries_total;
  s->first_cost = h->last_stats.first_cost;
  s->last_cost = h->last_stats.last_cost;
  s->final_lambda = h->last_stats.final_lambda;
  s->num_pairs = (int)h->last_sets.size();
  size_t tot = 0; for (const CorrSet& cs : h->last_sets) tot += cs.corr.size();
  s->num_correspondences = (uint64_t)tot;
  s->num_variables = 6 * ((int)h->clouds.size() + (h->has_fixed ? 1 : 0) - 1);
  s->t_transform = h->t_transform; s->t_search = h->t_search; s->t_inner = h->t_inner;
  s->t_acc = h->last_stats.t_acc; s->t_cost = h->last_stats.t_cost;
}
int orc_icp_last_tries(orc_icp* h, int* tries, int cap) {
  const int n = (int)h->last_stats.tries_per_iteration.size();
  for (int i = 0; i < std::min(n, cap); ++i) tries[i] = h->last_stats.tries_per_iteration[i];
  return n;
}
int orc_icp_last_pair_info(orc_icp* h, int k, int* src, int* tgt, uint64_t* count) {
  if (k < 0 || k >= (int)h->last_sets.size()) return 1;
  *src = h->last_sets[k].src; *tgt = h->last_sets[k].tgt; *count = h->last_sets[k].corr.size();
  return 0;
}
int orc_icp_last_pair_corr(orc_icp* h, int k, int* q, int* m, float* d2) {
  if (k < 0 || k >= (int)h->last_sets.size()) return 1;
  const std::vector<Corr>& v = h->last_sets[k].corr;
  for (size_t i = 0; i < v.size(); ++i) { q[i] = v[i].q; m[i] = v[i].m; d2[i] = v[i].d2; }
  return 0;
}
// Effective normal equations of the first inner iteration: the UPPER triangle the solver reads (mirrored to full).
int orc_icp_last_normal_eq(orc_icp* h, double* H, double* b) {
  const int nv = (int)h->last_stats.b0.size();
  for (int c = 0; c < nv; ++c) for (int r = 0; r < nv; ++r) {
    const double v = (r <= c) ? h->last_stats.H0[(size_t)c * nv + r] : h->last_stats.H0[(size_t)r * nv + c];
    H[(size_t)c * nv + r] = v;
  }
  for (int i = 0; i < nv; ++i) b[i] = h->last_stats.b0[i];
  return nv;
}

void orc_transform_cloud(const float* xyz, const float* nrm, size_t n, const float T[16], float* oxyz, float* onrm) {
  Affine3f A; std::memcpy(A.m, T, sizeof(A.m));
  transform_cloud(xyz, nrm, n, A, oxyz, onrm);
}

uint64_t orc_find_correspondences(const float* src, size_t ns, const float* tgt, size_t nt, float max_dist,
                                  int use_kdtree, int* q, int* m, float* d2) {
  std::vector<Corr> v;
  find_correspondences(src, ns, tgt, nt, max_dist, use_kdtree != 0, &v);
  for (size_t i = 0; i < v.size(); ++i) { q[i] = v[i].q; m[i] = v[i].m; d2[i] = v[i].d2; }
  return v.size();
}

// Timing probe for the CPU baseline: kd-tree build on the target + nearest-within-radius for the given queries, one thread.
void orc_time_search(const float* src, size_t ns, const float* tgt, size_t nt, float max_dist, double* t_build, double* t_query,
                     uint64_t* matched) {
  const float r2 = (float)((double)max_dist * (double)max_dist);
  double t0 = omp_get_wtime();
  KdTree tree; tree.build(tgt, nt, 3, 15);
  double t1 = omp_get_wtime();
  uint64_t m = 0;
  for (size_t i = 0; i < ns; ++i) { float d2; if (tree.nearest_within(src + 3 * i, r2, &d2) >= 0) ++m; }
  double t2 = omp_get_wtime();
  *t_build = t1 - t0; *t_query = t2 - t1; *matched = m;
}

void orc_se3_exp_left_mul(const double x[6], const float q_in[4], const float t_in[3], float q_out[4], float t_out[3]) {
  SE3f T; T.q = {q_in[0], q_in[1], q_in[2], q_in[3]}; T.t = {t_in[0], t_in[1], t_in[2]};
  const SE3f r = se3_mul(se3d_exp_cast_float(x), T);
  q_out[0] = r.q.x; q_out[1] = r.q.y; q_out[2] = r.q.z; q_out[3] = r.q.w;
  t_out[0] = r.t.x; t_out[1] = r.t.y; t_out[2] = r.t.z;
}

int orc_ldlt_solve_upper(const double* A, int n, const double* b, double* x) {
  std::vector<double> Av(A, A + (size_t)n * n), bv(b, b + n), xv;
  ldlt_solve_upper(Av, n, bv, &xv);
  for (int i = 0; i < n; ++i) x[i] = xv[i];
  return 0;
}

}  // extern "C"
