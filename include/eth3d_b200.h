/* eth3d_b200.h — C ABI of libeth3d_b200.so: B200-native (sm_100a) hot paths of ETH3D/dataset-pipeline.
 *
 * The reference has no plugin/FFI layer: its hot paths are C++ classes linked into BaseLib. The entry points below
 * are the narrowest seams every caller (the three tools + the tests) already goes through; each declaration cites the
 * reference interface it replaces. Conventions:
 *   - plain pointers and sizes only; no C++/torch types; opaque handles;
 *   - every function returns an int status (B2_OK = 0). Nothing aborts or throws across the boundary (the reference
 *     uses glog CHECK/LOG(FATAL), e.g. icp_point_to_plane.cc:142 — here that is B2_ERR_STATE);
 *   - 4x4 transforms are 16 floats COLUMN-major (= Eigen::Affine3f / Matrix4f storage);
 *   - se(3) tangent order is [translation(3); rotation(3)], updates are left-multiplicative with negated x
 *     (icp_point_to_plane_impl.h:162-168,235);
 *   - host pointers unless a parameter is named *_dev; a handle is thread-compatible, not thread-safe;
 *   - the library fails loudly (B2_ERR_CUDA / B2_ERR_NO_DEVICE) when no sm_100 device is present: there is NO CPU
 *     fallback.
 */
#ifndef ETH3D_B200_H_
#define ETH3D_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_ABI_VERSION 1

enum {
  B2_OK = 0,
  B2_ERR_ARG = 1,        /* null pointer, bad size, bad index */
  B2_ERR_STATE = 2,      /* call order violated (e.g. run with no movable cloud; reference CHECK at icp_point_to_plane.cc:142) */
  B2_ERR_CUDA = 3,       /* a CUDA call failed; b2_last_error() has the text */
  B2_ERR_NO_DEVICE = 4,  /* no CUDA device / not sm_100 */
  B2_ERR_ALLOC = 5,
  B2_ERR_COMM = 6        /* the allreduce hook reported failure */
};

/* Text of the last error on this thread ("" if none). */
const char* b2_last_error(void);
int b2_abi_version(void);
/* Device index the library runs on, name, SM count. B2_ERR_NO_DEVICE when there is none. */
int b2_device_info(int* device, char* name, size_t name_cap, int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------------------------
 * Path A — multi-scan point-to-plane ICP.
 * Replaces icp::PointToPlaneICP (src/icp/icp_point_to_plane.h:39-57) = FindCorrespondencesFast
 * (icp_point_to_plane.cc:42-105) + AlignMeshes (:169-342) + PointToPlaneICPImpl::compute
 * (icp_point_to_plane_impl.h:115-293).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct b2_icp b2_icp;

/* Data-parallel exchange hook (multi-GPU). Called once per normal-equation pass with a DEVICE buffer of `count`
 * doubles that must be sum-reduced in place across all ranks, ordered on `stream` (a cudaStream_t). Return 0 on
 * success. NULL = single GPU. The host language supplies NCCL (ncclAllReduce(sum, double) over NVLink). */
typedef int (*b2_allreduce_fn)(void* user, double* buf_dev, size_t count, void* stream);

typedef struct b2_icp_config {
  int32_t device;               /* CUDA device ordinal; -1 = current device */
  int32_t inner_max_iterations; /* 0 = reference value 150 (icp_point_to_plane.cc:312) */
  int32_t keep_correspondences; /* !=0: keep per-pair (query,match,d2) lists for b2_icp_get_pair_correspondences */
  int32_t rank, world_size;     /* pair-direction sharding: this handle searches/accumulates directions k with
                                   k % world_size == rank; 0/1 = everything */
  b2_allreduce_fn allreduce;    /* required when world_size > 1 */
  void* allreduce_user;
  void* stream;                 /* cudaStream_t to run on; NULL = a stream owned by the handle */
} b2_icp_config;

typedef struct b2_icp_stats {
  int32_t inner_iterations;     /* LM iterations of the last outer iteration (impl.h:119) */
  int32_t lm_tries_total;       /* LM tries = cost evaluations of the last outer iteration (impl.h:219) */
  int32_t num_pairs;            /* non-empty correspondence sets (all ranks) */
  int32_t num_variables;        /* 6 * (impl clouds - 1) */
  uint64_t num_correspondences; /* over all pairs, all ranks */
  uint64_t local_correspondences; /* on this rank */
  double first_cost, last_cost, final_lambda;
  int32_t passes;               /* streaming passes over the packed correspondences (this outer iteration) */
  int32_t kernel_launches;      /* kernels of this library launched (this outer iteration) */
  /* device time (ms, CUDA events on the handle's stream) of the last outer iteration */
  float ms_index, ms_search, ms_pack, ms_inner, ms_total;
  float ms_accum_kernel_avg;    /* average duration of one accumulate-pass kernel */
} b2_icp_stats;

void b2_icp_default_config(b2_icp_config* cfg);
int b2_icp_create(const b2_icp_config* cfg, b2_icp** out);
int b2_icp_destroy(b2_icp* h);

/* AddPointCloud (icp_point_to_plane.h:43-46, .cc:109-135). xyz / normals: n points, `stride_bytes` between
 * consecutive points in each array (12 for packed float3; 48 for both pointing into a pcl::PointNormal array at
 * offsets 0 and 16). The data is copied to the device; the caller keeps ownership. fixed != 0: the cloud is
 * transformed to the global frame and concatenated to the fixed cloud, *out_id = -1 (reference behaviour). */
int b2_icp_add_cloud(b2_icp* h, const float* xyz, const float* normals, size_t n, size_t stride_bytes,
                     const float global_T_cloud[16], int fixed, int* out_id);
/* Same, from DEVICE memory (packed float3 arrays) — used when scans are already resident in HBM. */
int b2_icp_add_cloud_dev(b2_icp* h, const float* xyz_dev, const float* normals_dev, size_t n,
                         const float global_T_cloud[16], int fixed, int* out_id);

/* Run (icp_point_to_plane.h:50-55, .cc:137-163): up to max_num_iterations outer iterations, stops when every cloud
 * moved <= convergence_threshold. *converged = 1/0. print_progress prints the reference's progress lines to stdout. */
int b2_icp_run(b2_icp* h, float max_correspondence_distance, int initial_iteration, int max_num_iterations,
               float convergence_threshold_max_movement, int print_progress, int* converged);

/* GetResultGlobalTCloud (icp_point_to_plane.h:57). */
int b2_icp_get_pose(b2_icp* h, int cloud_id, float global_T_cloud[16]);
/* Not in the reference API: overwrite a pose (parity tests feed both implementations identical poses). */
int b2_icp_set_pose(b2_icp* h, int cloud_id, const float global_T_cloud[16]);

/* Introspection for parity dumps (state of the LAST outer iteration). */
int b2_icp_last_stats(b2_icp* h, b2_icp_stats* out);
int b2_icp_get_lm_tries(b2_icp* h, int32_t* tries, int cap, int* count);
int b2_icp_get_pair_info(b2_icp* h, int k, int* src_impl_index, int* tgt_impl_index, uint64_t* count);
/* Requires cfg.keep_correspondences. Ascending query index, like pcl::Correspondences out of
 * FindCorrespondencesFast: (index_query, index_match, distance = squared distance). */
int b2_icp_get_pair_correspondences(b2_icp* h, int k, int32_t* index_query, int32_t* index_match, float* distance);
/* Effective normal equations (what the solver reads: upper triangle, mirrored) of the FIRST inner iteration.
 * H: nv*nv doubles column-major, b: nv doubles. */
int b2_icp_get_normal_equations(b2_icp* h, double* H, double* b, double* cost, int* nv);

/* Pure host function (no device needed): the pair-directions AlignMeshes searches for n_movable clouds (+ the fixed cloud), in
 * the reference's ik order (icp_point_to_plane.cc:208-309, before the bbox gate), as impl-cloud indices (0 = fixed cloud if any),
 * and the rank that owns each under the round-robin sharding used by b2_icp_run. *count receives the number of directions. */
int b2_icp_plan_directions(int n_movable, int has_fixed, int world_size, int32_t* src_impl, int32_t* tgt_impl, int32_t* owner,
                           int cap, int* count);

/* Stand-alone correspondence search = FindCorrespondencesFast (icp_point_to_plane.cc:42-105) on two point sets that
 * are already in a common frame. Outputs sized for n_src; *count receives the number of correspondences. */
int b2_find_correspondences(const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt,
                            float max_correspondence_distance, int32_t* index_query, int32_t* index_match,
                            float* distance, uint64_t* count);

/* ------------------------------------------------------------------------------------------------------------------
 * Path A — kNN two-pass normal estimation.
 * Replaces pcl::NormalEstimationTwoPassOMP (src/geometry/two_pass_normal_3d_omp.h:53-99, .hpp:47-119) as driven by
 * icp_scan_aligner.cc:323-330 and normal_estimator.cc:177-194: setInputCloud, setKSearch(k), setViewPoint, compute.
 * out_nxyz_curv: n x 4 floats (normal_x, normal_y, normal_z, curvature); NaN where fewer than 3 neighbours
 * (two_pass_normal_3d.h:100-105). *is_dense = 0 if any NaN was written. out_knn_idx (nullable): n x k neighbour
 * indices sorted by (distance, index), -1 padded.
 * ------------------------------------------------------------------------------------------------------------------ */
int b2_normals_estimate(const float* xyz, size_t n, size_t stride_bytes, int k, const float viewpoint[3],
                        float* out_nxyz_curv, int32_t* out_knn_idx, int* is_dense);

#ifdef __cplusplus
}
#endif
#endif /* ETH3D_B200_H_ */
