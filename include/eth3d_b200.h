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

#define B2_ABI_VERSION 2

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
/* Device memory comes from a memory pool private to this library (one per device) that keeps what handles release, so that
 * create / destroy cycles do not re-map memory. b2_trim() returns everything the pools hold unused to the driver (call it after
 * destroying handles when the host process needs the memory back). */
int b2_trim(void);

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

/* NCCL communicator owned by the library (one process per GPU). NCCL is bound with dlopen("libnccl.so.2"), i.e. the copy a
 * PyTorch host already loaded, or the system one. Rank 0 calls b2_comm_unique_id and ships the 128 bytes to the other ranks by
 * its own means (file, MPI, torch.distributed broadcast); then every rank calls b2_comm_create. */
typedef struct b2_comm b2_comm;
int b2_comm_unique_id(unsigned char id[128]);
int b2_comm_create(int rank, int world_size, const unsigned char id[128], int device, b2_comm** out);
int b2_comm_destroy(b2_comm* c);
int b2_comm_allreduce_f64(b2_comm* c, double* buf_dev, size_t count, void* stream);   /* in-place sum, ordered on stream */
enum { B2_F64 = 0, B2_F32 = 1, B2_I32 = 2 };
int b2_comm_allreduce(b2_comm* c, void* buf_dev, size_t count, int dtype /* B2_F64 | B2_F32 | B2_I32 */, void* stream);   /* in-place sum */
int b2_comm_broadcast(b2_comm* c, void* buf_dev, size_t bytes, int root, void* stream);   /* in place, from rank `root` */
int b2_comm_info(b2_comm* c, int* rank, int* world_size);

typedef struct b2_icp_config {
  int32_t device;               /* CUDA device ordinal; -1 = current device */
  int32_t inner_max_iterations; /* 0 = reference value 150 (icp_point_to_plane.cc:312) */
  int32_t keep_correspondences; /* !=0: keep per-pair (query,match,d2) lists for b2_icp_get_pair_correspondences */
  int32_t rank, world_size;     /* pair-direction sharding: this handle searches/accumulates directions k with
                                   k % world_size == rank; 0/1 = everything */
  b2_allreduce_fn allreduce;    /* world_size > 1 needs this hook or `comm` */
  void* allreduce_user;
  void* stream;                 /* cudaStream_t to run on; NULL = a stream owned by the handle */
  b2_comm* comm;                /* library-owned NCCL communicator (preferred): ncclAllReduce(sum, double) on the handle's stream */
  float index_distance_hint;    /* > 0: the max_correspondence_distance b2_icp_run will be called with (ICPScanAligner knows it from
                                   its flags before it adds a cloud). b2_icp_add_cloud then builds a cloud's search index while the
                                   NEXT cloud's host-to-device copy is in flight instead of inside the first b2_icp_run. 0 = no hint.
                                   A wrong hint costs a rebuild, never correctness. */
  int32_t shard_uploads;        /* != 0 (needs `comm`): b2_icp_add_cloud becomes COLLECTIVE — every rank calls it for every movable cloud
                                   in the same order with the same n; only rank (cloud_id % world_size) reads its host buffers and
                                   copies them to its GPU, the other ranks receive the cloud by ncclBroadcast over NVLink (their
                                   xyz / normals arguments are ignored and may be NULL). Fixed clouds are uploaded by every rank. */
  int32_t search_ahead;         /* != 0 (the default of b2_icp_default_config; needs index_distance_hint, one rank): b2_icp_add_cloud also
                                   searches the pair-directions among the clouds added so far, at the hinted radius and the poses they
                                   were added with, on auxiliary streams behind the uploads that follow. The first b2_icp_run adopts
                                   those results if radius, poses and indexes are still the same, and searches as usual otherwise:
                                   results never differ, only when the work happens. */
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
  float ms_index, ms_search, ms_pack, ms_inner, ms_total;   /* ms_index = the per-iteration part (global-frame rows, boxes, AABBs) */
  float ms_accum_kernel_avg;    /* average duration of one accumulate-pass kernel (K5) */
  float ms_search_kernel_avg;   /* average duration of one correspondence-search kernel (K3), one pair-direction */
  int32_t search_launches;      /* pair-directions searched on this rank */
  uint64_t search_algorithmic_bytes; /* sum over those launches of 12*Q + 8*Q_matched + 12*T */
  float ms_index_build;         /* one-time static index builds that fell into this outer iteration (0 once every cloud is indexed) */
  int32_t sparse_grids;         /* clouds whose occupied-cell index uses the hash layout (grid above 2^31 cells) instead of the rank bitmap */
  uint64_t search_work[5];      /* B2_K3_WORK=1 only: candidates tested, level-1 box tests, level-2 box tests, cells scanned, queue items */
  int32_t searches_ahead;       /* pair-directions of this outer iteration whose search had been done behind the uploads (search_ahead);
                                   they are not in search_launches / ms_search / search_algorithmic_bytes */
  int32_t packs_overlapped;     /* correspondence sets of this outer iteration that were packed while later sets were still being
                                   searched (their time is inside ms_search; ms_pack is what remained after the last search) */
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

/* Not in the reference API: scheduling switches for A/B measurements (results never depend on them). Names: "pack_overlap" (1: sets are
   packed while later sets are searched; 0: after all searches), "lpt_order" (1: K3 tiles issued longest-first). */
int b2_icp_set_option(b2_icp* h, const char* name, int value);

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

/* Pure host function: the rank whose host buffers are read for movable cloud `cloud_id` (0-based, in AddPointCloud order) when
 * cfg.shard_uploads is set; the other ranks receive the cloud by broadcast. -1 for bad arguments. */
int b2_icp_upload_owner(int cloud_id, int world_size);

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
 * indices sorted by (distance, index), -1 padded. out_nxyz_curv may be NULL when out_knn_idx is given (neighbour lists only).
 * 1 <= k <= 2048 (lists of 56 and more neighbours are searched by one warp per query instead of one thread).
 * ------------------------------------------------------------------------------------------------------------------ */
int b2_normals_estimate(const float* xyz, size_t n, size_t stride_bytes, int k, const float viewpoint[3],
                        float* out_nxyz_curv, int32_t* out_knn_idx, int* is_dense);
/* The same over several GPUs (one process per GPU): every rank passes the whole cloud and gets the whole result; rank r
 * answers the r-th slice of the Morton-sorted queries and one sum-allreduce over `comm` merges the outputs. */
int b2_normals_estimate_dist(const float* xyz, size_t n, size_t stride_bytes, int k, const float viewpoint[3], b2_comm* comm, int device,
                             float* out_nxyz_curv, int* is_dense);
/* setRadiusSearch mode (two_pass_normal_3d_omp.hpp:66 with search_parameter_ = radius; normal_estimator.cc:181-182): the neighbours of
 * a point are ALL points with squared distance < (float)((double)radius*radius) (itself included), taken in the order radiusSearch
 * returns them (by distance, ties to the lower index). comm may be NULL (single GPU, device = -1 for the default).
 * out_neighbor_count (nullable): n counts. */
int b2_normals_estimate_radius(const float* xyz, size_t n, size_t stride_bytes, float radius, const float viewpoint[3], b2_comm* comm, int device,
                               float* out_nxyz_curv, int32_t* out_neighbor_count, int* is_dense);

/* ------------------------------------------------------------------------------------------------------------------
 * Point-cloud tools next to the hot paths, built on the exact kNN search of the normals.
 * ------------------------------------------------------------------------------------------------------------------ */
/* pcl::LocalStatisticalOutlierRemoval<PointT>::applyFilterIndices (src/geometry/local_statistical_outlier_removal.hpp:72-176) with
 * indices_ = the whole cloud, as PointCloudCleaner runs it (src/exe/point_cloud_cleaner.cc:80-96: setMeanK(knn),
 * setDistanceFactorThresh(factor), filter): pass 1, per point the mean distance to its mean_k nearest other points (double sum of float
 * sqrt, result float); pass 2, a point is an outlier when that mean exceeds distance_factor_threshold x the mean of its neighbours' means
 * (only means > 0 count). negative != 0 inverts the test (setNegative). Non-finite points are never part of the search and are reported
 * as removed. out_indices (capacity n): ascending indices of the kept points; out_removed_indices (nullable, capacity n): the others;
 * out_mean_distances (nullable, n): the pass-1 means, 0 for non-finite points. B2_ERR_STATE when fewer than mean_k + 1 finite points. */
int b2_lsor_filter(const float* xyz, size_t n, size_t stride_bytes, int mean_k, double distance_factor_threshold, int negative,
                   int32_t* out_indices, size_t* out_count, int32_t* out_removed_indices, size_t* out_removed_count, float* out_mean_distances);
/* igl::AABB<MatrixXf,3>::squared_distance (thirdparty/igl/AABB.cpp:344-430) per point: the minimum over the triangles of
 * igl::point_simplex_squared_distance (thirdparty/igl/point_simplex_squared_distance.cpp:44-135), fp32, libigl's operation order.
 * points: n x 3; vertices: num_vertices x 3; faces: num_faces x 3 vertex indices. */
int b2_mesh_squared_distance(const float* points, size_t n, const float* vertices, size_t num_vertices, const uint32_t* faces, size_t num_faces,
                             float* out_squared_distance);
/* SplatCreator's per-point body (src/exe/splat_creator.cc:146-215): splat radius = min(distance to the 4th nearest other point,
 * max_splat_size); right = normal.unitOrthogonal(), up = normal x right; corners = point + radius * (+-right +- up) in the order top right,
 * bottom right, bottom left, top left (out_corners: n x 4 x 3, zeros for points with a NaN normal); out_added[i] = 1 when the point or
 * one of its corners is farther than sqrt(squared_distance_threshold) from the mesh (the splat the tool appends: faces (2,1,0), (0,3,2) of
 * its four vertices). Splats are reported in point order (the reference's `omp parallel for` appends them in arrival order).
 * xyz / normals: n points, stride_bytes apart in each array. out_radius, out_splat_count nullable. */
int b2_splat_create(const float* xyz, const float* normals, size_t n, size_t stride_bytes, const float* vertices, size_t num_vertices,
                    const uint32_t* faces, size_t num_faces, float max_splat_size, float squared_distance_threshold, float* out_corners,
                    uint8_t* out_added, float* out_radius, size_t* out_splat_count);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-resolution point cloud — the producer of Path B's point scales and neighbour indices.
 * Replaces opt::MergeClosePoints (src/opt/multi_scale_point_cloud.cc:44-124), the scale loop of
 * opt::CreateMultiScalePointCloud (:263-368) and opt::Problem::DeterminePointNeighbors (src/opt/problem.cc:706-786) as driven by
 * Problem::AddPointCloud... (problem.cc:161-362). Host buffers in and out; colours are the grey values of PreprocessScans (:186-212).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct b2_ms_stats {
  uint64_t neighbor_pairs;   /* (i, j > i) pairs closer than the merge distance, summed over the scales */
  int32_t rounds;            /* dependency rounds of the centre selection, summed over the scales */
  int32_t scales;
  float ms_device;           /* device time of the merges (CUDA events) */
} b2_ms_stats;
/* MergeClosePoints: in index order a point not yet merged becomes a centre and averages ALL points with squared distance
 * < (float)((double)merge_distance^2) (radiusSearch order: by distance, ties to the lower index): position = fp32 sum / count; colour =
 * mean over the scan contributing most points (the first to reach the maximal count); max_radius = maximum; every merged point is
 * marked done. Outputs sized for n points; *out_n = merged points. num_scans <= 32. stats nullable. */
int b2_ms_merge_close_points(const float* xyz, size_t n, const float* colors, const uint8_t* scan_indices, const float* max_radius, int num_scans,
                             float merge_distance, float* out_xyz, float* out_colors, uint8_t* out_scan_indices, float* out_max_radius, size_t* out_n,
                             b2_ms_stats* stats);
/* CreateMultiScalePointCloud after ComputeMinMaxPointRadius: min_radius / max_radius per point (+inf / -inf where no image observes the
 * point). radius_0 = min(min_radius) * min_radius_bias, doubled per scale until radius >= 0.99 * max(max_radius); scale s merges, at
 * distance merge_distance_factor * radius_s, the survivors of scale s-1 (radius_s <= their max_radius) followed by the input points whose
 * min_radius was passed. Outputs concatenated over the scales (out_capacity points in total); out_radius / out_counts hold max_scales. */
int b2_ms_create(const float* xyz, size_t n, const float* colors, const uint8_t* scan_indices, const float* min_radius, const float* max_radius,
                 int num_scans, float min_radius_bias, float merge_distance_factor, int max_scales, size_t out_capacity, int* out_scale_count,
                 float* out_radius, uint64_t* out_counts, float* out_xyz, float* out_colors, uint8_t* out_scan_indices, b2_ms_stats* stats);
/* DeterminePointNeighbors: the candidate_count nearest other points (nearestKSearch of candidate_count + 1 incl. the point itself, among
 * the points of the same scan when limit_neighbors_to_same_scan_index), shuffled by std::shuffle with one std::mt19937(0) for the whole
 * call (libstdc++ 9 algorithm, the reference's toolchain), first neighbor_count kept. out: n x neighbor_count (size_t in the reference).
 * B2_ERR_STATE where the reference CHECKs (too few points per scan; a point that is not its own nearest neighbour). */
int b2_ms_point_neighbors(const float* xyz, size_t n, const uint8_t* scan_indices, int scan_count, int limit_neighbors_to_same_scan_index,
                          int candidate_count, int neighbor_count, uint64_t* out_neighbor_indices);

/* ------------------------------------------------------------------------------------------------------------------
 * Path B — dense photometric image<->scan alignment (tool ImageRegistrator).
 * Replaces, behind one handle, the pieces opt::Optimizer drives on an opt::Problem (src/opt/optimizer.h:36-57,
 * optimizer.cc:49-190): VisibilityEstimator::CreateObservationsForAllImages + DetermineIfAllNeighborsAreObserved
 * (visibility_estimator.h:46-48, .cc:49-91,199-256), ColorOptimizer::Apply (color_optimizer.cc:40-123),
 * CostCalculator::ComputeCost (cost_calculator.cc:44-100), IntrinsicsAndPoseOptimizer::Apply
 * (intrinsics_and_pose_optimizer.h:48-51, .cc:48-259), OcclusionGeometry::RenderDepthMap splat path
 * (occlusion_geometry.cc:404-464) and the image / mask / intrinsics pyramids (image.cc:106-154, intrinsics.cc:45-79).
 * Scope: every camera model of src/camera, camera rigs, depth residuals off (the reference default, parameters.h:54); occlusion depth
 * from a mesh, from splats, from a caller-supplied depth map, or none (all visible).
 * Variable layout of H/b/delta: [ParameterCount() per intrinsics | 6 per further rig camera | 6 per image (translation, rotation)], ascending ids.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct b2_reg b2_reg;

typedef struct b2_reg_params {   /* mirror of opt::Parameters (src/opt/parameters.h:40-67) as far as Path B reads it */
  int32_t point_neighbor_count;                 /* 5 */
  float fixed_residuals_weight, variable_residuals_weight;   /* 1, 1 */
  int32_t robust_weighting_type;                /* 0 none, 1 Huber (default), 2 Tukey (robust_weighting.h:42-46) */
  float robust_weighting_parameter;             /* 30*sqrt(5)/sqrt(2) */
  float maximum_valid_intensity;                /* 252 */
  float occlusion_depth_threshold;              /* 0.01 */
  int32_t min_occlusion_check_image_scale;      /* 0 */
  int32_t max_initial_image_area_in_pixels;     /* 200*160 */
  float splat_radius;                           /* 0.03 */
  int32_t image_scale_count_override;           /* 0 = Problem::InitializeImages rule (problem.cc:478-494) */
  int32_t device;                               /* -1 = current */
  float min_occlusion_depth, max_occlusion_depth;   /* 0.05, 100: near / far clip of the mesh depth pass (parameters.h:60-61) */
  int32_t mask_occlusion_boundaries;            /* 1: RenderDepthMap's default (occlusion_geometry.h:86) */
} b2_reg_params;

typedef struct b2_reg_stats {
  uint64_t observations;            /* over all images and point scales (last create_observations) */
  uint64_t residual_evaluations;    /* observations through pass 1+2 of the last accumulate (SURVEY §8d definition) */
  int32_t kernel_launches;          /* kernels of this library launched by the last call */
  float ms_last_call;               /* device time of the last call (CUDA events) */
  float ms_jacobian_kernel, ms_accumulate_kernel;   /* of the last accumulate */
} b2_reg_stats;

/* Stand-alone camera model evaluation: constructs the camera as the reference's constructors do (including the radius cut-off
 * search, camera_base_impl.h:410-462; cutoffs[0] = radius_cutoff_squared() of the camera, cutoffs[1] = of a fisheye camera's inner
 * model) and evaluates on the device: op 0 nothing, 1 NormalizedToImage (camera_base_impl.h:155-164; in n x 2, out n x 2),
 * 2 ImageDerivativeByWorld (:350-356; in n x 3, out n x 6 row-major 2x3), 3 ImageDerivativeByIntrinsics (:362-408; in n x 3,
 * out n x 2*num_params row-major). */
int b2_camera_eval(int camera_model, int width, int height, const float* params, int num_params, int op, const float* in, size_t n, float* out,
                   float cutoffs[2]);

void b2_reg_default_params(b2_reg_params* p);
int b2_reg_create(const b2_reg_params* p, b2_reg** out);
int b2_reg_destroy(b2_reg* h);
/* camera_model = camera::CameraBase::Type (camera_base.h:67-84), all 15 models: 0 FOV (fx fy cx cy omega), 1 POLYNOMIAL (+ k1 k2 k3),
 * 2 POLYNOMIAL_TANGENTIAL / 3 FISHEYE_POLYNOMIAL_TANGENTIAL (+ k1 k2 p1 p2), 4 PINHOLE (fx fy cx cy), 5 BENCHMARK = ETH3D's
 * THIN_PRISM_FISHEYE / 14 THIN_PRISM (+ k1 k2 p1 p2 k3 k4 sx1 sy1), 6 FISHEYE_POLYNOMIAL_4 / 11 POLYNOMIAL_4 (+ k1 k2 k3 k4), 10 FULL_OPENCV
 * (+ k1 k2 p1 p2 k3 k4 k5 k6), and the single-focal-length models 7 SIMPLE_PINHOLE (f cx cy), 8 RADIAL / 12 RADIAL_FISHEYE (+ k1 k2),
 * 9 SIMPLE_RADIAL / 13 SIMPLE_RADIAL_FISHEYE (+ k) — GetParameters order; num_params must equal the model's ParameterCount(). COLMAP's
 * -0.5 px shift (colmap_model.cc:833) is the caller's job. The radius cut-off search of the reference's camera constructors
 * (camera_base_impl.h:410-462, camera_base_impl_radial.h:143-171, camera_simple_radial.cc:53-57) runs on the device whenever intrinsics change. */
int b2_reg_add_intrinsics(b2_reg* h, int camera_model, int width, int height, const float* params, int num_params, int* out_id);
/* gray: width*height uint8 (cv::imread GRAYSCALE); mask: same size or NULL (values 0/1/2, image.h:43-47);
 * image_T_global: Sophus::SE3f::data() order qx qy qz qw tx ty tz (what io::ReadColmapImages fills, colmap_model.cc:117-124). */
int b2_reg_add_image(b2_reg* h, int intrinsics_id, const uint8_t* gray, const uint8_t* mask, const float image_T_global[7], int* out_id);
/* Camera mask of an intrinsics (opt::Intrinsics::camera_mask, intrinsics.h:104; loaded next to the image masks, image.cc:62-72):
 * width*height uint8 with the image-mask values; observations on non-zero pixels are discarded (visibility_estimator.cc:492-501).
 * Before b2_reg_initialize. */
int b2_reg_set_camera_mask(b2_reg* h, int intrinsics_id, const uint8_t* mask);
/* Camera rigs (opt::Rig, rig.h:40-73; opt::RigImages, rig_images.h:38-64). image_T_rig: 7 floats per camera (qx qy qz qw tx ty tz),
 * camera 0 is the reference (identity); the other extrinsics are optimisation variables (6 each, after all intrinsics blocks and
 * before the image poses, CountAndIndexVariables :442-473). b2_reg_add_rig_images binds one already added image per camera (all
 * present) recorded at the same time; the dependent images lose their own pose variables and get image_T_rig[c] * pose(reference),
 * as AssignRigs leaves them (rig.cc:216-250; the averaging that produces the initial extrinsics there is the caller's job). */
int b2_reg_add_rig(b2_reg* h, int num_cameras, const float* image_T_rig, int* out_rig_id);
int b2_reg_add_rig_images(b2_reg* h, int rig_id, const int32_t* image_ids, int* out_rig_images_id);
int b2_reg_get_rigs(b2_reg* h, float* image_T_rig_all /* 7 per camera, rigs in id order */);
int b2_reg_set_rigs(b2_reg* h, const float* image_T_rig_all);
/* First variable of an intrinsics block (kind 0), a rig's extrinsics block (1) or the pose block an image uses (2; a dependent rig
 * image reports its reference image's block). id == count gives the end of that group. */
int b2_reg_variable_index(b2_reg* h, int kind, int id, int* out_index);
/* Multi-GPU: one process per GPU, images dealt round-robin (b2_reg_image_owner: image_id % world_size). Every rank
 * makes the SAME calls with the same arguments (a rank may pass gray = mask = NULL to b2_reg_add_image for images it does not
 * own); points, descriptors and the state (intrinsics, rigs, all poses) are replicated, pyramids / depth maps / observation sets
 * exist only on the owner. Exchanges, all sum-allreduces on the handle's stream: the descriptor sums and counts of
 * b2_reg_color_update (5 floats + 1 int per point and scale), [H | b | sums] of b2_reg_accumulate / b2_reg_apply, the four
 * residual sums of every cost evaluation. Call before b2_reg_add_image; comm = NULL returns to single-GPU operation. */
int b2_reg_set_comm(b2_reg* h, b2_comm* comm);
int b2_reg_image_owner(int image_id, int world_size);
/* Problem::InitializeImages + LoadImages pyramids (image.cc:106-154, intrinsics.cc:45-50). *image_scale_count receives
 * Problem::image_scale_count(). Any image size: a level with even parents is cv::resize INTER_AREA's integer 2x2 mean, one with an odd
 * parent its general area filter (fractional coverage), both bit-identical to OpenCV; image / mask levels TRUNCATE the halved size
 * (image.cc:116,137) while camera levels ROUND it (camera_base_impl.h:72), as in the reference. B2_ERR_ARG when a level would be empty
 * (the reference: LOG(FATAL) "Resizing failed"). */
int b2_reg_initialize(b2_reg* h, int* image_scale_count);
/* One scale of the multi-resolution point cloud (problem.h points()/point_radius()/neighbor indices) + grey colours from which the
 * fixed descriptors are derived (problem.cc:550-572). neighbor_indices: n * point_neighbor_count. */
int b2_reg_add_point_scale(b2_reg* h, const float* xyz, size_t n, float point_radius, const uint64_t* neighbor_indices,
                           const float* colors, int* out_scale);
int b2_reg_set_splat_points(b2_reg* h, const float* xyz, size_t n);              /* OcclusionGeometry::SetSplatPoints */
/* OcclusionGeometry::AddMesh with edges (occlusion_geometry.cc:87-130,466-645): triangle mesh in the global frame. The depth pass
 * (the reference's OpenGL render, :213-245) is a CUDA z-buffer, followed by MaskOutOcclusionBoundaries (:284-402). Splats, when
 * set, take precedence (as in RenderDepthMap, :196-211). */
int b2_reg_set_mesh(b2_reg* h, const float* vertices, size_t num_vertices, const uint32_t* faces, size_t num_faces);
int b2_reg_set_depth_map(b2_reg* h, int image_id, int width, int height, const float* depth);   /* given occlusion depth */
int b2_reg_set_image_scale(b2_reg* h, int image_scale);                           /* Problem::SetImageScale */
int b2_reg_num_variables(b2_reg* h, int* nv);
/* RenderDepthMap at the occlusion-check scale; out may be NULL to query the size. */
int b2_reg_render_depth(b2_reg* h, int image_id, int* width, int* height, int* image_scale, float* out);
int b2_reg_create_observations(b2_reg* h, int border_size);
int b2_reg_num_observations(b2_reg* h, int image_id, int point_scale, uint64_t* count);
int b2_reg_get_observations(b2_reg* h, int image_id, int point_scale, uint64_t* point_index, float* x, float* y, float* image_scale,
                            uint8_t* all_neighbors_observed);
/* ComputePointIntensityAndJacobians of every observation of (image, scale): intensity[n], j_intrinsics[np*n] (np = the image's
 * camera model parameter count), j_pose[6n]. */
int b2_reg_get_point_jacobians(b2_reg* h, int image_id, int point_scale, float* intensity, float* j_intrinsics, float* j_pose);
/* As above plus j_rig_extrinsics[6n]: for a dependent rig image j_pose is the derivative by the rig REFERENCE image's pose and
 * j_rig_extrinsics by this camera's image_T_rig (intrinsics_and_pose_optimizer.cc:1107-1143); zeros for other images. */
int b2_reg_get_point_jacobians_rig(b2_reg* h, int image_id, int point_scale, float* intensity, float* j_intrinsics, float* j_pose,
                                   float* j_rig_extrinsics);
int b2_reg_color_update(b2_reg* h);
int b2_reg_get_descriptors(b2_reg* h, int point_scale, float* fixed_desc, float* variable_desc, int32_t* observation_counts);
/* sums: fixed_sum, n_fixed, variable_sum, n_variable, 0, 0 (the six accumulators of cost_calculator.cc:48-53). */
int b2_reg_cost(b2_reg* h, double* cost, double sums[6]);
/* Normal equations of IntrinsicsAndPoseOptimizer::Apply (:102-185): H nv*nv column-major (the Upper view the solver reads,
 * mirrored), b, sums, *cost = "initial residual". */
int b2_reg_accumulate(b2_reg* h, double* H, double* b, double sums[6], double* cost);
/* intrinsics_params: the parameter vectors of all intrinsics concatenated in id order (4 or 12 floats each) — also the order of
 * the intrinsics blocks in the optimizer's variable vector, followed by 6 per image. */
int b2_reg_get_state(b2_reg* h, float* intrinsics_params, float* poses /*7 per image*/);
int b2_reg_set_state(b2_reg* h, const float* intrinsics_params, const float* poses);
/* ComputeResidualForState for state (+) delta with the current observations' visibility lists frozen (:385-440). */
int b2_reg_cost_for_delta(b2_reg* h, const double* delta, double* cost);
/* IntrinsicsAndPoseOptimizer::Apply: LM with multiplicative damping, <=10 tries, last try always applied, signed max_change. */
int b2_reg_apply(b2_reg* h, float* lambda, float* max_change, int* applied_update, int* lm_tries);
/* Optimizer::RunOnCurrentScale (optimizer.h:44-50). */
int b2_reg_run_on_current_scale(b2_reg* h, int max_num_iterations, float max_change_convergence_threshold,
                                int iterations_without_new_optimum_threshold, int print_progress, double* optimum_cost,
                                int* converged, int* iterations);
int b2_reg_last_stats(b2_reg* h, b2_reg_stats* out);
/* GroundTruthCreator (src/exe/ground_truth_creator.cc) on the images, intrinsics and occlusion geometry of the handle.
 * A scan point is visible in an image when it lies in front of the camera, projects inside the highest-resolution image
 * (intrinsics.model(0)), is not behind the occlusion depth map rendered at intrinsics.min_image_scale (+ occlusion_depth_threshold) and
 * does not fall on a kEvalObs (= 2) pixel of the image mask (:66-79, :163-174).
 * b2_reg_gt_accumulate_observations = AccumulateScanObservationsForImage (:44-82): observation_counts[i] += 1 for every visible point
 * (exact counts; the reference's unsynchronised increments under `omp parallel for` are a race, not a semantic).
 * b2_reg_gt_create = CreateGroundTruthForImage (:84-215) without the file I/O. out_occlusion_depth (nullable, w x h floats); out_gt_depth
 * (nullable): per-pixel minimum depth of the visible points with observation_counts >= 2, +inf elsewhere; inout_scan_rendering_bgr
 * (nullable, w x h x 3, initialised by the caller with the image as cv::imread gives it): squares of 2 * scan_point_radius + 1 pixels in
 * the point's colour (rgb: n x 3), points painted in index order, later over earlier. Points = all scans concatenated in scan order. */
int b2_reg_gt_accumulate_observations(b2_reg* h, int image_id, const float* xyz, size_t n, int32_t* observation_counts);
int b2_reg_gt_create(b2_reg* h, int image_id, const float* xyz, const uint8_t* rgb, size_t n, const int32_t* observation_counts, int scan_point_radius,
                     float* out_occlusion_depth, float* out_gt_depth, uint8_t* inout_scan_rendering_bgr);
/* ComputeMinMaxPointRadius (src/opt/multi_scale_point_cloud.cc:126-184) over all images, as CreateMultiScalePointCloud calls it (:232-255):
 * a point visible in an image (visibility_estimator.cc:296-364: in front, inside, not occluded, not masked, not saturated, at the
 * occlusion-check image scale) gets the radius that spans 0.5 px at the finest image scale (ImageToNormalized through the undistortion
 * lookup for distorted models); min_radius[i] = min(..., r), max_radius[i] = max(..., r / min_scaling_factor). In/out arrays: the caller
 * initialises them (+inf / -inf). After b2_reg_initialize; single GPU. Feeds b2_ms_create. */
int b2_reg_min_max_point_radius(b2_reg* h, const float* xyz, size_t n, double min_scaling_factor, float* min_radius, float* max_radius);

#ifdef __cplusplus
}
#endif
#endif /* ETH3D_B200_H_ */
