/* ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
 * C interface of the CPU restatement (liboracle.so), loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs through oracle/oracle.py. Nothing under
 * dataset_pipeline_b200/ may include, link or dlopen this. */
#ifndef ORC_API_H_
#define ORC_API_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_icp orc_icp;

typedef struct orc_icp_stats {
  int32_t inner_iterations;     /* accumulate passes of the last AlignMeshes */
  int32_t lm_tries_total;       /* cost passes of the last AlignMeshes */
  int32_t num_pairs;            /* non-empty correspondence sets */
  int32_t num_variables;
  uint64_t num_correspondences;
  double first_cost, last_cost, final_lambda;
  double t_transform, t_search, t_inner; /* cumulative wall seconds since create */
  double t_acc, t_cost;                  /* wall seconds in the accumulate / cost passes of the last AlignMeshes */
} orc_icp_stats;

orc_icp* orc_icp_create(void);
void orc_icp_destroy(orc_icp*);
void orc_icp_set_options(orc_icp*, int use_kdtree, int inner_max_iterations);
/* sampling knob for the bounded CPU-baseline timing only: search every stride-th source point */
void orc_icp_set_query_stride(orc_icp*, size_t stride);
/* icp_point_to_plane.cc:109-135. xyz/nrm: n x 3 floats. T: column-major 4x4 (Eigen::Affine3f). Returns id (-1 fixed). */
int orc_icp_add_cloud(orc_icp*, const float* xyz, const float* nrm, size_t n, const float T_colmajor[16], int fixed);
/* icp_point_to_plane.cc:137-163. Returns 0, or 1 when no movable cloud was added (reference aborts). */
int orc_icp_run(orc_icp*, float max_dist, int initial_iteration, int max_iters, float thr, int* converged);
int orc_icp_get_pose(orc_icp*, int id, float T_colmajor[16]);
int orc_icp_set_pose(orc_icp*, int id, const float T_colmajor[16]);
void orc_icp_last_stats(orc_icp*, orc_icp_stats*);
int orc_icp_last_tries(orc_icp*, int* tries, int cap);
int orc_icp_last_pair_info(orc_icp*, int k, int* src_impl_index, int* tgt_impl_index, uint64_t* count);
int orc_icp_last_pair_corr(orc_icp*, int k, int* q, int* m, float* d2);
int orc_icp_last_normal_eq(orc_icp*, double* H_colmajor, double* b);

void orc_transform_cloud(const float* xyz, const float* nrm, size_t n, const float T[16], float* oxyz, float* onrm);
uint64_t orc_find_correspondences(const float* src, size_t ns, const float* tgt, size_t nt, float max_dist,
                                  int use_kdtree, int* q, int* m, float* d2);
void orc_time_search(const float* src, size_t ns, const float* tgt, size_t nt, float max_dist, double* t_build, double* t_query,
                     uint64_t* matched);
void orc_se3_exp_left_mul(const double x[6], const float q_in[4], const float t_in[3], float q_out[4], float t_out[3]);
int orc_ldlt_solve_upper(const double* A_colmajor, int n, const double* b, double* x);

/* normals (orc_normals.cc) */
int orc_normals_knn(const float* xyz, size_t n, int k, const float viewpoint[3], float* out_nxyz_curv /* n x 4 */,
                    int* out_knn_idx /* n x k or NULL */);

/* radius mode: every point with d2 < (float)((double)r*r), sorted by (distance, index); out_count (nullable) = neighbours found */
int orc_normals_radius(const float* xyz, size_t n, float radius, const float viewpoint[3], float* out_nxyz_curv, int* out_count);

/* ---- Path B: photometric image<->scan alignment (orc_reg.cc), pinhole cameras, no rigs, no depth residuals ---- */
typedef struct orc_reg orc_reg;
typedef struct orc_reg_params {   /* mirror of opt::Parameters (src/opt/parameters.h:40-67) restricted to what Path B reads */
  int32_t point_neighbor_count;
  float fixed_residuals_weight, variable_residuals_weight;
  int32_t robust_weighting_type;        /* 0 none, 1 huber, 2 tukey */
  float robust_weighting_parameter;
  float maximum_valid_intensity, occlusion_depth_threshold;
  int32_t min_occlusion_check_image_scale, max_initial_image_area_in_pixels;
  float splat_radius;
  int32_t image_scale_count_override;   /* >0: force Problem::image_scale_count_ (tests do this through friend helpers) */
  float min_occlusion_depth, max_occlusion_depth;   /* 0.05, 100 (parameters.h:60-61) */
  int32_t mask_occlusion_boundaries;    /* RenderDepthMap default true (occlusion_geometry.h:86) */
} orc_reg_params;
void orc_reg_default_params(orc_reg_params*);
orc_reg* orc_reg_create(const orc_reg_params*);
void orc_reg_destroy(orc_reg*);
int orc_reg_add_intrinsics(orc_reg*, int w, int h, const float fx_fy_cx_cy[4]);   /* pinhole */
/* type = camera::CameraBase::Type value: 4 pinhole (4 params), 14 thin prism (12), 5 benchmark = thin-prism fisheye (12);
 * params in GetParameters order: fx fy cx cy k1 k2 p1 p2 k3 k4 sx1 sy1. Returns -1 for an unsupported type. */
int orc_reg_add_intrinsics_model(orc_reg*, int type, int w, int h, const float* params);
/* image_T_global as Sophus::SE3f::data(): qx qy qz qw tx ty tz */
int orc_reg_add_image(orc_reg*, int intrinsics_id, const uint8_t* gray, const uint8_t* mask_or_null, const float image_T_global[7]);
int orc_reg_set_camera_mask(orc_reg*, int intrinsics_id, const uint8_t* mask);   /* intrinsics.h:104; before initialize */
int orc_reg_initialize(orc_reg*);
/* rigs (rig.h:40-73, rig_images.h:38-64): image_T_rig 7 floats per camera, camera 0 = reference; add_rig_images binds one image per
 * camera (all present) and sets the dependent images' poses to image_T_rig[c] * pose(reference). -1 on bad arguments. */
int orc_reg_add_rig(orc_reg*, int num_cameras, const float* image_T_rig);
int orc_reg_add_rig_images(orc_reg*, int rig_id, const int* image_ids);
void orc_reg_get_rigs(orc_reg*, float* image_T_rig_all);
void orc_reg_set_rigs(orc_reg*, const float* image_T_rig_all);
/* first variable of an intrinsics block (kind 0), of a rig's extrinsics block (1), of the pose block an image uses (2) */
int orc_reg_variable_index(orc_reg*, int kind, int id);
void orc_reg_point_jacobians_rig(orc_reg*, int image, int point_scale, uint64_t obs_index, float* intensity, float* jK, float jP[6], float* jR);
int orc_reg_add_point_scale(orc_reg*, const float* xyz, size_t n, float radius, const uint64_t* neighbor_indices, const float* colors);
void orc_reg_set_splat_points(orc_reg*, const float* xyz, size_t n);
void orc_reg_set_mesh(orc_reg*, const float* vertices, size_t nv, const uint32_t* faces, size_t nf);
uint64_t orc_reg_mesh_edges(orc_reg*, uint32_t* v1, uint32_t* v2, uint32_t* f1, uint32_t* f2, uint8_t* flags);
int orc_reg_set_depth_map(orc_reg*, int image, int w, int h, const float* depth);
void orc_reg_set_image_scale(orc_reg*, int image_scale);
int orc_reg_image_scale_count(orc_reg*);
int orc_reg_num_variables(orc_reg*);
int orc_reg_render_depth(orc_reg*, int image, int* w, int* h, float* out_or_null);
void orc_reg_create_observations(orc_reg*, int border);
/* GroundTruthCreator (src/exe/ground_truth_creator.cc:44-215): visibility counts per scan point, then per image the occlusion depth, the
 * ground-truth depth map and the scan rendering (BGR, in/out). */
void orc_reg_gt_accumulate_observations(orc_reg*, int image, const float* xyz, size_t n, int32_t* counts);
void orc_reg_gt_create(orc_reg*, int image, const float* xyz, const uint8_t* rgb, size_t n, const int32_t* counts, int scan_point_radius,
                       float* occlusion_depth, float* gt_depth, uint8_t* rendering_bgr);
/* ComputeMinMaxPointRadius over all images (multi_scale_point_cloud.cc:126-184, 232-255); min/max in-out (+inf / -inf initially) */
void orc_reg_min_max_point_radius(orc_reg*, const float* xyz, size_t n, double min_scaling_factor, float* min_radius, float* max_radius);
uint64_t orc_reg_num_observations(orc_reg*, int image, int point_scale);
void orc_reg_get_observations(orc_reg*, int image, int point_scale, uint64_t* idx, float* x, float* y, float* s, uint8_t* nbrs_observed);
void orc_reg_color_update(orc_reg*);
void orc_reg_get_descriptors(orc_reg*, int point_scale, float* fixed, float* variable, int* counts);
double orc_reg_cost(orc_reg*, double sums[6]);
double orc_reg_accumulate(orc_reg*, double* H_colmajor, double* b, double sums[6]);
void orc_reg_get_state(orc_reg*, float* intr_params, float* poses);
void orc_reg_set_state(orc_reg*, const float* intr_params, const float* poses);
double orc_reg_cost_for_delta(orc_reg*, const double* delta);
int orc_reg_apply(orc_reg*, float* lambda, float* max_change, int* applied);
int orc_reg_run_on_current_scale(orc_reg*, int max_it, float max_change_thr, int no_opt_thr, double* optimum_cost, int* converged);
void orc_reg_point_jacobians(orc_reg*, int image, int point_scale, uint64_t obs_index, float* intensity, float* jK /* np */, float jP[6]);
int orc_interp_bilinear(const uint8_t* img, int w, int h, float x, float y, float* v, float* dx, float* dy);
void orc_interp_trilinear(const uint8_t* img0, int w0, int h0, const uint8_t* img1, float x, float y, float z, float* v, float* dx, float* dy, float* dz);
float orc_robust(int type, float p, float r, int weight);

/* ---- multi-resolution point cloud (orc_multiscale.cc): MergeClosePoints, CreateMultiScalePointCloud's scale loop,
 * Problem::DeterminePointNeighbors ---- */
uint64_t orc_ms_merge_close_points(const float* xyz, size_t n, const float* colors, const uint8_t* scan, const float* max_radius, int num_scans,
                                   float merge_distance, float* oxyz, float* ocol, uint8_t* oscan, float* omaxr);
int orc_ms_create(const float* xyz, size_t n, const float* colors, const uint8_t* scan, const float* min_radius, const float* max_radius,
                  int num_scans, float min_radius_bias, float merge_distance_factor, int max_scales, float* out_radius, uint64_t* out_counts,
                  float* oxyz, float* ocol, uint8_t* oscan);
int orc_ms_point_neighbors(const float* xyz, size_t n, const uint8_t* scan, int scan_count, int limit_to_same_scan, int candidate_count,
                           int neighbor_count, uint64_t* out);

/* ---- point-cloud tools (orc_cleaner.cc): LocalStatisticalOutlierRemoval (PointCloudCleaner), SplatCreator ---- */
int orc_lsor_filter(const float* xyz, size_t n, int mean_k, double distance_factor_threshold, int negative, int32_t* out_indices,
                    uint64_t* out_count, int32_t* out_removed, uint64_t* out_removed_count, float* out_distances);
void orc_mesh_squared_distance(const float* points, size_t n, const float* vertices, const uint32_t* faces, size_t nf, float* out);
uint64_t orc_splat_create(const float* xyz, const float* normals, size_t n, const float* vertices, const uint32_t* faces, size_t nf,
                          float max_splat_size, float squared_distance_threshold, float* corners, uint8_t* added, float* radius);

/* ---- camera models (orc_camera.h) ---- */
int orc_cam_param_count(int type);   /* -1 unsupported */
/* constructs the camera (runs the cut-off search as the reference constructors do); out[0] = radius_cutoff_squared of the
 * camera itself, out[1] = of the inner model of a fisheye camera (inf otherwise) */
int orc_cam_cutoff(int type, int w, int h, const float* params, float out[2]);
/* op: 0 Distort(n) -> 2, 1 NormalizedToImage(n) -> 2, 2 ImageDerivativeByWorld(p) -> 6, 3 ImageDerivativeByIntrinsics(p) -> 2*np,
 * 4 Undistort(distorted) -> 2 (camera_base_impl.h:252-255 / camera_base_impl_fisheye.h:80-91), 5 DistortedDerivativeByNormalized(n) -> 4.
 * in: n x 2 (ops 0,1,4,5) or n x 3 (ops 2,3) floats. */
int orc_cam_eval(int type, int w, int h, const float* params, int op, const float* in, size_t n, float* out);
void orc_image_pyramid_level(const uint8_t* src, int w, int h, uint8_t* dst);

#ifdef __cplusplus
}
#endif
#endif
