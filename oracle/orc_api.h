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

#ifdef __cplusplus
}
#endif
#endif
