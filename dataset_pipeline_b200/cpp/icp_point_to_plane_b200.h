// Drop-in C++ shim with the reference's class shape (icp::PointToPlaneICP, /root/reference/src/icp/icp_point_to_plane.h:39-57)
// on top of the C ABI (include/eth3d_b200.h). Header-only, no PCL/Eigen needed: it is templated on the point-cloud and
// transform types so the reference's callers (exe/icp_scan_aligner.cc:289,336,343,351; opt/test/test_icp.cc) compile against it
// with pcl::PointCloud<pcl::PointNormal>::Ptr and Eigen::Affine3f unchanged, and plain structs work in tests here.
//
// Requirements on the types:
//   CloudPtr : ->size(), ->points.data() (or ->data()) of a 48-byte record with x,y,z at offset 0 and normal_x..z at offset 16
//              (pcl::PointNormal layout, /root/reference SURVEY.md §8 "Layouts").
//   Affine   : .data() -> 16 floats column-major (Eigen::Affine3f::data()).
#pragma once
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/eth3d_b200.h"

namespace icp_b200 {

struct PointNormal48 {   // layout-compatible with pcl::PointNormal (48 bytes)
  float x, y, z, pad0;
  float normal_x, normal_y, normal_z, pad1;
  float curvature, pad2[3];
};
static_assert(sizeof(PointNormal48) == 48, "pcl::PointNormal is 48 bytes");

struct Affine3fPOD {     // column-major 4x4, like Eigen::Affine3f
  float m[16];
  const float* data() const { return m; }
  float* data() { return m; }
};

class PointToPlaneICP {
 public:
  PointToPlaneICP() {
    b2_icp_config cfg;
    b2_icp_default_config(&cfg);
    check(b2_icp_create(&cfg, &h_));
  }
  explicit PointToPlaneICP(const b2_icp_config& cfg) { check(b2_icp_create(&cfg, &h_)); }
  ~PointToPlaneICP() { if (h_) b2_icp_destroy(h_); }
  PointToPlaneICP(const PointToPlaneICP&) = delete;
  PointToPlaneICP& operator=(const PointToPlaneICP&) = delete;

  // int AddPointCloud(PointCloud<PointNormal>::Ptr, const Eigen::Affine3f& global_T_cloud, bool fixed)
  template <typename CloudPtr, typename Affine>
  int AddPointCloud(const CloudPtr& cloud, const Affine& global_T_cloud, bool fixed) {
    const auto* rec = cloud->points.data();
    static_assert(sizeof(*rec) == 48, "expected a 48-byte PointNormal record");
    const float* base = reinterpret_cast<const float*>(rec);
    int id = 0;
    check(b2_icp_add_cloud(h_, base, base + 4, cloud->size(), 48, global_T_cloud.data(), fixed ? 1 : 0, &id));
    return id;
  }

  // bool Run(float max_correspondence_distance, int initial_iteration, int max_num_iterations,
  //          float convergence_threshold_max_movement, bool print_progress)
  bool Run(float max_correspondence_distance, int initial_iteration, int max_num_iterations,
           float convergence_threshold_max_movement, bool print_progress) {
    int converged = 0;
    check(b2_icp_run(h_, max_correspondence_distance, initial_iteration, max_num_iterations, convergence_threshold_max_movement,
                     print_progress ? 1 : 0, &converged));
    return converged != 0;
  }

  // Eigen::Affine3f GetResultGlobalTCloud(int cloud_index)
  template <typename Affine = Affine3fPOD>
  Affine GetResultGlobalTCloud(int cloud_index) {
    Affine T;
    float m[16];
    check(b2_icp_get_pose(h_, cloud_index, m));
    std::memcpy(T.data(), m, sizeof(m));
    return T;
  }

  b2_icp* handle() { return h_; }

 private:
  static void check(int rc) {
    // The reference aborts via glog CHECK / LOG(FATAL); a C++ host gets an exception instead.
    if (rc != B2_OK) throw std::runtime_error(std::string("eth3d_b200: ") + b2_last_error());
  }
  b2_icp* h_ = nullptr;
};

}  // namespace icp_b200
