// Drop-in C++ shims with the reference's class shapes for the PCL-protocol seams of the point-cloud tools, on top of the C ABI
// (include/eth3d_b200.h). Header-only, no PCL needed: templated on the point type / cloud pointer so the reference's callers compile
// against them with pcl::PointCloud<PointT>::Ptr unchanged, and plain structs work in tests here.
//
//   pcl_b200::LocalStatisticalOutlierRemoval<PointT>   pcl::LocalStatisticalOutlierRemoval<PointT>
//       (/root/reference/src/geometry/local_statistical_outlier_removal.h, .hpp:44-176) as PointCloudCleaner uses it
//       (/root/reference/src/exe/point_cloud_cleaner.cc:86-93): setInputCloud, setMeanK, setDistanceFactorThresh, setNegative, filter
//   pcl_b200::NormalEstimationTwoPassOMP<PointInT, PointOutT>   pcl::NormalEstimationTwoPassOMP
//       (/root/reference/src/geometry/two_pass_normal_3d_omp.h:53-99) as icp_scan_aligner.cc:323-330 / normal_estimator.cc:177-194 use it:
//       setInputCloud, setSearchMethod (ignored: the index is internal), setKSearch / setRadiusSearch, setViewPoint, compute
//
// Requirements on the types: PointT has x, y, z as its first three floats (every PCL point type); PointOutT has normal_x, normal_y,
// normal_z, curvature members; CloudPtr offers ->size() and ->points (a contiguous std::vector<PointT>).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/eth3d_b200.h"

namespace pcl_b200 {

inline void check(int rc) {
  if (rc != B2_OK) throw std::runtime_error(std::string("eth3d_b200: ") + b2_last_error());
}

template <typename PointT>
class LocalStatisticalOutlierRemoval {
 public:
  explicit LocalStatisticalOutlierRemoval(bool extract_removed_indices = false) : extract_removed_(extract_removed_indices) {}

  template <typename CloudPtr>
  void setInputCloud(const CloudPtr& cloud) { points_ = cloud->points.data(); n_ = cloud->size(); }
  void setMeanK(int nr_k) { mean_k_ = nr_k; }
  int getMeanK() const { return mean_k_; }
  void setDistanceFactorThresh(double factor) { factor_ = factor; }
  double getDistanceFactorThresh() const { return factor_; }
  void setNegative(bool negative) { negative_ = negative; }

  // applyFilterIndices: the indices of the points that stay
  void filter(std::vector<int>& indices) {
    indices.resize(n_);
    removed_.resize(extract_removed_ ? n_ : 0);
    size_t kept = 0, removed = 0;
    static_assert(sizeof(int) == sizeof(int32_t), "int is 32 bits");
    check(b2_lsor_filter(reinterpret_cast<const float*>(points_), n_, sizeof(PointT), mean_k_, factor_, negative_ ? 1 : 0,
                         reinterpret_cast<int32_t*>(indices.data()), &kept, extract_removed_ ? reinterpret_cast<int32_t*>(removed_.data()) : nullptr,
                         &removed, nullptr));
    indices.resize(kept);
    if (extract_removed_) removed_.resize(removed);
  }
  // applyFilter: the filtered cloud (copyPointCloud(*input_, indices, output), local_statistical_outlier_removal.hpp:65-66)
  template <typename Cloud>
  void filter(Cloud& output) {
    std::vector<int> indices;
    filter(indices);
    output.points.resize(indices.size());
    for (size_t i = 0; i < indices.size(); ++i) output.points[i] = points_[indices[i]];
  }
  const std::vector<int>& getRemovedIndices() const { return removed_; }

 private:
  const PointT* points_ = nullptr;
  size_t n_ = 0;
  int mean_k_ = 1;
  double factor_ = 3.0;
  bool negative_ = false, extract_removed_ = false;
  std::vector<int> removed_;
};

template <typename PointInT, typename PointOutT>
class NormalEstimationTwoPassOMP {
 public:
  explicit NormalEstimationTwoPassOMP(unsigned int /*nr_threads*/ = 0) {}
  template <typename CloudPtr>
  void setInputCloud(const CloudPtr& cloud) { points_ = cloud->points.data(); n_ = cloud->size(); }
  template <typename TreePtr>
  void setSearchMethod(const TreePtr&) {}
  void setKSearch(int k) { k_ = k; radius_ = 0.f; }
  void setRadiusSearch(double radius) { radius_ = (float)radius; k_ = 0; }
  void setViewPoint(float vpx, float vpy, float vpz) { vp_[0] = vpx; vp_[1] = vpy; vp_[2] = vpz; }

  // Feature::compute: output.points[i] gets normal_x, normal_y, normal_z, curvature; output.is_dense as PCL sets it
  template <typename CloudOut>
  void compute(CloudOut& output) {
    std::vector<float> out(4 * n_);
    int dense = 1;
    if (radius_ > 0.f) check(b2_normals_estimate_radius(reinterpret_cast<const float*>(points_), n_, sizeof(PointInT), radius_, vp_, nullptr, -1, out.data(), nullptr, &dense));
    else check(b2_normals_estimate(reinterpret_cast<const float*>(points_), n_, sizeof(PointInT), k_, vp_, out.data(), nullptr, &dense));
    output.points.resize(n_);
    for (size_t i = 0; i < n_; ++i) {
      PointOutT& p = output.points[i];
      p.normal_x = out[4 * i]; p.normal_y = out[4 * i + 1]; p.normal_z = out[4 * i + 2]; p.curvature = out[4 * i + 3];
    }
    output.is_dense = dense != 0;
  }

 private:
  const PointInT* points_ = nullptr;
  size_t n_ = 0;
  int k_ = 0;
  float radius_ = 0.f;
  float vp_[3] = {0.f, 0.f, 0.f};
};

}  // namespace pcl_b200
