// Compile check of the header-only C++ shims (tests/test_cabi.py): stand-ins for the PCL types, the calls the reference tools make.
#include <memory>
#include "dataset_pipeline_b200/cpp/point_cloud_tools_b200.h"
#include "dataset_pipeline_b200/cpp/icp_point_to_plane_b200.h"
struct PointXYZRGB { float x, y, z, pad; unsigned int rgba; float pad2[3]; };
struct Normal { float normal_x, normal_y, normal_z, curvature; };
template <typename P> struct Cloud { std::vector<P> points; bool is_dense = true; size_t size() const { return points.size(); } };
int main() {
  auto cloud = std::make_shared<Cloud<PointXYZRGB>>();
  cloud->points.resize(100);
  pcl_b200::LocalStatisticalOutlierRemoval<PointXYZRGB> sor;
  sor.setInputCloud(cloud); sor.setMeanK(8); sor.setDistanceFactorThresh(2.0);
  Cloud<PointXYZRGB> filtered;
  try { sor.filter(filtered); } catch (const std::exception&) {}
  pcl_b200::NormalEstimationTwoPassOMP<PointXYZRGB, Normal> ne;
  ne.setInputCloud(cloud); ne.setSearchMethod(nullptr); ne.setKSearch(16); ne.setViewPoint(0, 0, 0);
  Cloud<Normal> normals;
  try { ne.compute(normals); } catch (const std::exception&) {}
  return 0;
}
