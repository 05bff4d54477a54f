// Internal seam between the exact kNN search of K7 (b2_normals.cu) and the stages that consume neighbour lists on the device
// (b2_cleaner.cu: LocalStatisticalOutlierRemoval, SplatCreator).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <functional>

namespace b2 {

enum { kKnnStatNone = 0, kKnnStatMeanDistance = 1, kKnnStatLastD2 = 2 };

struct KnnHook {
  int stat_mode = kKnnStatNone;   // per-point statistic written next to the lists (indexed by the ORIGINAL point index)
  bool need_idx = false;          // n x k neighbour lists (original indices, (d2, index) ascending, -1 padded)
  // Runs on the search stream after the kNN kernel, before the scratch is released. xyz_dev: the packed n x 3 input.
  std::function<int(cudaStream_t st, const float* xyz_dev, const int* idx_dev, const float* stat_dev, size_t n, int k)> run;
};

int knn_with_hook(const float* xyz, size_t n, size_t stride_bytes, int k, const KnnHook& hook);

}  // namespace b2
