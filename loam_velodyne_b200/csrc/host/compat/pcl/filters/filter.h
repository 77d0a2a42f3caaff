// pcl::removeNaNFromPointCloud for builds without PCL (used by the ROS adapters' message handlers upstream).
#pragma once
#include <cmath>
#include <vector>
#include <pcl/point_cloud.h>

namespace pcl {

template <typename PointT>
void removeNaNFromPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, std::vector<int>& index) {
  const std::size_t n = in.points.size();
  if (in.is_dense) {
    if (&in != &out) out = in;
    index.resize(n);
    for (std::size_t j = 0; j < n; ++j) index[j] = static_cast<int>(j);
    return;
  }
  std::vector<PointT> kept;
  kept.reserve(n);
  index.clear();
  for (std::size_t i = 0; i < n; ++i) {
    const PointT& p = in.points[i];
    if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) {
      kept.push_back(p);
      index.push_back(static_cast<int>(i));
    }
  }
  out.header = in.header;
  out.points.swap(kept);
  out.width = static_cast<std::uint32_t>(out.points.size());
  out.height = 1;
  out.is_dense = true;
}

}  // namespace pcl
