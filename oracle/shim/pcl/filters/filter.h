// TEST INFRASTRUCTURE (oracle shim) -- not part of the shipped product.
// pcl::removeNaNFromPointCloud restated from PCL's published behaviour (pcl/filters/impl/filter.hpp):
// dense clouds are copied verbatim and `index` becomes 0..n-1 (an O(n) pass -- the reference calls this inside
// its per-point loop, BasicLaserOdometry.cpp:252, and pays that cost); otherwise non-finite x/y/z are dropped.
#pragma once
#include <cmath>
#include <vector>
#include <pcl/point_cloud.h>

namespace pcl {

template <typename PointT>
void removeNaNFromPointCloud(const PointCloud<PointT>& cloud_in, PointCloud<PointT>& cloud_out, std::vector<int>& index) {
  if (&cloud_in != &cloud_out) {
    cloud_out.header = cloud_in.header;
    cloud_out.points.resize(cloud_in.points.size());
  }
  index.resize(cloud_in.points.size());
  std::size_t j = 0;
  if (cloud_in.is_dense) {
    if (&cloud_in != &cloud_out) cloud_out = cloud_in;
    for (j = 0; j < cloud_out.points.size(); ++j) index[j] = static_cast<int>(j);
  } else {
    for (std::size_t i = 0; i < cloud_in.points.size(); ++i) {
      if (!std::isfinite(cloud_in.points[i].x) || !std::isfinite(cloud_in.points[i].y) ||
          !std::isfinite(cloud_in.points[i].z))
        continue;
      cloud_out.points[j] = cloud_in.points[i];
      index[j] = static_cast<int>(i);
      j++;
    }
    if (j != cloud_in.points.size()) {
      cloud_out.points.resize(j);
      index.resize(j);
    }
    cloud_out.height = 1;
    cloud_out.width = static_cast<std::uint32_t>(j);
    cloud_out.is_dense = true;
  }
}

}  // namespace pcl
