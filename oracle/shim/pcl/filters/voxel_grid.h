// TEST INFRASTRUCTURE (oracle shim) -- not part of the shipped product.
// pcl::VoxelGrid<PointXYZI> restated from PCL's published algorithm (pcl/filters/impl/voxel_grid.hpp,
// `applyFilter`, downsample_all_data = true, min_points_per_voxel = 0); PCL is not vendored by the reference
// and its version is unpinned (CMakeLists.txt:15), so this boundary is "parity unpinned":
//   bbox over all (finite) points -> min_b = floor(min * inv_leaf), max_b likewise -> div_b = max_b - min_b + 1
//   idx = (floor(x*inv)-min_b.x) + (floor(y*inv)-min_b.y)*div.x + (floor(z*inv)-min_b.z)*div.x*div.y
//   std::sort by idx (unstable) -> one output point per run = arithmetic mean of x,y,z,intensity, ascending idx
//   if dx*dy*dz overflows int32 the input is returned unchanged (PCL prints a warning).
// Call sites: BasicScanRegistration.cpp:246-250; BasicLaserMapping.cpp:98-99,259-262,519-527,580-588.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include <pcl/point_cloud.h>

namespace pcl {

template <typename PointT>
class VoxelGrid {
 public:
  typedef typename PointCloud<PointT>::Ptr PointCloudPtr;
  typedef typename PointCloud<PointT>::ConstPtr PointCloudConstPtr;

  VoxelGrid() { leaf_[0] = leaf_[1] = leaf_[2] = 0.f; inv_[0] = inv_[1] = inv_[2] = 0.f; }

  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  void setLeafSize(float lx, float ly, float lz) {
    leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz;
    inv_[0] = 1.0f / lx; inv_[1] = 1.0f / ly; inv_[2] = 1.0f / lz;
  }
  const float* getLeafSize() const { return leaf_; }

  void filter(PointCloud<PointT>& output) {
    output.points.clear();
    output.width = 0;
    output.height = 1;
    output.is_dense = true;
    if (!input_ || input_->points.empty()) return;
    const std::vector<PointT>& in = input_->points;
    output.header = input_->header;

    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    const bool dense = input_->is_dense;
    for (std::size_t i = 0; i < in.size(); i++) {
      if (!dense && (!std::isfinite(in[i].x) || !std::isfinite(in[i].y) || !std::isfinite(in[i].z))) continue;
      mn[0] = std::min(mn[0], in[i].x); mx[0] = std::max(mx[0], in[i].x);
      mn[1] = std::min(mn[1], in[i].y); mx[1] = std::max(mx[1], in[i].y);
      mn[2] = std::min(mn[2], in[i].z); mx[2] = std::max(mx[2], in[i].z);
    }
    std::int64_t dx = static_cast<std::int64_t>((mx[0] - mn[0]) * inv_[0]) + 1;
    std::int64_t dy = static_cast<std::int64_t>((mx[1] - mn[1]) * inv_[1]) + 1;
    std::int64_t dz = static_cast<std::int64_t>((mx[2] - mn[2]) * inv_[2]) + 1;
    if ((dx * dy * dz) > static_cast<std::int64_t>(std::numeric_limits<std::int32_t>::max())) {
      output = *input_;  // "Leaf size is too small for the input dataset. Integer indices would overflow."
      return;
    }
    int min_b[3], max_b[3], div_b[3], mul[3];
    for (int a = 0; a < 3; a++) {
      min_b[a] = static_cast<int>(std::floor(mn[a] * inv_[a]));
      max_b[a] = static_cast<int>(std::floor(mx[a] * inv_[a]));
      div_b[a] = max_b[a] - min_b[a] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];

    struct IdxPt {
      unsigned int idx;
      unsigned int pt;
      bool operator<(const IdxPt& o) const { return idx < o.idx; }
    };
    std::vector<IdxPt> iv;
    iv.reserve(in.size());
    for (std::size_t i = 0; i < in.size(); i++) {
      if (!dense && (!std::isfinite(in[i].x) || !std::isfinite(in[i].y) || !std::isfinite(in[i].z))) continue;
      int ijk0 = static_cast<int>(std::floor(in[i].x * inv_[0]) - static_cast<float>(min_b[0]));
      int ijk1 = static_cast<int>(std::floor(in[i].y * inv_[1]) - static_cast<float>(min_b[1]));
      int ijk2 = static_cast<int>(std::floor(in[i].z * inv_[2]) - static_cast<float>(min_b[2]));
      int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
      iv.push_back(IdxPt{static_cast<unsigned int>(idx), static_cast<unsigned int>(i)});
    }
    std::sort(iv.begin(), iv.end());

    std::size_t first = 0;
    while (first < iv.size()) {
      std::size_t last = first + 1;
      while (last < iv.size() && iv[last].idx == iv[first].idx) ++last;
      // centroid of all fields (x, y, z, intensity), accumulated in float in sorted order
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      for (std::size_t k = first; k < last; k++) {
        const PointT& p = in[iv[k].pt];
        sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
      }
      const float n = static_cast<float>(last - first);
      PointT o;
      o.x = sx / n; o.y = sy / n; o.z = sz / n; o.intensity = si / n;
      output.points.push_back(o);
      first = last;
    }
    output.width = static_cast<std::uint32_t>(output.points.size());
  }

 private:
  PointCloudConstPtr input_;
  float leaf_[3], inv_[3];
};

}  // namespace pcl
