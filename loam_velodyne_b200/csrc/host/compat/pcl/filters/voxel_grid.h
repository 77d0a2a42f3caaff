// pcl::VoxelGrid<PointXYZI> facade for builds without PCL.  BasicLaserMapping::downSizeFilterCorner()/Surf()/Map()
// return `pcl::VoxelGrid<pcl::PointXYZI>&` and the ROS adapter only calls setLeafSize on them
// (LaserMapping.cpp:121,135,149 upstream), so this facade just carries the leaf size; the filtering itself runs on
// the GPU through loam_b200_voxel_grid (include/loam_b200.h).
#pragma once
#include <pcl/point_cloud.h>

#define LOAM_B200_COMPAT_PCL 1

namespace pcl {

template <typename PointT>
class VoxelGrid {
 public:
  VoxelGrid() { leaf_[0] = leaf_[1] = leaf_[2] = 0.f; }
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  float leafX() const { return leaf_[0]; }
  const float* getLeafSizeArray() const { return leaf_; }

 private:
  float leaf_[3];
};

}  // namespace pcl
