// Drop-in for the reference's include/loam_velodyne/BasicTransformMaintenance.h: fuses the high-rate odometry pose with the
// low-rate mapping correction (upstream BasicTransformMaintenance.cpp:45-178).  O(1) host scalar math -- nothing here
// touches the GPU; it is provided so that a node built on the three Basic* drop-ins finds the fourth class as well.
// The association itself is the routine BasicLaserMapping uses for its pose prediction (host/pose_algebra.h).
#pragma once

#include "Twist.h"

namespace loam {

class BasicTransformMaintenance {
 public:
  void updateOdometry(double pitch, double yaw, double roll, double x, double y, double z);
  void updateMappingTransform(Twist const& transformAftMapped, Twist const& transformBefMapped);
  void updateMappingTransform(double pitch, double yaw, double roll, double x, double y, double z, double twist_rot_x,
                              double twist_rot_y, double twist_rot_z, double twist_pos_x, double twist_pos_y,
                              double twist_pos_z);

  void transformAssociateToMap();

  // result accessor: rot_x, rot_y, rot_z, x, y, z
  auto const& transformMapped() const { return _transformMapped; }

 private:
  float _transformSum[6]{};
  float _transformIncre[6]{};
  float _transformMapped[6]{};
  float _transformBefMapped[6]{};
  float _transformAftMapped[6]{};
};

}  // namespace loam
