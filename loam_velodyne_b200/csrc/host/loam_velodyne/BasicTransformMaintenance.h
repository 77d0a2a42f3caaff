// Pose fusion between the odometry rate and the mapping rate -- the fourth ROS-free class of the reference
// (include/loam_velodyne/BasicTransformMaintenance.h, src/lib/BasicTransformMaintenance.cpp:45-178), offered with the same
// public calls so that a node built on the three GPU-backed Basic* drop-ins finds it too.  Host scalar arithmetic only:
// the association is the routine the mapping stage uses for its pose prediction (host/pose_algebra.h).
#pragma once

#include "Twist.h"

namespace loam {

class BasicTransformMaintenance {
 public:
  /** fused pose after transformAssociateToMap(): rot_x, rot_y, rot_z, x, y, z */
  auto const& transformMapped() const { return _mapped; }

  /** compose odometry pose, odometry pose at the last mapping update and mapped pose of that update */
  void transformAssociateToMap();

  /** latest odometry pose (angles in rad) */
  void updateOdometry(double pitch, double yaw, double roll, double x, double y, double z);

  /** latest mapping result: mapped pose (aft) and the odometry pose it was computed from (bef) */
  void updateMappingTransform(Twist const& transformAftMapped, Twist const& transformBefMapped);
  void updateMappingTransform(double pitch,
                              double yaw,
                              double roll,
                              double x,
                              double y,
                              double z,
                              double twist_rot_x,
                              double twist_rot_y,
                              double twist_rot_z,
                              double twist_pos_x,
                              double twist_pos_y,
                              double twist_pos_z);

 private:
  Twist _odometry, _odometryAtMapping, _mappedAtMapping;  // inputs, narrowed to float like the reference's arrays
  float _mapped[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
};

}  // namespace loam
