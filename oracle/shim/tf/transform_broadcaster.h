// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h.  tf::StampedTransform / TransformBroadcaster as the adapters use
// them (LaserOdometry.cpp:59-60,314-317; LaserMapping.cpp:45-46; TransformMaintenance.cpp): broadcasts are swallowed.
#pragma once
#include <string>
#include <tf/transform_datatypes.h>
namespace tf {
struct StampedTransform {
  ros::Time stamp_;
  std::string frame_id_, child_frame_id_;
  Quaternion rotation_;
  Vector3 origin_;
  void setRotation(const Quaternion& q) { rotation_ = q; }
  void setOrigin(const Vector3& v) { origin_ = v; }
};
struct TransformBroadcaster {
  void sendTransform(const StampedTransform&) {}
};
}  // namespace tf
