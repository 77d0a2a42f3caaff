// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h.  nav_msgs::Odometry with the members the reference's LaserOdometry /
// LaserMapping / TransformMaintenance adapters read and write (header, child_frame_id, pose.pose, twist.twist).
#pragma once
#include <string>
#include <boost/shared_ptr.hpp>
#include <geometry_msgs/Quaternion.h>
#include <sensor_msgs/PointCloud2.h>
namespace nav_msgs {
struct Odometry {
  std_msgs_shim::Header header;
  std::string child_frame_id;
  struct { struct { geometry_msgs::Point position; geometry_msgs::Quaternion orientation; } pose; } pose;
  struct { struct { geometry_msgs::Vector3 linear, angular; } twist; } twist;
  typedef boost::shared_ptr<Odometry const> ConstPtr;
};
}  // namespace nav_msgs
