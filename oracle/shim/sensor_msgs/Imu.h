// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h
#pragma once
#include <sensor_msgs/PointCloud2.h>
#include <geometry_msgs/Quaternion.h>
namespace sensor_msgs {
struct Imu {
  std_msgs_shim::Header header;
  geometry_msgs_shim::Quaternion orientation;
  geometry_msgs_shim::Vector3 linear_acceleration;
  typedef boost::shared_ptr<Imu const> ConstPtr;
};
}  // namespace sensor_msgs
