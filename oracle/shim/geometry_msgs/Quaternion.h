// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h
#pragma once
namespace geometry_msgs {
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
}  // namespace geometry_msgs
namespace geometry_msgs_shim { using geometry_msgs::Quaternion; using geometry_msgs::Vector3; }
