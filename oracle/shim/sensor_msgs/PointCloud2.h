// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h
#pragma once
#include <ros/ros.h>
namespace std_msgs_shim { struct Header { ros::Time stamp; std::string frame_id; }; }
namespace sensor_msgs {
struct PointCloud2 {
  std_msgs_shim::Header header;
  typedef boost::shared_ptr<PointCloud2 const> ConstPtr;
};
typedef boost::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
}  // namespace sensor_msgs
