// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h.  The message <-> cloud conversions are never executed by the oracle
// (it enters at MultiScanRegistration::process with a pcl cloud); they only have to exist.
#pragma once
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <typename PointT>
void toROSMsg(const PointCloud<PointT>&, sensor_msgs::PointCloud2&) {}
template <typename PointT>
void fromROSMsg(const sensor_msgs::PointCloud2&, PointCloud<PointT>&) {}
}  // namespace pcl
