// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h.  Quaternion -> roll / pitch / yaw as tf::Matrix3x3::getRPY defines
// them (only reached through ScanRegistration::handleIMUMessage, which the oracle never calls: the configs carry no IMU).
#pragma once
#include <cmath>
#include <sensor_msgs/Imu.h>
namespace tf {
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
inline void quaternionMsgToTF(const geometry_msgs_shim::Quaternion& m, Quaternion& q) { q.x = m.x; q.y = m.y; q.z = m.z; q.w = m.w; }
struct Matrix3x3 {
  Quaternion q;
  explicit Matrix3x3(const Quaternion& q_) : q(q_) {}
  void getRPY(double& roll, double& pitch, double& yaw) const {
    roll = std::atan2(2.0 * (q.w * q.x + q.y * q.z), 1.0 - 2.0 * (q.x * q.x + q.y * q.y));
    double s = 2.0 * (q.w * q.y - q.z * q.x);
    s = s > 1.0 ? 1.0 : (s < -1.0 ? -1.0 : s);
    pitch = std::asin(s);
    yaw = std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
  }
};
}  // namespace tf
