// TEST INFRASTRUCTURE (oracle/shim): see ros/ros.h.  Quaternion -> roll / pitch / yaw as tf::Matrix3x3::getRPY defines
// them (only reached through ScanRegistration::handleIMUMessage, which the oracle never calls: the configs carry no IMU).
#pragma once
#include <cmath>
#include <sensor_msgs/Imu.h>
namespace tf {
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
  Quaternion() {}
  Quaternion(double x_, double y_, double z_, double w_) : x(x_), y(y_), z(z_), w(w_) {}
};
struct Vector3 {
  double x = 0, y = 0, z = 0;
  Vector3() {}
  Vector3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};
// roll about x, pitch about y, yaw about z (tf::createQuaternionMsgFromRollPitchYaw)
inline geometry_msgs::Quaternion createQuaternionMsgFromRollPitchYaw(double roll, double pitch, double yaw) {
  const double cr = std::cos(roll / 2), sr = std::sin(roll / 2), cp = std::cos(pitch / 2), sp = std::sin(pitch / 2),
               cy = std::cos(yaw / 2), sy = std::sin(yaw / 2);
  geometry_msgs::Quaternion q;
  q.x = sr * cp * cy - cr * sp * sy;
  q.y = cr * sp * cy + sr * cp * sy;
  q.z = cr * cp * sy - sr * sp * cy;
  q.w = cr * cp * cy + sr * sp * sy;
  return q;
}
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
