// loam::BasicTransformMaintenance (drop-in, host only): see loam_velodyne/BasicTransformMaintenance.h.
#include "loam_velodyne/BasicTransformMaintenance.h"

#include "pose_algebra.h"

namespace loam {

namespace {
// the reference keeps every pose as six floats: narrow first, then build the cached sines / cosines from the floats
Twist narrowed(double rx, double ry, double rz, double x, double y, double z) {
  Twist w;
  w.rot_x = (float)rx;
  w.rot_y = (float)ry;
  w.rot_z = (float)rz;
  w.pos.x() = (float)x;
  w.pos.y() = (float)y;
  w.pos.z() = (float)z;
  return w;
}
}  // namespace

void BasicTransformMaintenance::updateOdometry(double pitch, double yaw, double roll, double x, double y, double z) {
  _odometry = narrowed(pitch, yaw, roll, x, y, z);
}

void BasicTransformMaintenance::updateMappingTransform(double pitch, double yaw, double roll, double x, double y, double z,
                                                       double twist_rot_x, double twist_rot_y, double twist_rot_z,
                                                       double twist_pos_x, double twist_pos_y, double twist_pos_z) {
  _mappedAtMapping = narrowed(pitch, yaw, roll, x, y, z);
  _odometryAtMapping = narrowed(twist_rot_x, twist_rot_y, twist_rot_z, twist_pos_x, twist_pos_y, twist_pos_z);
}

void BasicTransformMaintenance::updateMappingTransform(Twist const& aft, Twist const& bef) {
  _mappedAtMapping = aft;
  _odometryAtMapping = bef;
}

void BasicTransformMaintenance::transformAssociateToMap() {
  Twist increment, fused;
  hostmath::associateToMap(_odometry, _odometryAtMapping, _mappedAtMapping, increment, fused);
  _mapped[0] = fused.rot_x.rad();
  _mapped[1] = fused.rot_y.rad();
  _mapped[2] = fused.rot_z.rad();
  _mapped[3] = fused.pos.x();
  _mapped[4] = fused.pos.y();
  _mapped[5] = fused.pos.z();
}

}  // namespace loam
