// loam::BasicTransformMaintenance (drop-in, host only): see loam_velodyne/BasicTransformMaintenance.h.
#include "loam_velodyne/BasicTransformMaintenance.h"

#include "pose_algebra.h"

namespace loam {

namespace {
Twist twistOf(const float t[6]) {
  Twist w;
  w.rot_x = t[0];
  w.rot_y = t[1];
  w.rot_z = t[2];
  w.pos.x() = t[3];
  w.pos.y() = t[4];
  w.pos.z() = t[5];
  return w;
}
}  // namespace

void BasicTransformMaintenance::updateOdometry(double pitch, double yaw, double roll, double x, double y, double z) {
  const double v[6] = {pitch, yaw, roll, x, y, z};
  for (int i = 0; i < 6; i++) _transformSum[i] = (float)v[i];
}

void BasicTransformMaintenance::updateMappingTransform(double pitch, double yaw, double roll, double x, double y, double z,
                                                       double twist_rot_x, double twist_rot_y, double twist_rot_z,
                                                       double twist_pos_x, double twist_pos_y, double twist_pos_z) {
  const double a[6] = {pitch, yaw, roll, x, y, z};
  const double b[6] = {twist_rot_x, twist_rot_y, twist_rot_z, twist_pos_x, twist_pos_y, twist_pos_z};
  for (int i = 0; i < 6; i++) {
    _transformAftMapped[i] = (float)a[i];
    _transformBefMapped[i] = (float)b[i];
  }
}

void BasicTransformMaintenance::updateMappingTransform(Twist const& aft, Twist const& bef) {
  updateMappingTransform(aft.rot_x.rad(), aft.rot_y.rad(), aft.rot_z.rad(), aft.pos.x(), aft.pos.y(), aft.pos.z(),
                         bef.rot_x.rad(), bef.rot_y.rad(), bef.rot_z.rad(), bef.pos.x(), bef.pos.y(), bef.pos.z());
}

void BasicTransformMaintenance::transformAssociateToMap() {
  Twist incre, mapped;
  hostmath::associateToMap(twistOf(_transformSum), twistOf(_transformBefMapped), twistOf(_transformAftMapped), incre, mapped);
  _transformIncre[3] = incre.pos.x();
  _transformIncre[4] = incre.pos.y();
  _transformIncre[5] = incre.pos.z();
  _transformMapped[0] = mapped.rot_x.rad();
  _transformMapped[1] = mapped.rot_y.rad();
  _transformMapped[2] = mapped.rot_z.rad();
  _transformMapped[3] = mapped.pos.x();
  _transformMapped[4] = mapped.pos.y();
  _transformMapped[5] = mapped.pos.z();
}

}  // namespace loam
