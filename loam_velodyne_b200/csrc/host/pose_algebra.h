// Pose association shared by BasicLaserMapping::transformAssociateToMap (upstream BasicLaserMapping.cpp:103-167) and
// BasicTransformMaintenance::transformAssociateToMap (upstream BasicTransformMaintenance.cpp:83-178): the odometry pose
// `sum`, the odometry pose at the last mapping update `bef` and the mapped pose of that update `aft` are composed into
// the current map-frame pose `out` (closed-form products of ZXY Euler rotations, as published with LOAM); `incre` receives
// the translation increment in the sensor frame.  fp32 throughout, sines / cosines as cached by loam::Angle.
#pragma once

#include <cmath>

#include "host_math.h"
#include "loam_velodyne/Twist.h"

namespace loam {
namespace hostmath {

inline void associateToMap(const Twist& sum, const Twist& bef, const Twist& aft, Twist& incre, Twist& out) {
  incre.pos = bef.pos - sum.pos;
  hostmath::rotateYXZ(incre.pos, -(sum.rot_y), -(sum.rot_x), -(sum.rot_z));

  const float sbcx = sum.rot_x.sin(), cbcx = sum.rot_x.cos();
  const float sbcy = sum.rot_y.sin(), cbcy = sum.rot_y.cos();
  const float sbcz = sum.rot_z.sin(), cbcz = sum.rot_z.cos();
  const float sblx = bef.rot_x.sin(), cblx = bef.rot_x.cos();
  const float sbly = bef.rot_y.sin(), cbly = bef.rot_y.cos();
  const float sblz = bef.rot_z.sin(), cblz = bef.rot_z.cos();
  const float salx = aft.rot_x.sin(), calx = aft.rot_x.cos();
  const float saly = aft.rot_y.sin(), caly = aft.rot_y.cos();
  const float salz = aft.rot_z.sin(), calz = aft.rot_z.cos();

  const float srx = -sbcx * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz) -
                    cbcx * sbcy * (calx * calz * (cbly * sblz - cblz * sblx * sbly) -
                                   calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                    cbcx * cbcy * (calx * salz * (cblz * sbly - cbly * sblx * sblz) -
                                   calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx);
  out.rot_x = -std::asin(srx);

  const float srycrx = sbcx * (cblx * cblz * (caly * salz - calz * salx * saly) -
                               cblx * sblz * (caly * calz + salx * saly * salz) + calx * saly * sblx) -
                       cbcx * cbcy * ((caly * calz + salx * saly * salz) * (cblz * sbly - cbly * sblx * sblz) +
                                      (caly * salz - calz * salx * saly) * (sbly * sblz + cbly * cblz * sblx) -
                                      calx * cblx * cbly * saly) +
                       cbcx * sbcy * ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) +
                                      (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) +
                                      calx * cblx * saly * sbly);
  const float crycrx = sbcx * (cblx * sblz * (calz * saly - caly * salx * salz) -
                               cblx * cblz * (saly * salz + caly * calz * salx) + calx * caly * sblx) +
                       cbcx * cbcy * ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) +
                                      (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) +
                                      calx * caly * cblx * cbly) -
                       cbcx * sbcy * ((saly * salz + caly * calz * salx) * (cbly * sblz - cblz * sblx * sbly) +
                                      (calz * saly - caly * salx * salz) * (cbly * cblz + sblx * sbly * sblz) -
                                      calx * caly * cblx * sbly);
  out.rot_y = std::atan2(srycrx / out.rot_x.cos(), crycrx / out.rot_x.cos());

  const float srzcrx = (cbcz * sbcy - cbcy * sbcx * sbcz) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) -
                                                             calx * calz * (sbly * sblz + cbly * cblz * sblx) +
                                                             cblx * cbly * salx) -
                       (cbcy * cbcz + sbcx * sbcy * sbcz) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) -
                                                             calx * salz * (cbly * cblz + sblx * sbly * sblz) +
                                                             cblx * salx * sbly) +
                       cbcx * sbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
  const float crzcrx = (cbcy * sbcz - cbcz * sbcx * sbcy) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) -
                                                             calx * salz * (cbly * cblz + sblx * sbly * sblz) +
                                                             cblx * salx * sbly) -
                       (sbcy * sbcz + cbcy * cbcz * sbcx) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) -
                                                             calx * calz * (sbly * sblz + cbly * cblz * sblx) +
                                                             cblx * cbly * salx) +
                       cbcx * cbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
  out.rot_z = std::atan2(srzcrx / out.rot_x.cos(), crzcrx / out.rot_x.cos());

  Vector3 v = incre.pos;
  hostmath::rotateZXY(v, out.rot_z, out.rot_x, out.rot_y);
  out.pos = aft.pos - v;
}

}  // namespace hostmath
}  // namespace loam
