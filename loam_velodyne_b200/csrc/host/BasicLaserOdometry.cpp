// Host side of the laser-odometry drop-in.  Control flow mirrors upstream BasicLaserOdometry::process
// (src/lib/BasicLaserOdometry.cpp:196-666); per-point work happens in loam_b200_odom_iterate.
#include "loam_velodyne/BasicLaserOdometry.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <cassert>
#include <cmath>

#include "b200_runtime.h"
#include "host_math.h"
#include "loam_velodyne/BasicScanRegistration.h"

namespace loam {

using hostmath::rad2deg;

static void fillOdomPoseOf(const Twist& t, float scanPeriod, int iter, loam_b200_odom_pose& p) {
  p.rot[0] = t.rot_x.rad(); p.rot[1] = t.rot_y.rad(); p.rot[2] = t.rot_z.rad();
  // upstream re-evaluates std::sin / std::cos of the float angle per Jacobian row (:504-509): same values as the
  // Angle caches
  p.sin_[0] = t.rot_x.sin(); p.sin_[1] = t.rot_y.sin(); p.sin_[2] = t.rot_z.sin();
  p.cos_[0] = t.rot_x.cos(); p.cos_[1] = t.rot_y.cos(); p.cos_[2] = t.rot_z.cos();
  p.pos[0] = t.pos.x(); p.pos[1] = t.pos.y(); p.pos[2] = t.pos.z();
  p.inv_scan_period = 1.f / scanPeriod;
  p.iter = iter;
}

enum { C_SHARP = 0, C_LESS_SHARP, C_FLAT, C_LESS_FLAT, C_FULL, C_LAST_CORNER, C_LAST_SURF, C_NUM };

BasicLaserOdometry::BasicLaserOdometry(float scanPeriod, size_t maxIterations)
    : _scanPeriod(scanPeriod), _frameCount(0), _maxIterations(maxIterations), _systemInited(false), _deltaTAbort(0.1),
      _deltaRAbort(0.1), _c(new b200::DualCloud[C_NUM]), _gpu(new b200::Context()),
      _solver(new b200::GaussNewtonSolver()) {
  static const int slots[C_NUM] = {LOAM_B200_C_ODOM_SHARP, LOAM_B200_C_ODOM_LESS_SHARP, LOAM_B200_C_ODOM_FLAT,
                                   LOAM_B200_C_ODOM_LESS_FLAT, LOAM_B200_C_ODOM_FULL, LOAM_B200_C_ODOM_LAST_CORNER,
                                   LOAM_B200_C_ODOM_LAST_SURF};
  for (int i = 0; i < C_NUM; i++) _c[i].bind(_gpu, slots[i]);
  // the scan-to-scan loop is a chain of small latency-bound kernels: when the three stages share a GPU (in-process
  // pipeline) its stream goes ahead of the registration's and the map update's bulk kernels
  _gpu->setPriority(1);
}

BasicLaserOdometry::~BasicLaserOdometry() {
  delete _solver;
  delete[] _c;
  delete _gpu;
}

pcl::PointCloud<pcl::PointXYZI>::Ptr& BasicLaserOdometry::cornerPointsSharp() { return _c[C_SHARP].hostPtrMutable(); }
pcl::PointCloud<pcl::PointXYZI>::Ptr& BasicLaserOdometry::cornerPointsLessSharp() { return _c[C_LESS_SHARP].hostPtrMutable(); }
pcl::PointCloud<pcl::PointXYZI>::Ptr& BasicLaserOdometry::surfPointsFlat() { return _c[C_FLAT].hostPtrMutable(); }
pcl::PointCloud<pcl::PointXYZI>::Ptr& BasicLaserOdometry::surfPointsLessFlat() { return _c[C_LESS_FLAT].hostPtrMutable(); }
pcl::PointCloud<pcl::PointXYZI>::Ptr& BasicLaserOdometry::laserCloud() { return _c[C_FULL].hostPtrMutable(); }
pcl::PointCloud<pcl::PointXYZI>::Ptr const& BasicLaserOdometry::lastCornerCloud() { return _c[C_LAST_CORNER].hostPtr(); }
pcl::PointCloud<pcl::PointXYZI>::Ptr const& BasicLaserOdometry::lastSurfaceCloud() { return _c[C_LAST_SURF].hostPtr(); }
b200::DualCloud& BasicLaserOdometry::deviceCloud(int which) {
  return _c[which == 0 ? C_LAST_CORNER : which == 1 ? C_LAST_SURF : C_FULL];
}

void BasicLaserOdometry::adopt(BasicScanRegistration& reg) {
  // reg clouds: 0 full, 1 sharp, 2 less sharp, 3 flat, 4 less flat
  static const int from[5] = {1, 2, 3, 4, 0};
  static const int to[5] = {C_SHARP, C_LESS_SHARP, C_FLAT, C_LESS_FLAT, C_FULL};
  int dstSlots[5], srcSlots[5];
  for (int i = 0; i < 5; i++) {
    b200::DualCloud& src = reg.deviceCloud(from[i]);
    src.ensureDevice();
    dstSlots[i] = _c[to[i]].slot();
    srcSlots[i] = src.slot();
  }
  (void)dstSlots;
  // one launch: the five clouds and this sweep's query array (sharp + flat points)
  _gpu->check(loam_b200_odom_adopt(_gpu->get(), reg.deviceContext()->get(), srcSlots), "loam_b200_odom_adopt");
  for (int i = 0; i < 5; i++) _c[to[i]].deviceWritten((int)reg.deviceCloud(from[i]).size());
  updateIMU(reg.imuTransform());
}

void BasicLaserOdometry::transformLaserCloudToEnd() {
  if (_c[C_FULL].size() == 0) return;
  _c[C_FULL].ensureDevice();
  loam_b200_odom_pose p;
  fillOdomPoseOf(_transform, _scanPeriod, 0, p);
  _gpu->check(loam_b200_cloud_transform_to_end(_gpu->get(), _c[C_FULL].slot(), &p), "loam_b200_cloud_transform_to_end");
  _c[C_FULL].deviceWritten((int)_c[C_FULL].size());
  if (hasIMU()) applyImuToEnd(_c[C_FULL].hostMutable());
}

bool BasicLaserOdometry::hasIMU() const {
  return _imuRollStart.rad() != 0.f || _imuPitchStart.rad() != 0.f || _imuYawStart.rad() != 0.f ||
         _imuRollEnd.rad() != 0.f || _imuPitchEnd.rad() != 0.f || _imuYawEnd.rad() != 0.f ||
         _imuShiftFromStart.x() != 0.f || _imuShiftFromStart.y() != 0.f || _imuShiftFromStart.z() != 0.f;
}

void BasicLaserOdometry::updateIMU(pcl::PointCloud<pcl::PointXYZ> const& imuTrans) {
  assert(4 == imuTrans.size());
  _imuPitchStart = imuTrans.points[0].x;
  _imuYawStart = imuTrans.points[0].y;
  _imuRollStart = imuTrans.points[0].z;
  _imuPitchEnd = imuTrans.points[1].x;
  _imuYawEnd = imuTrans.points[1].y;
  _imuRollEnd = imuTrans.points[1].z;
  _imuShiftFromStart = imuTrans.points[2];
  _imuVeloFromStart = imuTrans.points[3];
}

size_t BasicLaserOdometry::transformToEnd(pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud) {
  const size_t n = cloud->points.size();
  if (n == 0) return 0;
  // an arbitrary caller-owned cloud: host buffer in, host buffer out
  b200::pack(*cloud, _bufA);
  loam_b200_odom_pose p;
  fillOdomPoseOf(_transform, _scanPeriod, 0, p);
  _gpu->check(loam_b200_transform_to_end(_gpu->get(), _bufA.data(), (int)n, &p), "loam_b200_transform_to_end");
  b200::unpack(_bufA.data(), n, *cloud);
  if (hasIMU()) applyImuToEnd(*cloud);
  return n;
}

// IMU terms of upstream :78-83 (all identities without IMU data)
void BasicLaserOdometry::applyImuToEnd(pcl::PointCloud<pcl::PointXYZI>& cloud) {
  for (auto& pt : cloud.points) {
    pt.x += -_imuShiftFromStart.x();
    pt.y += -_imuShiftFromStart.y();
    pt.z += -_imuShiftFromStart.z();
    hostmath::rotateZXY(pt, _imuRollStart, _imuPitchStart, _imuYawStart);
    hostmath::rotateYXZ(pt, -_imuYawEnd, -_imuPitchEnd, -_imuRollEnd);
  }
}

void BasicLaserOdometry::uploadLast() {
  _c[C_LAST_CORNER].dropNonFinite();  // upstream :252 (there after the tree was built; here before, so indices stay valid)
  _c[C_LAST_CORNER].ensureDevice();
  _c[C_LAST_SURF].ensureDevice();
  _gpu->check(loam_b200_odom_rebuild_last(_gpu->get()), "loam_b200_odom_rebuild_last");
}

void BasicLaserOdometry::process() {
  if (!_systemInited) {
    _c[C_LESS_SHARP].swap(_c[C_LAST_CORNER]);
    _c[C_LESS_FLAT].swap(_c[C_LAST_SURF]);
    uploadLast();
    _transformSum.rot_x += _imuPitchStart;
    _transformSum.rot_z += _imuRollStart;
    _systemInited = true;
    return;
  }

  _frameCount++;
  _transform.pos -= _imuVeloFromStart * _scanPeriod;
  _lastIterations = 0;
  // `bool isDegenerate = false` is a local of process() upstream (:212): a projection of an earlier sweep is never applied
  _solver->isDegenerate = false;
  static const bool trace = std::getenv("LOAM_B200_TRACE") != nullptr;
  auto tnow = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tr0 = tnow();
  double tr1 = tr0, tr2 = tr0;

  size_t lastCornerCloudSize = _c[C_LAST_CORNER].size();
  size_t lastSurfaceCloudSize = _c[C_LAST_SURF].size();

  if (lastCornerCloudSize > 10 && lastSurfaceCloudSize > 100) {
    // upstream :230 / :252: removeNaNFromPointCloud on the sharp cloud and on the last corner cloud.  PCL only filters
    // clouds flagged !is_dense; clouds adopted on the device are dense by construction (the front end drops non-finite
    // points, MultiScanRegistration.cpp:187-192), so this only acts on caller-filled host clouds
    _c[C_SHARP].dropNonFinite();
    _c[C_SHARP].ensureDevice();
    _c[C_FLAT].ensureDevice();
    _gpu->check(loam_b200_odom_prepare(_gpu->get()), "loam_b200_odom_prepare");
    tr1 = tnow();

    if (b200::deviceResidentLoops()) {
      // optional: the whole iteration loop (:246-622) on the device (loam_b200_odom_solve, csrc/lmstep.cuh)
      const float rot[3] = {_transform.rot_x.rad(), _transform.rot_y.rad(), _transform.rot_z.rad()};
      const float pos[3] = {_transform.pos.x(), _transform.pos.y(), _transform.pos.z()};
      loam_b200_lm_result res;
      _gpu->check(loam_b200_odom_solve(_gpu->get(), rot, pos, 1.f / _scanPeriod, (int)_maxIterations, _deltaTAbort,
                                       _deltaRAbort, &res),
                  "loam_b200_odom_solve");
      _lastIterations = (size_t)res.iterations;
      _transform.rot_x = res.rot[0];
      _transform.rot_y = res.rot[1];
      _transform.rot_z = res.rot[2];
      _transform.pos.x() = res.pos[0];
      _transform.pos.y() = res.pos[1];
      _transform.pos.z() = res.pos[2];
    } else
    for (size_t iterCount = 0; iterCount < _maxIterations; iterCount++) {
      _lastIterations = iterCount + 1;
      loam_b200_odom_pose pose;
      fillOdomPoseOf(_transform, _scanPeriod, (int)iterCount, pose);
      loam_b200_normal_eq ne;
      _gpu->check(loam_b200_odom_iterate(_gpu->get(), &pose, &ne), "loam_b200_odom_iterate");
      if (ne.n_selected < 10) continue;

      float x[6];
      _solver->solve(ne, iterCount == 0, 10.f, x);

      _transform.rot_x = _transform.rot_x.rad() + x[0];
      _transform.rot_y = _transform.rot_y.rad() + x[1];
      _transform.rot_z = _transform.rot_z.rad() + x[2];
      _transform.pos.x() += x[3];
      _transform.pos.y() += x[4];
      _transform.pos.z() += x[5];

      if (!std::isfinite(_transform.rot_x.rad())) _transform.rot_x = Angle();
      if (!std::isfinite(_transform.rot_y.rad())) _transform.rot_y = Angle();
      if (!std::isfinite(_transform.rot_z.rad())) _transform.rot_z = Angle();
      if (!std::isfinite(_transform.pos.x())) _transform.pos.x() = 0.0;
      if (!std::isfinite(_transform.pos.y())) _transform.pos.y() = 0.0;
      if (!std::isfinite(_transform.pos.z())) _transform.pos.z() = 0.0;

      const float deltaR = std::sqrt(std::pow(rad2deg(x[0]), 2) + std::pow(rad2deg(x[1]), 2) + std::pow(rad2deg(x[2]), 2));
      const float deltaT = std::sqrt(std::pow(x[3] * 100, 2) + std::pow(x[4] * 100, 2) + std::pow(x[5] * 100, 2));
      if (deltaR < _deltaRAbort && deltaT < _deltaTAbort) break;
    }
  }

  tr2 = tnow();
  Angle rx, ry, rz;
  accumulateRotation(_transformSum.rot_x, _transformSum.rot_y, _transformSum.rot_z, -_transform.rot_x,
                     -_transform.rot_y.rad() * 1.05, -_transform.rot_z, rx, ry, rz);

  Vector3 v(_transform.pos.x() - _imuShiftFromStart.x(), _transform.pos.y() - _imuShiftFromStart.y(),
            _transform.pos.z() * 1.05 - _imuShiftFromStart.z());
  hostmath::rotateZXY(v, rz, rx, ry);
  Vector3 trans = _transformSum.pos - v;

  pluginIMURotation(rx, ry, rz, _imuPitchStart, _imuYawStart, _imuRollStart, _imuPitchEnd, _imuYawEnd, _imuRollEnd, rx,
                    ry, rz);

  _transformSum.rot_x = rx;
  _transformSum.rot_y = ry;
  _transformSum.rot_z = rz;
  _transformSum.pos = trans;

  // transformToEnd(less sharp / less flat) in HBM, then they become the "last" clouds (:651-655)
  loam_b200_odom_pose endPose;
  fillOdomPoseOf(_transform, _scanPeriod, 0, endPose);
  _c[C_LESS_SHARP].ensureDevice();
  _c[C_LESS_FLAT].ensureDevice();
  _gpu->check(loam_b200_cloud_transform_to_end2(_gpu->get(), _c[C_LESS_SHARP].slot(), _c[C_LESS_FLAT].slot(), &endPose),
              "loam_b200_cloud_transform_to_end2");
  for (int which : {C_LESS_SHARP, C_LESS_FLAT}) {
    if (_c[which].size() == 0) continue;
    _c[which].deviceWritten((int)_c[which].size());
    if (hasIMU()) applyImuToEnd(_c[which].hostMutable());
  }
  _c[C_LESS_SHARP].swap(_c[C_LAST_CORNER]);
  _c[C_LESS_FLAT].swap(_c[C_LAST_SURF]);

  lastCornerCloudSize = _c[C_LAST_CORNER].size();
  lastSurfaceCloudSize = _c[C_LAST_SURF].size();
  const double tr3 = tnow();
  if (lastCornerCloudSize > 10 && lastSurfaceCloudSize > 100) uploadLast();
  if (trace)
    fprintf(stderr, "[odom] prepare %.0f us, loop %.0f us (%zu it), to-end %.0f us, rebuild issue %.0f us\n", tr1 - tr0, tr2 - tr1,
            _lastIterations, tr3 - tr2, tnow() - tr3);
}

// Euler-angle composition helpers: closed-form products of ZXY rotations, as published with LOAM
// (upstream BasicLaserOdometry.cpp:91-179).
void BasicLaserOdometry::pluginIMURotation(const Angle& bcx, const Angle& bcy, const Angle& bcz, const Angle& blx,
                                           const Angle& bly, const Angle& blz, const Angle& alx, const Angle& aly,
                                           const Angle& alz, Angle& acx, Angle& acy, Angle& acz) {
  const float sbcx = bcx.sin(), cbcx = bcx.cos(), sbcy = bcy.sin(), cbcy = bcy.cos(), sbcz = bcz.sin(), cbcz = bcz.cos();
  const float sblx = blx.sin(), cblx = blx.cos(), sbly = bly.sin(), cbly = bly.cos(), sblz = blz.sin(), cblz = blz.cos();
  const float salx = alx.sin(), calx = alx.cos(), saly = aly.sin(), caly = aly.cos(), salz = alz.sin(), calz = alz.cos();

  const float srx = -sbcx * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly) -
                    cbcx * cbcz * (calx * saly * (cbly * sblz - cblz * sblx * sbly) -
                                   calx * caly * (sbly * sblz + cbly * cblz * sblx) + cblx * cblz * salx) -
                    cbcx * sbcz * (calx * caly * (cblz * sbly - cbly * sblx * sblz) -
                                   calx * saly * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sblz);
  acx = -std::asin(srx);

  const float srycrx = (cbcy * sbcz - cbcz * sbcx * sbcy) * (calx * saly * (cbly * sblz - cblz * sblx * sbly) -
                                                             calx * caly * (sbly * sblz + cbly * cblz * sblx) +
                                                             cblx * cblz * salx) -
                       (cbcy * cbcz + sbcx * sbcy * sbcz) * (calx * caly * (cblz * sbly - cbly * sblx * sblz) -
                                                             calx * saly * (cbly * cblz + sblx * sbly * sblz) +
                                                             cblx * salx * sblz) +
                       cbcx * sbcy * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly);
  const float crycrx = (cbcz * sbcy - cbcy * sbcx * sbcz) * (calx * caly * (cblz * sbly - cbly * sblx * sblz) -
                                                             calx * saly * (cbly * cblz + sblx * sbly * sblz) +
                                                             cblx * salx * sblz) -
                       (sbcy * sbcz + cbcy * cbcz * sbcx) * (calx * saly * (cbly * sblz - cblz * sblx * sbly) -
                                                             calx * caly * (sbly * sblz + cbly * cblz * sblx) +
                                                             cblx * cblz * salx) +
                       cbcx * cbcy * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly);
  acy = std::atan2(srycrx / acx.cos(), crycrx / acx.cos());

  const float srzcrx = sbcx * (cblx * cbly * (calz * saly - caly * salx * salz) -
                               cblx * sbly * (caly * calz + salx * saly * salz) + calx * salz * sblx) -
                       cbcx * cbcz * ((caly * calz + salx * saly * salz) * (cbly * sblz - cblz * sblx * sbly) +
                                      (calz * saly - caly * salx * salz) * (sbly * sblz + cbly * cblz * sblx) -
                                      calx * cblx * cblz * salz) +
                       cbcx * sbcz * ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) +
                                      (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) +
                                      calx * cblx * salz * sblz);
  const float crzcrx = sbcx * (cblx * sbly * (caly * salz - calz * salx * saly) -
                               cblx * cbly * (saly * salz + caly * calz * salx) + calx * calz * sblx) +
                       cbcx * cbcz * ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) +
                                      (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) +
                                      calx * calz * cblx * cblz) -
                       cbcx * sbcz * ((saly * salz + caly * calz * salx) * (cblz * sbly - cbly * sblx * sblz) +
                                      (caly * salz - calz * salx * saly) * (cbly * cblz + sblx * sbly * sblz) -
                                      calx * calz * cblx * sblz);
  acz = std::atan2(srzcrx / acx.cos(), crzcrx / acx.cos());
}

void BasicLaserOdometry::accumulateRotation(Angle cx, Angle cy, Angle cz, Angle lx, Angle ly, Angle lz, Angle& ox,
                                            Angle& oy, Angle& oz) {
  const float srx = lx.cos() * cx.cos() * ly.sin() * cz.sin() - cx.cos() * cz.cos() * lx.sin() -
                    lx.cos() * ly.cos() * cx.sin();
  ox = -std::asin(srx);

  const float srycrx = lx.sin() * (cy.cos() * cz.sin() - cz.cos() * cx.sin() * cy.sin()) +
                       lx.cos() * ly.sin() * (cy.cos() * cz.cos() + cx.sin() * cy.sin() * cz.sin()) +
                       lx.cos() * ly.cos() * cx.cos() * cy.sin();
  const float crycrx = lx.cos() * ly.cos() * cx.cos() * cy.cos() -
                       lx.cos() * ly.sin() * (cz.cos() * cy.sin() - cy.cos() * cx.sin() * cz.sin()) -
                       lx.sin() * (cy.sin() * cz.sin() + cy.cos() * cz.cos() * cx.sin());
  oy = std::atan2(srycrx / ox.cos(), crycrx / ox.cos());

  const float srzcrx = cx.sin() * (lz.cos() * ly.sin() - ly.cos() * lx.sin() * lz.sin()) +
                       cx.cos() * cz.sin() * (ly.cos() * lz.cos() + lx.sin() * ly.sin() * lz.sin()) +
                       lx.cos() * cx.cos() * cz.cos() * lz.sin();
  const float crzcrx = lx.cos() * lz.cos() * cx.cos() * cz.cos() -
                       cx.cos() * cz.sin() * (ly.cos() * lz.sin() - lz.cos() * lx.sin() * ly.sin()) -
                       cx.sin() * (ly.sin() * lz.sin() + ly.cos() * lz.cos() * lx.sin());
  oz = std::atan2(srzcrx / ox.cos(), crzcrx / ox.cos());
}

}  // namespace loam
