// loam::BasicLaserOdometry -- drop-in for upstream include/loam_velodyne/BasicLaserOdometry.h:13-48.
// Same constructor, process(), updateIMU(), mutable cloud Ptr accessors (the ROS adapter fills them,
// LaserOdometry.cpp:182-238 upstream), transformSum()/transform()/lastCornerCloud()/lastSurfaceCloud(), setters,
// getters and transformToEnd().  The Gauss-Newton iterations run on the GPU (loam_b200_odom_iterate); the 6x6
// solve, degeneracy projection, convergence test and pose accumulation stay on the host (O(1) per iteration).
#pragma once

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "Twist.h"

namespace loam {

namespace b200 { class Context; class DualCloud; struct GaussNewtonSolver; }
class BasicScanRegistration;

class BasicLaserOdometry {
 public:
  explicit BasicLaserOdometry(float scanPeriod = 0.1, size_t maxIterations = 25);
  ~BasicLaserOdometry();
  BasicLaserOdometry(const BasicLaserOdometry&) = delete;
  BasicLaserOdometry& operator=(const BasicLaserOdometry&) = delete;

  void process();
  void updateIMU(pcl::PointCloud<pcl::PointXYZ> const& imuTrans);

  // input clouds: the caller fills them through these handles (LaserOdometry.cpp:182-238 upstream); handing one
  // out marks the GPU copy stale, so process() uploads it again
  pcl::PointCloud<pcl::PointXYZI>::Ptr& cornerPointsSharp();
  pcl::PointCloud<pcl::PointXYZI>::Ptr& cornerPointsLessSharp();
  pcl::PointCloud<pcl::PointXYZI>::Ptr& surfPointsFlat();
  pcl::PointCloud<pcl::PointXYZI>::Ptr& surfPointsLessFlat();
  pcl::PointCloud<pcl::PointXYZI>::Ptr& laserCloud();

  auto const& transformSum() { return _transformSum; }
  auto const& transform() { return _transform; }
  // result clouds live in HBM; downloaded on first use
  pcl::PointCloud<pcl::PointXYZI>::Ptr const& lastCornerCloud();
  pcl::PointCloud<pcl::PointXYZI>::Ptr const& lastSurfaceCloud();

  void setScanPeriod(float val) { _scanPeriod = val; }
  void setMaxIterations(size_t val) { _maxIterations = val; }
  void setDeltaTAbort(float val) { _deltaTAbort = val; }
  void setDeltaRAbort(float val) { _deltaRAbort = val; }

  auto frameCount() const { return _frameCount; }
  auto scanPeriod() const { return _scanPeriod; }
  auto maxIterations() const { return _maxIterations; }
  auto deltaTAbort() const { return _deltaTAbort; }
  auto deltaRAbort() const { return _deltaRAbort; }

  size_t transformToEnd(pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud);

  // ---- extensions (not part of the reference API) ----
  // iterations executed by the last process() call
  size_t lastIterationCount() const { return _lastIterations; }
  // take this sweep's five clouds + imuTransform from a scan registration object without a host round trip
  // (what ScanRegistration::publishResult -> LaserOdometry::*Handler moves over ROS topics upstream)
  void adopt(BasicScanRegistration& reg);
  // transformToEnd(laserCloud()) on the GPU copy (LaserOdometry::publishResult, LaserOdometry.cpp:326 upstream)
  void transformLaserCloudToEnd();
  b200::Context* deviceContext() { return _gpu; }
  b200::DualCloud& deviceCloud(int which);  // 0 last corner, 1 last surface, 2 laserCloud

 private:
  void pluginIMURotation(const Angle& bcx, const Angle& bcy, const Angle& bcz, const Angle& blx, const Angle& bly,
                         const Angle& blz, const Angle& alx, const Angle& aly, const Angle& alz, Angle& acx, Angle& acy,
                         Angle& acz);
  void accumulateRotation(Angle cx, Angle cy, Angle cz, Angle lx, Angle ly, Angle lz, Angle& ox, Angle& oy, Angle& oz);
  bool hasIMU() const;
  void uploadLast();
  void applyImuToEnd(pcl::PointCloud<pcl::PointXYZI>& cloud);

  float _scanPeriod;
  long _frameCount;
  size_t _maxIterations;
  bool _systemInited;
  float _deltaTAbort, _deltaRAbort;

  b200::DualCloud* _c;  // [7]: sharp, less sharp, flat, less flat, full, last corner, last surface

  Twist _transform, _transformSum;
  Angle _imuRollStart, _imuPitchStart, _imuYawStart, _imuRollEnd, _imuPitchEnd, _imuYawEnd;
  Vector3 _imuShiftFromStart, _imuVeloFromStart;

  b200::Context* _gpu;
  b200::GaussNewtonSolver* _solver;
  std::vector<float> _bufA, _bufB;
  size_t _lastIterations = 0;
};

}  // namespace loam
