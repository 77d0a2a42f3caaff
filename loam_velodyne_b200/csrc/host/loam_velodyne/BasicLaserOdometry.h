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

namespace b200 { class Context; struct GaussNewtonSolver; }

class BasicLaserOdometry {
 public:
  explicit BasicLaserOdometry(float scanPeriod = 0.1, size_t maxIterations = 25);
  ~BasicLaserOdometry();
  BasicLaserOdometry(const BasicLaserOdometry&) = delete;
  BasicLaserOdometry& operator=(const BasicLaserOdometry&) = delete;

  void process();
  void updateIMU(pcl::PointCloud<pcl::PointXYZ> const& imuTrans);

  auto& cornerPointsSharp() { return _cornerPointsSharp; }
  auto& cornerPointsLessSharp() { return _cornerPointsLessSharp; }
  auto& surfPointsFlat() { return _surfPointsFlat; }
  auto& surfPointsLessFlat() { return _surfPointsLessFlat; }
  auto& laserCloud() { return _laserCloud; }

  auto const& transformSum() { return _transformSum; }
  auto const& transform() { return _transform; }
  auto const& lastCornerCloud() { return _lastCornerCloud; }
  auto const& lastSurfaceCloud() { return _lastSurfaceCloud; }

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

  // extension: iterations executed by the last process() call
  size_t lastIterationCount() const { return _lastIterations; }

 private:
  void pluginIMURotation(const Angle& bcx, const Angle& bcy, const Angle& bcz, const Angle& blx, const Angle& bly,
                         const Angle& blz, const Angle& alx, const Angle& aly, const Angle& alz, Angle& acx, Angle& acy,
                         Angle& acz);
  void accumulateRotation(Angle cx, Angle cy, Angle cz, Angle lx, Angle ly, Angle lz, Angle& ox, Angle& oy, Angle& oz);
  bool hasIMU() const;
  void uploadLast();

  float _scanPeriod;
  long _frameCount;
  size_t _maxIterations;
  bool _systemInited;
  float _deltaTAbort, _deltaRAbort;

  pcl::PointCloud<pcl::PointXYZI>::Ptr _lastCornerCloud, _lastSurfaceCloud;
  pcl::PointCloud<pcl::PointXYZI>::Ptr _cornerPointsSharp, _cornerPointsLessSharp, _surfPointsFlat, _surfPointsLessFlat,
      _laserCloud;

  Twist _transform, _transformSum;
  Angle _imuRollStart, _imuPitchStart, _imuYawStart, _imuRollEnd, _imuPitchEnd, _imuYawEnd;
  Vector3 _imuShiftFromStart, _imuVeloFromStart;

  b200::Context* _gpu;
  b200::GaussNewtonSolver* _solver;
  std::vector<float> _bufA, _bufB;
  size_t _lastIterations = 0;
};

}  // namespace loam
