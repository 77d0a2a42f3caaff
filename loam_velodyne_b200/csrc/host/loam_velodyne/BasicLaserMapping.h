// loam::BasicLaserMapping -- drop-in for upstream include/loam_velodyne/BasicLaserMapping.h:77-111.
// Same constructor, process(Time), updateIMU, updateOdometry (both overloads), mutable cloud references
// (laserCloud / laserCloudCornerLast / laserCloudSurfLast), downSizeFilter*() handles, setters / getters,
// transformAftMapped / transformBefMapped / laserCloudSurroundDS / hasFreshMap.
// The scan-to-map optimisation (optimizeTransformTobeMapped, BasicLaserMapping.cpp:626-926 upstream) runs its
// per-point work on the GPU: BVH build over the surrounding-map clouds, fused 5-NN + fit + Jacobian + normal-equation
// reduction per iteration; the host keeps the 6x6 solve and the O(1) pose algebra.
#pragma once

#include <vector>

#include <pcl/filters/voxel_grid.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "Twist.h"
#include "time_utils.h"

namespace loam {

namespace b200 { class Context; struct GaussNewtonSolver; }

typedef struct IMUState2 {
  Time stamp;
  Angle roll;
  Angle pitch;
  static void interpolate(const IMUState2& start, const IMUState2& end, const float& ratio, IMUState2& result) {
    const float invRatio = 1 - ratio;
    result.roll = start.roll.rad() * invRatio + end.roll.rad() * ratio;
    result.pitch = start.pitch.rad() * invRatio + end.pitch.rad() * ratio;
  }
} IMUState2;

class BasicLaserMapping {
 public:
  explicit BasicLaserMapping(const float& scanPeriod = 0.1, const size_t& maxIterations = 10);
  ~BasicLaserMapping();
  BasicLaserMapping(const BasicLaserMapping&) = delete;
  BasicLaserMapping& operator=(const BasicLaserMapping&) = delete;

  bool process(Time const& laserOdometryTime);
  void updateIMU(IMUState2 const& newState);
  void updateOdometry(double pitch, double yaw, double roll, double x, double y, double z);
  void updateOdometry(Twist const& twist);

  auto& laserCloud() { return *_laserCloudFullRes; }
  auto& laserCloudCornerLast() { return *_laserCloudCornerLast; }
  auto& laserCloudSurfLast() { return *_laserCloudSurfLast; }

  void setScanPeriod(float val) { _scanPeriod = val; }
  void setMaxIterations(size_t val) { _maxIterations = val; }
  void setDeltaTAbort(float val) { _deltaTAbort = val; }
  void setDeltaRAbort(float val) { _deltaRAbort = val; }

  auto& downSizeFilterCorner() { return _downSizeFilterCorner; }
  auto& downSizeFilterSurf() { return _downSizeFilterSurf; }
  auto& downSizeFilterMap() { return _downSizeFilterMap; }

  auto frameCount() const { return _frameCount; }
  auto scanPeriod() const { return _scanPeriod; }
  auto maxIterations() const { return _maxIterations; }
  auto deltaTAbort() const { return _deltaTAbort; }
  auto deltaRAbort() const { return _deltaRAbort; }

  auto const& transformAftMapped() const { return _transformAftMapped; }
  auto const& transformBefMapped() const { return _transformBefMapped; }
  auto const& laserCloudSurroundDS() const { return *_laserCloudSurroundDS; }

  bool hasFreshMap() const { return _downsizedMapCreated; }

  // extensions (not in the reference API): pre-seed the cube map with points given in the map frame, using the same
  // cube-index arithmetic as the insertion step (upstream BasicLaserMapping.cpp:540-553); introspection for tests.
  void seedMap(pcl::PointCloud<pcl::PointXYZI> const& cornerPoints, pcl::PointCloud<pcl::PointXYZI> const& surfPoints);
  size_t lastIterationCount() const { return _lastIterations; }
  auto const& transformTobeMapped() const { return _transformTobeMapped; }
  auto const& cornerStackDS() const { return *_laserCloudCornerStackDS; }
  auto const& surfStackDS() const { return *_laserCloudSurfStackDS; }
  auto const& cornerFromMap() const { return *_laserCloudCornerFromMap; }
  auto const& surfFromMap() const { return *_laserCloudSurfFromMap; }
  void collectMap(pcl::PointCloud<pcl::PointXYZI>& corner, pcl::PointCloud<pcl::PointXYZI>& surf) const;

 private:
  typedef pcl::PointCloud<pcl::PointXYZI> Cloud;
  void optimizeTransformTobeMapped();
  void transformAssociateToMap();
  void transformUpdate();
  void pointAssociateToMap(const pcl::PointXYZI& pi, pcl::PointXYZI& po);
  void pointAssociateTobeMapped(const pcl::PointXYZI& pi, pcl::PointXYZI& po);
  void transformFullResToMap();
  bool createDownsizedMap();
  size_t toIndex(int i, int j, int k) const { return i + _laserCloudWidth * j + _laserCloudWidth * _laserCloudHeight * k; }
  bool cubeIndexOf(const pcl::PointXYZI& p, size_t& index) const;
  void shiftCubes(int axis, int direction);

  Time _laserOdometryTime;
  float _scanPeriod;
  const int _stackFrameNum;
  const int _mapFrameNum;
  long _frameCount;
  long _mapFrameCount;
  size_t _maxIterations;
  float _deltaTAbort, _deltaRAbort;

  int _laserCloudCenWidth, _laserCloudCenHeight, _laserCloudCenDepth;
  const size_t _laserCloudWidth, _laserCloudHeight, _laserCloudDepth, _laserCloudNum;

  Cloud::Ptr _laserCloudCornerLast, _laserCloudSurfLast, _laserCloudFullRes;
  Cloud::Ptr _laserCloudCornerStack, _laserCloudSurfStack, _laserCloudCornerStackDS, _laserCloudSurfStackDS;
  Cloud::Ptr _laserCloudSurround, _laserCloudSurroundDS, _laserCloudCornerFromMap, _laserCloudSurfFromMap;

  std::vector<Cloud::Ptr> _laserCloudCornerArray, _laserCloudSurfArray, _laserCloudCornerDSArray, _laserCloudSurfDSArray;
  std::vector<size_t> _laserCloudValidInd, _laserCloudSurroundInd;

  Twist _transformSum, _transformIncre, _transformTobeMapped, _transformBefMapped, _transformAftMapped;
  std::vector<IMUState2> _imuHistory;

  pcl::VoxelGrid<pcl::PointXYZI> _downSizeFilterCorner, _downSizeFilterSurf, _downSizeFilterMap;
  bool _downsizedMapCreated = false;

  b200::Context* _gpu;
  b200::GaussNewtonSolver* _solver;
  std::vector<float> _bufA, _bufB;
  size_t _lastIterations = 0;
};

}  // namespace loam
