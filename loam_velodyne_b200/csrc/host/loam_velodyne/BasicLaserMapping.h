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

namespace b200 { class Context; class DualCloud; struct GaussNewtonSolver; }
class BasicLaserOdometry;

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

  // input clouds (filled by the caller, LaserMapping.cpp:177-200 upstream) / registered full-resolution output;
  // the authoritative copies live in HBM, these handles download on first use and mark the GPU copy stale
  pcl::PointCloud<pcl::PointXYZI>& laserCloud();
  pcl::PointCloud<pcl::PointXYZI>& laserCloudCornerLast();
  pcl::PointCloud<pcl::PointXYZI>& laserCloudSurfLast();

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
  pcl::PointCloud<pcl::PointXYZI> const& laserCloudSurroundDS() const;

  bool hasFreshMap() const { return _downsizedMapCreated; }

  // extensions (not in the reference API): pre-seed the cube map with points given in the map frame, using the same
  // cube-index arithmetic as the insertion step (upstream BasicLaserMapping.cpp:540-553); introspection for tests.
  void seedMap(pcl::PointCloud<pcl::PointXYZI> const& cornerPoints, pcl::PointCloud<pcl::PointXYZI> const& surfPoints);
  size_t lastIterationCount() const { return _lastIterations; }
  // sizes of the down-sized feature stacks of the last process() (the queries of its optimisation loop)
  size_t cornerStackSize() const { return (size_t)_mapSizes[2]; }
  size_t surfStackSize() const { return (size_t)_mapSizes[3]; }
  // host wall seconds of the last process(): begin_sweep, LM loop, end_sweep, surround map
  const double* lastPhaseSeconds() const { return _phase; }
  auto const& transformTobeMapped() const { return _transformTobeMapped; }
  pcl::PointCloud<pcl::PointXYZI> const& cornerStackDS() const;
  pcl::PointCloud<pcl::PointXYZI> const& surfStackDS() const;
  pcl::PointCloud<pcl::PointXYZI> const& cornerFromMap() const;
  pcl::PointCloud<pcl::PointXYZI> const& surfFromMap() const;
  void collectMap(pcl::PointCloud<pcl::PointXYZI>& corner, pcl::PointCloud<pcl::PointXYZI>& surf) const;
  // take last corner / last surface / full-resolution clouds and transformSum from an odometry object without a
  // host round trip (LaserOdometry::publishResult -> LaserMapping::*Handler upstream)
  void adopt(BasicLaserOdometry& odom);
  b200::Context* deviceContext() { return _gpu; }
  // multi-GPU (one process per GPU): this object evaluates the rank-th of `world` slices of the correspondences and
  // all-reduces the 6x6 normal equations over NCCL every iteration (ncclId: 128 bytes from
  // loam_b200_comm_unique_id on rank 0; nullptr = slice only, the caller reduces)
  void enableSharding(int rank, int world, const unsigned char* ncclId);
  // multi-GPU with the MAP sharded by cube slabs (one process per GPU; csrc/shard.cuh, include/loam_b200.h "peer"):
  // exportPeerHandle() gives this rank's 64-byte inbox handle, enableCubeSharding() takes all of them in rank order.
  // Afterwards this object stores only the slabs it owns (+ 2 m halo): seedMap() keeps the points that belong here,
  // process() evaluates the queries that fall into its slabs and the per-iteration normal equations are all-reduced
  // inside the iteration kernel over NVLink peer memory.  slabMetres: slab width (whole metres, default 10).
  void exportPeerHandle(unsigned char out64[64]);
  void enableCubeSharding(int rank, int world, const unsigned char* handles, int slabMetres = 10);
  // the same for several objects of ONE process (tests / single-process multi-GPU): objs[r] becomes rank r
  static void enableCubeShardingLocal(BasicLaserMapping** objs, int world, int slabMetres = 10);
  /** Test hook (not in the reference): keep a copy of the points of the cubes in view (upstream's internal
   *  _laserCloudCornerFromMap / _laserCloudSurfFromMap) after every process(); off by default, the persistent GPU map
   *  only needs their sizes. */
  void retainFromMapClouds(bool on);

 private:
  typedef pcl::PointCloud<pcl::PointXYZI> Cloud;
  void optimizeTransformTobeMapped();
  void transformAssociateToMap();
  void transformUpdate();
  void pointAssociateToMap(const pcl::PointXYZI& pi, pcl::PointXYZI& po);
  bool createDownsizedMap();

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

  // clouds in HBM slots: corner last, surf last, full res, corner / surf stack DS, corner / surf from map, surround DS;
  // the 21 x 11 x 21 cube arrays of the reference are two flat point pools on the GPU (see csrc/mappool.cuh)
  b200::DualCloud* _c;
  std::vector<size_t> _laserCloudValidInd, _laserCloudSurroundInd;
  int _mapSizes[4] = {0, 0, 0, 0};

  Twist _transformSum, _transformIncre, _transformTobeMapped, _transformBefMapped, _transformAftMapped;
  std::vector<IMUState2> _imuHistory;

  pcl::VoxelGrid<pcl::PointXYZI> _downSizeFilterCorner, _downSizeFilterSurf, _downSizeFilterMap;
  bool _downsizedMapCreated = false;
  bool _retainFromMap = false;
  bool _sharded = false;
  int _shardRank = 0, _shardWorld = 1, _shardSlab = 0;  // cube-sharded map (0 = off)

  b200::Context* _gpu;
  b200::GaussNewtonSolver* _solver;
  std::vector<float> _bufA, _bufB;
  size_t _lastIterations = 0;
  double _phase[4] = {0, 0, 0, 0};
};

}  // namespace loam
