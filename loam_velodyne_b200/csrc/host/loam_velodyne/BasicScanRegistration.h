// loam::BasicScanRegistration -- drop-in for upstream include/loam_velodyne/BasicScanRegistration.h:135-164.
// Same public surface (processScanlines, configure, updateIMUData, projectPointToStartOfSweep and the const-ref
// accessors) so `class ScanRegistration : protected BasicScanRegistration` (upstream ScanRegistration.h:55) keeps
// compiling; the feature extraction itself (extractFeatures / setScanBuffersFor / setRegionBuffersFor /
// markAsPicked, BasicScanRegistration.cpp:155-254,284-386) runs on the GPU through loam_b200_extract_features.
#pragma once
#include <cmath>
#include <cstdint>

#include <utility>
#include <vector>

#include <pcl/point_cloud.h>

#include "Angle.h"
#include "Vector3.h"
#include "time_utils.h"

namespace loam {

namespace b200 { class Context; class DualCloud; }

typedef std::pair<size_t, size_t> IndexRange;

enum PointLabel {
  CORNER_SHARP = 2,
  CORNER_LESS_SHARP = 1,
  SURFACE_LESS_FLAT = 0,
  SURFACE_FLAT = -1
};

class RegistrationParams {
 public:
  RegistrationParams(const float& scanPeriod_ = 0.1, const int& imuHistorySize_ = 200, const int& nFeatureRegions_ = 6,
                     const int& curvatureRegion_ = 5, const int& maxCornerSharp_ = 2, const int& maxSurfaceFlat_ = 4,
                     const float& lessFlatFilterSize_ = 0.2, const float& surfaceCurvatureThreshold_ = 0.1);
  float scanPeriod;
  int imuHistorySize;
  int nFeatureRegions;
  int curvatureRegion;
  int maxCornerSharp;
  int maxCornerLessSharp;
  int maxSurfaceFlat;
  float lessFlatFilterSize;
  float surfaceCurvatureThreshold;
};

typedef struct IMUState {
  Time stamp;
  Angle roll;
  Angle pitch;
  Angle yaw;
  Vector3 position;
  Vector3 velocity;
  Vector3 acceleration;
  static void interpolate(const IMUState& start, const IMUState& end, const float& ratio, IMUState& result);
} IMUState;

namespace b200 {
/** Ring layout of a multi-beam sensor for processUnorderedSweep(): plain data with the same three numbers as the
 *  reference's loam::MultiScanMapper (MultiScanRegistration.h:49-103).  That class itself stays where upstream declares
 *  it (the ROS-bound MultiScanRegistration.h, which compiles unchanged against this header): declaring it here as well
 *  made upstream's MultiScanRegistration.cpp fail with a redefinition (round-1 review). */
struct RingLayout {
  float lowerBoundDeg, upperBoundDeg;
  uint16_t nScanRings;
  RingLayout(float lower = -15.f, float upper = 15.f, uint16_t rings = 16)
      : lowerBoundDeg(lower), upperBoundDeg(upper), nScanRings(rings) {}
  static RingLayout Velodyne_VLP_16() { return RingLayout(-15.f, 15.f, 16); }
  static RingLayout Velodyne_HDL_32() { return RingLayout(-30.67f, 10.67f, 32); }
  static RingLayout Velodyne_HDL_64E() { return RingLayout(-24.9f, 2.f, 64); }
};
}  // namespace b200

class BasicScanRegistration {
 public:
  BasicScanRegistration();
  ~BasicScanRegistration();
  BasicScanRegistration(const BasicScanRegistration&) = delete;
  BasicScanRegistration& operator=(const BasicScanRegistration&) = delete;

  void processScanlines(const Time& scanTime, std::vector<pcl::PointCloud<pcl::PointXYZI>> const& laserCloudScans);
  bool configure(const RegistrationParams& config = RegistrationParams());
  void updateIMUData(Vector3& acc, IMUState& newState);
  void projectPointToStartOfSweep(pcl::PointXYZI& point, float relTime);

  auto const& imuTransform() { return _imuTrans; }
  auto const& sweepStart() { return _sweepStart; }
  // The clouds live in HBM after processScanlines; these accessors download them on first use (then cache).
  pcl::PointCloud<pcl::PointXYZI> const& laserCloud();
  pcl::PointCloud<pcl::PointXYZI> const& cornerPointsSharp();
  pcl::PointCloud<pcl::PointXYZI> const& cornerPointsLessSharp();
  pcl::PointCloud<pcl::PointXYZI> const& surfacePointsFlat();
  pcl::PointCloud<pcl::PointXYZI> const& surfacePointsLessFlat();
  auto const& config() { return _config; }

  // ---- extensions (not part of the reference API) ----
  // same as processScanlines for a sweep that is already packed: n_rings rings of ring_sizes[r] consecutive
  // (x, y, z, intensity) float quadruples; skips the per-ring pcl clouds
  void processPackedSweep(const Time& scanTime, const float* xyzi, const int* ringSizes, int nRings);
  // ... or already resident in GPU memory (device pointer to the packed points)
  /** Extension: the ring-binning front end of MultiScanRegistration::process (MultiScanRegistration.cpp:160-238) on the
   *  GPU, followed by the regular extraction: `xyz` = n unordered sensor-frame points (3 floats each, arrival order;
   *  host memory, or device memory when `onDevice`).  Equivalent to building the per-ring clouds on the host and
   *  calling processScanlines().  Sweeps with IMU data need the per-point de-skew of projectPointToStartOfSweep in
   *  arrival order (host scalar code upstream): this entry point then throws std::runtime_error, use processScanlines. */
  void processUnorderedSweep(const Time& scanTime, const float* xyz, int n, b200::RingLayout rings, bool onDevice = false);
  /** Same, taking upstream's loam::MultiScanMapper (or anything with its three getters) directly: the one-line change in
   *  MultiScanRegistration::process is `processUnorderedSweep(scanTime, xyz, n, _scanMapper)`. */
  template <typename Mapper>
  void processUnorderedSweep(const Time& scanTime, const float* xyz, int n, Mapper& mapper, bool onDevice = false) {
    processUnorderedSweep(scanTime, xyz, n,
                          b200::RingLayout(mapper.getLowerBound(), mapper.getUpperBound(), mapper.getNumberOfScanRings()), onDevice);
  }
  void processDeviceSweep(const Time& scanTime, const void* deviceXyzi, const int* ringSizes, int nRings);
  // indices (into laserCloud()) of the picked features and the per-point labels of the last sweep
  std::vector<int> const& sharpIndices();
  std::vector<int> const& lessSharpIndices();
  std::vector<int> const& flatIndices();
  std::vector<signed char> const& pointLabels();
  // device-side hand-off to BasicLaserOdometry::adopt
  b200::Context* deviceContext() { return _gpu; }
  b200::DualCloud& deviceCloud(int which);  // 0 full, 1 sharp, 2 less sharp, 3 flat, 4 less flat

 private:
  bool hasIMUData() const { return !_imuHistory.empty(); }
  void reset(const Time& scanTime);
  void interpolateIMUStateFor(const float& relTime, IMUState& outputState);
  void setIMUTransformFor(const float& relTime);
  void transformToStartIMU(pcl::PointXYZI& point);
  void updateIMUTransform();

  void runExtraction(int n);
  void fetchIndices();

  RegistrationParams _config;
  std::vector<IndexRange> _scanIndices;
  b200::DualCloud* _clouds;  // [5]: full, sharp, less sharp, flat, less flat

  Time _sweepStart, _scanTime;
  IMUState _imuStart, _imuCur;
  Vector3 _imuPositionShift;
  size_t _imuIdx = 0;
  std::vector<IMUState> _imuHistory;  // bounded FIFO of imuHistorySize states
  pcl::PointCloud<pcl::PointXYZ> _imuTrans = {4, 1};

  b200::Context* _gpu;
  std::vector<float> _packed;
  std::vector<int> _ringStart, _ringEnd, _sharpIdx, _lessSharpIdx, _flatIdx;
  std::vector<signed char> _labels;
  bool _indicesFetched = false, _labelsFetched = false;
  int _lastN = 0;
};

}  // namespace loam
