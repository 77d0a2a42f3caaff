// Host side of the scan-registration drop-in: buffers in, one GPU call, clouds out.
#include "loam_velodyne/BasicScanRegistration.h"

#include "b200_runtime.h"
#include "host_math.h"

#include <stdexcept>

namespace loam {

RegistrationParams::RegistrationParams(const float& scanPeriod_, const int& imuHistorySize_, const int& nFeatureRegions_,
                                       const int& curvatureRegion_, const int& maxCornerSharp_,
                                       const int& maxSurfaceFlat_, const float& lessFlatFilterSize_,
                                       const float& surfaceCurvatureThreshold_)
    : scanPeriod(scanPeriod_), imuHistorySize(imuHistorySize_), nFeatureRegions(nFeatureRegions_),
      curvatureRegion(curvatureRegion_), maxCornerSharp(maxCornerSharp_), maxCornerLessSharp(10 * maxCornerSharp_),
      maxSurfaceFlat(maxSurfaceFlat_), lessFlatFilterSize(lessFlatFilterSize_),
      surfaceCurvatureThreshold(surfaceCurvatureThreshold_) {}

void IMUState::interpolate(const IMUState& start, const IMUState& end, const float& ratio, IMUState& result) {
  const float inv = 1 - ratio;
  result.roll = start.roll.rad() * inv + end.roll.rad() * ratio;
  result.pitch = start.pitch.rad() * inv + end.pitch.rad() * ratio;
  if (start.yaw.rad() - end.yaw.rad() > M_PI)
    result.yaw = start.yaw.rad() * inv + (end.yaw.rad() + 2 * M_PI) * ratio;
  else if (start.yaw.rad() - end.yaw.rad() < -M_PI)
    result.yaw = start.yaw.rad() * inv + (end.yaw.rad() - 2 * M_PI) * ratio;
  else
    result.yaw = start.yaw.rad() * inv + end.yaw.rad() * ratio;
  result.velocity = start.velocity * inv + end.velocity * ratio;
  result.position = start.position * inv + end.position * ratio;
}

BasicScanRegistration::BasicScanRegistration() : _clouds(new b200::DualCloud[5]), _gpu(new b200::Context()) {
  static const int slots[5] = {LOAM_B200_C_REG_FULL, LOAM_B200_C_REG_SHARP, LOAM_B200_C_REG_LESS_SHARP,
                               LOAM_B200_C_REG_FLAT, LOAM_B200_C_REG_LESS_FLAT};
  for (int i = 0; i < 5; i++) _clouds[i].bind(_gpu, slots[i]);
  _gpu->setPriority(-1);  // never the bottleneck of the three stages
}
BasicScanRegistration::~BasicScanRegistration() {
  delete[] _clouds;
  delete _gpu;
}

bool BasicScanRegistration::configure(const RegistrationParams& config) {
  _config = config;
  return true;
}

void BasicScanRegistration::reset(const Time& scanTime) {
  _scanTime = scanTime;
  _imuIdx = 0;
  if (hasIMUData()) interpolateIMUStateFor(0, _imuStart);
  _sweepStart = scanTime;
  for (int i = 0; i < 5; i++) _clouds[i].clear();
  _scanIndices.clear();
  _indicesFetched = _labelsFetched = false;
  _sharpIdx.clear();
  _lessSharpIdx.clear();
  _flatIdx.clear();
  _labels.clear();
}

b200::DualCloud& BasicScanRegistration::deviceCloud(int which) { return _clouds[which]; }
pcl::PointCloud<pcl::PointXYZI> const& BasicScanRegistration::laserCloud() { return _clouds[0].host(); }
pcl::PointCloud<pcl::PointXYZI> const& BasicScanRegistration::cornerPointsSharp() { return _clouds[1].host(); }
pcl::PointCloud<pcl::PointXYZI> const& BasicScanRegistration::cornerPointsLessSharp() { return _clouds[2].host(); }
pcl::PointCloud<pcl::PointXYZI> const& BasicScanRegistration::surfacePointsFlat() { return _clouds[3].host(); }
pcl::PointCloud<pcl::PointXYZI> const& BasicScanRegistration::surfacePointsLessFlat() { return _clouds[4].host(); }

// ring ranges exactly as upstream builds _scanIndices (BasicScanRegistration.cpp:35-42): inclusive (first, last),
// an empty ring in the middle is (c, c - 1), an empty leading ring (0, 0)
static void appendRange(std::vector<IndexRange>& ranges, size_t& cloudSize, size_t ringSize) {
  IndexRange range(cloudSize, 0);
  cloudSize += ringSize;
  range.second = cloudSize > 0 ? cloudSize - 1 : 0;
  ranges.push_back(range);
}

void BasicScanRegistration::processScanlines(const Time& scanTime,
                                             std::vector<pcl::PointCloud<pcl::PointXYZI>> const& laserCloudScans) {
  reset(scanTime);
  size_t cloudSize = 0;
  for (size_t i = 0; i < laserCloudScans.size(); i++) appendRange(_scanIndices, cloudSize, laserCloudScans[i].size());
  // pack the rings straight into the upload buffer; the host-side concatenated cloud is rebuilt lazily by laserCloud()
  _packed.resize(cloudSize * 4 + 4);
  size_t o = 0;
  for (auto const& ring : laserCloudScans)
    for (auto const& p : ring.points) {
      _packed[o++] = p.x; _packed[o++] = p.y; _packed[o++] = p.z; _packed[o++] = p.intensity;
    }
  if (cloudSize > 0)
    _gpu->check(loam_b200_cloud_upload(_gpu->get(), LOAM_B200_C_REG_FULL, _packed.data(), (int)cloudSize), "loam_b200_cloud_upload");
  runExtraction((int)cloudSize);
  updateIMUTransform();
}

void BasicScanRegistration::processPackedSweep(const Time& scanTime, const float* xyzi, const int* ringSizes, int nRings) {
  reset(scanTime);
  size_t cloudSize = 0;
  for (int i = 0; i < nRings; i++) appendRange(_scanIndices, cloudSize, (size_t)ringSizes[i]);
  if (cloudSize > 0)
    _gpu->check(loam_b200_cloud_upload(_gpu->get(), LOAM_B200_C_REG_FULL, xyzi, (int)cloudSize), "loam_b200_cloud_upload");
  runExtraction((int)cloudSize);
  updateIMUTransform();
}

void BasicScanRegistration::processUnorderedSweep(const Time& scanTime, const float* xyz, int n, b200::RingLayout rings,
                                                  bool onDevice) {
  // upstream de-skews every point with projectPointToStartOfSweep(point, relTime) in arrival order before binning
  // (MultiScanRegistration.cpp:230); that sequential host interpolation is not part of the device front end, so a sweep
  // with IMU data must take the reference's own route (host ring binning + processScanlines) instead of being
  // silently left skewed
  if (hasIMUData())
    throw std::runtime_error("BasicScanRegistration::processUnorderedSweep: IMU data present -- the device front end "
                             "does not de-skew; bin on the host (MultiScanRegistration::process) and call processScanlines");
  reset(scanTime);
  const int nRings = (int)rings.nScanRings;
  std::vector<int32_t> ringSizes((size_t)nRings, 0);
  int kept = 0;
  _gpu->check(loam_b200_reg_bin(_gpu->get(), xyz, n, onDevice ? 1 : 0, rings.lowerBoundDeg, rings.upperBoundDeg, nRings,
                                _config.scanPeriod, ringSizes.data(), &kept),
              "loam_b200_reg_bin");
  size_t cloudSize = 0;
  for (int i = 0; i < nRings; i++) appendRange(_scanIndices, cloudSize, (size_t)ringSizes[i]);
  runExtraction((int)cloudSize);
  updateIMUTransform();
}

void BasicScanRegistration::processDeviceSweep(const Time& scanTime, const void* deviceXyzi, const int* ringSizes, int nRings) {
  reset(scanTime);
  size_t cloudSize = 0;
  for (int i = 0; i < nRings; i++) appendRange(_scanIndices, cloudSize, (size_t)ringSizes[i]);
  if (cloudSize > 0)
    _gpu->check(loam_b200_cloud_upload_device(_gpu->get(), LOAM_B200_C_REG_FULL, deviceXyzi, (int)cloudSize),
                "loam_b200_cloud_upload_device");
  runExtraction((int)cloudSize);
  updateIMUTransform();
}

void BasicScanRegistration::runExtraction(int n) {
  _lastN = n;
  const int nRings = (int)_scanIndices.size();
  if (nRings == 0 || n == 0) return;
  _clouds[0].deviceWritten(n);
  _ringStart.resize(nRings);
  _ringEnd.resize(nRings);
  for (int r = 0; r < nRings; r++) {
    _ringStart[r] = (int)_scanIndices[r].first;
    _ringEnd[r] = (int)_scanIndices[r].second;
  }
  loam_b200_reg_params prm;
  prm.nFeatureRegions = _config.nFeatureRegions;
  prm.curvatureRegion = _config.curvatureRegion;
  prm.maxCornerSharp = _config.maxCornerSharp;
  prm.maxCornerLessSharp = _config.maxCornerLessSharp;
  prm.maxSurfaceFlat = _config.maxSurfaceFlat;
  prm.lessFlatFilterSize = _config.lessFlatFilterSize;
  prm.surfaceCurvatureThreshold = _config.surfaceCurvatureThreshold;
  int counts[4] = {0, 0, 0, 0};
  _gpu->check(loam_b200_reg_run(_gpu->get(), _ringStart.data(), _ringEnd.data(), nRings, &prm, counts), "loam_b200_reg_run");
  for (int k = 0; k < 4; k++) _clouds[1 + k].deviceWritten(counts[k]);
}

void BasicScanRegistration::fetchIndices() {
  if (_indicesFetched) return;
  _indicesFetched = true;
  if (_lastN == 0 || !_gpu->created()) return;
  std::vector<int>* dst[3] = {&_sharpIdx, &_lessSharpIdx, &_flatIdx};
  for (int w = 1; w <= 3; w++) {
    const int n = (int)_clouds[w].size();
    dst[w - 1]->resize(n + 1);
    int got = 0;
    _gpu->check(loam_b200_reg_indices(_gpu->get(), w, dst[w - 1]->data(), n, &got), "loam_b200_reg_indices");
    dst[w - 1]->resize(got);
  }
}
std::vector<int> const& BasicScanRegistration::sharpIndices() { fetchIndices(); return _sharpIdx; }
std::vector<int> const& BasicScanRegistration::lessSharpIndices() { fetchIndices(); return _lessSharpIdx; }
std::vector<int> const& BasicScanRegistration::flatIndices() { fetchIndices(); return _flatIdx; }
std::vector<signed char> const& BasicScanRegistration::pointLabels() {
  if (!_labelsFetched) {
    _labelsFetched = true;
    _labels.resize(_lastN);
    if (_lastN > 0 && _gpu->created())
      _gpu->check(loam_b200_reg_labels(_gpu->get(), reinterpret_cast<int8_t*>(_labels.data()), _lastN), "loam_b200_reg_labels");
  }
  return _labels;
}

// ---- IMU plumbing (host scalar code; upstream BasicScanRegistration.cpp:82-152,258-281).  The hot path's configs
// carry no IMU, in which case everything below degenerates to zeros exactly like upstream.
void BasicScanRegistration::updateIMUData(Vector3& acc, IMUState& newState) {
  if (!_imuHistory.empty()) {
    hostmath::rotateZXY(acc, newState.roll, newState.pitch, newState.yaw);
    const IMUState& prev = _imuHistory.back();
    const float dt = (float)toSec(newState.stamp - prev.stamp);
    newState.position = prev.position + (prev.velocity * dt) + (0.5f * acc * dt * dt);
    newState.velocity = prev.velocity + acc * dt;
  }
  _imuHistory.push_back(newState);
  const size_t cap = _config.imuHistorySize > 0 ? (size_t)_config.imuHistorySize : 200;
  if (_imuHistory.size() > cap) _imuHistory.erase(_imuHistory.begin());
}

void BasicScanRegistration::projectPointToStartOfSweep(pcl::PointXYZI& point, float relTime) {
  if (hasIMUData()) {
    setIMUTransformFor(relTime);
    transformToStartIMU(point);
  }
}

void BasicScanRegistration::setIMUTransformFor(const float& relTime) {
  interpolateIMUStateFor(relTime, _imuCur);
  const float relSweepTime = (float)(toSec(_scanTime - _sweepStart) + relTime);
  _imuPositionShift = _imuCur.position - _imuStart.position - _imuStart.velocity * relSweepTime;
}

void BasicScanRegistration::transformToStartIMU(pcl::PointXYZI& point) {
  hostmath::rotateZXY(point, _imuCur.roll, _imuCur.pitch, _imuCur.yaw);
  point.x += _imuPositionShift.x();
  point.y += _imuPositionShift.y();
  point.z += _imuPositionShift.z();
  hostmath::rotateYXZ(point, -_imuStart.yaw, -_imuStart.pitch, -_imuStart.roll);
}

void BasicScanRegistration::interpolateIMUStateFor(const float& relTime, IMUState& outputState) {
  double timeDiff = toSec(_scanTime - _imuHistory[_imuIdx].stamp) + relTime;
  while (_imuIdx < _imuHistory.size() - 1 && timeDiff > 0) {
    _imuIdx++;
    timeDiff = toSec(_scanTime - _imuHistory[_imuIdx].stamp) + relTime;
  }
  if (_imuIdx == 0 || timeDiff > 0) {
    outputState = _imuHistory[_imuIdx];
  } else {
    const float ratio = (float)(-timeDiff / toSec(_imuHistory[_imuIdx].stamp - _imuHistory[_imuIdx - 1].stamp));
    IMUState::interpolate(_imuHistory[_imuIdx], _imuHistory[_imuIdx - 1], ratio, outputState);
  }
}

void BasicScanRegistration::updateIMUTransform() {
  _imuTrans[0].x = _imuStart.pitch.rad();
  _imuTrans[0].y = _imuStart.yaw.rad();
  _imuTrans[0].z = _imuStart.roll.rad();
  _imuTrans[1].x = _imuCur.pitch.rad();
  _imuTrans[1].y = _imuCur.yaw.rad();
  _imuTrans[1].z = _imuCur.roll.rad();
  Vector3 shift = _imuPositionShift;
  hostmath::rotateYXZ(shift, -_imuStart.yaw, -_imuStart.pitch, -_imuStart.roll);
  _imuTrans[2].x = shift.x();
  _imuTrans[2].y = shift.y();
  _imuTrans[2].z = shift.z();
  Vector3 velo = _imuCur.velocity - _imuStart.velocity;
  hostmath::rotateYXZ(velo, -_imuStart.yaw, -_imuStart.pitch, -_imuStart.roll);
  _imuTrans[3].x = velo.x();
  _imuTrans[3].y = velo.y();
  _imuTrans[3].z = velo.z();
}

}  // namespace loam
