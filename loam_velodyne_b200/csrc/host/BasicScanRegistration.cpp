// Host side of the scan-registration drop-in: buffers in, one GPU call, clouds out.
#include "loam_velodyne/BasicScanRegistration.h"

#include "b200_runtime.h"
#include "host_math.h"

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

BasicScanRegistration::BasicScanRegistration() : _gpu(new b200::Context()) {}
BasicScanRegistration::~BasicScanRegistration() { delete _gpu; }

bool BasicScanRegistration::configure(const RegistrationParams& config) {
  _config = config;
  return true;
}

void BasicScanRegistration::reset(const Time& scanTime) {
  _scanTime = scanTime;
  _imuIdx = 0;
  if (hasIMUData()) interpolateIMUStateFor(0, _imuStart);
  _sweepStart = scanTime;
  _laserCloud.clear();
  _cornerPointsSharp.clear();
  _cornerPointsLessSharp.clear();
  _surfacePointsFlat.clear();
  _surfacePointsLessFlat.clear();
  _scanIndices.clear();
}

void BasicScanRegistration::processScanlines(const Time& scanTime,
                                             std::vector<pcl::PointCloud<pcl::PointXYZI>> const& laserCloudScans) {
  reset(scanTime);

  // ring-ordered full-resolution cloud + inclusive index range per ring (upstream BasicScanRegistration.cpp:35-42)
  size_t cloudSize = 0;
  for (size_t i = 0; i < laserCloudScans.size(); i++) {
    _laserCloud += laserCloudScans[i];
    IndexRange range(cloudSize, 0);
    cloudSize += laserCloudScans[i].size();
    range.second = cloudSize > 0 ? cloudSize - 1 : 0;
    _scanIndices.push_back(range);
  }

  const int n = (int)_laserCloud.size();
  const int nRings = (int)_scanIndices.size();
  if (nRings > 0) {
    b200::pack(_laserCloud, _packed);
    _ringStart.resize(nRings);
    _ringEnd.resize(nRings);
    for (int r = 0; r < nRings; r++) {
      _ringStart[r] = (int)_scanIndices[r].first;
      // an empty ring in the middle has second = first - 1 (size_t arithmetic upstream never underflows because
      // cloudSize > 0 there); an empty leading ring is (0, 0)
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

    const size_t capSharp = (size_t)nRings * prm.nFeatureRegions * prm.maxCornerSharp;
    const size_t capLess = (size_t)nRings * prm.nFeatureRegions * prm.maxCornerLessSharp;
    const size_t capFlat = (size_t)nRings * prm.nFeatureRegions * prm.maxSurfaceFlat;
    _sharpIdx.resize(capSharp + 1);
    _lessSharpIdx.resize(capLess + 1);
    _flatIdx.resize(capFlat + 1);
    _labels.resize((size_t)n + 1);
    _lessFlatDS.resize((size_t)n * 4 + 4);
    loam_b200_features out;
    out.sharp_idx = _sharpIdx.data();           out.sharp_cap = (int)capSharp;
    out.less_sharp_idx = _lessSharpIdx.data();  out.less_sharp_cap = (int)capLess;
    out.flat_idx = _flatIdx.data();             out.flat_cap = (int)capFlat;
    out.label = reinterpret_cast<int8_t*>(_labels.data());
    out.less_flat_ds = _lessFlatDS.data();      out.less_flat_cap = n;
    out.n_sharp = out.n_less_sharp = out.n_flat = out.n_less_flat = 0;
    _gpu->check(loam_b200_extract_features(_gpu->get(), _packed.data(), n, _ringStart.data(), _ringEnd.data(), nRings,
                                           &prm, &out),
                "loam_b200_extract_features");
    _sharpIdx.resize(out.n_sharp);
    _lessSharpIdx.resize(out.n_less_sharp);
    _flatIdx.resize(out.n_flat);
    _labels.resize(n);
    for (int idx : _sharpIdx) _cornerPointsSharp.push_back(_laserCloud[idx]);
    for (int idx : _lessSharpIdx) _cornerPointsLessSharp.push_back(_laserCloud[idx]);
    for (int idx : _flatIdx) _surfacePointsFlat.push_back(_laserCloud[idx]);
    b200::unpack(_lessFlatDS.data(), (size_t)out.n_less_flat, _surfacePointsLessFlat);
  }
  updateIMUTransform();
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
