// Host side of the laser-mapping drop-in.  Control flow mirrors upstream BasicLaserMapping::process /
// optimizeTransformTobeMapped (src/lib/BasicLaserMapping.cpp:266-599, 626-926): pose prediction, the rolling
// 21 x 11 x 21 grid of 50 m cubes, field-of-view cube selection and map bookkeeping stay here; voxel filters, tree
// builds, the per-iteration correspondence / Jacobian / normal-equation work and the bulk point transforms are GPU
// calls through include/loam_b200.h.
#include "loam_velodyne/BasicLaserMapping.h"

#include <cmath>

#include "b200_runtime.h"
#include "host_math.h"

namespace loam {

using hostmath::rad2deg;

BasicLaserMapping::BasicLaserMapping(const float& scanPeriod, const size_t& maxIterations)
    : _scanPeriod(scanPeriod), _stackFrameNum(1), _mapFrameNum(5), _frameCount(0), _mapFrameCount(0),
      _maxIterations(maxIterations), _deltaTAbort(0.05), _deltaRAbort(0.05), _laserCloudCenWidth(10),
      _laserCloudCenHeight(5), _laserCloudCenDepth(10), _laserCloudWidth(21), _laserCloudHeight(11),
      _laserCloudDepth(21), _laserCloudNum(_laserCloudWidth * _laserCloudHeight * _laserCloudDepth),
      _laserCloudCornerLast(new Cloud()), _laserCloudSurfLast(new Cloud()), _laserCloudFullRes(new Cloud()),
      _laserCloudCornerStack(new Cloud()), _laserCloudSurfStack(new Cloud()), _laserCloudCornerStackDS(new Cloud()),
      _laserCloudSurfStackDS(new Cloud()), _laserCloudSurround(new Cloud()), _laserCloudSurroundDS(new Cloud()),
      _laserCloudCornerFromMap(new Cloud()), _laserCloudSurfFromMap(new Cloud()), _gpu(new b200::Context()),
      _solver(new b200::GaussNewtonSolver()) {
  _frameCount = _stackFrameNum - 1;
  _mapFrameCount = _mapFrameNum - 1;
  _laserCloudCornerArray.resize(_laserCloudNum);
  _laserCloudSurfArray.resize(_laserCloudNum);
  _laserCloudCornerDSArray.resize(_laserCloudNum);
  _laserCloudSurfDSArray.resize(_laserCloudNum);
  for (size_t i = 0; i < _laserCloudNum; i++) {
    _laserCloudCornerArray[i].reset(new Cloud());
    _laserCloudSurfArray[i].reset(new Cloud());
    _laserCloudCornerDSArray[i].reset(new Cloud());
    _laserCloudSurfDSArray[i].reset(new Cloud());
  }
  _downSizeFilterCorner.setLeafSize(0.2, 0.2, 0.2);
  _downSizeFilterSurf.setLeafSize(0.4, 0.4, 0.4);
}

BasicLaserMapping::~BasicLaserMapping() {
  delete _solver;
  delete _gpu;
}

// Pose prediction: compose (Sum, BefMapped, AftMapped) into TobeMapped, closed-form ZXY Euler algebra as published
// with LOAM (upstream :103-167).
void BasicLaserMapping::transformAssociateToMap() {
  _transformIncre.pos = _transformBefMapped.pos - _transformSum.pos;
  hostmath::rotateYXZ(_transformIncre.pos, -(_transformSum.rot_y), -(_transformSum.rot_x), -(_transformSum.rot_z));

  const float sbcx = _transformSum.rot_x.sin(), cbcx = _transformSum.rot_x.cos();
  const float sbcy = _transformSum.rot_y.sin(), cbcy = _transformSum.rot_y.cos();
  const float sbcz = _transformSum.rot_z.sin(), cbcz = _transformSum.rot_z.cos();
  const float sblx = _transformBefMapped.rot_x.sin(), cblx = _transformBefMapped.rot_x.cos();
  const float sbly = _transformBefMapped.rot_y.sin(), cbly = _transformBefMapped.rot_y.cos();
  const float sblz = _transformBefMapped.rot_z.sin(), cblz = _transformBefMapped.rot_z.cos();
  const float salx = _transformAftMapped.rot_x.sin(), calx = _transformAftMapped.rot_x.cos();
  const float saly = _transformAftMapped.rot_y.sin(), caly = _transformAftMapped.rot_y.cos();
  const float salz = _transformAftMapped.rot_z.sin(), calz = _transformAftMapped.rot_z.cos();

  const float srx = -sbcx * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz) -
                    cbcx * sbcy * (calx * calz * (cbly * sblz - cblz * sblx * sbly) -
                                   calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                    cbcx * cbcy * (calx * salz * (cblz * sbly - cbly * sblx * sblz) -
                                   calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx);
  _transformTobeMapped.rot_x = -std::asin(srx);

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
  _transformTobeMapped.rot_y = std::atan2(srycrx / _transformTobeMapped.rot_x.cos(), crycrx / _transformTobeMapped.rot_x.cos());

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
  _transformTobeMapped.rot_z = std::atan2(srzcrx / _transformTobeMapped.rot_x.cos(), crzcrx / _transformTobeMapped.rot_x.cos());

  Vector3 v = _transformIncre.pos;
  hostmath::rotateZXY(v, _transformTobeMapped.rot_z, _transformTobeMapped.rot_x, _transformTobeMapped.rot_y);
  _transformTobeMapped.pos = _transformAftMapped.pos - v;
}

void BasicLaserMapping::transformUpdate() {
  if (!_imuHistory.empty()) {
    size_t imuIdx = 0;
    while (imuIdx < _imuHistory.size() - 1 && toSec(_laserOdometryTime - _imuHistory[imuIdx].stamp) + _scanPeriod > 0)
      imuIdx++;
    IMUState2 imuCur;
    if (imuIdx == 0 || toSec(_laserOdometryTime - _imuHistory[imuIdx].stamp) + _scanPeriod > 0) {
      imuCur = _imuHistory[imuIdx];
    } else {
      const float ratio = (float)((toSec(_imuHistory[imuIdx].stamp - _laserOdometryTime) - _scanPeriod) /
                                  toSec(_imuHistory[imuIdx].stamp - _imuHistory[imuIdx - 1].stamp));
      IMUState2::interpolate(_imuHistory[imuIdx], _imuHistory[imuIdx - 1], ratio, imuCur);
    }
    _transformTobeMapped.rot_x = 0.998 * _transformTobeMapped.rot_x.rad() + 0.002 * imuCur.pitch.rad();
    _transformTobeMapped.rot_z = 0.998 * _transformTobeMapped.rot_z.rad() + 0.002 * imuCur.roll.rad();
  }
  _transformBefMapped = _transformSum;
  _transformAftMapped = _transformTobeMapped;
}

void BasicLaserMapping::pointAssociateToMap(const pcl::PointXYZI& pi, pcl::PointXYZI& po) {
  po.x = pi.x;
  po.y = pi.y;
  po.z = pi.z;
  po.intensity = pi.intensity;
  hostmath::rotateZXY(po, _transformTobeMapped.rot_z, _transformTobeMapped.rot_x, _transformTobeMapped.rot_y);
  po.x += _transformTobeMapped.pos.x();
  po.y += _transformTobeMapped.pos.y();
  po.z += _transformTobeMapped.pos.z();
}

void BasicLaserMapping::pointAssociateTobeMapped(const pcl::PointXYZI& pi, pcl::PointXYZI& po) {
  po.x = pi.x - _transformTobeMapped.pos.x();
  po.y = pi.y - _transformTobeMapped.pos.y();
  po.z = pi.z - _transformTobeMapped.pos.z();
  po.intensity = pi.intensity;
  hostmath::rotateYXZ(po, -_transformTobeMapped.rot_y, -_transformTobeMapped.rot_x, -_transformTobeMapped.rot_z);
}

void BasicLaserMapping::transformFullResToMap() {
  const size_t n = _laserCloudFullRes->size();
  if (n == 0) return;
  b200::pack(*_laserCloudFullRes, _bufA);
  loam_b200_pose p;
  b200::fillPose(_transformTobeMapped, p);
  _gpu->check(loam_b200_transform_to_map(_gpu->get(), _bufA.data(), (int)n, &p), "loam_b200_transform_to_map");
  b200::unpack(_bufA.data(), n, *_laserCloudFullRes);
}

bool BasicLaserMapping::createDownsizedMap() {
  _mapFrameCount++;
  if (_mapFrameCount < _mapFrameNum) return false;
  _mapFrameCount = 0;
  _laserCloudSurround->clear();
  for (auto ind : _laserCloudSurroundInd) {
    *_laserCloudSurround += *_laserCloudCornerArray[ind];
    *_laserCloudSurround += *_laserCloudSurfArray[ind];
  }
  // upstream filters the surround map with the CORNER filter (:261-262); _downSizeFilterMap is never used
  b200::voxelFilter(*_gpu, *_laserCloudSurround, b200::leafOf(_downSizeFilterCorner), *_laserCloudSurroundDS, _bufA, _bufB);
  return true;
}

// cube index of a map-frame point: truncation toward zero plus the negative-side correction (upstream :540-553)
bool BasicLaserMapping::cubeIndexOf(const pcl::PointXYZI& p, size_t& index) const {
  const double CUBE_SIZE = 50.0, CUBE_HALF = CUBE_SIZE / 2;
  int cubeI = int((p.x + CUBE_HALF) / CUBE_SIZE) + _laserCloudCenWidth;
  int cubeJ = int((p.y + CUBE_HALF) / CUBE_SIZE) + _laserCloudCenHeight;
  int cubeK = int((p.z + CUBE_HALF) / CUBE_SIZE) + _laserCloudCenDepth;
  if (p.x + CUBE_HALF < 0) cubeI--;
  if (p.y + CUBE_HALF < 0) cubeJ--;
  if (p.z + CUBE_HALF < 0) cubeK--;
  if (cubeI >= 0 && cubeI < (int)_laserCloudWidth && cubeJ >= 0 && cubeJ < (int)_laserCloudHeight && cubeK >= 0 &&
      cubeK < (int)_laserCloudDepth) {
    index = cubeI + _laserCloudWidth * cubeJ + _laserCloudWidth * _laserCloudHeight * cubeK;
    return true;
  }
  return false;
}

// Roll the cube grid by one cell along `axis` (0 = width, 1 = height, 2 = depth).  direction +1 moves every cube
// to the next higher index and empties the lowest slab (the map centre index grows); -1 the opposite
// (upstream :311-441 spells the six cases out).
void BasicLaserMapping::shiftCubes(int axis, int direction) {
  const int dims[3] = {(int)_laserCloudWidth, (int)_laserCloudHeight, (int)_laserCloudDepth};
  const int n = dims[axis];
  const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
  int ijk[3];
  for (ijk[a1] = 0; ijk[a1] < dims[a1]; ijk[a1]++) {
    for (ijk[a2] = 0; ijk[a2] < dims[a2]; ijk[a2]++) {
      if (direction > 0) {
        for (int t = n - 1; t >= 1; t--) {
          ijk[axis] = t;
          const size_t a = toIndex(ijk[0], ijk[1], ijk[2]);
          ijk[axis] = t - 1;
          const size_t b = toIndex(ijk[0], ijk[1], ijk[2]);
          std::swap(_laserCloudCornerArray[a], _laserCloudCornerArray[b]);
          std::swap(_laserCloudSurfArray[a], _laserCloudSurfArray[b]);
        }
        ijk[axis] = 0;
      } else {
        for (int t = 0; t < n - 1; t++) {
          ijk[axis] = t;
          const size_t a = toIndex(ijk[0], ijk[1], ijk[2]);
          ijk[axis] = t + 1;
          const size_t b = toIndex(ijk[0], ijk[1], ijk[2]);
          std::swap(_laserCloudCornerArray[a], _laserCloudCornerArray[b]);
          std::swap(_laserCloudSurfArray[a], _laserCloudSurfArray[b]);
        }
        ijk[axis] = n - 1;
      }
      const size_t c = toIndex(ijk[0], ijk[1], ijk[2]);
      _laserCloudCornerArray[c]->clear();
      _laserCloudSurfArray[c]->clear();
    }
  }
}

bool BasicLaserMapping::process(Time const& laserOdometryTime) {
  _frameCount++;
  if (_frameCount < _stackFrameNum) return false;
  _frameCount = 0;
  _laserOdometryTime = laserOdometryTime;

  pcl::PointXYZI pointSel;
  transformAssociateToMap();

  for (auto const& pt : _laserCloudCornerLast->points) {
    pointAssociateToMap(pt, pointSel);
    _laserCloudCornerStack->push_back(pointSel);
  }
  for (auto const& pt : _laserCloudSurfLast->points) {
    pointAssociateToMap(pt, pointSel);
    _laserCloudSurfStack->push_back(pointSel);
  }

  pcl::PointXYZI pointOnYAxis;
  pointOnYAxis.x = 0.0;
  pointOnYAxis.y = 10.0;
  pointOnYAxis.z = 0.0;
  pointAssociateToMap(pointOnYAxis, pointOnYAxis);

  const double CUBE_SIZE = 50.0, CUBE_HALF = CUBE_SIZE / 2;
  int centerCubeI = int((_transformTobeMapped.pos.x() + CUBE_HALF) / CUBE_SIZE) + _laserCloudCenWidth;
  int centerCubeJ = int((_transformTobeMapped.pos.y() + CUBE_HALF) / CUBE_SIZE) + _laserCloudCenHeight;
  int centerCubeK = int((_transformTobeMapped.pos.z() + CUBE_HALF) / CUBE_SIZE) + _laserCloudCenDepth;
  if (_transformTobeMapped.pos.x() + CUBE_HALF < 0) centerCubeI--;
  if (_transformTobeMapped.pos.y() + CUBE_HALF < 0) centerCubeJ--;
  if (_transformTobeMapped.pos.z() + CUBE_HALF < 0) centerCubeK--;

  // keep the sensor at least 3 cubes away from every face of the grid
  while (centerCubeI < 3) { shiftCubes(0, +1); centerCubeI++; _laserCloudCenWidth++; }
  while (centerCubeI >= (int)_laserCloudWidth - 3) { shiftCubes(0, -1); centerCubeI--; _laserCloudCenWidth--; }
  while (centerCubeJ < 3) { shiftCubes(1, +1); centerCubeJ++; _laserCloudCenHeight++; }
  while (centerCubeJ >= (int)_laserCloudHeight - 3) { shiftCubes(1, -1); centerCubeJ--; _laserCloudCenHeight--; }
  while (centerCubeK < 3) { shiftCubes(2, +1); centerCubeK++; _laserCloudCenDepth++; }
  while (centerCubeK >= (int)_laserCloudDepth - 3) { shiftCubes(2, -1); centerCubeK--; _laserCloudCenDepth--; }

  // 5 x 5 x 5 neighbourhood; a cube is "valid" when one of its corners lies within 30..150 degrees of the sensor's
  // up axis (law-of-cosines test against a point 10 m up the y axis, upstream :456-489)
  _laserCloudValidInd.clear();
  _laserCloudSurroundInd.clear();
  const pcl::PointXYZI sensorPos = (pcl::PointXYZI)_transformTobeMapped.pos;
  for (int i = centerCubeI - 2; i <= centerCubeI + 2; i++) {
    for (int j = centerCubeJ - 2; j <= centerCubeJ + 2; j++) {
      for (int k = centerCubeK - 2; k <= centerCubeK + 2; k++) {
        if (i < 0 || i >= (int)_laserCloudWidth || j < 0 || j >= (int)_laserCloudHeight || k < 0 ||
            k >= (int)_laserCloudDepth)
          continue;
        const float centerX = 50.0f * (i - _laserCloudCenWidth);
        const float centerY = 50.0f * (j - _laserCloudCenHeight);
        const float centerZ = 50.0f * (k - _laserCloudCenDepth);
        bool isInLaserFOV = false;
        for (int ii = -1; ii <= 1; ii += 2) {
          for (int jj = -1; jj <= 1; jj += 2) {
            for (int kk = -1; kk <= 1; kk += 2) {
              const float cx = centerX + 25.0f * ii, cy = centerY + 25.0f * jj, cz = centerZ + 25.0f * kk;
              const float d1x = sensorPos.x - cx, d1y = sensorPos.y - cy, d1z = sensorPos.z - cz;
              const float squaredSide1 = d1x * d1x + d1y * d1y + d1z * d1z;
              const float d2x = pointOnYAxis.x - cx, d2y = pointOnYAxis.y - cy, d2z = pointOnYAxis.z - cz;
              const float squaredSide2 = d2x * d2x + d2y * d2y + d2z * d2z;
              const float check1 = 100.0f + squaredSide1 - squaredSide2 - 10.0f * std::sqrt(3.0f) * std::sqrt(squaredSide1);
              const float check2 = 100.0f + squaredSide1 - squaredSide2 + 10.0f * std::sqrt(3.0f) * std::sqrt(squaredSide1);
              if (check1 < 0 && check2 > 0) isInLaserFOV = true;
            }
          }
        }
        const size_t cubeIdx = i + _laserCloudWidth * j + _laserCloudWidth * _laserCloudHeight * k;
        if (isInLaserFOV) _laserCloudValidInd.push_back(cubeIdx);
        _laserCloudSurroundInd.push_back(cubeIdx);
      }
    }
  }

  // surrounding-map clouds for the optimisation
  _laserCloudCornerFromMap->clear();
  _laserCloudSurfFromMap->clear();
  for (auto const& ind : _laserCloudValidInd) {
    *_laserCloudCornerFromMap += *_laserCloudCornerArray[ind];
    *_laserCloudSurfFromMap += *_laserCloudSurfArray[ind];
  }

  // feature stacks back into the (predicted) sensor frame, then voxel-filtered
  for (auto& pt : *_laserCloudCornerStack) pointAssociateTobeMapped(pt, pt);
  for (auto& pt : *_laserCloudSurfStack) pointAssociateTobeMapped(pt, pt);
  b200::voxelFilter(*_gpu, *_laserCloudCornerStack, b200::leafOf(_downSizeFilterCorner), *_laserCloudCornerStackDS, _bufA, _bufB);
  const size_t laserCloudCornerStackNum = _laserCloudCornerStackDS->size();
  b200::voxelFilter(*_gpu, *_laserCloudSurfStack, b200::leafOf(_downSizeFilterSurf), *_laserCloudSurfStackDS, _bufA, _bufB);
  const size_t laserCloudSurfStackNum = _laserCloudSurfStackDS->size();
  _laserCloudCornerStack->clear();
  _laserCloudSurfStack->clear();

  optimizeTransformTobeMapped();

  // insert the down-sized stack points into their cubes with the optimised pose
  for (size_t i = 0; i < laserCloudCornerStackNum; i++) {
    pointAssociateToMap(_laserCloudCornerStackDS->points[i], pointSel);
    size_t cubeInd;
    if (cubeIndexOf(pointSel, cubeInd)) _laserCloudCornerArray[cubeInd]->push_back(pointSel);
  }
  for (size_t i = 0; i < laserCloudSurfStackNum; i++) {
    pointAssociateToMap(_laserCloudSurfStackDS->points[i], pointSel);
    size_t cubeInd;
    if (cubeIndexOf(pointSel, cubeInd)) _laserCloudSurfArray[cubeInd]->push_back(pointSel);
  }

  // down-size every cube in the field of view
  for (auto const& ind : _laserCloudValidInd) {
    b200::voxelFilter(*_gpu, *_laserCloudCornerArray[ind], b200::leafOf(_downSizeFilterCorner), *_laserCloudCornerDSArray[ind], _bufA, _bufB);
    b200::voxelFilter(*_gpu, *_laserCloudSurfArray[ind], b200::leafOf(_downSizeFilterSurf), *_laserCloudSurfDSArray[ind], _bufA, _bufB);
    _laserCloudCornerArray[ind].swap(_laserCloudCornerDSArray[ind]);
    _laserCloudSurfArray[ind].swap(_laserCloudSurfDSArray[ind]);
  }

  transformFullResToMap();
  _downsizedMapCreated = createDownsizedMap();
  return true;
}

void BasicLaserMapping::updateIMU(IMUState2 const& newState) {
  _imuHistory.push_back(newState);
  if (_imuHistory.size() > 200) _imuHistory.erase(_imuHistory.begin());
}

void BasicLaserMapping::updateOdometry(double pitch, double yaw, double roll, double x, double y, double z) {
  _transformSum.rot_x = pitch;
  _transformSum.rot_y = yaw;
  _transformSum.rot_z = roll;
  _transformSum.pos.x() = float(x);
  _transformSum.pos.y() = float(y);
  _transformSum.pos.z() = float(z);
}

void BasicLaserMapping::updateOdometry(Twist const& twist) { _transformSum = twist; }

void BasicLaserMapping::optimizeTransformTobeMapped() {
  _lastIterations = 0;
  if (_laserCloudCornerFromMap->size() <= 10 || _laserCloudSurfFromMap->size() <= 100) return;

  // the reference rebuilds both k-d trees over the concatenated surrounding map every sweep (:636-637)
  b200::pack(*_laserCloudCornerFromMap, _bufA);
  _gpu->check(loam_b200_tree_build(_gpu->get(), LOAM_B200_TREE_MAP_CORNER, _bufA.data(), (int)_laserCloudCornerFromMap->size()),
              "loam_b200_tree_build(corner map)");
  b200::pack(*_laserCloudSurfFromMap, _bufA);
  _gpu->check(loam_b200_tree_build(_gpu->get(), LOAM_B200_TREE_MAP_SURF, _bufA.data(), (int)_laserCloudSurfFromMap->size()),
              "loam_b200_tree_build(surf map)");
  b200::pack(*_laserCloudCornerStackDS, _bufA);
  b200::pack(*_laserCloudSurfStackDS, _bufB);
  _gpu->check(loam_b200_map_set_queries(_gpu->get(), _bufA.data(), (int)_laserCloudCornerStackDS->size(), _bufB.data(),
                                        (int)_laserCloudSurfStackDS->size()),
              "loam_b200_map_set_queries");

  for (size_t iterCount = 0; iterCount < _maxIterations; iterCount++) {
    _lastIterations = iterCount + 1;
    loam_b200_pose pose;
    b200::fillPose(_transformTobeMapped, pose);
    loam_b200_normal_eq ne;
    _gpu->check(loam_b200_map_iterate(_gpu->get(), &pose, &ne), "loam_b200_map_iterate");
    if (ne.n_selected < 50) continue;

    float x[6];
    _solver->solve(ne, iterCount == 0, 100.f, x);

    _transformTobeMapped.rot_x += x[0];
    _transformTobeMapped.rot_y += x[1];
    _transformTobeMapped.rot_z += x[2];
    _transformTobeMapped.pos.x() += x[3];
    _transformTobeMapped.pos.y() += x[4];
    _transformTobeMapped.pos.z() += x[5];

    const float deltaR = std::sqrt(std::pow(rad2deg(x[0]), 2) + std::pow(rad2deg(x[1]), 2) + std::pow(rad2deg(x[2]), 2));
    const float deltaT = std::sqrt(std::pow(x[3] * 100, 2) + std::pow(x[4] * 100, 2) + std::pow(x[5] * 100, 2));
    if (deltaR < _deltaRAbort && deltaT < _deltaTAbort) break;
  }
  transformUpdate();
}

void BasicLaserMapping::seedMap(Cloud const& cornerPoints, Cloud const& surfPoints) {
  size_t idx;
  for (auto const& p : cornerPoints.points)
    if (cubeIndexOf(p, idx)) _laserCloudCornerArray[idx]->push_back(p);
  for (auto const& p : surfPoints.points)
    if (cubeIndexOf(p, idx)) _laserCloudSurfArray[idx]->push_back(p);
}

void BasicLaserMapping::collectMap(Cloud& corner, Cloud& surf) const {
  corner.clear();
  surf.clear();
  for (auto const& c : _laserCloudCornerArray) corner += *c;
  for (auto const& c : _laserCloudSurfArray) surf += *c;
}

}  // namespace loam
