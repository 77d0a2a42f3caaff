// Host-side O(1) rotation helpers for pose bookkeeping (upstream src/lib/math_utils.h:129-275 semantics:
// elementary rotations applied in the named order, float arithmetic, cached Angle sin/cos).
#pragma once
#include <cmath>

#include "loam_velodyne/Angle.h"
#include "loam_velodyne/Vector3.h"

namespace loam {
namespace hostmath {

struct XYZ {
  float x, y, z;
};
inline XYZ get(const Vector3& v) { return {v.x(), v.y(), v.z()}; }
inline void put(Vector3& v, const XYZ& p) { v.x() = p.x; v.y() = p.y; v.z() = p.z; }
template <typename P>
inline XYZ get(const P& p) { return {p.x, p.y, p.z}; }
template <typename P>
inline void put(P& q, const XYZ& p) { q.x = p.x; q.y = p.y; q.z = p.z; }

inline void aboutX(XYZ& p, const Angle& a) {
  const float y = p.y;
  p.y = a.cos() * y - a.sin() * p.z;
  p.z = a.sin() * y + a.cos() * p.z;
}
inline void aboutY(XYZ& p, const Angle& a) {
  const float x = p.x;
  p.x = a.cos() * x + a.sin() * p.z;
  p.z = a.cos() * p.z - a.sin() * x;
}
inline void aboutZ(XYZ& p, const Angle& a) {
  const float x = p.x;
  p.x = a.cos() * x - a.sin() * p.y;
  p.y = a.sin() * x + a.cos() * p.y;
}

template <typename T>
inline void rotateZXY(T& v, const Angle& angZ, const Angle& angX, const Angle& angY) {
  XYZ p = get(v);
  aboutZ(p, angZ);
  aboutX(p, angX);
  aboutY(p, angY);
  put(v, p);
}
template <typename T>
inline void rotateYXZ(T& v, const Angle& angY, const Angle& angX, const Angle& angZ) {
  XYZ p = get(v);
  aboutY(p, angY);
  aboutX(p, angX);
  aboutZ(p, angZ);
  put(v, p);
}

inline float rad2deg(float r) { return (float)(r * 180.0 / M_PI); }

}  // namespace hostmath
}  // namespace loam
