// loam::Vector3 -- value type of the drop-in API (upstream include/loam_velodyne/Vector3.h:13-69).  Upstream derives
// from Eigen::Vector4f; this build carries its own four floats so the hot path has no Eigen dependency, and keeps
// the accessors / conversions / arithmetic the Basic* classes and their adapters use.
#pragma once
#include <pcl/point_types.h>

namespace loam {

class Vector3 {
 public:
  Vector3() { v_[0] = v_[1] = v_[2] = v_[3] = 0.f; }
  Vector3(float x, float y, float z) { v_[0] = x; v_[1] = y; v_[2] = z; v_[3] = 0.f; }
  Vector3(const pcl::PointXYZI& p) { v_[0] = p.x; v_[1] = p.y; v_[2] = p.z; v_[3] = 0.f; }

  Vector3& operator=(const pcl::PointXYZ& p) { v_[0] = p.x; v_[1] = p.y; v_[2] = p.z; return *this; }
  Vector3& operator=(const pcl::PointXYZI& p) { v_[0] = p.x; v_[1] = p.y; v_[2] = p.z; return *this; }

  float x() const { return v_[0]; }
  float y() const { return v_[1]; }
  float z() const { return v_[2]; }
  float& x() { return v_[0]; }
  float& y() { return v_[1]; }
  float& z() { return v_[2]; }
  float operator()(int i) const { return v_[i]; }
  float& operator()(int i) { return v_[i]; }

  Vector3& operator+=(const Vector3& o) { for (int i = 0; i < 4; i++) v_[i] += o.v_[i]; return *this; }
  Vector3& operator-=(const Vector3& o) { for (int i = 0; i < 4; i++) v_[i] -= o.v_[i]; return *this; }
  Vector3& operator*=(float s) { for (int i = 0; i < 4; i++) v_[i] *= s; return *this; }
  Vector3& operator/=(float s) { for (int i = 0; i < 4; i++) v_[i] /= s; return *this; }

  operator pcl::PointXYZI() const {
    pcl::PointXYZI p;
    p.x = v_[0]; p.y = v_[1]; p.z = v_[2]; p.intensity = 0.f;
    return p;
  }

 private:
  float v_[4];
};

inline Vector3 operator+(Vector3 a, const Vector3& b) { a += b; return a; }
inline Vector3 operator-(Vector3 a, const Vector3& b) { a -= b; return a; }
inline Vector3 operator*(Vector3 a, float s) { a *= s; return a; }
inline Vector3 operator*(float s, Vector3 a) { a *= s; return a; }
inline Vector3 operator/(Vector3 a, float s) { a /= s; return a; }

}  // namespace loam
