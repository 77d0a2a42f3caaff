// loam::Angle -- value type of the drop-in API (upstream include/loam_velodyne/Angle.h:16-67): a float angle in
// radians that caches float std::cos / std::sin at construction; unary minus flips the sign of the sine only.
#pragma once
#include <cmath>

namespace loam {

class Angle {
 public:
  Angle() : rad_(0.f), cos_(1.f), sin_(0.f) {}
  Angle(float radValue) : rad_(radValue), cos_(std::cos(radValue)), sin_(std::sin(radValue)) {}

  void operator+=(const float& r) { *this = Angle(rad_ + r); }
  void operator+=(const Angle& o) { *this = Angle(rad_ + o.rad_); }
  void operator-=(const float& r) { *this = Angle(rad_ - r); }
  void operator-=(const Angle& o) { *this = Angle(rad_ - o.rad_); }

  Angle operator-() const {
    Angle a;
    a.rad_ = -rad_;
    a.cos_ = cos_;
    a.sin_ = -sin_;
    return a;
  }

  float rad() const { return rad_; }
  float deg() const { return float(rad_ * 180 / M_PI); }
  float cos() const { return cos_; }
  float sin() const { return sin_; }

 private:
  float rad_, cos_, sin_;
};

}  // namespace loam
