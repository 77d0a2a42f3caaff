// loam::Twist -- three Angles + a position (upstream include/loam_velodyne/Twist.h:15-27).
#pragma once
#include "Angle.h"
#include "Vector3.h"

namespace loam {

class Twist {
 public:
  Twist() {}
  Angle rot_x;
  Angle rot_y;
  Angle rot_z;
  Vector3 pos;
};

}  // namespace loam
