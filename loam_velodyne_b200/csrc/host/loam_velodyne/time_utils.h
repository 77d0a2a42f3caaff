// loam::Time / toSec (upstream include/loam_velodyne/time_utils.h).
#pragma once
#include <chrono>

namespace loam {
using Time = std::chrono::system_clock::time_point;
inline double toSec(Time::duration d) { return std::chrono::duration<double>(d).count(); }
}  // namespace loam
