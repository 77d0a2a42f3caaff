// TEST INFRASTRUCTURE (oracle shim) -- not part of the shipped product.
// Minimal pcl/point_types.h: the two point types the reference's hot path uses, with PCL's memory layout
// (PointXYZ 16 B; PointXYZI 32 B, intensity at offset 16; both expose `data[4]`, which
// nanoflann_pcl.h:150 relies on).  PCL itself is not vendored by the reference (CMakeLists.txt:15).
#pragma once
#include <cmath>
#include <Eigen/Core>

#define pcl_isfinite(x) std::isfinite(x)

namespace pcl {

struct alignas(16) PointXYZ {
  union {
    float data[4];
    struct { float x, y, z; };
  };
  PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
  PointXYZ(float x_, float y_, float z_) : data{x_, y_, z_, 1.f} {}
};

struct alignas(16) PointXYZI {
  union {
    float data[4];
    struct { float x, y, z; };
  };
  union {
    struct { float intensity; };
    float data_c[4];
  };
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};

static_assert(sizeof(PointXYZ) == 16, "PointXYZ layout");
static_assert(sizeof(PointXYZI) == 32, "PointXYZI layout");

}  // namespace pcl
