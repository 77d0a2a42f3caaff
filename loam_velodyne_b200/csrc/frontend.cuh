// Ring-binning front end: MultiScanRegistration::process (MultiScanRegistration.cpp:160-238) with
// MultiScanMapper::getRingForAngle (:63-66).  An unordered sensor-frame xyz cloud in arrival order becomes the
// ring-ordered (x, y, z, ring + relTime) cloud BasicScanRegistration::processScanlines expects.
//
// The reference walks the cloud once with one piece of sequential state: `halfPassed` flips at the first kept point
// whose (branch-A adjusted) azimuth is more than pi past the start, and every later point takes branch B.  The flag is
// monotone, so the loop is two data-parallel passes around a min-reduction:
//   pass 1: per point: axis swap, finite / near-zero rejection, ring from the vertical angle, branch-A azimuth;
//           atomicMin of the first index whose branch-A azimuth trips the flag;
//   pass 2: per point: final azimuth (branch A up to and including that index, branch B after it), relTime,
//           intensity = ring + relTime, sort key = ring (255 = rejected);
//   then ONE stable 8-bit radix pass (cluster sort up to 131 k points, onesweep above) and a gather.
// Angles are evaluated in double and rounded once to float (the reference calls the float libm: equal up to the rare
// cases where a float libm result is not the correctly rounded one).
#pragma once

#include "ctx.cuh"

namespace loamb {

struct BinParams {
  float lower, factor;  // MultiScanMapper: _lowerBound, _factor = (nScanRings - 1) / (upper - lower)
  int n_rings;
  float scan_period;
};

__device__ __forceinline__ float atan2_f(float y, float x) { return (float)atan2((double)y, (double)x); }

__device__ __forceinline__ void bin_start_end(const float* __restrict__ xyz, int n, float& startOri, float& endOri) {
  startOri = -atan2_f(xyz[1], xyz[0]);
  // `-atan2 + 2 * float(M_PI)` is a float expression upstream (:166-167); the corrections below use the double constant
  endOri = -atan2_f(xyz[3 * (size_t)(n - 1) + 1], xyz[3 * (size_t)(n - 1) + 0]) + 2 * float(3.14159265358979323846);
  const double PI = 3.14159265358979323846;
  if ((double)(endOri - startOri) > 3 * PI) {
    endOri = (float)((double)endOri - 2 * PI);
  } else if ((double)(endOri - startOri) < PI) {
    endOri = (float)((double)endOri + 2 * PI);
  }
}

// ring id or -1 (rejected); ori_a = azimuth after the branch-A adjustment
__device__ __forceinline__ int bin_point(const float* __restrict__ xyz, int i, const BinParams& p, float startOri, float& px,
                                         float& py, float& pz, float& ori_raw) {
  px = xyz[3 * (size_t)i + 1];
  py = xyz[3 * (size_t)i + 2];
  pz = xyz[3 * (size_t)i + 0];
  if (!isfinite(px) || !isfinite(py) || !isfinite(pz)) return -1;
  if ((double)(px * px + py * py + pz * pz) < 0.0001) return -1;
  const float angle = (float)atan((double)(py / sqrtf(px * px + pz * pz)));
  // int(((angle * 180 / M_PI) - lower) * factor + 0.5), :63-66
  const double deg = (double)(angle * 180.f) / 3.14159265358979323846;  // float * int, then / double
  const int ring = (int)((deg - (double)p.lower) * (double)p.factor + 0.5);
  if (ring >= p.n_rings || ring < 0) return -1;
  ori_raw = -atan2_f(px, pz);
  return ring;
}

__device__ __forceinline__ float bin_branch_a(float ori, float startOri) {
  const double PI = 3.14159265358979323846;
  if ((double)ori < (double)startOri - PI / 2) {
    ori = (float)((double)ori + 2 * PI);
  } else if ((double)ori > (double)startOri + PI * 3 / 2) {
    ori = (float)((double)ori - 2 * PI);
  }
  return ori;
}

__global__ void bin_pass1_kernel(const float* __restrict__ xyz, int n, BinParams p, int* __restrict__ first_half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float startOri, endOri;
  bin_start_end(xyz, n, startOri, endOri);
  float px, py, pz, ori;
  if (bin_point(xyz, i, p, startOri, px, py, pz, ori) < 0) return;
  ori = bin_branch_a(ori, startOri);
  if ((double)(ori - startOri) > 3.14159265358979323846) atomicMin(first_half, i);
}

__global__ void bin_pass2_kernel(const float* __restrict__ xyz, int n, BinParams p, const int* __restrict__ first_half,
                                 float4* __restrict__ pts, unsigned* __restrict__ keys, int* __restrict__ vals,
                                 int* __restrict__ ring_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double PI = 3.14159265358979323846;
  int ring = -1;
  if (i < n) {
    float startOri, endOri;
    bin_start_end(xyz, n, startOri, endOri);
    float px, py, pz, ori;
    ring = bin_point(xyz, i, p, startOri, px, py, pz, ori);
    if (ring >= 0) {
      if (i <= *first_half) {  // halfPassed was still false when the loop reached this point
        ori = bin_branch_a(ori, startOri);
      } else {
        ori = (float)((double)ori + 2 * PI);
        if ((double)ori < (double)endOri - PI * 3 / 2) {
          ori = (float)((double)ori + 2 * PI);
        } else if ((double)ori > (double)endOri + PI / 2) {
          ori = (float)((double)ori - 2 * PI);
        }
      }
      const float relTime = p.scan_period * (ori - startOri) / (endOri - startOri);
      pts[i] = make_float4(px, py, pz, (float)ring + relTime);
    }
    keys[i] = ring >= 0 ? (unsigned)ring : 255u;
    vals[i] = i;
  }
  // points per ring, warp-aggregated
  const unsigned peers = __match_any_sync(0xffffffffu, ring);
  if (ring >= 0 && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&ring_count[ring], __popc(peers));
}

__global__ void bin_gather_kernel(const float4* __restrict__ pts, const unsigned* __restrict__ keys,
                                  const int* __restrict__ order, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || keys[i] == 255u) return;  // rejected points sort behind every ring
  out[i] = pts[order[i]];
}

}  // namespace loamb
