// Cube bookkeeping shared by the map code: the reference keeps 21 x 11 x 21 cubes of 50 m as separate host clouds
// (BasicLaserMapping.cpp:300-509, 536-593); here a point's cube is a pure function of its position and the grid centre
// (:540-553).  This header holds the cube index arithmetic, the cube classification + stream compaction used by the
// surround cloud (createDownsizedMap) and by the from-map debug clouds, and the stack round-trip transform.  The map
// itself (persistent, cell-sorted, incrementally maintained) lives in mapstore.cuh.
#pragma once

#include "lbvh.cuh"
#include "mapping_lm.cuh"
#include "voxel.cuh"

namespace loamb {

constexpr int CUBE_W = 21, CUBE_H = 11, CUBE_D = 21;
constexpr int CUBE_NUM = CUBE_W * CUBE_H * CUBE_D;
constexpr unsigned char CLS_KEEP = 0xFE, CLS_DROP = 0xFF;

struct CubeGrid {
  int cen_w, cen_h, cen_d;
};

// cube index of a map-frame point, exactly BasicLaserMapping.cpp:540-553 (double arithmetic, truncation toward zero
// plus the negative-side decrement); -1 when outside the grid
__device__ __forceinline__ int cube_index(const float4& p, const CubeGrid& g, int& ci, int& cj, int& ck) {
  const double HALF = 25.0, SIZE = 50.0;
  ci = (int)(((double)p.x + HALF) / SIZE) + g.cen_w;
  cj = (int)(((double)p.y + HALF) / SIZE) + g.cen_h;
  ck = (int)(((double)p.z + HALF) / SIZE) + g.cen_d;
  if ((double)p.x + HALF < 0) ci--;
  if ((double)p.y + HALF < 0) cj--;
  if ((double)p.z + HALF < 0) ck--;
  if (ci >= 0 && ci < CUBE_W && cj >= 0 && cj < CUBE_H && ck >= 0 && ck < CUBE_D)
    return ci + CUBE_W * cj + CUBE_W * CUBE_H * ck;
  return -1;
}

// rank_of_cube: CUBE_NUM bytes, rank in the valid list or CLS_KEEP
__global__ void classify_kernel(const float4* __restrict__ p, int n, CubeGrid g,
                                const unsigned char* __restrict__ rank_of_cube, unsigned char* __restrict__ cls,
                                const int* __restrict__ n_dev = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (n_dev && i >= *n_dev) {  // launch bound above the live count: never selected
    cls[i] = CLS_DROP;
    return;
  }
  int ci, cj, ck;
  const int c = cube_index(p[i], g, ci, cj, ck);
  cls[i] = c < 0 ? CLS_DROP : rank_of_cube[c];
}

// ---- generic two-pass stream compaction on a predicate over cls[]
// MODE 0: cls < limit (valid cubes)      MODE 1: cls == CLS_KEEP
template <int MODE>
__device__ __forceinline__ bool cls_pred(unsigned char c, int limit) {
  return MODE == 0 ? ((int)c < limit) : (c == CLS_KEEP);
}

template <int MODE>
__global__ void __launch_bounds__(SCAN_BS)
compact_count_kernel(const unsigned char* __restrict__ cls, int n, int limit, unsigned* __restrict__ local_pos,
                     unsigned* __restrict__ block_sums) {
  __shared__ unsigned ws[32];
  const int i = blockIdx.x * SCAN_BS + threadIdx.x;
  const unsigned h = (i < n && cls_pred<MODE>(cls[i], limit)) ? 1u : 0u;
  unsigned x = h;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned w = ws[threadIdx.x];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
      if (threadIdx.x >= o) w += y;
    }
    ws[threadIdx.x] = w;
  }
  __syncthreads();
  const unsigned incl = x + ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0u);
  if (i < n) local_pos[i] = (incl - h) | (h << 31);
  if (threadIdx.x == SCAN_BS - 1) block_sums[blockIdx.x] = incl;
}

// scatter selected points (and optionally their source index / class) to dst + dst_offset
__global__ void compact_scatter_kernel(const float4* __restrict__ src, const unsigned char* __restrict__ cls, int n,
                                       const unsigned* __restrict__ local_pos, const unsigned* __restrict__ block_off,
                                       float4* __restrict__ dst, unsigned char* __restrict__ dst_cls, int dst_offset,
                                       const int* __restrict__ dst_offset_dev = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned pv = local_pos[i];
  if (!(pv >> 31)) return;
  if (dst_offset_dev) dst_offset += *dst_offset_dev;
  const unsigned d = (pv & 0x7fffffffu) + block_off[i / SCAN_BS] + (unsigned)dst_offset;
  dst[d] = src[i];
  if (dst_cls) dst_cls[d] = cls[i];
}

// stack points: pointAssociateToMap with the predicted pose then pointAssociateTobeMapped back into the sensor
// frame (BasicLaserMapping.cpp:282-292 and :512-516; the round trip is not an identity in fp32 and is reproduced)
__global__ void stack_roundtrip_kernel(const float4* __restrict__ in, int n, MapIterArgs a, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = stack_roundtrip(a, in[i]);
}

}  // namespace loamb
