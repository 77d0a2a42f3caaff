// Device-resident surrounding map: the reference keeps 21 x 11 x 21 cubes of 50 m as
// separate host clouds, concatenates the cubes in the field of view every sweep, inserts the new stack points and
// voxel-filters every visible cube (BasicLaserMapping.cpp:300-509, 536-593).  Here each map kind (corner / surface)
// is ONE flat point pool in HBM; a point's cube is a pure function of its position and the grid centre
// (:540-553), so rolling the grid is just a change of three integers, "concatenate the valid cubes" is a stream
// compaction and "filter every valid cube" is one sort by (cube rank, voxel z, y, x) + run means:
//   classify  -> per point: rank of its cube in this sweep's valid list, KEEP (cube not visible), or DROP (left grid)
//   compact   -> from-map cloud (rank < n_valid) for the k-NN BVH
//   filter    -> key = rank<<24 | vz<<16 | vy<<8 | vx over (valid old points + inserted points) -> radix sort ->
//                centroid per run (pcl::VoxelGrid grouping: absolute lattice cell floor(x / leaf) per cube, output in
//                ascending voxel index per cube) ; pool' = filtered ++ kept
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

// voxel key inside a valid cube: rank << 24 | vz << 16 | vy << 8 | vx, voxel coordinates relative to a base half a
// metre below the cube's lower corner so they stay within 8 bits (<= 253 cells at the 0.2 m leaf)
__global__ void cube_voxel_key_kernel(const float4* __restrict__ p, const unsigned char* __restrict__ cls, int n,
                                      CubeGrid g, float inv_leaf, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = p[i];
  int ci, cj, ck;
  cube_index(q, g, ci, cj, ck);
  const float bx = 50.0f * (float)(ci - g.cen_w) - 25.5f;
  const float by = 50.0f * (float)(cj - g.cen_h) - 25.5f;
  const float bz = 50.0f * (float)(ck - g.cen_d) - 25.5f;
  const int vx = (int)floorf(q.x * inv_leaf) - (int)floorf(bx * inv_leaf);
  const int vy = (int)floorf(q.y * inv_leaf) - (int)floorf(by * inv_leaf);
  const int vz = (int)floorf(q.z * inv_leaf) - (int)floorf(bz * inv_leaf);
  keys[i] = ((unsigned)cls[i] << 24) | ((unsigned)(vz & 255) << 16) | ((unsigned)(vy & 255) << 8) | (unsigned)(vx & 255);
  vals[i] = i;
}

// stack points: pointAssociateToMap with the predicted pose then pointAssociateTobeMapped back into the sensor
// frame (BasicLaserMapping.cpp:282-292 and :512-516; the round trip is not an identity in fp32 and is reproduced)
struct ToSensorArgs {
  float srx, crx, sry, cry, srz, crz;  // of the SAME pose; negated angles flip the sine only (Angle.h:47-53)
  float tx, ty, tz;
};
__global__ void stack_roundtrip_kernel(const float4* __restrict__ in, int n, MapIterArgs a, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = stack_roundtrip(a, in[i]);
}

// new map points: pointAssociateToMap(stackDS) with the optimised pose (:536-577)
__global__ void insert_points_kernel(const float4* __restrict__ stack_ds, int n, MapIterArgs a,
                                     float4* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = stack_ds[i];
  float x, y, z;
  associate_to_map(a, q, x, y, z);
  dst[i] = make_float4(x, y, z, q.w);
}

}  // namespace loamb
