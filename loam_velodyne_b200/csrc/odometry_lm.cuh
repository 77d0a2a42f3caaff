// Fused scan-to-scan Gauss-Newton iteration and the bulk transforms (O6, M1).
// One thread per sharp / flat feature point: transformToStart (BasicLaserOdometry.cpp:40-53) -> on every 5th
// iteration 1-NN in the last corner / surface cloud (BVH walk, d^2 < 25 gate) plus the +-2.5-ring linear scans in
// the ring-ordered last cloud (:253-302, :370-435; loop bounds reproduce the reference) ->
// point-to-line / point-to-plane residual and weight (:304-361, :437-481) -> Jacobian row (:497-553) -> the same
// warp-shuffle / last-CTA reduction as the mapping kernel.
#pragma once

#include "mapping_lm.cuh"

namespace loamb {

struct OdomIterArgs {
  float rx, ry, rz;     // _transform rotation (rad)
  float tx, ty, tz;     // _transform translation
  float inv_sp;         // 1.f / scanPeriod
  int iter;
  int n_last_corner, n_last_surf;
  // Jacobian terms (BasicLaserOdometry.cpp:514-543), every product formed on the host in the reference's order
  float g1a, g1b, g1c, k1, k2, k3;      // arx group multiplying coeff.x
  float t1, t2, t3, k4, k5, k6;         // arx group multiplying coeff.y
  float u1, u2, u3, k7, k8, k9;         // arx group multiplying coeff.z
  float e1, e2, e3, e4, e5, k10;        // ary group multiplying coeff.x
  float f1, f2, f3, k11;                // ary group multiplying coeff.z
  float g1;                             // arz
  float h1, h2, k12, k13;
  float atx_y, aty_y, atz_x, atz_y, atz_z;  // crx*srz, crx*crz, crx*sry, srx, crx*cry
};

// pose -> kernel arguments (BasicLaserOdometry.cpp:514-543), host and device
LOAMB_HD inline void odom_args_from(const float rot[3], const float sin_[3], const float cos_[3], const float pos[3],
                                    float inv_scan_period, int iter, OdomIterArgs& a) {
  const float srx = sin_[0], crx = cos_[0], sry = sin_[1], cry = cos_[1], srz = sin_[2],
              crz = cos_[2];
  const float tx = pos[0], ty = pos[1], tz = pos[2];
  a.rx = rot[0]; a.ry = rot[1]; a.rz = rot[2];
  a.tx = tx; a.ty = ty; a.tz = tz;
  a.inv_sp = inv_scan_period;
  a.iter = iter;
  // BasicLaserOdometry.cpp:514-543 with s = 1 (every `s *` is an exact multiplication by one)
  a.g1a = -crx * sry * srz;  a.g1b = crx * crz * sry;  a.g1c = srx * sry;
  a.k1 = tx * crx * sry * srz;  a.k2 = ty * crx * crz * sry;  a.k3 = tz * srx * sry;
  a.t1 = srx * srz;  a.t2 = crz * srx;  a.t3 = crx;
  a.k4 = ty * crz * srx;  a.k5 = tz * crx;  a.k6 = tx * srx * srz;
  a.u1 = crx * cry * srz;  a.u2 = crx * cry * crz;  a.u3 = cry * srx;
  a.k7 = tz * cry * srx;  a.k8 = ty * crx * cry * crz;  a.k9 = tx * crx * cry * srz;
  a.e1 = -crz * sry - cry * srx * srz;
  a.e2 = cry * crz * srx - sry * srz;
  a.e3 = crx * cry;
  a.e4 = crz * sry + cry * srx * srz;
  a.e5 = sry * srz - cry * crz * srx;
  a.k10 = tz * crx * cry;
  a.f1 = cry * crz - srx * sry * srz;
  a.f2 = cry * srz + crz * srx * sry;
  a.f3 = crx * sry;
  a.k11 = tz * crx * sry;
  a.g1 = -cry * srz - crz * srx * sry;
  a.h1 = -crx * crz;  a.h2 = crx * srz;
  a.k12 = ty * crx * srz;  a.k13 = tx * crx * crz;
  a.atx_y = crx * srz;  a.aty_y = crx * crz;
  a.atz_x = crx * sry;  a.atz_y = srx;  a.atz_z = crx * cry;
}

// cos / sin evaluated in double and rounded once: agrees with a correctly rounded float libm (the reference
// calls std::cos / std::sin on float, Angle.h:23-26) except in rare double-rounding cases.
__device__ __forceinline__ void sincos_f(float a, float& s, float& c) {
  double sd, cd;
  sincos((double)a, &sd, &cd);
  s = (float)sd;
  c = (float)cd;
}

// ---- device-resident Gauss-Newton loop (lmstep.cuh): state block + the step the last CTA of an iteration runs
// what the iteration kernels of the device-resident loop work on: written once per sweep by the init kernel, so the
// kernels inside the (pre-instantiated) loop graph take nothing but the state pointer
struct OdomLoopIo {
  TreeView corner_tree, surf_tree;
  const float4* last_corner;
  const float4* last_surf;
  const float4* queries;
  int* ind;
  const int* ring_off_corner;
  const int* ring_off_surf;
  int n_sharp, n_flat, sharp_blocks, n_blocks;
};
struct OdomLmState {
  LmHeader h;
  GnState gn;
  float inv_sp, delta_t_abort, delta_r_abort;
  int max_iter, n_last_corner, n_last_surf;
  OdomIterArgs args;  // arguments of iteration h.iter
  OdomLoopIo io;
};

__device__ inline void odom_lm_refresh_args(OdomLmState* st) {
  float sn[3], cs[3];
  for (int i = 0; i < 3; i++) sincos_f(st->h.rot[i], sn[i], cs[i]);
  OdomIterArgs a;
  odom_args_from(st->h.rot, sn, cs, st->h.pos, st->inv_sp, st->h.iter, a);
  a.n_last_corner = st->n_last_corner;
  a.n_last_surf = st->n_last_surf;
  st->args = a;
}

__global__ void odom_lm_init_kernel(OdomLmState* st, float rx, float ry, float rz, float tx, float ty, float tz, float inv_sp,
                                    float delta_t_abort, float delta_r_abort, int max_iter, int n_last_corner,
                                    int n_last_surf, OdomLoopIo io, int mb_seq) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->io = io;
  st->h.mb_seq = mb_seq;
  st->h.rot[0] = rx; st->h.rot[1] = ry; st->h.rot[2] = rz;
  st->h.pos[0] = tx; st->h.pos[1] = ty; st->h.pos[2] = tz;
  st->h.iter = 0;
  st->h.done = max_iter <= 0 ? 1 : 0;
  st->h.iters_run = 0;
  st->gn.degenerate = 0;
  st->inv_sp = inv_sp;
  st->delta_t_abort = delta_t_abort;
  st->delta_r_abort = delta_r_abort;
  st->max_iter = max_iter;
  st->n_last_corner = n_last_corner;
  st->n_last_surf = n_last_surf;
  odom_lm_refresh_args(st);
}

// One warp, after the normal equations of iteration h.iter are complete in s_r[0..31] (shared memory)
// (BasicLaserOdometry.cpp:484-488 skip, :559-622 solve / update / NaN reset / convergence); lmstep_warp.cuh
__device__ inline void odom_lm_step_warp(OdomLmState* st, const float* s_r) {
  const int lane = threadIdx.x & 31;
  const int iter = st->h.iter;
  float rot[3], pos[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { rot[i] = st->h.rot[i]; pos[i] = st->h.pos[i]; }
  bool converged = false;
  if ((int)(s_r[27] + 0.5f) >= 10) {
    float x[6];
    gn_solve_warp(s_r, iter == 0, 10.f, &st->gn, x);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      rot[i] = rot[i] + x[i];
      pos[i] += x[3 + i];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (!isfinite(rot[i])) rot[i] = 0.f;
      if (!isfinite(pos[i])) pos[i] = 0.f;
    }
    float deltaR, deltaT;
    gn_deltas(x, deltaR, deltaT);
    converged = deltaR < st->delta_r_abort && deltaT < st->delta_t_abort;
  }
  // arguments of the next iteration: sin / cos of the three angles on lanes 0..2, the rest on lane 0
  float sn, cs;
  sincos_f(lane == 1 ? rot[1] : (lane == 2 ? rot[2] : rot[0]), sn, cs);
  float s3[3], c3[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    s3[i] = __shfl_sync(0xffffffffu, sn, i);
    c3[i] = __shfl_sync(0xffffffffu, cs, i);
  }
  if (lane == 0) {
    OdomIterArgs a;
    odom_args_from(rot, s3, c3, pos, st->inv_sp, iter + 1, a);
    a.n_last_corner = st->n_last_corner;
    a.n_last_surf = st->n_last_surf;
    st->args = a;
#pragma unroll
    for (int i = 0; i < 3; i++) { st->h.rot[i] = rot[i]; st->h.pos[i] = pos[i]; }
    st->h.iters_run = iter + 1;
    st->h.iter = iter + 1;
    if (converged || iter + 1 >= st->max_iter) st->h.done = 1;
    __threadfence();
  }
  __syncwarp();
}

__device__ __forceinline__ void rot_zxy(float& x, float& y, float& z, float sz, float cz, float sx, float cx,
                                        float sy, float cy) {
  const float x1 = cz * x - sz * y;
  const float y1 = sz * x + cz * y;
  const float y2 = cx * y1 - sx * z;
  const float z2 = sx * y1 + cx * z;
  const float x3 = cy * x1 + sy * z2;
  const float z3 = cy * z2 - sy * x1;
  x = x3; y = y2; z = z3;
}
__device__ __forceinline__ void rot_yxz(float& x, float& y, float& z, float sy, float cy, float sx, float cx,
                                        float sz, float cz) {
  const float x1 = cy * x + sy * z;
  const float z1 = cy * z - sy * x;
  const float y2 = cx * y - sx * z1;
  const float z2 = sx * y + cx * z1;
  const float x3 = cz * x1 - sz * y2;
  const float y3 = sz * x1 + cz * y2;
  x = x3; y = y3; z = z2;
}

__device__ __forceinline__ void transform_to_start(const OdomIterArgs& a, const float4& pi, float& x, float& y,
                                                   float& z) {
  const float s = a.inv_sp * (pi.w - (float)(int)pi.w);
  x = pi.x - s * a.tx;
  y = pi.y - s * a.ty;
  z = pi.z - s * a.tz;
  float sx, cx, sy, cy, sz, cz;
  sincos_f(-s * a.rx, sx, cx);
  sincos_f(-s * a.ry, sy, cy);
  sincos_f(-s * a.rz, sz, cz);
  rot_zxy(x, y, z, sz, cz, sx, cx, sy, cy);
}

__device__ __forceinline__ float sqdiff3(const float4& a, float bx, float by, float bz) {
  const float dx = a.x - bx, dy = a.y - by, dz = a.z - bz;
  return dx * dx + dy * dy + dz * dz;
}

// ---- ring offsets of a last-sweep cloud: off[r] = first index whose ring id (int(intensity)) is >= r, r = 0 .. 256;
// off[257] != 0 flags a cloud that is not ring-ordered or has ring ids outside [0, 255].  For a ring-ordered cloud the
// reference's scan loops with their ring `break` (:262-276, :281-296) visit exactly an index range given by two of
// these offsets, which removes the only dependency between the steps of the scan.
constexpr int RING_OFF_WORDS = 258;
// blockIdx.y selects the cloud: one launch covers the last corner and the last surface cloud (off1 = nullptr: one cloud)
__global__ void ring_offsets_kernel(const float4* __restrict__ last0, int n0, int* __restrict__ off0,
                                    const float4* __restrict__ last1 = nullptr, int n1 = 0, int* __restrict__ off1 = nullptr) {
  const float4* __restrict__ last = blockIdx.y == 0 ? last0 : last1;
  const int n = blockIdx.y == 0 ? n0 : n1;
  int* __restrict__ off = blockIdx.y == 0 ? off0 : off1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = (int)last[i].w;
  const int prev = i > 0 ? (int)last[i - 1].w : -1;
  if (r < 0 || r > 255 || prev > r) {
    off[257] = 1;
    return;
  }
  for (int rr = max(prev, -1) + 1; rr <= r; rr++) off[rr] = i;
  if (i == n - 1)
    for (int rr = r + 1; rr <= 256; rr++) off[rr] = n;
}

// ---- correspondence search, every 5th iteration (BasicLaserOdometry.cpp:250-302, :368-435): one WARP per feature
// point.  Lane 0 walks the BVH for the closest point (d^2 < 25 gate); then the whole warp scans the ring-ordered last
// cloud forwards and backwards 32 candidates at a time.  The reference's sequential loops are reproduced exactly:
// a chunk stops at the first candidate that trips the ring `break` (ballot + ffs), candidates are compared by
// (distance, scan order) so the strict `<` updates of the serial loop pick the same index, and the forward loops
// keep the reference's bound by the CURRENT feature count (:262, :378).
struct ScanBest {
  float d;
  int ord;  // position in the reference's visiting order (forward first, then backward)
  int idx;
};
// candidate of the serial loop `if (d < best) { best = d; idx = j; }` with best starting at 25: only d < 25 can ever
// win, and among equal distances the one visited first
__device__ __forceinline__ void scan_best_min(ScanBest& a, float d, int ord, int idx) {
  if (d < 25.f && (d < a.d || (d == a.d && ord < a.ord))) { a.d = d; a.ord = ord; a.idx = idx; }
}
__device__ __forceinline__ void scan_best_warp(ScanBest& a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float d = __shfl_xor_sync(0xffffffffu, a.d, o);
    const int ord = __shfl_xor_sync(0xffffffffu, a.ord, o);
    const int idx = __shfl_xor_sync(0xffffffffu, a.idx, o);
    scan_best_min(a, d, ord, idx);
  }
}

// DEVLOOP (device-resident loop, lmstep.cuh): arguments of the current iteration from the state block (staged in
// shared memory; the by-value argument of the per-iteration API stays in the constant bank); no work once converged
template <bool DEVLOOP>
__global__ void __launch_bounds__(LM_THREADS)
odom_search_kernel(TreeView corner_tree, TreeView surf_tree, const float4* last_corner, const float4* last_surf,
                   const float4* queries, int n_sharp, int n_flat, OdomIterArgs a_param, int* ind,
                   const OdomLmState* __restrict__ lm = nullptr, const int* ring_off_corner = nullptr,
                   const int* ring_off_surf = nullptr) {
  __shared__ OdomIterArgs s_args;
  if (DEVLOOP) {
    if (lm->h.done || lm->h.iter % 5 != 0) return;  // the loop graph launches this kernel every iteration
    if (threadIdx.x < (int)(sizeof(OdomIterArgs) / 4))
      reinterpret_cast<float*>(&s_args)[threadIdx.x] = reinterpret_cast<const float*>(&lm->args)[threadIdx.x];
    __syncthreads();
    const OdomLoopIo& io = lm->io;  // uniform loads
    corner_tree = io.corner_tree; surf_tree = io.surf_tree;
    last_corner = io.last_corner; last_surf = io.last_surf; queries = io.queries;
    n_sharp = io.n_sharp; n_flat = io.n_flat; ind = io.ind;
    ring_off_corner = io.ring_off_corner; ring_off_surf = io.ring_off_surf;
  }
  const OdomIterArgs& a = DEVLOOP ? s_args : a_param;
  const int lane = threadIdx.x & 31;
  const int qi = (blockIdx.x * LM_THREADS + threadIdx.x) >> 5;  // one warp per query
  if (qi >= n_sharp + n_flat) return;
  const bool is_corner = qi < n_sharp;
  const float4* last = is_corner ? last_corner : last_surf;
  const int n_last = is_corner ? a.n_last_corner : a.n_last_surf;
  const float4 po = queries[qi];
  float sx, sy, sz;
  transform_to_start(a, po, sx, sy, sz);  // every lane evaluates the same values
  int i1 = -1;
  if (lane == 0) {
    KnnResult<1> nn;
    knn_walk<1>(is_corner ? corner_tree : surf_tree, sx, sy, sz, 25.0f, nn);
    i1 = nn.idx[0];
  }
  i1 = __shfl_sync(0xffffffffu, i1, 0);
  int i2 = -1, i3 = -1;
  if (i1 >= 0) {
    const int scan = (int)last[i1].w;
    ScanBest b2{25.f, 0x7fffffff, -1}, b3{25.f, 0x7fffffff, -1};
    int ord = 0;
    const int* ring_off = is_corner ? ring_off_corner : ring_off_surf;
    if (ring_off && ring_off[257] == 0 && scan >= 0 && scan <= 252) {
      // ring-ordered cloud: both scans are plain index ranges (no `break` to resolve between steps, the loads pipeline).
      // Visiting order of the serial loops = forward ascending, then backward descending: `ord` encodes it for ties.
      const int fend = min(min(is_corner ? n_sharp : n_flat, n_last), ring_off[scan + 3]);  // ring <= scan + 2.5
#pragma unroll 4
      for (int j = i1 + 1 + lane; j < fend; j += 32) {
        const float4 p = last[j];
        const int r = (int)p.w;
        const float d = sqdiff3(p, sx, sy, sz);
        if (is_corner) {
          if (r > scan) scan_best_min(b2, d, j - (i1 + 1), j);
        } else {
          if (r <= scan) scan_best_min(b2, d, j - (i1 + 1), j);
          else scan_best_min(b3, d, j - (i1 + 1), j);
        }
      }
      const int lo = ring_off[max(scan - 2, 0)];  // ring >= scan - 2.5
#pragma unroll 4
      for (int j = i1 - 1 - lane; j >= lo; j -= 32) {
        const float4 p = last[j];
        const int r = (int)p.w;
        const float d = sqdiff3(p, sx, sy, sz);
        const int o2 = 0x20000000 + (i1 - 1 - j);
        if (is_corner) {
          if (r < scan) scan_best_min(b2, d, o2, j);
        } else {
          if (r >= scan) scan_best_min(b2, d, o2, j);
          else scan_best_min(b3, d, o2, j);
        }
      }
    } else {
    // Both scans visit 32 x SCAN_UNROLL candidates per step: the loads of a step are independent (issued together),
    // the ring `break` is then resolved chunk by chunk in visiting order, so the result is that of the serial loop.
    // (One chunk per step made the kernel a chain of ~60 dependent global loads per direction on the surface cloud:
    // 70 us at 14 % of the warp slots, profiles/r1_v6_odom_search.md.)
    constexpr int SCAN_UNROLL = 4;
    // forward: j = i1 + 1 .. fend - 1 while ring <= scan + 2.5
    const int fend = min(is_corner ? n_sharp : n_flat, n_last);
    for (int j0 = i1 + 1; j0 < fend; j0 += 32 * SCAN_UNROLL) {
      float4 p[SCAN_UNROLL];
#pragma unroll
      for (int u = 0; u < SCAN_UNROLL; u++) {
        const int j = j0 + u * 32 + lane;
        p[u] = j < fend ? last[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      bool stop_all = false;
#pragma unroll
      for (int u = 0; u < SCAN_UNROLL; u++) {
        if (stop_all) continue;
        const int j = j0 + u * 32 + lane;
        const bool in = j < fend;
        const int r = (int)p[u].w;
        const bool brk = in && (double)r > (double)scan + 2.5;
        const float d = sqdiff3(p[u], sx, sy, sz);
        const unsigned bm = __ballot_sync(0xffffffffu, brk);
        const int stop = bm ? __ffs(bm) - 1 : 32;
        if (in && lane < stop) {
          if (is_corner) {
            if (r > scan) scan_best_min(b2, d, ord + u * 32 + lane, j);
          } else {
            if (r <= scan) scan_best_min(b2, d, ord + u * 32 + lane, j);
            else scan_best_min(b3, d, ord + u * 32 + lane, j);
          }
        }
        if (bm) stop_all = true;
      }
      ord += 32 * SCAN_UNROLL;
      if (stop_all) break;
    }
    // backward: j = i1 - 1 .. 0 while ring >= scan - 2.5
    for (int j0 = i1 - 1; j0 >= 0; j0 -= 32 * SCAN_UNROLL) {
      float4 p[SCAN_UNROLL];
#pragma unroll
      for (int u = 0; u < SCAN_UNROLL; u++) {
        const int j = j0 - u * 32 - lane;
        p[u] = j >= 0 ? last[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      bool stop_all = false;
#pragma unroll
      for (int u = 0; u < SCAN_UNROLL; u++) {
        if (stop_all) continue;
        const int j = j0 - u * 32 - lane;
        const bool in = j >= 0;
        const int r = (int)p[u].w;
        const bool brk = in && (double)r < (double)scan - 2.5;
        const float d = sqdiff3(p[u], sx, sy, sz);
        const unsigned bm = __ballot_sync(0xffffffffu, brk);
        const int stop = bm ? __ffs(bm) - 1 : 32;
        if (in && lane < stop) {
          if (is_corner) {
            if (r < scan) scan_best_min(b2, d, ord + u * 32 + lane, j);
          } else {
            if (r >= scan) scan_best_min(b2, d, ord + u * 32 + lane, j);
            else scan_best_min(b3, d, ord + u * 32 + lane, j);
          }
        }
        if (bm) stop_all = true;
      }
      ord += 32 * SCAN_UNROLL;
      if (stop_all) break;
    }
    }
    scan_best_warp(b2);
    i2 = b2.idx;
    if (!is_corner) {
      scan_best_warp(b3);
      i3 = b3.idx;
    }
  }
  if (lane == 0) {
    ind[qi * 3 + 0] = i1;
    ind[qi * 3 + 1] = i2;
    ind[qi * 3 + 2] = i3;
  }
}

template <bool DEVLOOP>
__global__ void __launch_bounds__(LM_THREADS)
odom_iterate_kernel(TreeView corner_tree, TreeView surf_tree, const float4* last_corner, const float4* last_surf,
                    const float4* queries, int n_sharp, int n_flat, int sharp_blocks, OdomIterArgs a_param, int* ind,
                    float* __restrict__ partials,
                    float* __restrict__ result, unsigned int* ticket, float4* __restrict__ dbg_coeff,
                    int8_t* __restrict__ dbg_sel, const OdomLmState* __restrict__ lm = nullptr,
                    ResultMailbox mb = ResultMailbox{nullptr, 0}) {
  __shared__ OdomIterArgs s_args;
  unsigned n_blocks = gridDim.x;
  if (DEVLOOP) {
    if (lm->h.done) return;
    if (threadIdx.x < (int)(sizeof(OdomIterArgs) / 4))
      reinterpret_cast<float*>(&s_args)[threadIdx.x] = reinterpret_cast<const float*>(&lm->args)[threadIdx.x];
    __syncthreads();
    const OdomLoopIo& io = lm->io;  // uniform loads; the grid of the loop graph is sized for a capacity
    n_blocks = (unsigned)io.n_blocks;
    if (blockIdx.x >= n_blocks) return;
    last_corner = io.last_corner; last_surf = io.last_surf; queries = io.queries;
    n_sharp = io.n_sharp; n_flat = io.n_flat; sharp_blocks = io.sharp_blocks; ind = io.ind;
  }
  const OdomIterArgs& a = DEVLOOP ? s_args : a_param;
  float acc[29];
#pragma unroll
  for (int k = 0; k < 29; k++) acc[k] = 0.f;

  const bool is_corner = (int)blockIdx.x < sharp_blocks;
  const int local = is_corner ? blockIdx.x * LM_THREADS + threadIdx.x
                              : (blockIdx.x - sharp_blocks) * LM_THREADS + threadIdx.x;
  const int qi = is_corner ? local : n_sharp + local;
  const bool active = is_corner ? (local < n_sharp) : (local < n_flat);
  if (active) {
    const float4 po = queries[qi];
    float sx, sy, sz;
    transform_to_start(a, po, sx, sy, sz);
    // correspondences of the last search iteration (odom_search_kernel)
    const int i1 = ind[qi * 3 + 0], i2 = ind[qi * 3 + 1], i3 = ind[qi * 3 + 2];

    float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
    bool sel = false;
    if (is_corner) {
      if (i2 >= 0) {
        const float4 t1 = last_corner[i1], t2 = last_corner[i2];
        float la, lb, lc, ld2;
        line_residual(sx, sy, sz, t1.x, t1.y, t1.z, t2.x, t2.y, t2.z, la, lb, lc, ld2);
        float s = 1.f;
        if (a.iter >= 5) s = 1.f - 1.8f * fabsf(ld2);
        coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
        sel = (double)s > 0.1 && ld2 != 0.f;
      }
    } else {
      if (i2 >= 0 && i3 >= 0) {
        const float4 t1 = last_surf[i1], t2 = last_surf[i2], t3 = last_surf[i3];
        float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
        float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
        float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
        float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
        const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
        pa /= ps; pb /= ps; pc /= ps; pd /= ps;
        const float pd2 = pa * sx + pb * sy + pc * sz + pd;
        float s = 1.f;
        if (a.iter >= 5) s = 1.f - 1.8f * fabsf(pd2) / sqrtf(sqrtf(sx * sx + sy * sy + sz * sz));
        coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
        sel = (double)s > 0.1 && pd2 != 0.f;
      }
    }
    if (dbg_coeff) {
      dbg_coeff[qi] = coeff;
      dbg_sel[qi] = sel ? 1 : 0;
    }
    if (sel) {
      const float px = po.x, py = po.y, pz = po.z;
      const float G1 = a.g1a * px + a.g1b * py + a.g1c * pz + a.k1 - a.k2 - a.k3;
      const float G2 = a.t1 * px - a.t2 * py + a.t3 * pz + a.k4 - a.k5 - a.k6;
      const float G3 = a.u1 * px - a.u2 * py - a.u3 * pz + a.k7 + a.k8 - a.k9;
      const float H1 = a.e1 * px + a.e2 * py - a.e3 * pz + a.tx * a.e4 + a.ty * a.e5 + a.k10;
      const float H2 = a.f1 * px + a.f2 * py - a.f3 * pz + a.k11 - a.ty * a.f2 - a.tx * a.f1;
      const float I1 = a.g1 * px + a.f1 * py + a.tx * a.f2 - a.ty * a.f1;
      const float I2 = a.h1 * px - a.h2 * py + a.k12 + a.k13;
      const float I3 = a.e2 * px + a.e4 * py + a.tx * a.e5 - a.ty * a.e4;
      float row[6];
      row[0] = G1 * coeff.x + G2 * coeff.y + G3 * coeff.z;
      row[1] = H1 * coeff.x + H2 * coeff.z;
      row[2] = I1 * coeff.x + I2 * coeff.y + I3 * coeff.z;
      row[3] = (-a.f1) * coeff.x + a.atx_y * coeff.y - a.e4 * coeff.z;
      row[4] = (-a.f2) * coeff.x - a.aty_y * coeff.y - a.e5 * coeff.z;
      row[5] = a.atz_x * coeff.x - a.atz_y * coeff.y - a.atz_z * coeff.z;
      const float b = (float)(-0.05 * (double)coeff.w);
      accumulate_row(acc, row, b, is_corner);
    }
  }
  reduce_normal_equations(acc, partials, result, ticket, mb, n_blocks);
}

// one warp: the 32 sums are staged through shared memory by one coalesced load, the warp solves (lmstep_warp.cuh)
__global__ void odom_lm_step_kernel(OdomLmState* st, const float* __restrict__ result, unsigned long long handle = 0ull,
                                    float* mailbox_host = nullptr) {
  __shared__ float s_r[NEQ];
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  s_r[threadIdx.x] = __ldcg(&result[threadIdx.x]);
  __syncwarp();
  if (!st->h.done) odom_lm_step_warp(st, s_r);  // uniform over the warp
  if (threadIdx.x == 0) lm_loop_control(st->h, handle, mailbox_host);
}

// BasicLaserOdometry::transformToEnd without IMU terms (BasicLaserOdometry.cpp:57-87): in place on a device cloud.
// sy,cy,... are the host-cached sin/cos of the full transform.
struct ToEndArgs {
  float rx, ry, rz, tx, ty, tz, inv_sp;
  float srx, crx, sry, cry, srz, crz;
};
// (p1, n1): an optional second cloud handled by the same launch (less sharp + less flat at the end of an odometry sweep)
__global__ void transform_to_end_kernel(float4* __restrict__ p0, int n0, ToEndArgs a, float4* __restrict__ p1 = nullptr, int n1 = 0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float4* __restrict__ p = p0;
  if (i >= n0) {
    i -= n0;
    p = p1;
    if (i >= n1) return;
  }
  float4 q = p[i];
  const float s = a.inv_sp * (q.w - (float)(int)q.w);
  float x = q.x - s * a.tx, y = q.y - s * a.ty, z = q.z - s * a.tz;
  q.w = (float)(int)q.w;
  float sx, cx, sy, cy, sz, cz;
  sincos_f(-s * a.rx, sx, cx);
  sincos_f(-s * a.ry, sy, cy);
  sincos_f(-s * a.rz, sz, cz);
  rot_zxy(x, y, z, sz, cz, sx, cx, sy, cy);
  rot_yxz(x, y, z, a.sry, a.cry, a.srx, a.crx, a.srz, a.crz);
  // += pos - imuShiftFromStart (zero without IMU); the two IMU rotations are identities (Angle(): cos 1, sin 0)
  q.x = x + a.tx;
  q.y = y + a.ty;
  q.z = z + a.tz;
  p[i] = q;
}

// Several device-to-device cloud copies in ONE launch (the hand-offs between the stage objects move 3-7 clouds per sweep;
// as separate cudaMemcpyAsync calls they cost more host time than the copies take on the GPU).
constexpr int GATHER_MAX_SEG = 8;
struct GatherSegs {
  const float4* src[GATHER_MAX_SEG];
  float4* dst[GATHER_MAX_SEG];
  int end[GATHER_MAX_SEG];  // exclusive prefix end of segment k in the flattened index space
  int n_seg;
};
__global__ void gather_segments_kernel(GatherSegs g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.end[g.n_seg - 1]) return;
  int k = 0;
#pragma unroll
  for (int s = 0; s < GATHER_MAX_SEG - 1; s++)
    if (s < g.n_seg - 1 && i >= g.end[s]) k = s + 1;
  const int local = i - (k > 0 ? g.end[k - 1] : 0);
  g.dst[k][local] = g.src[k][local];
}

__global__ void transform_to_map_kernel(float4* __restrict__ p, int n, MapIterArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 q = p[i];
  float x, y, z;
  associate_to_map(a, q, x, y, z);
  q.x = x; q.y = y; q.z = z;
  p[i] = q;
}

}  // namespace loamb
