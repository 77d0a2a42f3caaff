// Morton-sorted linear BVH in HBM + exact k-NN walk.
// Replaces nanoflann's k-d tree (include/loam_velodyne/nanoflann.hpp:916-1043 build, :1354-1412 search) with the
// same observable semantics: exact k nearest neighbours, ascending squared distance, fp32 L2 accumulated
// x -> y -> z (L2_Simple_Adaptor::evalMetric, nanoflann.hpp:372-379), a candidate only enters when strictly
// closer than the current k-th (KNNResultSet::addPoint, :115-139; equal distances keep first-seen order).
//
// Build (per sweep, like the reference's per-sweep setInputCloud, BasicLaserMapping.cpp:636-637):
//   bbox reduce -> 30-bit Morton key per point -> LSD radix sort (key, index) -> gather points in Morton order
//   (xyz + original index in w) -> leaves of LEAF_SIZE consecutive points -> Karras topology over the leaves ->
//   bottom-up box refit.  Nodes are 64 B (both child boxes + links): one aligned 64 B read per visit.
#pragma once


#include "ctx.cuh"

namespace loamb {

constexpr int LEAF_SIZE = 8;
constexpr int KNN_MAX = 8;

// ---------------------------------------------------------------- bbox
__device__ __forceinline__ unsigned enc_f(float f) {  // order-preserving float -> uint
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec_f(unsigned u) {
  unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(v);
#else
  float f;
  memcpy(&f, &v, 4);
  return f;
#endif
}

__global__ void bbox_init_kernel(unsigned* bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

// n_dev (optional): the live point count sits in device memory (produced by a compaction earlier on the stream);
// m is then only the launch bound
__global__ void bbox_kernel(const float4* __restrict__ p, int m, unsigned* bb, const int* __restrict__ n_dev = nullptr) {
  if (n_dev) m = min(m, *n_dev);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const float4 q = p[i];
    mn[0] = fminf(mn[0], q.x); mx[0] = fmaxf(mx[0], q.x);
    mn[1] = fminf(mn[1], q.y); mx[1] = fmaxf(mx[1], q.y);
    mn[2] = fminf(mn[2], q.z); mx[2] = fmaxf(mx[2], q.z);
  }
  for (int a = 0; a < 3; a++) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&bb[a], enc_f(mn[a]));
      atomicMax(&bb[3 + a], enc_f(mx[a]));
    }
  }
}

__device__ __forceinline__ unsigned expand10(unsigned v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__device__ __forceinline__ unsigned morton_key(const float4& q, float lx, float ly, float lz, float hx, float hy, float hz) {
  const float ex = hx - lx, ey = hy - ly, ez = hz - lz;
  const float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-6f));
  const float sc = 1023.0f / ext;  // cubic cells
  const unsigned ix = (unsigned)fminf(fmaxf((q.x - lx) * sc, 0.f), 1023.f);
  const unsigned iy = (unsigned)fminf(fmaxf((q.y - ly) * sc, 0.f), 1023.f);
  const unsigned iz = (unsigned)fminf(fmaxf((q.z - lz) * sc, 0.f), 1023.f);
  return (expand10(ix) << 2) | (expand10(iy) << 1) | expand10(iz);
}

__global__ void morton_kernel(const float4* __restrict__ p, int m, const unsigned* __restrict__ bb,
                              unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  keys[i] = morton_key(p[i], dec_f(bb[0]), dec_f(bb[1]), dec_f(bb[2]), dec_f(bb[3]), dec_f(bb[4]), dec_f(bb[5]));
  vals[i] = i;
}

// ---------------------------------------------------------------- LSD radix sort, 8 bits per pass ("onesweep")
// One histogram launch for all passes, then ONE launch per 8-bit digit: a CTA takes the next tile (atomic ticket),
// ranks its keys per digit (stable: warp w owns consecutive keys, 32 at a time, match.any gives the rank among equal
// digits), publishes the tile's digit counts and resolves its global offsets by decoupled look-back over the earlier
// tiles' (aggregate | inclusive prefix) words -- no grid-wide barrier and no separate scan kernel.  The first version
// (histogram / scan / scatter per pass, later one cooperative kernel with 12 grid syncs) cost ~100 us per sort whatever
// the size and was half of the GPU time of a sweep (profiles/r1_v3_launch_list_summary.md).
// A pass whose digit is the same for every key (e.g. the top byte of a 22-bit cell key) degenerates to a tile copy.
constexpr int RS_THREADS = 256;
constexpr int RS_MAX_PASSES = 4;
constexpr unsigned RS_FLAG_AGG = 1u << 30, RS_FLAG_PREFIX = 2u << 30, RS_COUNT_MASK = (1u << 30) - 1u;
// scratch header (unsigned words): [RS_MAX_PASSES][256] global digit counts, then RS_MAX_PASSES tile tickets (+ pad)
constexpr int RS_HEADER = RS_MAX_PASSES * 256 + 8;

__host__ __device__ inline int rs_items_for(int m) { return m < 200000 ? 4 : 16; }

// histogram of every pass's digit in one read of the keys; also clears this tile's look-back words
template <int ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
onesweep_hist_kernel(const unsigned* __restrict__ keys, int m, const int* __restrict__ n_dev, int passes,
                     unsigned* __restrict__ header, unsigned* __restrict__ status, int n_tiles) {
  __shared__ unsigned h[RS_MAX_PASSES][256];
  if (n_dev) m = min(m, *n_dev);
  const int tile = blockIdx.x;
  for (int p = 0; p < passes; p++) {
    h[p][threadIdx.x] = 0;
    status[((size_t)p * n_tiles + tile) * 256 + threadIdx.x] = 0u;
  }
  __syncthreads();
  const int base = tile * (RS_THREADS * ITEMS);
  if (base >= m) return;
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const int i = base + it * RS_THREADS + threadIdx.x;
    if (i < m) {
      const unsigned k = keys[i];
      for (int p = 0; p < passes; p++) atomicAdd(&h[p][(k >> (8 * p)) & 255u], 1u);
    }
  }
  __syncthreads();
  for (int p = 0; p < passes; p++) {
    const unsigned v = h[p][threadIdx.x];
    if (v) atomicAdd(&header[p * 256 + threadIdx.x], v);
  }
}

struct OnesweepSmem {
  unsigned wcount[RS_THREADS / 32][256];
  unsigned gbase[256];
  unsigned dsum[RS_THREADS / 32];
  int tile;
  int uniform;
};

template <int ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
onesweep_pass_kernel(const unsigned* __restrict__ keys_in, const int* __restrict__ vals_in, int m,
                     const int* __restrict__ n_dev, int shift, const unsigned* __restrict__ digit_counts,
                     unsigned* status, unsigned* ticket, unsigned* __restrict__ keys_out, int* __restrict__ vals_out) {
  constexpr int NW = RS_THREADS / 32;
  constexpr int TILE = RS_THREADS * ITEMS;
  __shared__ OnesweepSmem sm;
  if (n_dev) m = min(m, *n_dev);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    sm.tile = (int)atomicAdd(ticket, 1u);
    sm.uniform = 0;
  }
  for (int d = lane; d < 256; d += 32) sm.wcount[warp][d] = 0;
  // global digit base = exclusive scan of the 256 digit counts (thread d <-> digit d)
  const unsigned dc = digit_counts[threadIdx.x];
  unsigned digit_excl;
  {
    unsigned x = dc;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) sm.dsum[warp] = x;
    __syncthreads();
    if (dc == (unsigned)m && m > 0) sm.uniform = 1;  // every key has this digit: the pass is the identity
    unsigned woff = 0;
    for (int w = 0; w < warp; w++) woff += sm.dsum[w];
    digit_excl = (x - dc) + woff;
  }
  __syncthreads();
  const int tile = sm.tile;
  if ((long long)tile * TILE >= m) return;
  if (sm.uniform) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const int i = tile * TILE + it * RS_THREADS + threadIdx.x;
      if (i < m) {
        keys_out[i] = keys_in[i];
        vals_out[i] = vals_in[i];
      }
    }
    return;
  }
  const int wbase = tile * TILE + warp * (ITEMS * 32);
  unsigned mykeys[ITEMS];
  int myvals[ITEMS];
  unsigned short myrank[ITEMS];
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const int i = wbase + it * 32 + lane;
    const bool valid = i < m;
    mykeys[it] = valid ? keys_in[i] : 0xffffffffu;
    myvals[it] = valid ? vals_in[i] : 0;
  }
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const bool valid = wbase + it * 32 + lane < m;
    const unsigned d = valid ? ((mykeys[it] >> shift) & 255u) : 256u;  // invalid lanes only match each other
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const unsigned before = __popc(peers & ((1u << lane) - 1u));
    unsigned prior = 0;
    if (valid) prior = sm.wcount[warp][d];
    myrank[it] = (unsigned short)(prior + before);
    __syncwarp();
    if (valid && before == 0) sm.wcount[warp][d] = prior + __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    // thread d: per-warp exclusive offsets of digit d inside the tile, the tile total, then the look-back
    const int d = threadIdx.x;
    unsigned acc = 0;
#pragma unroll
    for (int wv = 0; wv < NW; wv++) {
      const unsigned c = sm.wcount[wv][d];
      sm.wcount[wv][d] = acc;
      acc += c;
    }
    volatile unsigned* st = status;
    unsigned excl = 0;
    if (tile == 0) {
      st[d] = RS_FLAG_PREFIX | acc;
    } else {
      st[(size_t)tile * 256 + d] = RS_FLAG_AGG | acc;
      int t = tile - 1;
      bool done = false;
      while (!done) {
        // eight earlier tiles in flight; consume in order until an inclusive prefix (or a not-yet-published word)
        unsigned v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (t - u >= 0) ? st[(size_t)(t - u) * 256 + d] : (2u << 30);
#pragma unroll
        for (int u = 0; u < 8; u++) {
          if (done) break;
          const unsigned flag = v[u] & ~RS_COUNT_MASK;
          if (flag == 0u) break;  // not published yet: poll again from tile t
          excl += v[u] & RS_COUNT_MASK;
          t--;
          if (flag == RS_FLAG_PREFIX) done = true;
        }
      }
      st[(size_t)tile * 256 + d] = RS_FLAG_PREFIX | (excl + acc);
    }
    sm.gbase[d] = digit_excl + excl;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < ITEMS; it++) {
    const int i = wbase + it * 32 + lane;
    if (i < m) {
      const unsigned k = mykeys[it];
      const unsigned d = (k >> shift) & 255u;
      const unsigned dst = sm.gbase[d] + sm.wcount[warp][d] + myrank[it];
      keys_out[dst] = k;
      vals_out[dst] = myvals[it];
    }
  }
}

// exclusive scan of `total` counters by a single CTA (block sums of compactions / voxel heads)
__global__ void __launch_bounds__(1024) radix_scan_kernel(unsigned* __restrict__ hist, int total,
                                                           int* __restrict__ last_out = nullptr) {
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned v = i < total ? hist[i] : 0u;
    unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned w = warp_sums[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const unsigned woff = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u;
    const unsigned incl = x + woff + carry;
    if (i < total) hist[i] = incl - v;
    // callers scan n + 1 entries so that the last one receives the grand total; optionally mirror it
    if (last_out && i == total - 1) *last_out = (int)(incl - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
}

// ---------------------------------------------------------------- gather + leaves
__global__ void gather_sorted_kernel(const float4* __restrict__ pts, const int* __restrict__ order, int m,
                                     float4* __restrict__ sorted, const int* __restrict__ n_dev = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) m = min(m, *n_dev);
  if (i >= m) return;
  const int o = order[i];
  float4 p = pts[o];
  p.w = __int_as_float(o);
  sorted[i] = p;
}

// The three tree phases as per-element bodies: the stand-alone kernels below run one element per thread, the fused
// cluster build (clustersort.cuh) strides over them between cluster barriers.  Data produced by another CTA earlier in
// the same launch is read with ld.global.cg (L1 is not coherent across SMs).
__device__ __forceinline__ void leaf_body(int l, const float4* __restrict__ sorted, int m, int n_leaf,
                                          float4* __restrict__ box_lo, float4* __restrict__ box_hi,
                                          int* __restrict__ flags) {
  const int b = l * LEAF_SIZE, e = min(b + LEAF_SIZE, m);
  float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  for (int i = b; i < e; i++) {
    const float4 p = __ldcg(&sorted[i]);
    lo.x = fminf(lo.x, p.x); lo.y = fminf(lo.y, p.y); lo.z = fminf(lo.z, p.z);
    hi.x = fmaxf(hi.x, p.x); hi.y = fmaxf(hi.y, p.y); hi.z = fmaxf(hi.z, p.z);
  }
  // leaves live after the n_leaf-1 internal nodes in the box arrays
  box_lo[n_leaf - 1 + l] = make_float4(lo.x, lo.y, lo.z, 0.f);
  box_hi[n_leaf - 1 + l] = make_float4(hi.x, hi.y, hi.z, 0.f);
  if (l < n_leaf - 1) flags[l] = 0;
}

__global__ void leaf_kernel(const float4* __restrict__ sorted, const unsigned* __restrict__ keys, int m, int n_leaf,
                            unsigned* __restrict__ leaf_key, float4* __restrict__ box_lo, float4* __restrict__ box_hi,
                            int* __restrict__ flags) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_leaf) return;
  leaf_key[l] = keys[l * LEAF_SIZE];
  leaf_body(l, sorted, m, n_leaf, box_lo, box_hi, flags);
}

// ---------------------------------------------------------------- Karras (2012) topology over leaf keys
__device__ __forceinline__ int delta_fn(const unsigned* __restrict__ k, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned a = __ldcg(&k[i]), b = __ldcg(&k[j]);
  if (a == b) return 32 + __clz((unsigned)i ^ (unsigned)j);
  return __clz(a ^ b);
}

__device__ __forceinline__ void karras_body(int i, const unsigned* __restrict__ leaf_key, int n_leaf,
                                            BvhNode* __restrict__ nodes, int* __restrict__ parent) {
  const int d = (delta_fn(leaf_key, n_leaf, i, i + 1) - delta_fn(leaf_key, n_leaf, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta_fn(leaf_key, n_leaf, i, i - d);
  int lmax = 2;
  while (delta_fn(leaf_key, n_leaf, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta_fn(leaf_key, n_leaf, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta_fn(leaf_key, n_leaf, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (delta_fn(leaf_key, n_leaf, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int c0 = (lo == gamma) ? ~gamma : gamma;            // leaf link = ~leaf id
  const int c1 = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  nodes[i].lo0.w = __int_as_float(c0);
  nodes[i].hi0.w = __int_as_float(c1);
  parent[(c0 < 0) ? (n_leaf - 1 + ~c0) : c0] = i;
  parent[(c1 < 0) ? (n_leaf - 1 + ~c1) : c1] = i;
  if (i == 0) parent[0] = -1;
}

__global__ void karras_kernel(const unsigned* __restrict__ leaf_key, int n_leaf, BvhNode* __restrict__ nodes,
                              int* __restrict__ parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_leaf - 1) return;
  karras_body(i, leaf_key, n_leaf, nodes, parent);
}

// bottom-up refit: one thread per leaf climbs; the second arrival at a node merges the children
__device__ __forceinline__ void refit_body(int l, int n_leaf, BvhNode* __restrict__ nodes, const int* __restrict__ parent,
                                           float4* __restrict__ box_lo, float4* __restrict__ box_hi,
                                           int* __restrict__ flags) {
  int cur = n_leaf - 1 + l;
  int p = __ldcg(&parent[cur]);
  while (p >= 0) {
    __threadfence();
    if (atomicAdd(&flags[p], 1) == 0) return;  // first arrival: sibling not ready yet
    __threadfence();
    const int c0 = __float_as_int(__ldcg(&nodes[p].lo0.w)), c1 = __float_as_int(__ldcg(&nodes[p].hi0.w));
    const int b0 = (c0 < 0) ? (n_leaf - 1 + ~c0) : c0, b1 = (c1 < 0) ? (n_leaf - 1 + ~c1) : c1;
    const float4 l0 = __ldcg(&box_lo[b0]), h0 = __ldcg(&box_hi[b0]);
    const float4 l1 = __ldcg(&box_lo[b1]), h1 = __ldcg(&box_hi[b1]);
    nodes[p].lo0 = make_float4(l0.x, l0.y, l0.z, __int_as_float(c0));
    nodes[p].hi0 = make_float4(h0.x, h0.y, h0.z, __int_as_float(c1));
    nodes[p].lo1 = make_float4(l1.x, l1.y, l1.z, 0.f);
    nodes[p].hi1 = make_float4(h1.x, h1.y, h1.z, 0.f);
    __stcg(&box_lo[p], make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f));
    __stcg(&box_hi[p], make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f));
    cur = p;
    p = __ldcg(&parent[cur]);
  }
}

__global__ void refit_kernel(int n_leaf, BvhNode* __restrict__ nodes, const int* __restrict__ parent,
                             float4* __restrict__ box_lo, float4* __restrict__ box_hi, int* __restrict__ flags) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_leaf) return;
  refit_body(l, n_leaf, nodes, parent, box_lo, box_hi, flags);
}

// ---------------------------------------------------------------- k-NN walk (one query per thread)
struct TreeView {
  const BvhNode* nodes;
  const float4* sorted;
  int m, n_leaf, root;
};

template <int K>
struct KnnResult {
  float d2[K];
  float x[K], y[K], z[K];
  int idx[K];
};

__device__ __forceinline__ float box_d2(float qx, float qy, float qz, const float4& lo, const float4& hi) {
  const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
  return dx * dx + dy * dy + dz * dz;
}

// Exact K nearest with d2 < max_d2.  Results ascending; unfilled slots keep idx = -1, d2 = max_d2 sentinel.
// STATS = true additionally counts visited internal nodes / leaves (walk_stats[0], walk_stats[1]) for the roofline's
// algorithmic-bytes figure (Q * [16 + D*64 + V*L*16]).
template <int K, bool STATS = false>
__device__ __forceinline__ void knn_walk(const TreeView& t, float qx, float qy, float qz, float max_d2,
                                         KnnResult<K>& r, unsigned* walk_stats = nullptr) {
#pragma unroll
  for (int i = 0; i < K; i++) { r.d2[i] = max_d2; r.idx[i] = -1; r.x[i] = 0.f; r.y[i] = 0.f; r.z[i] = 0.f; }
  if (t.m <= 0) return;
  int stack_n[64];
  float stack_d[64];
  int sp = 0;
  int node = t.root;
  while (true) {
    if (node < 0) {
      if (STATS) walk_stats[1]++;
      const int leaf = ~node;
      const int b = leaf * LEAF_SIZE;
      const int cnt = min(LEAF_SIZE, t.m - b);
#pragma unroll
      for (int i = 0; i < LEAF_SIZE; i++) {
        if (i < cnt) {
          const float4 p = __ldg(&t.sorted[b + i]);
          const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < r.d2[K - 1]) {
            // insert after every entry with d2 <= d (first-seen order among equals)
            float cd = d, cx = p.x, cy = p.y, cz = p.z;
            int ci = __float_as_int(p.w);
#pragma unroll
            for (int s = 0; s < K; s++) {
              if (cd < r.d2[s]) {
                const float td = r.d2[s], tx = r.x[s], ty = r.y[s], tz = r.z[s];
                const int ti = r.idx[s];
                r.d2[s] = cd; r.x[s] = cx; r.y[s] = cy; r.z[s] = cz; r.idx[s] = ci;
                cd = td; cx = tx; cy = ty; cz = tz; ci = ti;
              }
            }
          }
        }
      }
    } else {
      if (STATS) walk_stats[0]++;
      const BvhNode* nd = t.nodes + node;
      const float4 lo0 = __ldg(&nd->lo0), hi0 = __ldg(&nd->hi0), lo1 = __ldg(&nd->lo1), hi1 = __ldg(&nd->hi1);
      const float d0 = box_d2(qx, qy, qz, lo0, hi0), d1 = box_d2(qx, qy, qz, lo1, hi1);
      const int c0 = __float_as_int(lo0.w), c1 = __float_as_int(hi0.w);
      const bool first0 = d0 <= d1;
      const float dn = first0 ? d0 : d1, df = first0 ? d1 : d0;
      const int cn = first0 ? c0 : c1, cf = first0 ? c1 : c0;
      const float worst = r.d2[K - 1];
      if (df < worst && sp < 64) { stack_n[sp] = cf; stack_d[sp] = df; sp++; }
      if (dn < worst) { node = cn; continue; }
    }
    // pop
    bool found = false;
    while (sp > 0) {
      sp--;
      if (stack_d[sp] < r.d2[K - 1]) { node = stack_n[sp]; found = true; break; }
    }
    if (!found) break;
  }
}

template <int K>
__global__ void knn_kernel(TreeView t, const float4* __restrict__ q, int nq, float max_d2, int* __restrict__ idx_out,
                           float* __restrict__ d2_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const float4 p = q[i];
  KnnResult<K> r;
  knn_walk<K>(t, p.x, p.y, p.z, max_d2, r);
#pragma unroll
  for (int j = 0; j < K; j++) {
    idx_out[i * K + j] = r.idx[j];
    d2_out[i * K + j] = r.idx[j] >= 0 ? r.d2[j] : FLT_MAX;
  }
}

}  // namespace loamb
