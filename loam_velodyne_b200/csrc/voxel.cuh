// Device pcl::VoxelGrid<PointXYZI> (centroid of all fields per occupied voxel, output in ascending voxel index).
// Semantics follow PCL's published applyFilter (bbox -> min_b/div_b -> linear voxel index -> sort -> run means,
// "leaf size too small" guard returns the input); call sites in the reference: BasicLaserMapping.cpp:519-527
// (feature stacks, 0.2 / 0.4 m) and :580-588 (per-cube map maintenance).  Runs are averaged in (voxel, input order).
#pragma once

#include "lbvh.cuh"

namespace loamb {

__global__ void voxel_key_kernel(const float4* __restrict__ p, int n, float inv, int minb0, int minb1, int minb2,
                                 int div0, int div1, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = p[i];
  const int i0 = (int)(floorf(q.x * inv) - (float)minb0);
  const int i1 = (int)(floorf(q.y * inv) - (float)minb1);
  const int i2 = (int)(floorf(q.z * inv) - (float)minb2);
  keys[i] = (unsigned)(i0 + i1 * div0 + i2 * div0 * div1);
  vals[i] = i;
}

// heads[i] = 1 when sorted key i starts a run; per-block totals for the scan
constexpr int SCAN_BS = 1024;
__global__ void __launch_bounds__(SCAN_BS)
voxel_head_kernel(const unsigned* __restrict__ keys, int n, unsigned* __restrict__ pos, unsigned* __restrict__ block_sums) {
  __shared__ unsigned ws[32];
  const int i = blockIdx.x * SCAN_BS + threadIdx.x;
  const unsigned h = (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
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
  if (i < n) pos[i] = (incl - h) | (h << 31);  // exclusive rank inside the block + head flag
  if (threadIdx.x == SCAN_BS - 1) block_sums[blockIdx.x] = incl;
}

// one thread per run head: mean of x, y, z, intensity over the run, in sorted order
__global__ void voxel_centroid_kernel(const float4* __restrict__ p, const unsigned* __restrict__ keys,
                                      const int* __restrict__ vals, const unsigned* __restrict__ pos,
                                      const unsigned* __restrict__ block_off, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned pv = pos[i];
  if (!(pv >> 31)) return;
  const unsigned dst = (pv & 0x7fffffffu) + block_off[i / SCAN_BS];
  const unsigned k = keys[i];
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  int cnt = 0;
  for (int u = i; u < n && keys[u] == k; u++) {
    const float4 q = p[vals[u]];
    sx += q.x; sy += q.y; sz += q.z; si += q.w;
    cnt++;
  }
  const float fn = (float)cnt;
  out[dst] = make_float4(sx / fn, sy / fn, sz / fn, si / fn);
}

// d_in: n points on the device; d_out: capacity n.  *count receives the number of voxels (host sync inside).
inline int voxel_grid_device(loam_b200_ctx* c, const float4* d_in, int n, float leaf, float4* d_out, int* count) {
  *count = 0;
  if (n <= 0) return LOAM_B200_OK;
  SortScratch& s = c->sort;
  LB_CUDA(c, s.keys_a.reserve(n));
  LB_CUDA(c, s.keys_b.reserve(n));
  LB_CUDA(c, s.vals_a.reserve(n));
  LB_CUDA(c, s.vals_b.reserve(n));
  LB_CUDA(c, c->bbox.reserve(8));
  unsigned* bb = reinterpret_cast<unsigned*>(c->bbox.p);
  bbox_init_kernel<<<1, 32, 0, c->stream>>>(bb);
  LB_LAUNCH_CHECK(c);
  const int bbox_blocks = std::min((n + 255) / 256, c->sm_count * 8);
  bbox_kernel<<<bbox_blocks, 256, 0, c->stream>>>(d_in, n, bb);
  LB_LAUNCH_CHECK(c);
  unsigned hb[6];
  LB_CUDA(c, cudaMemcpyAsync(hb, bb, sizeof hb, cudaMemcpyDeviceToHost, c->stream));
  LB_CUDA(c, cudaStreamSynchronize(c->stream));
  float mn[3], mx[3];
  for (int a = 0; a < 3; a++) { mn[a] = dec_f(hb[a]); mx[a] = dec_f(hb[3 + a]); }
  const float inv = 1.0f / leaf;
  const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1,
                  dz = (long long)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > 2147483647ll) {
    // pcl: "Leaf size is too small for the input dataset" -> output = input
    LB_CUDA(c, cudaMemcpyAsync(d_out, d_in, (size_t)n * 16, cudaMemcpyDeviceToDevice, c->stream));
    *count = n;
    return LOAM_B200_OK;
  }
  const int minb0 = (int)floorf(mn[0] * inv), minb1 = (int)floorf(mn[1] * inv), minb2 = (int)floorf(mn[2] * inv);
  const int div0 = (int)floorf(mx[0] * inv) - minb0 + 1, div1 = (int)floorf(mx[1] * inv) - minb1 + 1;
  const int div2 = (int)floorf(mx[2] * inv) - minb2 + 1;
  voxel_key_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_in, n, inv, minb0, minb1, minb2, div0, div1, s.keys_a.p,
                                                           s.vals_a.p);
  LB_LAUNCH_CHECK(c);
  // only as many radix passes as the voxel index needs
  long long span = (long long)div0 * div1 * div2;
  int bits = 1;
  while ((1ll << bits) < span && bits < 32) bits++;
  int rc = LOAM_B200_OK;
  {
    const int n_tiles = (n + RS_TILE - 1) / RS_TILE;
    LB_CUDA(c, s.hist.reserve((size_t)256 * n_tiles));
    unsigned *ka = s.keys_a.p, *kb = s.keys_b.p;
    int *va = s.vals_a.p, *vb = s.vals_b.p;
    int passes = (bits + 7) / 8;
    if (passes & 1) passes++;
    for (int p = 0; p < passes; p++) {
      radix_hist_kernel<<<n_tiles, RS_THREADS, 0, c->stream>>>(ka, n, p * 8, s.hist.p, n_tiles);
      LB_LAUNCH_CHECK(c);
      radix_scan_kernel<<<1, 1024, 0, c->stream>>>(s.hist.p, 256 * n_tiles);
      LB_LAUNCH_CHECK(c);
      radix_scatter_kernel<<<n_tiles, RS_THREADS, 0, c->stream>>>(ka, va, n, p * 8, s.hist.p, n_tiles, kb, vb);
      LB_LAUNCH_CHECK(c);
      unsigned* tk = ka; ka = kb; kb = tk;
      int* tv = va; va = vb; vb = tv;
    }
  }
  if (rc) return rc;
  const int nb = (n + SCAN_BS - 1) / SCAN_BS;
  LB_CUDA(c, c->vox_key.reserve((size_t)n + nb + 8));
  unsigned* pos = c->vox_key.p;
  unsigned* bsum = c->vox_key.p + n;
  voxel_head_kernel<<<nb, SCAN_BS, 0, c->stream>>>(s.keys_a.p, n, pos, bsum);
  LB_LAUNCH_CHECK(c);
  radix_scan_kernel<<<1, 1024, 0, c->stream>>>(bsum, nb + 1);  // exclusive; entry nb receives the grand total
  LB_LAUNCH_CHECK(c);
  voxel_centroid_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_in, s.keys_a.p, s.vals_a.p, pos, bsum, n, d_out);
  LB_LAUNCH_CHECK(c);
  unsigned total = 0;
  LB_CUDA(c, cudaMemcpyAsync(&total, bsum + nb, sizeof total, cudaMemcpyDeviceToHost, c->stream));
  LB_CUDA(c, cudaStreamSynchronize(c->stream));
  *count = (int)total;
  return LOAM_B200_OK;
}

}  // namespace loamb
