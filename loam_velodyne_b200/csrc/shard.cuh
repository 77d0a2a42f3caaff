// Cell of a stored map point and the multi-GPU slab ownership of cells (used by the scan-to-map kernel and the map store).
#pragma once

namespace loamb {

// Cell of a stored point.  The reference's cube index truncates (x + 25) / 50 toward zero and then decrements when
// x + 25 < 0 (:540-553), which puts a coordinate that is EXACTLY a negative multiple of 50 below -25 one cube lower
// than floor() would; such a point is filed under the cell below so that "cube = f(cell)" holds for every point.
__host__ __device__ __forceinline__ int store_cell(float x) {
  int c = (int)floorf(x);
  if ((float)c == x && c + 25 < 0 && (c + 25) % 50 == 0) c--;
  return c;
}

// ---- multi-GPU: the map sharded by slabs of 1 m cells along x ------------------------------------
// Rank r OWNS the cells whose x index c satisfies (floor(c / slab) mod world) == r and additionally STORES a halo of
// SHARD_HALO cells on both sides of every slab it owns.  A scan-to-map query is evaluated by the rank that owns the cell
// of its transformed position; the reference only accepts a correspondence whose 5th neighbour is closer than 1 m
// (BasicLaserMapping.cpp:671, :760), and the search looks at the 3 x 3 x 3 cells around the query, so the owner sees
// every candidate the single-GPU search sees: identical neighbours, identical Jacobian rows.  The halo is two cells
// wide because the end-of-sweep voxel filter works on voxels that straddle cell faces (leaf 0.2 / 0.4 m on a lattice
// offset by 25.5 m): a voxel cut by the outer edge of the stored region lies entirely in the second halo cell, so every
// voxel that reaches the first halo cell (the one queries can see) is complete and its centroid equals the unsharded one.
// slab == 0: no cube sharding (everything owned and stored).
constexpr int SHARD_HALO = 2;
struct ShardSpec {
  int rank, world, slab;
};
__host__ __device__ __forceinline__ int shard_owner_of(int cell_x, int slab, int world) {
  const int sidx = cell_x >= 0 ? cell_x / slab : -((-cell_x + slab - 1) / slab);
  const int r = sidx % world;
  return r < 0 ? r + world : r;
}
__host__ __device__ __forceinline__ bool shard_owns(const ShardSpec& sh, int cell_x) {
  return sh.slab <= 0 || sh.world <= 1 || shard_owner_of(cell_x, sh.slab, sh.world) == sh.rank;
}
__host__ __device__ __forceinline__ bool shard_stores(const ShardSpec& sh, int cell_x) {
  if (sh.slab <= 0 || sh.world <= 1) return true;
  for (int d = -SHARD_HALO; d <= SHARD_HALO; d++)
    if (shard_owner_of(cell_x + d, sh.slab, sh.world) == sh.rank) return true;
  return false;
}

}  // namespace loamb
