"""Times the small-cloud paths (voxel filter, LBVH build) through the kernel ABI with CUDA events.

    python tools/small_cloud_bench.py [n ...]        (LOAM_B200_LIB selects an A/B build)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from loam_velodyne_b200 import api, synth


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [5000, 20000, 60000]
    scene = synth.make_scene()
    pts, rs = synth.make_sweep(scene, synth.Lidar.hdl64(), 3)
    ctx = api.Ctx(0)
    rng = np.random.RandomState(0)
    for n in sizes:
        cloud = pts[rng.choice(pts.shape[0], n, replace=False)].copy()
        for _ in range(3):
            ctx.voxel_grid(cloud, 0.4)
            ctx.tree_build(api.TREE_ODOM_SURF, cloud)
        ctx.profile(True)
        for _ in range(20):
            out = ctx.voxel_grid(cloud, 0.4)
            ctx.tree_build(api.TREE_ODOM_SURF, cloud)
        prof = ctx.profile_get()
        ctx.profile(False)
        print(f"n={n}: voxel_grid {prof['voxel'][0] / prof['voxel'][1] * 1e3:.1f} us ({out.shape[0]} voxels), "
              f"tree_build {prof['tree_build'][0] / prof['tree_build'][1] * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
