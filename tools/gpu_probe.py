"""Exploratory GPU-vs-reference comparison (development aid; the real parity tests live in tests/)."""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from loam_velodyne_b200 import api, synth
from oracle import pydriver


def cmp_cloud(name, a, b):
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}")
        return False
    if a.size == 0:
        print(f"  {name}: both empty")
        return True
    eq = np.array_equal(a, b)
    md = np.abs(a - b).max()
    print(f"  {name}: n={a.shape[0]} bit-exact={eq} maxdiff={md:.3e}")
    return eq


def main():
    which = sys.argv[1:] or ["feat", "knn", "voxel", "pipe"]
    ref = pydriver.load("reference")
    scene = synth.make_scene()
    ctx = api.Ctx(0)
    if "feat" in which:
        for lidar, nm in ((synth.Lidar.vlp16(), "vlp16"), (synth.Lidar.hdl64(), "hdl64")):
            pts, rs = synth.make_sweep(scene, lidar, 3)
            t0 = time.time()
            f = ctx.extract_features(pts, rs)
            t1 = time.time()
            r = ref.scanreg()
            r.process(pts, rs)
            t2 = time.time()
            print(f"features {nm}: gpu {1e3*(t1-t0):.2f} ms  ref {1e3*(t2-t1):.2f} ms")
            cmp_cloud("sharp", pts[f["sharp"]], r.cloud("sharp"))
            cmp_cloud("less_sharp", pts[f["less_sharp"]], r.cloud("less_sharp"))
            cmp_cloud("flat", pts[f["flat"]], r.cloud("flat"))
            cmp_cloud("less_flat", f["less_flat_ds"], r.cloud("less_flat"))
    if "knn" in which:
        corner, surf = synth.make_map(scene, 200_000)
        rng = np.random.RandomState(5)
        q = surf[rng.randint(0, surf.shape[0], 20000)].copy()
        q[:, :3] += rng.normal(0, 0.3, (q.shape[0], 3)).astype(np.float32)
        ctx.tree_build(api.TREE_MAP_SURF, surf)
        t0 = time.time()
        gi, gd = ctx.tree_knn(api.TREE_MAP_SURF, q, 5)
        t1 = time.time()
        ri, rd = ref.knn(surf, q, 5)
        t2 = time.time()
        print(f"knn 5 of {surf.shape[0]}: gpu {1e3*(t1-t0):.2f} ms ref {1e3*(t2-t1):.2f} ms; idx equal {np.array_equal(gi, ri)} "
              f"d2 equal {np.array_equal(gd, rd)} mismatching rows {(gi != ri).any(axis=1).sum()}")
        gi, gd = ctx.tree_knn(api.TREE_MAP_SURF, q, 1)
        ri, rd = ref.knn(surf, q, 1)
        print(f"knn 1: idx equal {np.array_equal(gi, ri)} d2 equal {np.array_equal(gd, rd)}")
    if "voxel" in which:
        pts, rs = synth.make_sweep(scene, synth.Lidar.hdl64(), 1)
        for leaf in (0.2, 0.4):
            g = ctx.voxel_grid(pts, leaf)
            r = ref.voxel_grid(pts, leaf)
            print(f"voxel leaf {leaf}: gpu {g.shape} ref {r.shape}", end=" ")
            if g.shape == r.shape:
                print("maxdiff", np.abs(g - r).max(), "exact", np.array_equal(g, r))
            else:
                print()
    if "pipe" in which:
        lidar = synth.Lidar.vlp16()
        corner, surf = synth.make_map(scene, 200_000)
        pr = ref.pipeline()
        pr.seed_map(corner, surf)
        pg = api.Pipeline()
        pg.seed_map(corner, surf)
        v = (0, 0, 1.0)
        yr = math.radians(5)
        for i in range(10):
            pts, rs = synth.make_sweep(scene, lidar, i, v=v, yaw_rate=yr)
            ok_r, od_r, aft_r, st_r = pr.sweep(pts, rs)
            ok_g, od_g, aft_g, st_g = pg.sweep(pts, rs)
            print(f"sweep {i}: odom diff {np.abs(od_r-od_g).max():.2e} aft diff {np.abs(aft_r-aft_g).max():.2e} "
                  f"ref ms {np.round(st_r*1e3,1)} gpu ms {np.round(st_g*1e3,1)} iters o{pg.odom.last_iterations()} m{pg.mapping.last_iterations()}")
            if i in (0, 1, 5):
                cmp_cloud("last_corner", pg.odom.cloud("last_corner"), pr.odom.cloud("last_corner"))
                cmp_cloud("last_surf", pg.odom.cloud("last_surf"), pr.odom.cloud("last_surf"))
                cmp_cloud("corner_stack_ds", pg.mapping.cloud("corner_stack_ds"), pr.mapping.cloud("corner_stack_ds"))
                cmp_cloud("surf_stack_ds", pg.mapping.cloud("surf_stack_ds"), pr.mapping.cloud("surf_stack_ds"))
        print("aft ref", aft_r, "\naft gpu", aft_g)


if __name__ == "__main__":
    main()
