"""Times the odometry correspondence search + iteration kernels through the kernel ABI (CUDA events)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loam_velodyne_b200 import api, synth


def main():
    scene = synth.make_scene()
    lidar = synth.Lidar.hdl64()
    ctx = api.Ctx(0)
    p0, r0 = synth.make_sweep(scene, lidar, 0, yaw_rate=math.radians(5.0))
    p1, r1 = synth.make_sweep(scene, lidar, 1, yaw_rate=math.radians(5.0))
    f0 = ctx.extract_features(p0, r0)
    last_c = p0[f0["less_sharp"]].copy()
    last_s = f0["less_flat_ds"].copy()
    last_c[:, 3] = np.floor(last_c[:, 3])
    last_s[:, 3] = np.floor(last_s[:, 3])
    f1 = ctx.extract_features(p1, r1)
    sharp, flat = p1[f1["sharp"]], p1[f1["flat"]]
    ctx.odom_set_last(last_c, last_s)
    ctx.odom_set_current(sharp, flat)
    tf = np.zeros(6, np.float32)
    print("last corner/surf", last_c.shape[0], last_s.shape[0], "queries", sharp.shape[0] + flat.shape[0])
    for it, label in ((0, "search + iterate"), (1, "iterate only")):
        for _ in range(5):
            ctx.odom_iterate(tf, it)
        ctx.profile(True)
        for _ in range(30):
            ctx.odom_iterate(tf, it)
        ms, n = ctx.profile_get()["odom_iter"]
        ctx.profile(False)
        print(f"{label}: {ms / n * 1e3:.1f} us")


if __name__ == "__main__":
    main()
