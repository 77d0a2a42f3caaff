"""Regenerates tests/golden/*.npz from oracle/_ref (the UNMODIFIED reference sources compiled against oracle/shim).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures are small on purpose (VLP-16 at 600 azimuth steps, 40 k-point map) so they can live in git; they pin
 * the feature clouds of BasicScanRegistration::processScanlines,
 * the odometry / mapping poses of a 6-sweep registration -> odometry -> mapping run,
 * k-NN, VoxelGrid and the dense solves on fixed inputs,
 * the ring-binning front end (MultiScanRegistration::process) on a jittered, defect-laden raw cloud.
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from loam_velodyne_b200 import synth  # noqa: E402
from oracle import pydriver  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = pydriver.load("reference")
    assert ref.kind == "reference"
    scene = synth.make_scene()
    lidar = synth.Lidar(16, 600, -15.0, 15.0)
    corner, surf = synth.make_map(scene, 40_000)
    pipe = ref.pipeline()
    pipe.seed_map(corner, surf)
    out = {"map_corner": corner, "map_surf": surf}
    for i in range(6):
        pts, rs = synth.make_sweep(scene, lidar, i, yaw_rate=math.radians(5.0))
        ok, odom, aft, _ = pipe.sweep(pts, rs)
        out[f"pts{i}"] = pts
        out[f"rings{i}"] = rs
        out[f"odom{i}"] = odom
        out[f"aft{i}"] = aft
        if i in (0, 3):
            for name in ("sharp", "less_sharp", "flat", "less_flat"):
                out[f"{name}{i}"] = pipe.scanreg.cloud(name)
    out["final_corner_cubes"] = pipe.mapping.cloud("corner_cubes")
    out["final_surf_cubes_n"] = np.array([pipe.mapping.cloud("surf_cubes").shape[0]])
    np.savez_compressed(os.path.join(HERE, "pipeline_vlp16_600.npz"), **out)

    rng = np.random.RandomState(11)
    q = surf[rng.randint(0, surf.shape[0], 400)].copy()
    q[:, :3] += rng.normal(0, 0.3, (400, 3)).astype(np.float32)
    idx5, d5 = ref.knn(surf, q, 5)
    idx1, d1 = ref.knn(corner, q, 1)
    vox = ref.voxel_grid(out["pts2"], 0.4)
    A = rng.normal(size=(6, 6)).astype(np.float32)
    A = (A @ A.T + 6 * np.eye(6)).astype(np.float32)
    b = rng.normal(size=6).astype(np.float32)
    x = ref.qr_solve6(A, b)
    ev, V = ref.eig_sym(A)
    A3 = A[:3, :3].copy()
    ev3, V3 = ref.eig_sym(A3)
    P5 = (surf[1000:1005, :3] + rng.normal(0, 0.01, (5, 3))).astype(np.float32)
    x53 = ref.lsq53(P5)
    np.savez_compressed(os.path.join(HERE, "pieces.npz"), knn_pts=surf, knn_corner=corner, knn_q=q, knn_idx5=idx5,
                        knn_d5=d5, knn_idx1=idx1, knn_d1=d1, vox_in=out["pts2"], vox_out=vox, A=A, b=b, x=x, ev=ev,
                        V=V, ev3=ev3, V3=V3, P5=P5, x53=x53)
    # ring-binning front end: MultiScanRegistration::process (the reference's ROS adapter compiled against oracle/shim/ros)
    lid = synth.Lidar(16, 600, -15.0, 15.0)
    pts_f, rs_f = synth.make_sweep(scene, lid, 2, yaw_rate=math.radians(5.0))
    raw = synth.raw_cloud_from_sweep(pts_f, rs_f, n_bad=25, seed=9, elev_jitter_deg=0.2)
    ms = ref.multiscan(-15.0, 15.0, 16)
    binned, sizes = ms.process(raw)
    np.savez_compressed(os.path.join(HERE, "multiscan_vlp16_600.npz"), raw=raw, binned=binned, sizes=sizes,
                        sharp=ms.cloud("sharp"), flat=ms.cloud("flat"))
    print("golden fixtures written:", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
