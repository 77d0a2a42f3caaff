"""Drives a few fused sweeps of the bench workload (for ncu launch lists / --set full captures)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loam_velodyne_b200 import api, synth

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    scene = synth.make_scene()
    lidar = synth.Lidar.hdl64()
    corner, surf = synth.make_map(scene, m)
    p = api.Pipeline()
    p.seed_map(corner, surf)
    for i in range(n):
        pts, rs = synth.make_sweep(scene, lidar, i, yaw_rate=math.radians(5.0))
        ok, od, aft, st = p.sweep(pts, rs)
        print(i, np.round(st * 1e3, 3), "iters", p.odom.last_iterations(), p.mapping.last_iterations(), p.mapping.last_phase_ms(), flush=True)

if __name__ == "__main__":
    main()
