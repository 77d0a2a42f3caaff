"""Debug helper: many create / stream / sync / destroy cycles of the streaming pipeline (looks for rare teardown races)."""
import faulthandler, math, os, sys, time
faulthandler.enable(all_threads=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from loam_velodyne_b200 import api, synth

def main():
    cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    scene = synth.make_scene()
    lidar = synth.Lidar.hdl64() if len(sys.argv) > 2 and sys.argv[2] == "hdl64" else synth.Lidar.vlp16()
    corner, surf = synth.make_map(scene, 1_000_000 if lidar.n_rings == 64 else 200_000)
    sweeps = [synth.make_sweep(scene, lidar, i, yaw_rate=math.radians(5.0)) for i in range(12)]
    dev = [torch.from_numpy(p).cuda() for p, _ in sweeps]
    torch.cuda.current_stream().synchronize()
    t0 = time.time()
    for c in range(cycles):
        p = api.Pipeline()
        p.seed_map(corner, surf)
        res = p.run_stream(sweeps, device_ptrs=[t.data_ptr() for t in dev]) if c % 2 == 0 else p.run_stream(sweeps)
        assert len(res) == len(sweeps)
        p.stage_seconds(reset=True)
        torch.cuda.synchronize()
        del p
        if c % 25 == 0:
            print("cycle", c, round(time.time() - t0, 1), flush=True)
    print("STRESS_OK", cycles)

main()
