"""Multi-GPU sharded stream (one process per GPU, launch with torchrun): every rank feeds the same sweeps.
--mode peer (default): the MAP is sharded by cube slabs with a 2 m halo, a rank evaluates the scan-to-map queries that fall
  into its slabs and the normal equations are all-reduced inside the iteration kernel over NVLink peer memory (CUDA IPC);
--mode nccl: replicated map, contiguous query slices, ncclAllReduce behind every iteration kernel (round-1 form).
--check also runs the unsharded pipeline on rank 0 and compares the trajectories."""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from loam_velodyne_b200 import api, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweeps", type=int, default=8)
    ap.add_argument("--map", type=int, default=200_000)
    ap.add_argument("--lidar", default="vlp16")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--mode", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--slab", type=int, default=10, help="slab width in metres (peer mode)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    api.set_device(local)
    nccl_id = None
    if a.mode == "nccl":
        nid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            nid = torch.tensor(list(api.nccl_unique_id()), dtype=torch.uint8, device=f"cuda:{local}")
        dist.broadcast(nid, 0)
        nccl_id = bytes(nid.cpu().tolist())

    scene = synth.make_scene()
    lidar = getattr(synth.Lidar, a.lidar)()
    corner, surf = synth.make_map(scene, a.map)
    sweeps = [synth.make_sweep(scene, lidar, i, yaw_rate=math.radians(5.0)) for i in range(a.sweeps)]
    p = api.Pipeline()
    if a.mode == "nccl":
        p.seed_map(corner, surf)
        p.mapping.enable_sharding(rank, world, nccl_id)
    else:
        mine = torch.tensor(list(p.mapping.peer_export()), dtype=torch.uint8, device=f"cuda:{local}")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        p.mapping.enable_cube_sharding(rank, world, b"".join(bytes(t.cpu().tolist()) for t in every), a.slab)
        p.seed_map(corner, surf)  # keeps the points of this rank's slabs (+ halo)
    ref = None
    if a.check and rank == 0:
        ref = api.Pipeline()
        ref.seed_map(corner, surf)
    worst = 0.0
    t0 = time.perf_counter()
    for pts, rs in sweeps:
        ok, od, aft, st = p.sweep(pts, rs)
        if ref is not None:
            _, od_r, aft_r, _ = ref.sweep(pts, rs)
            worst = max(worst, float(np.abs(aft - aft_r).max()), float(np.abs(od - od_r).max()))
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    # every rank must hold the same pose (identical solves on identical all-reduced sums)
    n_map = torch.tensor([float(p.mapping.cloud("corner_cubes").shape[0] + p.mapping.cloud("surf_cubes").shape[0])], device=f"cuda:{local}")
    sizes = [torch.zeros_like(n_map) for _ in range(world)]
    dist.all_gather(sizes, n_map)
    if rank == 0:
        print("map points held per rank:", [int(t.item()) for t in sizes])
    pose = torch.from_numpy(aft).cuda()
    poses = [torch.zeros_like(pose) for _ in range(world)]
    dist.all_gather(poses, pose)
    same = all(torch.equal(poses[0], q) for q in poses)
    if rank == 0:
        print(f"mode {a.mode} world {world} sweeps {a.sweeps} {el / a.sweeps * 1e3:.2f} ms/sweep identical_across_ranks {same} "
              f"max_pose_diff_vs_single_gpu {worst:.2e}")
        if a.check:
            assert same and worst <= 1e-4, (same, worst)
            print("SHARDED_OK")
    if a.mode == "peer":
        p.mapping.disable_cube_sharding()
        dist.barrier()
    del p
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
