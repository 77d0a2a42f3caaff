"""Host-side logic of the multi-GPU path on CPU: two gloo ranks each evaluate their query slice of one scan-to-map
iteration (through the oracle -- there is no GPU here), all-reduce the 32-float normal-equation message and must
reproduce the unsharded normal equations.  Exercises the slice arithmetic the CUDA path uses (api.shard_slice mirrors
map_iterate_impl) and the collective's message layout."""
import math
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import math, os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch
import torch.distributed as dist
from loam_velodyne_b200 import api, synth
from oracle import pydriver

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
orc = pydriver.load("restatement")
scene = synth.make_scene()
corner, surf = synth.make_map(scene, 60_000)
pts, rs = synth.make_sweep(scene, synth.Lidar(16, 900, -15.0, 15.0), 4, yaw_rate=math.radians(5.0))
s = orc.scanreg(); s.process(pts, rs)
cq = orc.voxel_grid(s.cloud("less_sharp"), 0.2)
sq = orc.voxel_grid(s.cloud("less_flat"), 0.4)
pos, yaw = synth.pose_at(0.5, np.array([0.0, 0.0, 1.0]), math.radians(5.0))
twist = np.array([0.0, yaw, 0.0, *pos], np.float32)
full = pydriver.map_iteration(orc, corner, surf, cq, sq, twist)
c0, c1 = api.shard_slice(cq.shape[0], rank, world)
s0, s1 = api.shard_slice(sq.shape[0], rank, world)
part = pydriver.map_iteration(orc, corner, surf, cq[c0:c1], sq[s0:s1], twist)
# 36-float message: 21 upper-triangle AtA + 6 AtB + n_selected, padded (SURVEY.md 8e)
msg = torch.zeros(36, dtype=torch.float32)
iu = np.triu_indices(6)
msg[:21] = torch.from_numpy(part["AtA"][iu])
msg[21:27] = torch.from_numpy(part["AtB"])
msg[27] = float(part["n_selected"])
dist.all_reduce(msg, op=dist.ReduceOp.SUM)
AtA = np.zeros((6, 6), np.float32); AtA[iu] = msg[:21].numpy(); AtA = AtA + AtA.T - np.diag(np.diag(AtA))
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
assert int(msg[27].item()) == full["n_selected"], (msg[27].item(), full["n_selected"])
assert rel(AtA, full["AtA"]) < 1e-5, rel(AtA, full["AtA"])
assert rel(msg[21:27].numpy(), full["AtB"]) < 1e-5
# every rank holds the same reduced message -> identical solves
gathered = [torch.zeros_like(msg) for _ in range(world)]
dist.all_gather(gathered, msg)
assert all(torch.equal(g, gathered[0]) for g in gathered)
# the slices tile the query range exactly
lo = [api.shard_slice(1001, r, world) for r in range(world)]
assert lo[0][0] == 0 and lo[-1][1] == 1001 and all(lo[i][1] == lo[i + 1][0] for i in range(world - 1))
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_normal_equations_gloo_world2(build_libs, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_shard_slice_properties():
    from loam_velodyne_b200 import api
    for n in (0, 1, 7, 1138, 16535):
        for world in (1, 2, 4, 8):
            sl = [api.shard_slice(n, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1


def test_bench_core_partition():
    """bench.py binds every rank to physical cores of its GPU's NUMA node, split between the ranks of that node."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    node0 = frozenset(list(range(0, 32)) + list(range(64, 96)))     # 32 cores x 2 hardware threads
    node1 = frozenset(list(range(32, 64)) + list(range(96, 128)))
    gpus = [node0] * 4 + [node1] * 4
    first = lambda c: c < 64                                       # cpu c + 64 is the sibling of cpu c
    allowed = set(range(128))
    seen = []
    for r in range(8):
        chunk = bench.partition_cores(r, 8, gpus, allowed, first)
        assert chunk is not None and len(chunk) == 8 and all(c < 64 for c in chunk)
        assert set(chunk) <= (node0 if r < 4 else node1)
        seen += chunk
    assert len(seen) == len(set(seen)) == 64                       # disjoint, every physical core used once
    assert bench.partition_cores(0, 1, gpus, allowed, first) == list(range(32))
    assert bench.partition_cores(0, 2, gpus, allowed, first) == list(range(16))
    assert bench.partition_cores(1, 2, gpus, allowed, first) == list(range(16, 32))
    # a container that only grants a few CPUs: no binding rather than squeezing the threads
    assert bench.partition_cores(0, 8, gpus, {0, 1, 2, 3, 4, 5, 6, 7}, first) is None
    # CPUs outside the GPU's node only: fall back to what is allowed
    assert bench.partition_cores(0, 1, gpus, set(range(32, 48)), first) == list(range(32, 48))


def test_slab_partition_covers_every_cell_once_and_halo_is_two_cells(build_libs):
    """Host logic of the cube-sharded map (csrc/shard.cuh): every 1 m cell has exactly one owner, owners cycle slab by slab,
    and a rank stores exactly its cells plus two halo cells on both sides of every slab."""
    from loam_velodyne_b200 import api
    for world in (1, 2, 3, 8):
        for slab in (1, 5, 10, 25):
            owners = [api.shard_owner(c, world, slab) for c in range(-130, 131)]
            assert set(owners) == set(range(world)) or 261 < slab * world
            for c, o in zip(range(-130, 131), owners):
                assert 0 <= o < world
                assert o == ((c // slab) % world)  # python floor division: slabs tile the negative axis too
            for rank in range(world):
                for c in range(-60, 61):
                    near = any(api.shard_owner(c + d, world, slab) == rank for d in (-2, -1, 0, 1, 2))
                    assert api.shard_stores(c + 0.5, rank, world, slab) == (near or world == 1)
    # a coordinate that is exactly a negative multiple of 50 below -25 is filed one cell lower (the reference's truncating
    # cube index, BasicLaserMapping.cpp:540-553): the partition follows the stored cell
    assert api.shard_stores(-75.0, api.shard_owner(-76, 8, 10), 8, 10)


def test_seed_points_partition_is_a_cover(build_libs):
    """Every seed point is stored by at least one rank and owned by exactly one (no map point is lost by sharding)."""
    from loam_velodyne_b200 import api
    rng = np.random.RandomState(4)
    xs = rng.uniform(-125, 125, 2000).astype(np.float32)
    for world, slab in ((2, 10), (8, 10), (8, 7)):
        stored = np.array([[api.shard_stores(x, r, world, slab) for r in range(world)] for x in xs])
        assert stored.any(axis=1).all()
        owners = np.array([api.shard_owner(int(np.floor(x)), world, slab) for x in xs])
        assert all(stored[i, owners[i]] for i in range(len(xs)))
        # storage overhead of the halo: (slab + 4) / slab on average
        assert abs(stored.sum() / len(xs) - min(world, (slab + 4) / slab)) < 0.15
