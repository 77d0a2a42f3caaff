"""k-NN bandwidth stress (BASELINE config 5 flavour): a map far larger than the 126 MB L2 and queries spread over all
of it, so the fixed-radius 5-NN of map_iterate_kernel has to come from HBM.  Prints the CUDA-event time of the fused
iteration kernel, its algorithmic bytes (queries + table probes + candidate points) and GB/s.

    python tools/knn_stress.py [map_points=20000000] [queries=2000000]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from loam_velodyne_b200 import api, synth


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    scene = synth.make_scene()
    corner, surf = synth.make_map(scene, m)
    rng = np.random.RandomState(1)
    q = surf[rng.randint(0, surf.shape[0], nq)].copy()
    q[:, :3] += rng.normal(0, 0.05, (nq, 3)).astype(np.float32)
    ctx = api.Ctx(0)
    ctx.tree_build(api.TREE_MAP_CORNER, corner)
    ctx.tree_build(api.TREE_MAP_SURF, surf)
    ctx.map_set_queries(np.zeros((0, 4), np.float32), q)
    twist = np.zeros(6, np.float32)
    ne, probes, cands = ctx.map_iterate_stats(twist)
    for _ in range(3):
        ctx.map_iterate(twist)
    ctx.profile(True)
    for _ in range(reps):
        ctx.map_iterate(twist)
    ms, n = ctx.profile_get()["map_iter"]
    ctx.profile(False)
    alg = nq * 16 + probes * 16 + cands * 16 + 36 * 4
    dur = ms / n * 1e-3
    print(f"map {surf.shape[0]} pts ({surf.shape[0] * 16 / 1e6:.0f} MB points + {2 ** int(np.ceil(np.log2(2 * surf.shape[0]))) * 16 / 1e6:.0f} MB table), "
          f"queries {nq}, selected {ne['n_selected']}")
    print(f"kernel {dur * 1e6:.1f} us  probes/query {probes / nq:.1f}  candidates/query {cands / nq:.1f}  "
          f"algorithmic {alg / 1e6:.1f} MB  -> {alg / dur / 1e9:.1f} GB/s  ({nq / dur / 1e6:.1f} M queries/s)")


if __name__ == "__main__":
    main()
