#!/usr/bin/env python
"""bench.py -- sweeps/sec of the scan-to-map cycle (registration -> odometry -> mapping) on synthetic ring-ordered
sweeps, BASELINE.json's metric, at N GPUs of one node.

    python bench.py --gpus 1 --steps K --warmup W            # CUDA arm (this repo)
    python bench.py --impl reference --gpus 1 ...            # the reference's own CPU implementation, same workload

A "step" is one full pass of the hot path over one sweep.  The default workload is BASELINE config 3: HDL-64E
64 x 2048 sweeps against a 1 M-point surrounding map.  One JSON line is printed by rank 0 (contract in the task
statement): metric / value / e2e / roofline / cpu_baseline / clocks / gpu_launches.

Multi-GPU (see DESIGN.md "Multi-GPU"): the path shards by map cubes / queries with one 36-float all-reduce per LM
iteration; that mode is strong-scaling of a single stream and latency-bound, so the default N > 1 run is
"replicas": every rank registers its own independent sweep stream against its own map (weak scaling, no data-path
collective), which is how a fleet of sensors would use an 8-GPU box.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENTRY_BYTES = 16          # bytes of one cell-table entry (csrc/gridnn.cuh)
TRAFFIC_FILE = "r2_map_iterate_traffic.json"       # committed ncu DRAM bytes per launch, config 3
TRAFFIC_FILE_HBM = "r2_map_iterate_hbm_traffic.json"  # ... config 5

WORKLOADS = {
    # name: (lidar factory name, map points, description)
    "vlp16_200k": ("vlp16", 200_000, "VLP-16 16x1800 sweeps, 200k-pt map (BASELINE config 2)"),
    "hdl64_1m": ("hdl64", 1_000_000, "HDL-64E 64x2048 sweeps, 1M-pt surrounding map (BASELINE config 3)"),
    "hdl64_10m": ("hdl64", 10_000_000, "HDL-64E 64x2048 sweeps, 10M-pt map (BASELINE config 4)"),
    "dense128_20m": ("dense128", 20_000_000, "128-ring x 4096 dense sweeps, 20M-pt map (BASELINE config 5)"),
}


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""

    def __init__(self, gpu_index):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.gpu_index = gpu_index
        self.t = None

    def _run_nvml(self):
        """NVML polling (a few ms per sample): the timed region of this bench lasts tens of milliseconds."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self._stop.is_set():
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for nm, b in bits.items():
                if r & b:
                    self.reasons.add(nm)
            self._stop.wait(0.004)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu_index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def _run(self):
        # keep the sampler off the cores the stage threads work on: last core of the set this process may use
        try:
            cores = sorted(os.sched_getaffinity(0))
            if len(cores) > 8:
                os.sched_setaffinity(threading.get_native_id(), {cores[-1]})
        except Exception:
            pass
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def partition_cores(local_rank, world, gpu_cpus, allowed, first_sibling=lambda c: True):
    """Pure part of the CPU binding: gpu_cpus[i] = CPUs NVML recommends for GPU i, allowed = CPUs this process may use.
    Returns the cores for `local_rank`: physical cores (one hardware thread each) of its GPU's NUMA node, split evenly
    between the ranks whose GPUs share that node; None when fewer than 4 cores would be left (main thread spins on the
    result mailbox, two helper threads issue work)."""
    mine = gpu_cpus[local_rank]
    cores = sorted(mine & set(allowed)) or sorted(allowed)
    cores = [c for c in cores if first_sibling(c)] or cores
    peers = [r for r in range(world) if gpu_cpus[r] == mine] or [local_rank]
    k, idx = len(peers), peers.index(local_rank) if local_rank in peers else 0
    per = max(len(cores) // k, 1)
    chunk = cores[idx * per:(idx + 1) * per] or cores
    return chunk if len(chunk) >= 4 else None


def pin_to_gpu_cores(local_rank, world):
    """Bind this rank (and the library's helper threads it will spawn) to CPU cores on its GPU's NUMA node: the pipeline
    hand-shakes with the GPU dozens of times per sweep through mapped host memory, so cross-socket latency and ranks
    stacked on the same cores show up directly in sweeps/s.  Returns a short description for the JSON line (None when
    the platform offers no affinity information)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64

        def cpus_of(i):
            mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), words)
            return frozenset(c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1)

        def first_sibling(c):  # one hardware thread per core: the main thread spins, a sibling would starve
            try:
                with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as fh:
                    txt = fh.read().strip()
                sib = []
                for part in txt.split(","):
                    lo, _, hi = part.partition("-")
                    sib += list(range(int(lo), int(hi or lo) + 1))
                return c == min(sib)
            except Exception:
                return True

        n_local = max(world, local_rank + 1)
        chunk = partition_cores(local_rank, world, [cpus_of(i) for i in range(n_local)], os.sched_getaffinity(0),
                                first_sibling)
        if not chunk:
            return None
        os.sched_setaffinity(0, chunk)
        return f"{len(chunk)} cores of the GPU's NUMA node (cpus {chunk[0]}-{chunk[-1]})"
    except Exception:
        return None


def make_workload(name, n_sweeps, rank=0):
    from loam_velodyne_b200 import synth
    lidar_name, m, _ = WORKLOADS[name]
    scene = synth.make_scene()
    lidar = getattr(synth.Lidar, lidar_name)()
    corner, surf = synth.make_map(scene, m)
    # every rank drives its own stream: same world, different heading rate so the sweeps differ
    yaw_rate = math.radians(5.0 + 0.5 * rank)
    sweeps = [synth.make_sweep(scene, lidar, i, yaw_rate=yaw_rate) for i in range(n_sweeps)]
    return lidar, corner, surf, sweeps


def measure_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_cuda(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from loam_velodyne_b200 import api
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    api.set_device(local_rank)
    pinned_cores = pin_to_gpu_cores(local_rank, world)
    n_total = args.warmup + args.steps
    sharded = world > 1 and args.mode == "sharded"
    lidar, corner, surf, sweeps = make_workload(args.workload, n_total, 0 if sharded else rank)
    dev = f"cuda:{local_rank}"

    streams = 1 if sharded else world  # independent sweep streams processed by the job

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        el = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        return float(el.item())

    # every sweep of the workload in HBM (arm "value") and in pinned host memory (arm "e2e": the library just sees host
    # pointers; contract: "from pinned host memory")
    d_sweeps = [torch.from_numpy(p).to(dev) for p, _ in sweeps]
    pinned = [torch.from_numpy(p).pin_memory() for p, _ in sweeps]
    sweeps = [(pinned[i].numpy(), sweeps[i][1]) for i in range(len(sweeps))]
    torch.cuda.synchronize()

    def all_gather_bytes(b):
        mine = torch.tensor(list(b), dtype=torch.uint8, device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        return b"".join(bytes(t.cpu().tolist()) for t in every)

    # what the timed windows run on: (corner map, surface map, host sweeps, device sweeps, cube-sharded?)
    work = {"corner": corner, "surf": surf, "sweeps": sweeps, "d_sweeps": d_sweeps, "cube_sharded": sharded,
            "warmup": args.warmup, "steps": args.steps}

    def fresh_pipeline():
        p = api.Pipeline()
        if work["cube_sharded"]:
            # the MAP sharded by cube slabs over the ranks, all-reduce fused into the iteration kernel (NVLink peer memory)
            p.mapping.enable_cube_sharding(rank, world, all_gather_bytes(p.mapping.peer_export()), args.slab)
        p.seed_map(work["corner"], work["surf"])
        return p

    def window(streaming, device_input):
        """One timed window on a FRESH pipeline (every step must see new data against the map the previous steps left):
        W warm-up steps, then exactly K steps between two device-synchronised barriers; the pipeline is drained and every
        helper thread joined inside the timed region.  Returns (max-over-ranks seconds, poses of all W + K sweeps, stage
        seconds, iteration counts)."""
        pipe = fresh_pipeline()
        sweeps, d_sweeps = work["sweeps"], work["d_sweeps"]
        poses = []
        acc = {"stage": np.zeros(5), "it_o": 0, "it_m": 0}

        def run(lo, hi, timed):
            if streaming:
                got = 0
                for i in range(lo, hi):
                    if device_input:
                        pipe.submit(ring_sizes=sweeps[i][1], device_ptr=d_sweeps[i].data_ptr())
                    else:
                        pipe.submit(sweeps[i][0], sweeps[i][1])
                    r = pipe.collect(wait=False)
                    if r is not None:
                        poses.append(r)
                        got += 1
                while got < hi - lo:
                    poses.append(pipe.collect(wait=True))
                    got += 1
            else:
                for i in range(lo, hi):
                    if device_input:
                        ok, odom, aft, st = pipe.sweep_device(d_sweeps[i].data_ptr(), sweeps[i][1])
                    else:
                        ok, odom, aft, st = pipe.sweep(*sweeps[i])
                    poses.append((ok, odom, aft))
                    if timed:
                        acc["stage"] += st
                        acc["it_o"] += pipe.odom.last_iterations()
                        acc["it_m"] += pipe.mapping.last_iterations()
            pipe.sync()

        run(0, work["warmup"], False)
        if streaming:
            pipe.stage_seconds(reset=True)
        barrier()
        t0 = time.perf_counter()
        run(work["warmup"], work["warmup"] + work["steps"], True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if streaming:
            ss = pipe.stage_seconds()
            acc["stage"] = np.concatenate([ss["busy"], ss["idle"], ss["handoff"]])
        el = max_over_ranks(t1 - t0)
        if work["cube_sharded"]:
            pipe.mapping.disable_cube_sharding()  # unmap the peers' inboxes on every rank before any rank frees its own
        barrier()
        del pipe
        return el, poses, acc["stage"], acc["it_o"], acc["it_m"]

    def arm(streaming, device_input, min_seconds, max_windows):
        """Repeat the K-step window until the timed windows add up to `min_seconds` (the same count on every rank);
        report the median window."""
        first = window(streaming, device_input)
        n_win = int(min(max(math.ceil(min_seconds / max(first[0], 1e-6)), 3), max_windows))
        runs = [first] + [window(streaming, device_input) for _ in range(n_win - 1)]
        els = sorted(r[0] for r in runs)
        return {"seconds": els[len(els) // 2], "min": els[0], "max": els[-1], "windows": len(els), "runs": runs}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L = api.lib()
    # ---- "value": the three stages streaming over consecutive sweeps, inputs already resident in HBM
    a_dev = arm(True, True, args.min_seconds, args.max_windows)
    # ---- "e2e": the same through host buffers (H2D of every sweep and D2H of the poses inside the timed region)
    launches_before = L.loam_b200_total_launch_count()
    a_host = arm(True, False, args.min_seconds, args.max_windows)
    launches_e2e = (L.loam_b200_total_launch_count() - launches_before) / float(a_host["windows"])
    # ---- one sweep at a time (registration -> odometry -> mapping strictly in sequence): the per-sweep latency
    s_dev = arm(False, True, 0.0, 3)
    s_host = arm(False, False, 0.0, 3)
    # ---- the three classes used the way the reference's ROS adapters use them: every hand-off through host pcl clouds and
    # the reference's own entry points, every cloud the adapters publish downloaded (ScanRegistration.cpp:187-199,
    # LaserOdometry.cpp:286-330, LaserMapping.cpp:281-307)
    adapters = None
    if world == 1:
        best, d2h_pts = None, 0
        for _ in range(2):
            pipe = fresh_pipeline()
            for i in range(args.warmup):
                pipe.sweep(*work["sweeps"][i], mode="hostclouds")
                pipe.mapping.cloud("full")
            pipe.sync()
            t0 = time.perf_counter()
            d2h_pts = 0
            for i in range(args.warmup, n_total):
                pipe.sweep(*work["sweeps"][i], mode="hostclouds")
                d2h_pts += pipe.mapping.cloud("full").shape[0]  # the registered full-resolution cloud LaserMapping publishes
            pipe.sync()
            el = time.perf_counter() - t0
            best = el if best is None else min(best, el)
            n_feat = sum(pipe.scanreg.cloud(k).shape[0] for k in ("sharp", "less_sharp", "flat", "less_flat"))
            n_last = pipe.odom.cloud("last_corner").shape[0] + pipe.odom.cloud("last_surf").shape[0]
            del pipe
        n_full = int(work["sweeps"][0][0].shape[0])
        adapters = {"value": round(args.steps / best, 3), "unit": "sweeps/s", "ms_per_step": round(1e3 * best / args.steps, 4),
                    "what": "loam_b200_pipeline_sweep_hostclouds + the registered cloud: processScanlines(vector<PointCloud>), "
                            "host pcl clouds between the three classes, every published cloud downloaded",
                    "h2d_bytes_per_step": 16 * (2 * n_full + n_feat + n_full + n_last + n_full),
                    "d2h_bytes_per_step": 16 * (n_full + n_feat + n_full + n_last + n_full) + 96}
    clocks = sampler.stop() if rank == 0 else None
    value = streams * args.steps / a_dev["seconds"]
    e2e_value = streams * args.steps / a_host["seconds"]

    # all four arms compute the same trajectory, bit for bit
    ref_poses = a_dev["runs"][0][1]
    if rank == 0:
        for nm, a in (("e2e", a_host), ("sequential device", s_dev), ("sequential host", s_host)):
            for (ok0, od0, aft0), (ok1, od1, aft1) in zip(ref_poses, a["runs"][0][1]):
                if not (np.array_equal(od0, od1) and np.array_equal(aft0, aft1)):
                    raise SystemExit(f"arm '{nm}' disagrees with the streaming device-input arm: {aft0} vs {aft1}")

    # ---- N > 1: ONE stream against a map that is sharded over the GPUs (BASELINE configs 4 / 5): cube slabs + 2 m halo,
    # fused all-reduce of the normal equations over NVLink peer memory; reported next to the replica figure
    shard_report = None
    if world > 1 and not sharded and not args.no_sharded:
        try:
            shard_report = sharded_section(args, api, torch, dev, rank, world, work, arm, ref_sweeps=None)
        except Exception as e:  # a lost peer turns into an error on every rank: keep the replica figures of the line
            print(f"bench.py: rank {rank}: cube-sharded section failed: {e!r}", file=sys.stderr)
            shard_report = {"error": repr(e)[:300]}

    # ---- kernel-level pass (rank 0, N = 1 semantics): north-star kernel roofline through the kernel ABI
    out = None
    if rank == 0:
        roof, _ = kernel_roofline(args, api, corner, surf, sweeps[args.warmup], None)
        # kernels launched by libloam_b200.so during warm-up + timed steps of one e2e window, scaled to the timed steps
        launches = int(round(launches_e2e * args.steps / float(n_total)))
        n_pts = int(sweeps[0][0].shape[0])
        _, _, stage_dev, it_o, it_m = s_dev["runs"][0]
        stage_host = s_host["runs"][0][2]
        stage_names = ["registration", "odometry", "full_to_end", "mapping", "total"]
        # per step: the packed sweep up (+ ring table); down: the two poses of the sweep + per-loop header words + counts
        h2d = n_pts * 16 + 64 * 8
        d2h = 2 * 6 * 4 + 2 * 9 * 4 + 20 * 4
        ms = lambda x: round(1e3 * x, 3)
        out = {
            "metric": "sweeps/sec scan-to-map", "value": round(value, 3), "unit": "sweeps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * a_dev["seconds"] / args.steps, 4),
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][2], "sweep_points": n_pts, "cpu_binding": pinned_cores,
                       "map_points": int(corner.shape[0] + surf.shape[0]),
                       "mode": ("sharded: one stream, map sharded by cube slabs (+2 m halo) over the GPUs, all-reduce of AtA/AtB fused "
                                "into the iteration kernel over NVLink peer memory" if sharded
                                else "replicas: one independent sweep stream and map per GPU, no data-path collective"
                                if world > 1 else "single"),
                       "execution": "registration / odometry / mapping as three concurrent single-threaded stages over "
                                    "consecutive sweeps (how the reference's three ROS nodes run; loam_b200_pipeline_submit / "
                                    "_collect); the strictly sequential per-sweep figures are under 'sequential'",
                       "timing": {"window_steps": args.steps, "windows": a_dev["windows"],
                                  "timed_seconds_total": round(sum(r[0] for r in a_dev["runs"]), 3),
                                  "reported": "median window; every window runs on a fresh pipeline (new data every step)",
                                  "window_ms_min_median_max": [ms(a_dev["min"]), ms(a_dev["seconds"]), ms(a_dev["max"])],
                                  "e2e_window_ms_min_median_max": [ms(a_host["min"]), ms(a_host["seconds"]), ms(a_host["max"])]},
                       "streaming_stage_ms_per_sweep": {
                           k: [round(1e3 * float(v) / args.steps, 4) for v in a_dev["runs"][0][2][3 * j:3 * j + 3]]
                           for j, k in enumerate(["busy_reg_odom_map", "waiting_reg_odom_map", "handoff_reg_odom_map"])},
                       "sequential": {"value_device_input": round(streams * args.steps / s_dev["seconds"], 3),
                                      "value_host_input": round(streams * args.steps / s_host["seconds"], 3),
                                      "unit": "sweeps/s", "latency_ms_per_sweep": round(1e3 * s_dev["seconds"] / args.steps, 4),
                                      "stage_ms_device_input": {k: round(1e3 * v / args.steps, 4) for k, v in zip(stage_names, stage_dev)},
                                      "stage_ms_host_input": {k: round(1e3 * v / args.steps, 4) for k, v in zip(stage_names, stage_host)}},
                       "l2_note": ("inputs change every step (new sweep, map updated every sweep); map = %d MB points + 2x that in cell table: %s the 126 MB L2"
                                   % (int(corner.shape[0] + surf.shape[0]) * 16 // 1000000,
                                      "resident in" if corner.shape[0] + surf.shape[0] <= 2_000_000 else "larger than")),
                       "odom_iters_per_sweep": round(it_o / args.steps, 2), "map_iters_per_sweep": round(it_m / args.steps, 2)},
            "e2e": {"value": round(e2e_value, 3), "unit": "sweeps/s", "ms_per_step": round(1e3 * a_host["seconds"] / args.steps, 4),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        }
        if shard_report is not None:
            out["sharded"] = shard_report
        if adapters is not None:
            out["e2e_adapters"] = adapters
        if world == 1 and not args.no_hbm_roofline:
            out["roofline_hbm"] = hbm_roofline(args, api)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, corner, surf, sweeps, ref_poses)
            out["pose_delta_vs_reference"] = out["cpu_baseline"].pop("pose_delta_vs_reference")
    if world > 1:
        dist.destroy_process_group()
    return out


def sharded_section(args, api, torch, dev, rank, world, work, arm, ref_sweeps):
    """One sweep stream on `world` GPUs with the map sharded by cube slabs (DESIGN.md "Multi-GPU"): every rank feeds the
    same sweeps (registration and odometry are replicated), holds its slabs of the map, evaluates the queries
    that fall into them; the per-iteration all-reduce is fused into the iteration kernel.  Timed like the main arms.
    BASELINE config 4 (HDL-64 vs 10 M) at every N > 1; config 5 (128 x 4096 vs 20 M) in addition on 8 GPUs."""

    def one(workload, warmup, steps, with_sequential):
        lidar_name, m, desc = WORKLOADS[workload]
        _, corner, surf, sweeps = make_workload(workload, warmup + steps, 0)  # the same stream on every rank
        d_sweeps = [torch.from_numpy(p).to(dev) for p, _ in sweeps]
        pinned = [torch.from_numpy(p).pin_memory() for p, _ in sweeps]
        sweeps = [(pinned[i].numpy(), sweeps[i][1]) for i in range(len(sweeps))]
        saved = dict(work)
        work.update({"corner": corner, "surf": surf, "sweeps": sweeps, "d_sweeps": d_sweeps, "cube_sharded": True,
                     "warmup": warmup, "steps": steps})
        try:
            a_stream = arm(True, True, min(args.min_seconds, 0.2), 12 if with_sequential else 3)
            a_seq = arm(False, True, 0.0, 3) if with_sequential else None
        finally:
            work.clear()
            work.update(saved)
        if rank != 0:
            return None
        ms = lambda x: round(1e3 * x, 3)
        rep = {"workload": desc, "map_points": int(corner.shape[0] + surf.shape[0]), "sweep_points": int(sweeps[0][0].shape[0]),
               "n_gpus": world, "steps": steps, "warmup": warmup,
               "value": round(steps / a_stream["seconds"], 3), "unit": "sweeps/s", "scaling": "strong",
               "window_ms_min_median_max": [ms(a_stream["min"]), ms(a_stream["seconds"]), ms(a_stream["max"])],
               "windows": a_stream["windows"]}
        if a_seq is not None:
            last = a_seq["runs"][0]
            rep.update({"sequential_value": round(steps / a_seq["seconds"], 3),
                        "sequential_mapping_ms_per_sweep": round(1e3 * float(last[2][3]) / steps, 4),
                        "map_iters_per_sweep": round(last[4] / steps, 2)})
        return rep

    rep = one(args.sharded_workload, args.warmup, args.steps, True)
    rep5 = None
    if world >= 8 or args.sharded_config5:
        try:
            rep5 = one("dense128_20m", 3, min(args.steps, 8), False)
        except Exception as e:  # keep the rest of the line if the largest workload fails on this box
            rep5 = {"error": repr(e)[:300]}
    if rank != 0:
        return None
    rep["what"] = ("ONE sweep stream, map sharded by %d m cube slabs (+ 2 m halo) over the GPUs, queries evaluated by the owner "
                   "of their cell, all-reduce of the 32 normal-equation sums fused into the iteration kernel (peer stores over "
                   "NVLink, CUDA IPC inboxes); registration and odometry replicated" % args.slab)
    if rep5 is not None:
        rep["config5"] = rep5
    return rep


def kernel_roofline(args, api, corner, surf, sweep, pipe):
    """North-star kernel (fused 5-NN + fit + Jacobian + reduction = map_iterate_kernel) exactly as the pipeline launches it
    (map_iterate_kernel<false, MapCellLookup, false> on the mapping stage's own context and persistent map), timed with
    CUDA events on that context's stream over 50 launches (loam_b200_map_kernel_profile)."""
    pipe = api.Pipeline()
    pipe.seed_map(corner, surf)
    for _ in range(3):
        pipe.sweep(*sweep)
    prof = pipe.mapping.kernel_profile(50)
    nq, probes, cands = prof["queries"], prof["probes_per_query"] * prof["queries"], prof["candidates_per_query"] * prof["queries"]
    # algorithmic bytes per launch (DESIGN.md "Roofline"): every query reads itself (16 B), the cell-table entries it
    # probes and every candidate point of the occupied cells (16 B each); the 36-float result
    alg_bytes = nq * 16 + probes * ENTRY_BYTES + cands * 16 + 36 * 4
    dur_s = prof["avg_us"] * 1e-6
    peak, how = measure_peaks()
    achieved = alg_bytes / dur_s / 1e9
    del pipe
    # DRAM bytes per launch from the committed ncu --set full capture of this kernel on this workload (a number printed
    # under a profiler is never measured here); only quoted for the workload it was captured on
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
    if args.workload == "hdl64_1m" and os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
        traffic, traffic_src = int(tj["dram_bytes_per_launch"]), tj["source"]
    return ({"bound": "hbm", "kernel": "map_iterate_kernel<MapCellLookup> (fused fixed-radius 5-NN + line/plane fit + Jacobian + "
                                       "6x6 reduction), the instantiation the pipeline launches, on the mapping stage's own context",
             "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
             "traffic": traffic, "traffic_source": traffic_src, "peak_source": how, "algorithmic_bytes_per_launch": int(alg_bytes),
             "avg_launch_us": round(dur_s * 1e6, 2), "queries": int(nq), "table_probes_per_query": round(probes / max(nq, 1), 2),
             "candidate_points_per_query": round(cands / max(nq, 1), 2),
             "note": "1M-pt map + cell table fit in the 126 MB L2, so DRAM traffic is structurally far below the algorithmic "
                     "bytes and the fraction of the HBM roofline small; the HBM-bound regime of the same kernel is 'roofline_hbm'"},
            None)


def hbm_roofline(args, api):
    """The same kernel where it IS bound by HBM (the k-NN bandwidth stress of BASELINE config 5): a 20 M-point map held by
    the pipeline's persistent store (320 MB of points + cell table, far beyond the 126 MB L2) and 2 M queries spread over
    all of it, so every launch has to fetch its candidates from HBM.  The pipeline's own instantiation
    (map_iterate_kernel<MapCellLookup> on the mapping stage's context), CUDA-event time over 10 launches."""
    from loam_velodyne_b200 import synth
    scene = synth.make_scene()
    corner, surf = synth.make_map(scene, 20_000_000)
    lidar = synth.Lidar.hdl64()
    pipe = api.Pipeline()
    pipe.seed_map(corner, surf)
    pipe.sweep(*synth.make_sweep(scene, lidar, 0, yaw_rate=math.radians(5.0)))  # builds the store (sort + cell table)
    rng = np.random.RandomState(1)
    nq = 2_000_000
    q = surf[rng.randint(0, surf.shape[0], nq)].copy()
    q[:, :3] += rng.normal(0, 0.05, (nq, 3)).astype(np.float32)
    if os.environ.get("LOAM_B200_STRESS_SORTED"):  # development: queries in cell order (z, y, x)
        cell = np.floor(q[:, :3]).astype(np.int64)
        q = np.ascontiguousarray(q[np.lexsort((cell[:, 0], cell[:, 1], cell[:, 2]))])
    prof = pipe.mapping.kernel_profile_queries(q, 10)
    probes, cands = prof["probes_per_query"] * nq, prof["candidates_per_query"] * nq
    alg_bytes = nq * 16 + probes * ENTRY_BYTES + cands * 16 + 36 * 4
    dur_s = prof["avg_us"] * 1e-6
    peak, how = measure_peaks()
    del pipe
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", TRAFFIC_FILE_HBM)
    if os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
        traffic, traffic_src = int(tj["dram_bytes_per_launch"]), tj["source"]
    return {"bound": "hbm", "workload": "k-NN bandwidth stress (BASELINE config 5): 2 M queries spread over a 20 M-pt map held by the "
                                        "pipeline's persistent store",
            "kernel": "map_iterate_kernel<MapCellLookup>", "achieved": round(alg_bytes / dur_s / 1e9, 2), "peak": peak, "unit": "GB/s",
            "frac": round(alg_bytes / dur_s / 1e9 / peak, 5), "traffic": traffic, "traffic_source": traffic_src, "peak_source": how,
            "algorithmic_bytes_per_launch": int(alg_bytes), "avg_launch_us": round(dur_s * 1e6, 2), "queries": int(nq),
            "table_probes_per_query": round(prof["probes_per_query"], 2), "candidate_points_per_query": round(prof["candidates_per_query"], 2),
            "queries_per_second": round(nq / dur_s, 0), "map_points": int(corner.shape[0] + surf.shape[0])}


def cpu_baseline(args, corner, surf, sweeps, gpu_poses=None):
    """The reference's CPU path (compiled reference when oracle/_ref travelled, else the restatement), -O3 build,
    one thread (every reference node is single-threaded), on a bounded sample of the same workload."""
    from oracle import pydriver
    drv = pydriver.best(fast=True)
    pipe = drv.pipeline()
    pipe.seed_map(corner, surf)
    n = min(len(sweeps), 2 + args.cpu_sweeps)
    times = []
    stage = np.zeros(5)
    delta = 0.0
    for i in range(n):
        ok, odom_c, aft_c, st = pipe.sweep(*sweeps[i])
        if gpu_poses is not None:  # the same sweeps through the CUDA path: parity on the benchmarked configuration
            _, odom_g, aft_g = gpu_poses[i]
            delta = max(delta, float(np.abs(odom_g - odom_c).max()), float(np.abs(aft_g - aft_c).max()))
        if i >= 2:
            times.append(st[4])
            stage += st
    v = len(times) / sum(times)
    return {"pose_delta_vs_reference": {"max_abs": delta, "unit": "m and rad", "sweeps": n, "tolerance": 1e-4,
                                        "what": "max over the cpu_baseline sweeps of |pose_gpu - pose_reference| (odometry "
                                                "transformSum and mapped transformAftMapped, 6 components each)"},
            "value": round(v, 3), "unit": "sweeps/s", "cores": 1, "host_cores_available": os.cpu_count(),
            "kind": "reference" if drv.kind == "reference" else "port",
            "sample": f"{len(times)} sweeps of the same workload after 2 warm-up sweeps, wall time inside the C++ pipeline driver",
            "stage_ms": {k: round(1e3 * s / len(times), 2) for k, s in zip(["registration", "odometry", "full_to_end", "mapping", "total"], stage)}}


def reference_pipelined(drv, corner, surf, sweeps, warmup, steps):
    """The reference as it is deployed: scan registration, odometry and mapping are three single-threaded ROS nodes that
    work on consecutive sweeps at the same time.  Three threads, each owning one of the reference's classes, connected by
    queues that carry the clouds the topics carry (ScanRegistration.cpp:187-199, LaserOdometry.cpp:286-330); ctypes
    releases the GIL inside the C++ calls.  Returns (sweeps/s between the completion of sweep warmup-1 and the last
    sweep, final mapped pose)."""
    import queue
    reg, odo, mp = drv.scanreg(), drv.odom(), drv.mapping()
    mp.seed(corner, surf)
    q1, q2 = queue.Queue(maxsize=2), queue.Queue(maxsize=2)
    n_total = warmup + steps
    done_t = [0.0] * n_total
    errors = []

    def stage_reg():
        try:
            for i in range(n_total):
                reg.process(*sweeps[i])
                q1.put(tuple(reg.cloud(k) for k in ("sharp", "less_sharp", "flat", "less_flat", "full")))
        except Exception as e:  # pragma: no cover
            errors.append(e)
            q1.put(None)

    def stage_odom():
        try:
            for i in range(n_total):
                item = q1.get()
                if item is None:
                    q2.put(None)
                    return
                odo.set_inputs(*item)
                odo.process()
                odo.full_to_end()
                q2.put((odo.cloud("last_corner"), odo.cloud("last_surf"), odo.cloud("full"), odo.twist("sum")))
        except Exception as e:  # pragma: no cover
            errors.append(e)
            q2.put(None)

    def stage_map():
        try:
            for i in range(n_total):
                item = q2.get()
                if item is None:
                    return
                mp.set_inputs(item[0], item[1], item[2])
                mp.update_odometry(item[3])
                mp.process()
                done_t[i] = time.perf_counter()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=f) for f in (stage_reg, stage_odom, stage_map)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return steps / (done_t[n_total - 1] - done_t[warmup - 1]), mp.twist("aft")


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, with all the
    threads it can use: its three stages are single-threaded and run as three concurrent nodes (reference_pipelined);
    the strictly sequential single-core figure is reported next to it."""
    if rank != 0:
        return None
    from oracle import pydriver
    drv = pydriver.best(fast=True)
    n_total = args.warmup + args.steps
    lidar, corner, surf, sweeps = make_workload(args.workload, n_total, 0)
    v, aft_pipelined = reference_pipelined(drv, corner, surf, sweeps, args.warmup, args.steps)
    # sequential run of the same sweeps on one core (the cpu_baseline figure of the cuda arm)
    pipe = drv.pipeline()
    pipe.seed_map(corner, surf)
    n_seq = min(n_total, args.warmup + 10)
    for i in range(args.warmup):
        pipe.sweep(*sweeps[i])
    t0 = time.perf_counter()
    aft_seq = None
    for i in range(args.warmup, n_seq):
        _, _, aft_seq, _ = pipe.sweep(*sweeps[i])
    v_seq = (n_seq - args.warmup) / (time.perf_counter() - t0)
    kind = "reference" if drv.kind == "reference" else "port"
    return {"impl": "reference", "metric": "sweeps/sec scan-to-map", "value": round(v, 3), "unit": "sweeps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / v, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][2], "sweep_points": int(sweeps[0][0].shape[0]),
                       "map_points": int(corner.shape[0] + surf.shape[0]),
                       "mode": "three single-threaded stages (registration / odometry / mapping) pipelined over consecutive "
                               "sweeps, as the reference's three ROS nodes run"},
            "cpu_baseline": {"value": round(v, 3), "unit": "sweeps/s", "cores": 3, "kind": kind,
                             "sample": f"{args.steps} sweeps through three concurrent stage threads, {os.cpu_count()} host cores "
                                       f"present (the reference cannot use more: every node is single-threaded)",
                             "sequential_one_core": round(v_seq, 3)},
            "e2e": {"value": round(v, 3), "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def guarded_single_gpu_run():
    """N = 1 only: the measurement runs in a child process and is repeated ONCE if that process dies without a result.
    (One builder run of this bench ended in a segmentation fault a few seconds after start-up that neither a rerun under
    faulthandler nor 200 create / stream / destroy cycles of the pipeline reproduced; a lost bench line costs a round, a
    rerun costs a minute.  The JSON line says whether a rerun was needed: "bench_reruns".  A usage error is not retried.)"""
    env = dict(os.environ)
    env["LOAM_B200_BENCH_CHILD"] = "1"
    cmd = [sys.executable, "-X", "faulthandler", os.path.abspath(__file__)] + sys.argv[1:]
    for attempt in range(2):
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, text=True)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        for l in r.stdout.splitlines():
            if not l.startswith("{"):
                print(l, file=sys.stderr)
        if r.returncode == 0 and lines:
            out = json.loads(lines[-1])
            if isinstance(out, dict) and "metric" in out:
                out["bench_reruns"] = attempt
            print(json.dumps(out))
            return 0
        print(f"bench.py: measurement process ended with status {r.returncode} and no result (attempt {attempt + 1})", file=sys.stderr)
        if r.returncode == 2:  # argparse
            return 2
    return 1


def main():
    if (int(os.environ.get("WORLD_SIZE", "1")) == 1 and "LOAM_B200_BENCH_CHILD" not in os.environ
            and "--only-hbm" not in sys.argv and "LOAM_B200_BENCH_NO_GUARD" not in os.environ):
        sys.exit(guarded_single_gpu_run())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="hdl64_1m", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sweeps", type=int, default=12, help="sweeps timed for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slab", type=int, default=10, help="slab width in metres of the cube-sharded map (N > 1)")
    ap.add_argument("--sharded-workload", default="hdl64_10m", choices=sorted(WORKLOADS),
                    help="N > 1: workload of the additional single-stream run on the cube-sharded map")
    ap.add_argument("--only-hbm", action="store_true", help="development: print only the roofline_hbm object")
    ap.add_argument("--no-hbm-roofline", action="store_true", help="skip the config-5 (20 M-point map) kernel measurement")
    ap.add_argument("--sharded-config5", action="store_true",
                    help="N > 1: also run BASELINE config 5 (128x4096 vs 20M) on the cube-sharded map (default: only on 8 GPUs)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the cube-sharded single-stream section")
    ap.add_argument("--min-seconds", type=float, default=0.6,
                    help="repeat the K-step timed window (fresh pipeline each) until the windows add up to this")
    ap.add_argument("--max-windows", type=int, default=100)
    ap.add_argument("--mode", default="replicas", choices=["replicas", "sharded"],
                    help="N > 1: 'replicas' = one independent sweep stream per GPU (weak scaling, default); 'sharded' = one "
                         "stream, every rank evaluates a slice of the scan-to-map correspondences, NCCL all-reduce of the "
                         "normal equations per LM iteration (strong scaling)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank, world, local_rank = rank_world()
    if args.only_hbm:  # development aid: just the config-5 kernel measurement
        import __graft_entry__ as ge
        ge.build()
        from loam_velodyne_b200 import api
        print(json.dumps(hbm_roofline(args, api)))
        return
    if args.impl == "reference":
        if args.steps > 40:
            args.steps = 40  # bounded sample: ~0.3 s per sweep on one core
        out = run_reference(args, rank, world)
    else:
        import __graft_entry__ as ge
        if rank == 0:
            ge.build()
        out = run_cuda(args, rank, world, local_rank)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
