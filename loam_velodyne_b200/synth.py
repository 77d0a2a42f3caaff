"""Deterministic synthetic world + ring-ordered sweep generator (SURVEY.md §8d "Synthetic inputs").

Everything is expressed in LOAM's camera convention (x left, y up, z forward), i.e. *after* the axis swap of
MultiScanRegistration.cpp:182-184, and is delivered ring-ordered exactly like the input of
BasicScanRegistration::processScanlines (BasicScanRegistration.cpp:28): per ring, points in azimuth order with
``intensity = ring + relTime`` (MultiScanRegistration.cpp:228-229).

Scene: ground plane y = -1.8 m, a grid of box buildings (vertical walls in two orthogonal directions, convex
vertical edges), vertical poles (cylinders r = 0.15 m) along the streets, and a perimeter wall so every ray
returns (fixed-size sweeps).  The sensor moves with a constant twist; each ray is cast from the pose interpolated
at its own firing time, so sweeps carry the motion distortion LOAM de-skews.

The map sampler puts points on the lattices the reference's voxel filters converge to (surface 0.4 m, corner
0.2 m; BasicLaserMapping.cpp:98-99) with +-2 cm jitter and pads with horizontal slabs until the requested
size is reached (configs: 200 k / 1 M / 10 M / 20 M points).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

SENSOR_HEIGHT = 1.8


@dataclasses.dataclass
class Lidar:
    n_rings: int
    n_az: int
    lower_deg: float
    upper_deg: float

    @staticmethod
    def vlp16(n_az: int = 1800) -> "Lidar":
        return Lidar(16, n_az, -15.0, 15.0)  # MultiScanRegistration.h:83

    @staticmethod
    def hdl64(n_az: int = 2048) -> "Lidar":
        return Lidar(64, n_az, -24.9, 2.0)  # MultiScanRegistration.h:89

    @staticmethod
    def dense128(n_az: int = 4096) -> "Lidar":
        return Lidar(128, n_az, -25.0, 15.0)


@dataclasses.dataclass
class Scene:
    boxes: np.ndarray  # (B, 6) xmin, xmax, ymin, ymax, zmin, zmax (buildings, solid)
    poles: np.ndarray  # (P, 4) cx, cz, radius, top_y
    bound: float  # perimeter half-extent (inner faces of an enclosing wall)
    bound_top: float


def make_scene(seed: int = 1, extent: float = 110.0, block: float = 30.0, street: float = 20.0,
               n_poles: int = 64) -> Scene:
    rng = np.random.RandomState(seed)
    pitch = block + street
    boxes = []
    k = int(math.ceil(extent / pitch)) + 1
    for i in range(-k, k + 1):
        for j in range(-k, k + 1):
            # street centre lines run through x = 0 and z = 0; blocks sit between them
            cx = (i + 0.5) * pitch
            cz = (j + 0.5) * pitch
            hx = block / 2 * rng.uniform(0.7, 1.0)
            hz = block / 2 * rng.uniform(0.7, 1.0)
            h = rng.uniform(6.0, 18.0)
            if abs(cx) + hx > extent - 2 or abs(cz) + hz > extent - 2:
                continue
            boxes.append([cx - hx, cx + hx, -SENSOR_HEIGHT, h, cz - hz, cz + hz])
    boxes = np.asarray(boxes, dtype=np.float64)
    poles = []
    while len(poles) < n_poles:
        # poles on the pavement edges of the two streets through the origin and of parallel streets
        along = rng.uniform(-extent + 5, extent - 5)
        lane = rng.randint(-2, 3) * pitch + rng.choice([-1.0, 1.0]) * (street / 2 - 1.5)
        if rng.rand() < 0.5:
            cx, cz = lane, along
        else:
            cx, cz = along, lane
        inside = np.any((boxes[:, 0] - 0.5 < cx) & (cx < boxes[:, 1] + 0.5) &
                        (boxes[:, 4] - 0.5 < cz) & (cz < boxes[:, 5] + 0.5))
        if inside or math.hypot(cx, cz) < 3.0:
            continue
        poles.append([cx, cz, 0.15, rng.uniform(4.0, 8.0)])
    return Scene(boxes, np.asarray(poles, dtype=np.float64), extent, 60.0)


def _raycast(scene: Scene, o: np.ndarray, d: np.ndarray) -> np.ndarray:
    """Nearest hit distance for rays o + t d (o, d: (N, 3) float64, d unit)."""
    n = o.shape[0]
    t = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        # ground
        tg = (-SENSOR_HEIGHT - o[:, 1]) / d[:, 1]
        tg = np.where((d[:, 1] < 0) & (tg > 0), tg, np.inf)
        t = np.minimum(t, tg)
        inv = 1.0 / d
        # enclosing room (we are inside): exit distance through the slabs
        b = scene.bound
        lo = np.array([-b, -SENSOR_HEIGHT - 1.0, -b])
        hi = np.array([b, scene.bound_top, b])
        t1 = (lo - o) * inv
        t2 = (hi - o) * inv
        texit = np.nanmin(np.maximum(t1, t2), axis=1)
        t = np.minimum(t, np.where(texit > 0, texit, np.inf))
        # buildings: slab test per box
        for bx in scene.boxes:
            lo = bx[[0, 2, 4]]
            hi = bx[[1, 3, 5]]
            t1 = (lo - o) * inv
            t2 = (hi - o) * inv
            tn = np.nanmax(np.minimum(t1, t2), axis=1)
            tf = np.nanmin(np.maximum(t1, t2), axis=1)
            hit = (tn <= tf) & (tn > 1e-6)
            t = np.where(hit & (tn < t), tn, t)
        # poles: infinite cylinder in (x, z), clipped in y
        a = d[:, 0] ** 2 + d[:, 2] ** 2
        for cx, cz, r, top in scene.poles:
            ox = o[:, 0] - cx
            oz = o[:, 2] - cz
            bq = ox * d[:, 0] + oz * d[:, 2]
            cq = ox * ox + oz * oz - r * r
            disc = bq * bq - a * cq
            ok = disc > 0
            sq = np.sqrt(np.where(ok, disc, 0.0))
            tc = (-bq - sq) / a
            y = o[:, 1] + tc * d[:, 1]
            hit = ok & (tc > 1e-6) & (y >= -SENSOR_HEIGHT) & (y <= top)
            t = np.where(hit & (tc < t), tc, t)
    return t


def pose_at(t: float, v: np.ndarray, yaw_rate: float):
    """Constant-twist trajectory: body velocity v (m/s, LOAM frame) and yaw rate (rad/s) about +y.
    Returns (position (3,), yaw)."""
    yaw = yaw_rate * t
    if abs(yaw_rate) < 1e-12:
        p = v * t
    else:
        # integrate R_y(yaw(t)) v
        s, c = math.sin(yaw), math.cos(yaw)
        w = yaw_rate
        # R_y(a) [vx,0,vz] = [c vx + s vz, 0, -s vx + c vz]
        px = (s * v[0] + (1 - c) * v[2]) / w
        pz = ((c - 1) * v[0] + s * v[2]) / w
        p = np.array([px, v[1] * t, pz])
    return p, yaw


def make_sweep(scene: Scene, lidar: Lidar, sweep_idx: int, scan_period: float = 0.1,
               v=(0.0, 0.0, 1.0), yaw_rate: float = math.radians(5.0), noise_sigma: float = 0.01,
               max_range: float = 0.0):
    """One ring-ordered sweep.  Returns (pts (N,4) float32, ring_sizes (R,) int32).

    Rays that return nothing (only possible with max_range > 0) are dropped, giving ragged rings like a real
    sensor; with the default perimeter wall every ray returns and N = n_rings * n_az."""
    R, A = lidar.n_rings, lidar.n_az
    v = np.asarray(v, dtype=np.float64)
    elev = np.radians(np.linspace(lidar.lower_deg, lidar.upper_deg, R))
    frac = np.arange(A, dtype=np.float64) / A  # firing time fraction within the sweep
    theta = 2.0 * math.pi * frac  # ori - startOri  (MultiScanRegistration.cpp:206-228)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    # ori = -atan2(x, z)  ->  x = -sin(ori) cos(el), z = cos(ori) cos(el)
    dl = np.stack([-np.sin(theta)[None, :] * ce, np.broadcast_to(se, (R, A)), np.cos(theta)[None, :] * ce], axis=-1)
    t0 = sweep_idx * scan_period
    # per-azimuth pose (all rings fire together at an azimuth)
    pos = np.empty((A, 3))
    yaw = np.empty(A)
    for a in range(A):
        pos[a], yaw[a] = pose_at(t0 + frac[a] * scan_period, v, yaw_rate)
    cy, sy = np.cos(yaw)[None, :], np.sin(yaw)[None, :]
    dw = np.stack([cy * dl[..., 0] + sy * dl[..., 2], dl[..., 1], -sy * dl[..., 0] + cy * dl[..., 2]], axis=-1)
    o = np.broadcast_to(pos[None, :, :], (R, A, 3)).reshape(-1, 3)
    t = _raycast(scene, o, dw.reshape(-1, 3)).reshape(R, A)
    rng = np.random.RandomState(1000 + sweep_idx)
    t = t + rng.normal(0.0, noise_sigma, size=t.shape)
    valid = np.isfinite(t) & (t > 0.5)
    if max_range > 0:
        valid &= t < max_range
    p = (t[..., None] * dl).astype(np.float32)
    inten = (np.arange(R, dtype=np.float32)[:, None] +
             (np.float32(scan_period) * frac.astype(np.float32))[None, :]).astype(np.float32)
    pts = np.concatenate([p, inten[..., None]], axis=-1)
    ring_sizes = valid.sum(axis=1).astype(np.int32)
    pts = pts[valid]  # row-major boolean mask keeps ring-major, azimuth order
    return np.ascontiguousarray(pts, dtype=np.float32), ring_sizes


def raw_cloud_from_sweep(pts, ring_sizes, n_bad: int = 0, seed: int = 0, elev_jitter_deg: float = 0.0):
    """Unordered sensor-frame cloud (n, 3) as a spinning lidar driver delivers it (the input of
    MultiScanRegistration::process): the ring-ordered sweep of make_sweep() re-ordered by firing time (all rings of an
    azimuth together), axes swapped back to the sensor frame (loam x <- y, y <- z, z <- x), intensity dropped.
    n_bad NaN / inf / zero points are inserted at deterministic places; elev_jitter_deg tilts every point a little
    (a real sensor's beams are not exactly equally spaced)."""
    p = np.asarray(pts, np.float32)
    rel = p[:, 3] - np.floor(p[:, 3])
    order = np.argsort(rel, kind="stable")  # firing time; ties (same azimuth) keep ring order
    q = p[order]
    x, y, z = q[:, 0].astype(np.float64), q[:, 1].astype(np.float64), q[:, 2].astype(np.float64)
    rng = np.random.RandomState(seed)
    if elev_jitter_deg > 0:
        h = np.sqrt(x * x + z * z)
        el = np.arctan2(y, h) + np.radians(rng.uniform(-elev_jitter_deg, elev_jitter_deg, size=y.shape))
        r = np.sqrt(h * h + y * y)
        y = r * np.sin(el)
        s = r * np.cos(el) / np.maximum(h, 1e-9)
        x, z = x * s, z * s
    raw = np.stack([z, x, y], axis=1).astype(np.float32)  # sensor x = loam z, sensor y = loam x, sensor z = loam y
    if n_bad > 0 and raw.shape[0] > 10:
        where = np.sort(rng.choice(np.arange(1, raw.shape[0] - 1), size=n_bad, replace=False))
        bad = np.zeros((n_bad, 3), np.float32)
        kinds = rng.randint(0, 4, size=n_bad)
        bad[kinds == 0] = np.nan
        bad[kinds == 1, 0] = np.inf
        bad[kinds == 2] = 0.0
        bad[kinds == 3] = np.float32(0.004)  # |p|^2 = 4.8e-5 < 1e-4
        raw = np.insert(raw, where, bad, axis=0)
    return np.ascontiguousarray(raw, dtype=np.float32)


def _lattice_rect(origin, u, v, lu, lv, step, rng, jitter):
    nu = max(int(lu / step), 1)
    nv = max(int(lv / step), 1)
    a, b = np.meshgrid((np.arange(nu) + 0.5) * step, (np.arange(nv) + 0.5) * step, indexing="ij")
    p = origin[None, :] + a.reshape(-1, 1) * u[None, :] + b.reshape(-1, 1) * v[None, :]
    return p + rng.uniform(-jitter, jitter, size=p.shape)


def make_map(scene: Scene, n_target: int, seed: int = 7, surf_step: float = 0.4, corner_step: float = 0.2,
             jitter: float = 0.02, window: float = 120.0):
    """Map clouds (corner (Mc,4), surf (Ms,4) float32, map frame) sized so Mc + Ms ~= n_target.

    Corner points: vertical building edges + pole flanks at corner_step; surface points: ground, walls, roofs at
    surf_step; then horizontal slabs above the scene (never hit by rays, pure k-NN ballast) until n_target."""
    rng = np.random.RandomState(seed)
    surf = []
    corner = []
    ex, ey, ez = np.eye(3)
    g = scene.bound
    surf.append(_lattice_rect(np.array([-g, -SENSOR_HEIGHT, -g]), ex, ez, 2 * g, 2 * g, surf_step, rng, jitter))
    for bx in scene.boxes:
        x0, x1, y0, y1, z0, z1 = bx
        surf.append(_lattice_rect(np.array([x0, y0, z0]), ex, ey, x1 - x0, y1 - y0, surf_step, rng, jitter))
        surf.append(_lattice_rect(np.array([x0, y0, z1]), ex, ey, x1 - x0, y1 - y0, surf_step, rng, jitter))
        surf.append(_lattice_rect(np.array([x0, y0, z0]), ez, ey, z1 - z0, y1 - y0, surf_step, rng, jitter))
        surf.append(_lattice_rect(np.array([x1, y0, z0]), ez, ey, z1 - z0, y1 - y0, surf_step, rng, jitter))
        surf.append(_lattice_rect(np.array([x0, y1, z0]), ex, ez, x1 - x0, z1 - z0, surf_step, rng, jitter))
        for cx in (x0, x1):
            for cz in (z0, z1):
                ys = np.arange(y0 + corner_step / 2, y1, corner_step)
                c = np.stack([np.full_like(ys, cx), ys, np.full_like(ys, cz)], axis=1)
                corner.append(c + rng.uniform(-jitter, jitter, size=c.shape))
    for cx, cz, r, top in scene.poles:
        ys = np.arange(-SENSOR_HEIGHT + corner_step / 2, top, corner_step)
        for ang in (0.0, 0.5 * math.pi, math.pi, 1.5 * math.pi):
            c = np.stack([np.full_like(ys, cx + r * math.cos(ang)), ys, np.full_like(ys, cz + r * math.sin(ang))], axis=1)
            corner.append(c + rng.uniform(-jitter, jitter, size=c.shape))
    # perimeter walls
    for sx in (-g, g):
        surf.append(_lattice_rect(np.array([sx, -SENSOR_HEIGHT, -g]), ez, ey, 2 * g, scene.bound_top + SENSOR_HEIGHT,
                                  surf_step, rng, jitter))
        surf.append(_lattice_rect(np.array([-g, -SENSOR_HEIGHT, sx]), ex, ey, 2 * g, scene.bound_top + SENSOR_HEIGHT,
                                  surf_step, rng, jitter))
    surf = np.concatenate(surf)
    corner = np.concatenate(corner)
    have = surf.shape[0] + corner.shape[0]
    if have > n_target:
        # thin the far field first: keep everything near the origin, subsample the rest deterministically
        keep_c = min(corner.shape[0], max(n_target // 10, 11))
        keep_s = n_target - keep_c
        ds = np.linalg.norm(surf[:, [0, 2]], axis=1)
        dc = np.linalg.norm(corner[:, [0, 2]], axis=1)
        surf = surf[np.argsort(ds, kind="stable")[:keep_s]]
        corner = corner[np.argsort(dc, kind="stable")[:keep_c]]
    else:
        # ballast slabs above the scene (y from 70 m up, within the 5x5x5 cube window) until the target is met
        y = 70.0
        slabs = []
        missing = n_target - have
        per_slab = int((2 * window / surf_step) ** 2)
        while missing > 0:
            s = _lattice_rect(np.array([-window, y, -window]), ex, ez, 2 * window, 2 * window, surf_step, rng, jitter)
            if s.shape[0] > missing:
                s = s[:missing]
            slabs.append(s)
            missing -= s.shape[0]
            y += 0.8
            if y > 120.0:
                y = 70.4
                surf_step_next = surf_step  # keep lattice; offsets differ through jitter
        if slabs:
            surf = np.concatenate([surf] + slabs)
    def pack(p):
        out = np.zeros((p.shape[0], 4), dtype=np.float32)
        out[:, :3] = p.astype(np.float32)
        return out
    return pack(corner), pack(surf)
