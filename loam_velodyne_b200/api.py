"""ctypes binding of libloam_b200.so (include/loam_b200.h and include/loam_b200_host.h).

Two layers, both thin:
  * ``Ctx``           -- the kernel-level C ABI (feature extraction, BVH k-NN, LM iterations, voxel grid, transforms)
  * ``ScanRegistration`` / ``LaserOdometry`` / ``LaserMapping`` / ``Pipeline`` -- handles over the C++ drop-in classes
    loam::BasicScanRegistration / BasicLaserOdometry / BasicLaserMapping (same method names as the reference classes).

There is no CPU implementation behind any of this: without the built library the import fails, and without a B200 every
compute call raises ``LoamB200Error``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LOAM_B200_LIB: development aid (A/B builds of the same ABI); the product library sits next to this file
LIB_PATH = os.environ.get("LOAM_B200_LIB") or os.path.join(HERE, "libloam_b200.so")

_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int)
_B = C.POINTER(C.c_int8)
_D = C.POINTER(C.c_double)


class LoamB200Error(RuntimeError):
    pass


class RegParams(C.Structure):
    _fields_ = [("nFeatureRegions", C.c_int), ("curvatureRegion", C.c_int), ("maxCornerSharp", C.c_int),
                ("maxCornerLessSharp", C.c_int), ("maxSurfaceFlat", C.c_int), ("lessFlatFilterSize", C.c_float),
                ("surfaceCurvatureThreshold", C.c_float)]

    @staticmethod
    def default():
        return RegParams(6, 5, 2, 20, 4, np.float32(0.2), np.float32(0.1))


class Features(C.Structure):
    _fields_ = [("sharp_idx", _I), ("sharp_cap", C.c_int), ("n_sharp", C.c_int),
                ("less_sharp_idx", _I), ("less_sharp_cap", C.c_int), ("n_less_sharp", C.c_int),
                ("flat_idx", _I), ("flat_cap", C.c_int), ("n_flat", C.c_int),
                ("label", _B),
                ("less_flat_ds", _F), ("less_flat_cap", C.c_int), ("n_less_flat", C.c_int)]


class Pose(C.Structure):
    _fields_ = [("rot", C.c_float * 3), ("sin_", C.c_float * 3), ("cos_", C.c_float * 3), ("pos", C.c_float * 3)]


class OdomPose(C.Structure):
    _fields_ = [("rot", C.c_float * 3), ("sin_", C.c_float * 3), ("cos_", C.c_float * 3), ("pos", C.c_float * 3),
                ("inv_scan_period", C.c_float), ("iter", C.c_int)]


class LmResult(C.Structure):
    _fields_ = [("rot", C.c_float * 3), ("pos", C.c_float * 3), ("iterations", C.c_int), ("converged", C.c_int)]


class NormalEq(C.Structure):
    _fields_ = [("AtA", C.c_float * 36), ("AtB", C.c_float * 6), ("n_selected", C.c_int),
                ("n_corner_selected", C.c_int)]


class MapWindow(C.Structure):
    _fields_ = [("cen", C.c_int * 3), ("valid_cubes", _I), ("n_valid", C.c_int), ("corner_leaf", C.c_float),
                ("surf_leaf", C.c_float)]


TREE_ODOM_CORNER, TREE_ODOM_SURF, TREE_MAP_CORNER, TREE_MAP_SURF = range(4)
K_FEATURES, K_TREE_BUILD, K_KNN, K_MAP_ITER, K_ODOM_ITER, K_TRANSFORM, K_VOXEL = range(7)
KERNEL_FAMILIES = ["features", "tree_build", "knn", "map_iter", "odom_iter", "transform", "voxel"]

_lib = None


def lib():
    """Load libloam_b200.so (built by ``__graft_entry__.build()`` / ``make -C loam_velodyne_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LoamB200Error(f"{LIB_PATH} is missing: build it with `make -C loam_velodyne_b200/csrc` "
                            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "loam_b200_strerror": (C.c_char_p, [C.c_int]),
        "loam_b200_last_error": (C.c_char_p, [vp]),
        "loam_b200_version": (C.c_int, []),
        "loam_b200_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "loam_b200_destroy": (C.c_int, [vp]),
        "loam_b200_sync": (C.c_int, [vp]),
        "loam_b200_bind_thread": (C.c_int, [C.c_int]),
        "loam_b200_set_priority": (C.c_int, [vp, C.c_int]),
        "loam_b200_allow_async_capture": (C.c_int, [vp, C.c_int]),
        "loam_b200_stream": (vp, [vp]),
        "loam_b200_extract_features": (C.c_int, [vp, _F, C.c_int, _I, _I, C.c_int, C.POINTER(RegParams),
                                                 C.POINTER(Features)]),
        "loam_b200_tree_build": (C.c_int, [vp, C.c_int, _F, C.c_int]),
        "loam_b200_tree_size": (C.c_int, [vp, C.c_int]),
        "loam_b200_tree_knn": (C.c_int, [vp, C.c_int, _F, C.c_int, C.c_int, C.c_float, _I, _F]),
        "loam_b200_map_set_queries": (C.c_int, [vp, _F, C.c_int, _F, C.c_int]),
        "loam_b200_map_iterate": (C.c_int, [vp, C.POINTER(Pose), C.POINTER(NormalEq)]),
        "loam_b200_map_iterate_debug": (C.c_int, [vp, C.POINTER(Pose), C.POINTER(NormalEq), _F, _B]),
        "loam_b200_map_iterate_stats": (C.c_int, [vp, C.POINTER(Pose), C.POINTER(NormalEq), C.POINTER(C.c_ulonglong),
                                          C.POINTER(C.c_ulonglong)]),
        "loam_b200_odom_set_last": (C.c_int, [vp, _F, C.c_int, _F, C.c_int]),
        "loam_b200_odom_set_current": (C.c_int, [vp, _F, C.c_int, _F, C.c_int]),
        "loam_b200_odom_solve": (C.c_int, [vp, _F, _F, C.c_float, C.c_int, C.c_float, C.c_float, C.POINTER(LmResult)]),
        "loam_b200_map_solve": (C.c_int, [vp, _F, _F, C.c_int, C.c_float, C.c_float, C.POINTER(LmResult)]),
        "loam_b200_odom_iterate": (C.c_int, [vp, C.POINTER(OdomPose), C.POINTER(NormalEq)]),
        "loam_b200_odom_iterate_debug": (C.c_int, [vp, C.POINTER(OdomPose), C.POINTER(NormalEq), _F, _B, _I]),
        "loam_b200_debug_gn_solve": (C.c_int, [vp, _F, C.c_int, C.c_int, C.c_float, _F, _I]),
        "loam_b200_transform_to_end": (C.c_int, [vp, _F, C.c_int, C.POINTER(OdomPose)]),
        "loam_b200_transform_to_map": (C.c_int, [vp, _F, C.c_int, C.POINTER(Pose)]),
        "loam_b200_voxel_grid": (C.c_int, [vp, _F, C.c_int, C.c_float, _F, C.c_int, _I]),
        "loam_b200_cloud_upload": (C.c_int, [vp, C.c_int, _F, C.c_int]),
        "loam_b200_cloud_upload_device": (C.c_int, [vp, C.c_int, vp, C.c_int]),
        "loam_b200_cloud_download": (C.c_int, [vp, C.c_int, _F, C.c_int, _I]),
        "loam_b200_cloud_size": (C.c_int, [vp, C.c_int]),
        "loam_b200_cloud_swap": (C.c_int, [vp, C.c_int, C.c_int]),
        "loam_b200_cloud_copy_many": (C.c_int, [vp, _I, vp, _I, C.c_int]),
        "loam_b200_cloud_copy": (C.c_int, [vp, C.c_int, vp, C.c_int]),
        "loam_b200_reg_bin": (C.c_int, [vp, _F, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, _I, _I]),
        "loam_b200_reg_run": (C.c_int, [vp, _I, _I, C.c_int, C.POINTER(RegParams), _I]),
        "loam_b200_reg_indices": (C.c_int, [vp, C.c_int, _I, C.c_int, _I]),
        "loam_b200_reg_labels": (C.c_int, [vp, _B, C.c_int]),
        "loam_b200_odom_prepare": (C.c_int, [vp]),
        "loam_b200_odom_adopt": (C.c_int, [vp, vp, _I]),
        "loam_b200_cloud_transform_to_end2": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(OdomPose)]),
        "loam_b200_odom_rebuild_last": (C.c_int, [vp]),
        "loam_b200_cloud_transform_to_end": (C.c_int, [vp, C.c_int, C.POINTER(OdomPose)]),
        "loam_b200_cloud_transform_to_map": (C.c_int, [vp, C.c_int, C.POINTER(Pose)]),
        "loam_b200_map_pool_append": (C.c_int, [vp, C.c_int, _F, C.c_int]),
        "loam_b200_map_begin_sweep": (C.c_int, [vp, C.POINTER(Pose), C.POINTER(MapWindow), _I]),
        "loam_b200_map_end_sweep": (C.c_int, [vp, C.POINTER(Pose)]),
        "loam_b200_map_end_sweep_async": (C.c_int, [vp, C.POINTER(Pose)]),
        "loam_b200_map_surround": (C.c_int, [vp, _I, _I, C.c_int, C.c_float]),
        "loam_b200_map_surround_async": (C.c_int, [vp, _I, _I, C.c_int, C.c_float]),
        "loam_b200_map_debug_from_map": (C.c_int, [vp, C.c_int]),
        "loam_b200_comm_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
        "loam_b200_comm_init": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]),
        "loam_b200_comm_destroy": (C.c_int, [vp]),
        "loam_b200_map_set_shard": (C.c_int, [vp, C.c_int, C.c_int]),
        "loam_b200_profile_enable": (C.c_int, [vp, C.c_int]),
        "loam_b200_profile_reset": (C.c_int, [vp]),
        "loam_b200_profile_get": (C.c_int, [vp, C.c_int, _D, C.POINTER(C.c_longlong)]),
        "loam_b200_launch_count": (C.c_longlong, [vp]),
        "loam_b200_total_launch_count": (C.c_longlong, []),
        # host handles
        "loam_b200_host_last_error": (C.c_char_p, []),
        "loam_b200_host_set_device": (None, [C.c_int]),
        "loam_b200_scanreg_create": (vp, []),
        "loam_b200_scanreg_destroy": (None, [vp]),
        "loam_b200_scanreg_configure": (C.c_int, [vp, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]),
        "loam_b200_transform_maintenance": (C.c_int, [_F, _F, _F, _F]),
        "loam_b200_host_gn_solve": (C.c_int, [_F, _F, C.c_int, C.c_float, _F, _I]),
        "loam_b200_scanreg_process_unordered": (C.c_int, [vp, _F, C.c_int, C.c_float, C.c_float, C.c_int]),
        "loam_b200_scanreg_process": (C.c_int, [vp, _F, _I, C.c_int]),
        "loam_b200_scanreg_cloud_size": (C.c_int, [vp, C.c_int]),
        "loam_b200_scanreg_cloud_copy": (C.c_int, [vp, C.c_int, _F]),
        "loam_b200_scanreg_index_size": (C.c_int, [vp, C.c_int]),
        "loam_b200_scanreg_index_copy": (C.c_int, [vp, C.c_int, _I]),
        "loam_b200_odom_create": (vp, [C.c_float, C.c_int]),
        "loam_b200_odom_destroy": (None, [vp]),
        "loam_b200_odom_set_inputs": (C.c_int, [vp, _F, C.c_int, _F, C.c_int, _F, C.c_int, _F, C.c_int, _F, C.c_int]),
        "loam_b200_odom_process": (C.c_int, [vp]),
        "loam_b200_odom_full_to_end": (C.c_int, [vp]),
        "loam_b200_odom_get_twist": (C.c_int, [vp, C.c_int, _F]),
        "loam_b200_odom_cloud_size": (C.c_int, [vp, C.c_int]),
        "loam_b200_odom_cloud_copy": (C.c_int, [vp, C.c_int, _F]),
        "loam_b200_odom_last_iterations": (C.c_int, [vp]),
        "loam_b200_map_create": (vp, [C.c_float, C.c_int]),
        "loam_b200_map_destroy": (None, [vp]),
        "loam_b200_map_seed": (C.c_int, [vp, _F, C.c_int, _F, C.c_int]),
        "loam_b200_map_set_inputs": (C.c_int, [vp, _F, C.c_int, _F, C.c_int, _F, C.c_int]),
        "loam_b200_map_update_odometry": (C.c_int, [vp, _F]),
        "loam_b200_map_process": (C.c_int, [vp]),
        "loam_b200_map_get_twist": (C.c_int, [vp, C.c_int, _F]),
        "loam_b200_map_cloud_size": (C.c_int, [vp, C.c_int]),
        "loam_b200_map_cloud_copy": (C.c_int, [vp, C.c_int, _F]),
        "loam_b200_map_last_iterations": (C.c_int, [vp]),
        "loam_b200_map_last_phase_seconds": (C.c_int, [vp, _D]),
        "loam_b200_map_retain_from_map": (C.c_int, [vp, C.c_int]),
        "loam_b200_host_nccl_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
        "loam_b200_map_enable_sharding": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]),
        "loam_b200_map_kernel_profile": (C.c_int, [vp, C.c_int, _D]),
        "loam_b200_map_kernel_profile_queries": (C.c_int, [vp, _F, C.c_int, C.c_int, _D]),
        "loam_b200_map_peer_export": (C.c_int, [vp, C.POINTER(C.c_ubyte)]),
        "loam_b200_map_enable_cube_sharding": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.c_int]),
        "loam_b200_map_enable_cube_sharding_local": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int]),
        "loam_b200_map_disable_cube_sharding": (C.c_int, [vp]),
        "loam_b200_peer_export": (C.c_int, [vp, C.POINTER(C.c_ubyte)]),
        "loam_b200_peer_connect": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.c_int]),
        "loam_b200_peer_connect_local": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int]),
        "loam_b200_peer_disconnect": (C.c_int, [vp]),
        "loam_b200_shard_owner": (C.c_int, [C.c_int, C.c_int, C.c_int]),
        "loam_b200_shard_stores": (C.c_int, [C.c_float, C.c_int, C.c_int, C.c_int]),
        "loam_b200_pipeline_create": (vp, [C.c_float, C.c_int, C.c_int]),
        "loam_b200_pipeline_destroy": (None, [vp]),
        "loam_b200_pipeline_seed_map": (C.c_int, [vp, _F, C.c_int, _F, C.c_int]),
        "loam_b200_pipeline_sweep": (C.c_int, [vp, _F, _I, C.c_int, _F, _F, _D]),
        "loam_b200_pipeline_sweep_device": (C.c_int, [vp, vp, _I, C.c_int, _F, _F, _D]),
        "loam_b200_pipeline_sweep_hostclouds": (C.c_int, [vp, _F, _I, C.c_int, _F, _F, _D]),
        "loam_b200_pipeline_submit": (C.c_int, [vp, _F, vp, _I, C.c_int]),
        "loam_b200_pipeline_collect": (C.c_int, [vp, C.c_int, _F, _F, _I]),
        "loam_b200_pipeline_sync": (C.c_int, [vp]),
        "loam_b200_pipeline_stage_seconds": (C.c_int, [vp, _D, C.c_int]),
        "loam_b200_pipeline_scanreg": (vp, [vp]),
        "loam_b200_pipeline_odom": (vp, [vp]),
        "loam_b200_pipeline_map": (vp, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(_F)


def _ip(a):
    return a.ctypes.data_as(_I)


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(-1, 4)


def make_pose(twist6) -> Pose:
    """Pose struct from (rot_x, rot_y, rot_z, x, y, z) with float sin/cos like loam::Angle caches them."""
    t = np.asarray(twist6, dtype=np.float32)
    p = Pose()
    for i in range(3):
        p.rot[i] = t[i]
        p.sin_[i] = np.sin(t[i], dtype=np.float32)
        p.cos_[i] = np.cos(t[i], dtype=np.float32)
        p.pos[i] = t[3 + i]
    return p


def make_odom_pose(twist6, scan_period=0.1, it=0) -> OdomPose:
    t = np.asarray(twist6, dtype=np.float32)
    p = OdomPose()
    for i in range(3):
        p.rot[i] = t[i]
        p.sin_[i] = np.sin(t[i], dtype=np.float32)
        p.cos_[i] = np.cos(t[i], dtype=np.float32)
        p.pos[i] = t[3 + i]
    p.inv_scan_period = np.float32(1.0) / np.float32(scan_period)
    p.iter = it
    return p


def ring_ranges(ring_sizes):
    """Inclusive (start, end) per ring exactly as processScanlines builds _scanIndices (BasicScanRegistration.cpp:38-41)."""
    sizes = np.asarray(ring_sizes, dtype=np.int64)
    ends = np.cumsum(sizes)
    starts = ends - sizes
    e = np.where(ends > 0, ends - 1, 0)
    return starts.astype(np.int32), e.astype(np.int32)


class Ctx:
    """Kernel-level context (include/loam_b200.h)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.loam_b200_create(C.byref(h), device)
        if rc != 0:
            raise LoamB200Error(f"loam_b200_create: {self.L.loam_b200_strerror(rc).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.loam_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            detail = self.L.loam_b200_last_error(self.h).decode()
            raise LoamB200Error(f"{what}: {self.L.loam_b200_strerror(rc).decode()} {detail}")

    def reg_bin(self, xyz, lower_deg, upper_deg, n_rings, scan_period=0.1):
        """Ring-binning front end (loam_b200_reg_bin) -> (ring-ordered n_kept x 4 cloud, ring sizes)."""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        sizes = np.zeros(n_rings, np.int32)
        kept = C.c_int(0)
        self._ck(self.L.loam_b200_reg_bin(self.h, _fp(a), a.shape[0], 0, lower_deg, upper_deg, n_rings, scan_period,
                                          _ip(sizes), C.byref(kept)), "reg_bin")
        out = np.empty((max(kept.value, 1), 4), np.float32)
        got = C.c_int(0)
        self._ck(self.L.loam_b200_cloud_download(self.h, 0, _fp(out), out.shape[0], C.byref(got)), "cloud_download")  # C_REG_FULL
        return out[:got.value].copy(), sizes

    def extract_features(self, pts, ring_sizes, params: RegParams | None = None):
        pts = _pts(pts)
        n = pts.shape[0]
        rs, re = ring_ranges(ring_sizes)
        R = rs.shape[0]
        prm = params or RegParams.default()
        cs, cl, cf = R * prm.nFeatureRegions * prm.maxCornerSharp, R * prm.nFeatureRegions * prm.maxCornerLessSharp, \
            R * prm.nFeatureRegions * prm.maxSurfaceFlat
        sharp = np.empty(cs + 1, np.int32)
        less = np.empty(cl + 1, np.int32)
        flat = np.empty(cf + 1, np.int32)
        label = np.empty(n + 1, np.int8)
        lf = np.empty((n + 1, 4), np.float32)
        out = Features(_ip(sharp), cs, 0, _ip(less), cl, 0, _ip(flat), cf, 0, label.ctypes.data_as(_B), _fp(lf), n, 0)
        self._ck(self.L.loam_b200_extract_features(self.h, _fp(pts), n, _ip(rs), _ip(re), R, C.byref(prm), C.byref(out)),
                 "extract_features")
        return {"sharp": sharp[:out.n_sharp].copy(), "less_sharp": less[:out.n_less_sharp].copy(),
                "flat": flat[:out.n_flat].copy(), "label": label[:n].copy(),
                "less_flat_ds": lf[:out.n_less_flat].copy()}

    def tree_build(self, slot, pts):
        pts = _pts(pts)
        self._ck(self.L.loam_b200_tree_build(self.h, slot, _fp(pts), pts.shape[0]), "tree_build")

    def tree_knn(self, slot, queries, k, max_d2=np.inf):
        q = _pts(queries)
        idx = np.empty((q.shape[0], k), np.int32)
        d2 = np.empty((q.shape[0], k), np.float32)
        self._ck(self.L.loam_b200_tree_knn(self.h, slot, _fp(q), q.shape[0], k, np.float32(max_d2), _ip(idx), _fp(d2)),
                 "tree_knn")
        return idx, d2

    def map_set_shard(self, rank, world):
        self._ck(self.L.loam_b200_map_set_shard(self.h, rank, world), "map_set_shard")

    def map_set_queries(self, corner, surf):
        c, s = _pts(corner), _pts(surf)
        self._ck(self.L.loam_b200_map_set_queries(self.h, _fp(c), c.shape[0], _fp(s), s.shape[0]), "map_set_queries")
        self._nq = c.shape[0] + s.shape[0]

    def map_iterate(self, twist6, debug=False):
        p = make_pose(twist6)
        ne = NormalEq()
        if debug:
            coeff = np.zeros((self._nq, 4), np.float32)
            sel = np.zeros(self._nq, np.int8)
            self._ck(self.L.loam_b200_map_iterate_debug(self.h, C.byref(p), C.byref(ne), _fp(coeff),
                                                        sel.ctypes.data_as(_B)), "map_iterate_debug")
            return _ne(ne), coeff, sel
        self._ck(self.L.loam_b200_map_iterate(self.h, C.byref(p), C.byref(ne)), "map_iterate")
        return _ne(ne)

    def map_iterate_stats(self, twist6):
        """(normal equations, BVH nodes visited, leaves visited) of one instrumented launch."""
        p = make_pose(twist6)
        ne = NormalEq()
        nodes = C.c_ulonglong(0)
        leaves = C.c_ulonglong(0)
        self._ck(self.L.loam_b200_map_iterate_stats(self.h, C.byref(p), C.byref(ne), C.byref(nodes), C.byref(leaves)),
                 "map_iterate_stats")
        return _ne(ne), nodes.value, leaves.value

    def odom_set_last(self, corner, surf):
        c, s = _pts(corner), _pts(surf)
        self._ck(self.L.loam_b200_odom_set_last(self.h, _fp(c), c.shape[0], _fp(s), s.shape[0]), "odom_set_last")

    def odom_set_current(self, sharp, flat):
        c, s = _pts(sharp), _pts(flat)
        self._ck(self.L.loam_b200_odom_set_current(self.h, _fp(c), c.shape[0], _fp(s), s.shape[0]), "odom_set_current")
        self._noq = c.shape[0] + s.shape[0]

    def odom_iterate(self, twist6, it, scan_period=0.1, debug=False):
        p = make_odom_pose(twist6, scan_period, it)
        ne = NormalEq()
        if debug:
            coeff = np.zeros((self._noq, 4), np.float32)
            sel = np.zeros(self._noq, np.int8)
            ind = np.zeros((self._noq, 3), np.int32)
            self._ck(self.L.loam_b200_odom_iterate_debug(self.h, C.byref(p), C.byref(ne), _fp(coeff),
                                                         sel.ctypes.data_as(_B), _ip(ind)), "odom_iterate_debug")
            return _ne(ne), coeff, sel, ind
        self._ck(self.L.loam_b200_odom_iterate(self.h, C.byref(p), C.byref(ne)), "odom_iterate")
        return _ne(ne)

    def debug_gn_solve(self, AtA, AtB, first_iteration=True, eigen_threshold=10.0):
        """The device loops' warp-parallel 6 x 6 step on n systems (AtA n x 6 x 6, AtB n x 6) -> (x n x 6, degenerate n)."""
        A = np.ascontiguousarray(AtA, dtype=np.float32).reshape(-1, 36)
        b = np.ascontiguousarray(AtB, dtype=np.float32).reshape(-1, 6)
        packed = np.ascontiguousarray(np.concatenate([A, b], axis=1), dtype=np.float32)
        n = packed.shape[0]
        x = np.zeros((n, 6), np.float32)
        deg = np.zeros(n, np.int32)
        self._ck(self.L.loam_b200_debug_gn_solve(self.h, _fp(packed), n, 1 if first_iteration else 0, eigen_threshold,
                                                 _fp(x), _ip(deg)), "debug_gn_solve")
        return x, deg

    def odom_solve(self, twist6, scan_period=0.1, max_iter=25, delta_t=0.1, delta_r=0.1):
        """Whole scan-to-scan Gauss-Newton loop on the device -> (twist6, iterations)."""
        t = np.ascontiguousarray(twist6, dtype=np.float32)
        rot, pos = t[:3].copy(), t[3:].copy()
        res = LmResult()
        self._ck(self.L.loam_b200_odom_solve(self.h, _fp(rot), _fp(pos), np.float32(1.0) / np.float32(scan_period), max_iter,
                                             delta_t, delta_r, C.byref(res)), "odom_solve")
        return np.array(list(res.rot) + list(res.pos), np.float32), res.iterations

    def map_solve(self, twist6, max_iter=10, delta_t=0.05, delta_r=0.05):
        """Whole scan-to-map Gauss-Newton loop on the device -> (twist6, iterations)."""
        t = np.ascontiguousarray(twist6, dtype=np.float32)
        rot, pos = t[:3].copy(), t[3:].copy()
        res = LmResult()
        self._ck(self.L.loam_b200_map_solve(self.h, _fp(rot), _fp(pos), max_iter, delta_t, delta_r, C.byref(res)),
                 "map_solve")
        return np.array(list(res.rot) + list(res.pos), np.float32), res.iterations

    def transform_to_end(self, pts, twist6, scan_period=0.1):
        a = _pts(pts).copy()
        p = make_odom_pose(twist6, scan_period, 0)
        self._ck(self.L.loam_b200_transform_to_end(self.h, _fp(a), a.shape[0], C.byref(p)), "transform_to_end")
        return a

    def transform_to_map(self, pts, twist6):
        a = _pts(pts).copy()
        p = make_pose(twist6)
        self._ck(self.L.loam_b200_transform_to_map(self.h, _fp(a), a.shape[0], C.byref(p)), "transform_to_map")
        return a

    def voxel_grid(self, pts, leaf):
        a = _pts(pts)
        out = np.empty_like(a)
        n = C.c_int(0)
        self._ck(self.L.loam_b200_voxel_grid(self.h, _fp(a), a.shape[0], np.float32(leaf), _fp(out), a.shape[0],
                                             C.byref(n)), "voxel_grid")
        return out[:n.value].copy()

    def profile(self, on=True):
        self.L.loam_b200_profile_enable(self.h, 1 if on else 0)
        self.L.loam_b200_profile_reset(self.h)

    def profile_get(self):
        out = {}
        for i, name in enumerate(KERNEL_FAMILIES):
            ms = C.c_double(0)
            n = C.c_longlong(0)
            self.L.loam_b200_profile_get(self.h, i, C.byref(ms), C.byref(n))
            out[name] = (ms.value, n.value)
        return out

    def launch_count(self):
        return self.L.loam_b200_launch_count(self.h)


def _ne(ne: NormalEq):
    return {"AtA": np.array(ne.AtA, dtype=np.float32).reshape(6, 6), "AtB": np.array(ne.AtB, dtype=np.float32),
            "n_selected": ne.n_selected, "n_corner_selected": ne.n_corner_selected}


# ---------------------------------------------------------------------------------------------------------------------
# handles over the C++ drop-in classes
class _Handle:
    def _ck(self, rc, what):
        if rc < 0:
            raise LoamB200Error(f"{what}: {self.L.loam_b200_host_last_error().decode()}")
        return rc

    def _cloud(self, size_fn, copy_fn, which):
        n = size_fn(self.h, which)
        out = np.empty((n, 4), np.float32)
        if n:
            copy_fn(self.h, which, _fp(out))
        return out


class ScanRegistration(_Handle):
    """loam::BasicScanRegistration (reference: include/loam_velodyne/BasicScanRegistration.h:135-164)."""
    NAMES = {"full": 0, "sharp": 1, "less_sharp": 2, "flat": 3, "less_flat": 4}

    def __init__(self, handle=None):
        self.L = lib()
        self.own = handle is None
        self.h = handle if handle is not None else self.L.loam_b200_scanreg_create()

    def __del__(self):
        if getattr(self, "own", False) and self.h:
            self.L.loam_b200_scanreg_destroy(self.h)
            self.h = None

    def configure(self, scan_period=0.1, n_regions=6, curv_region=5, max_sharp=2, max_flat=4, less_flat_leaf=0.2,
                  curv_thr=0.1):
        self._ck(self.L.loam_b200_scanreg_configure(self.h, scan_period, n_regions, curv_region, max_sharp, max_flat,
                                                    less_flat_leaf, curv_thr), "configure")

    def process(self, pts, ring_sizes):
        pts = _pts(pts)
        rs = np.ascontiguousarray(ring_sizes, dtype=np.int32)
        self._ck(self.L.loam_b200_scanreg_process(self.h, _fp(pts), _ip(rs), rs.shape[0]), "processScanlines")

    def process_unordered(self, xyz, lower_deg, upper_deg, n_rings):
        """MultiScanRegistration::process: unordered sensor-frame xyz (n x 3) -> ring binning on the GPU -> extraction."""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        self._ck(self.L.loam_b200_scanreg_process_unordered(self.h, _fp(a), a.shape[0], lower_deg, upper_deg, n_rings),
                 "processUnorderedSweep")

    def cloud(self, name):
        return self._cloud(self.L.loam_b200_scanreg_cloud_size, self.L.loam_b200_scanreg_cloud_copy, self.NAMES[name])

    def indices(self, name):
        w = self.NAMES[name]
        n = self.L.loam_b200_scanreg_index_size(self.h, w)
        out = np.empty(n, np.int32)
        if n:
            self.L.loam_b200_scanreg_index_copy(self.h, w, _ip(out))
        return out


class LaserOdometry(_Handle):
    """loam::BasicLaserOdometry (reference: include/loam_velodyne/BasicLaserOdometry.h:13-48)."""
    NAMES = {"last_corner": 0, "last_surf": 1, "full": 2}

    def __init__(self, scan_period=0.1, max_iter=25, handle=None):
        self.L = lib()
        self.own = handle is None
        self.h = handle if handle is not None else self.L.loam_b200_odom_create(scan_period, max_iter)

    def __del__(self):
        if getattr(self, "own", False) and self.h:
            self.L.loam_b200_odom_destroy(self.h)
            self.h = None

    def set_inputs(self, sharp, less_sharp, flat, less_flat, full):
        a = [_pts(x) for x in (sharp, less_sharp, flat, less_flat, full)]
        self._ck(self.L.loam_b200_odom_set_inputs(self.h, _fp(a[0]), a[0].shape[0], _fp(a[1]), a[1].shape[0],
                                                  _fp(a[2]), a[2].shape[0], _fp(a[3]), a[3].shape[0], _fp(a[4]),
                                                  a[4].shape[0]), "set_inputs")

    def process(self):
        self._ck(self.L.loam_b200_odom_process(self.h), "BasicLaserOdometry::process")

    def full_to_end(self):
        self._ck(self.L.loam_b200_odom_full_to_end(self.h), "transformToEnd")

    def twist(self, which):
        out = np.empty(6, np.float32)
        self.L.loam_b200_odom_get_twist(self.h, {"transform": 0, "sum": 1}[which], _fp(out))
        return out

    def cloud(self, name):
        return self._cloud(self.L.loam_b200_odom_cloud_size, self.L.loam_b200_odom_cloud_copy, self.NAMES[name])

    def last_iterations(self):
        return self.L.loam_b200_odom_last_iterations(self.h)


class LaserMapping(_Handle):
    """loam::BasicLaserMapping (reference: include/loam_velodyne/BasicLaserMapping.h:77-111)."""
    NAMES = {"full": 0, "surround_ds": 1, "corner_from_map": 2, "surf_from_map": 3, "corner_stack_ds": 4,
             "surf_stack_ds": 5, "corner_cubes": 6, "surf_cubes": 7}

    def __init__(self, scan_period=0.1, max_iter=10, handle=None):
        self.L = lib()
        self.own = handle is None
        self.h = handle if handle is not None else self.L.loam_b200_map_create(scan_period, max_iter)

    def __del__(self):
        if getattr(self, "own", False) and self.h:
            self.L.loam_b200_map_destroy(self.h)
            self.h = None

    def retain_from_map(self, on=True):
        """Test hook: keep the reference's internal from-map clouds ("corner_from_map" / "surf_from_map") every sweep."""
        self._ck(self.L.loam_b200_map_retain_from_map(self.h, 1 if on else 0), "retainFromMapClouds")

    def seed(self, corner, surf):
        c, s = _pts(corner), _pts(surf)
        self._ck(self.L.loam_b200_map_seed(self.h, _fp(c), c.shape[0], _fp(s), s.shape[0]), "seedMap")

    def set_inputs(self, corner_last, surf_last, full):
        a = [_pts(x) for x in (corner_last, surf_last, full)]
        self._ck(self.L.loam_b200_map_set_inputs(self.h, _fp(a[0]), a[0].shape[0], _fp(a[1]), a[1].shape[0], _fp(a[2]),
                                                 a[2].shape[0]), "set_inputs")

    def update_odometry(self, sum6):
        s = np.ascontiguousarray(sum6, dtype=np.float32)
        self.L.loam_b200_map_update_odometry(self.h, _fp(s))

    def process(self):
        return bool(self._ck(self.L.loam_b200_map_process(self.h), "BasicLaserMapping::process"))

    def twist(self, which):
        out = np.empty(6, np.float32)
        self.L.loam_b200_map_get_twist(self.h, {"aft": 0, "bef": 1, "tobe": 2}[which], _fp(out))
        return out

    def cloud(self, name):
        return self._cloud(self.L.loam_b200_map_cloud_size, self.L.loam_b200_map_cloud_copy, self.NAMES[name])

    def last_iterations(self):
        return self.L.loam_b200_map_last_iterations(self.h)

    def last_phase_ms(self):
        out = np.zeros(4, np.float64)
        self.L.loam_b200_map_last_phase_seconds(self.h, out.ctypes.data_as(_D))
        return dict(zip(["begin_sweep", "lm_loop", "end_sweep", "surround"], np.round(out * 1e3, 3)))

    def enable_sharding(self, rank, world, nccl_id: bytes | None):
        """Evaluate the rank-th of `world` query slices; with nccl_id (128 bytes, same on every rank) the normal
        equations are all-reduced over NCCL each iteration."""
        if nccl_id is None:
            self._ck(self.L.loam_b200_map_enable_sharding(self.h, rank, world, None), "enableSharding")
        else:
            buf = (C.c_ubyte * 128).from_buffer_copy(nccl_id)
            self._ck(self.L.loam_b200_map_enable_sharding(self.h, rank, world, buf), "enableSharding")


    def kernel_profile(self, reps=50):
        """The scan-to-map iteration kernel as this object launches it, timed with CUDA events (see the header)."""
        out = np.zeros(5, np.float64)
        self._ck(self.L.loam_b200_map_kernel_profile(self.h, reps, out.ctypes.data_as(_D)), "map_kernel_profile")
        return {"avg_us": float(out[0]), "queries": int(out[1]), "probes_per_query": float(out[2]),
                "candidates_per_query": float(out[3]), "n_selected": int(out[4])}

    def kernel_profile_queries(self, queries, reps=10):
        """... on caller-supplied surface queries in the map frame (k-NN bandwidth stress)."""
        q = _pts(queries)
        out = np.zeros(5, np.float64)
        self._ck(self.L.loam_b200_map_kernel_profile_queries(self.h, _fp(q), q.shape[0], reps, out.ctypes.data_as(_D)),
                 "map_kernel_profile_queries")
        return {"avg_us": float(out[0]), "queries": int(out[1]), "probes_per_query": float(out[2]),
                "candidates_per_query": float(out[3]), "n_selected": int(out[4])}

    # ---- multi-GPU with the map sharded by cube slabs (include/loam_b200_host.h)
    def peer_export(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        self._ck(self.L.loam_b200_map_peer_export(self.h, buf), "exportPeerHandle")
        return bytes(buf)

    def enable_cube_sharding(self, rank, world, handles: bytes, slab_metres=10):
        """handles: world x 64 bytes (peer_export of every rank, in rank order)."""
        buf = (C.c_ubyte * (64 * world)).from_buffer_copy(handles)
        self._ck(self.L.loam_b200_map_enable_cube_sharding(self.h, rank, world, buf, slab_metres), "enableCubeSharding")


    def disable_cube_sharding(self):
        """Unmap the peers' inboxes; every rank calls it, then the ranks synchronise, then the objects may be destroyed."""
        self._ck(self.L.loam_b200_map_disable_cube_sharding(self.h), "disableCubeSharding")


def enable_cube_sharding_local(mappings, slab_metres=10):
    """Several LaserMapping objects of this process become the ranks 0..n-1 of one cube-sharded map."""
    L = lib()
    arr = (C.c_void_p * len(mappings))(*[m.h for m in mappings])
    if L.loam_b200_map_enable_cube_sharding_local(arr, len(mappings), slab_metres) < 0:
        raise LoamB200Error(f"enableCubeShardingLocal: {L.loam_b200_host_last_error().decode()}")


def shard_stores(x, rank, world, slab_metres=10):
    return lib().loam_b200_shard_stores(float(x), rank, world, slab_metres) == 1


def shard_owner(cell_x, world, slab_metres=10):
    return lib().loam_b200_shard_owner(int(cell_x), slab_metres, world)


class Pipeline(_Handle):
    """registration -> odometry -> mapping chained in-process on one sweep (SURVEY.md §8b "who calls it")."""

    def __init__(self, scan_period=0.1, odom_iter=25, map_iter=10):
        self.L = lib()
        self.h = self.L.loam_b200_pipeline_create(scan_period, odom_iter, map_iter)
        self.scanreg = ScanRegistration(self.L.loam_b200_pipeline_scanreg(self.h))
        self.odom = LaserOdometry(handle=self.L.loam_b200_pipeline_odom(self.h))
        self.mapping = LaserMapping(handle=self.L.loam_b200_pipeline_map(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.loam_b200_pipeline_destroy(self.h)
            self.h = None

    def seed_map(self, corner, surf):
        c, s = _pts(corner), _pts(surf)
        self._ck(self.L.loam_b200_pipeline_seed_map(self.h, _fp(c), c.shape[0], _fp(s), s.shape[0]), "seed_map")

    def sweep(self, pts, ring_sizes, mode="fused"):
        """mode "fused": clouds stay in HBM between the three stages; "hostclouds": every hand-off goes through the
        reference's own entry points and host pcl clouds (as separate ROS nodes would use the classes)."""
        pts = _pts(pts)
        rs = np.ascontiguousarray(ring_sizes, dtype=np.int32)
        odom = np.empty(6, np.float32)
        aft = np.empty(6, np.float32)
        st = np.zeros(5, np.float64)
        fn = self.L.loam_b200_pipeline_sweep if mode == "fused" else self.L.loam_b200_pipeline_sweep_hostclouds
        ok = self._ck(fn(self.h, _fp(pts), _ip(rs), rs.shape[0], _fp(odom), _fp(aft), st.ctypes.data_as(_D)),
                      "pipeline_sweep")
        return bool(ok), odom, aft, st

    def sweep_device(self, device_ptr, ring_sizes):
        """Sweep already resident in GPU memory (device_ptr: address of n x 4 float32, e.g. tensor.data_ptr())."""
        rs = np.ascontiguousarray(ring_sizes, dtype=np.int32)
        odom = np.empty(6, np.float32)
        aft = np.empty(6, np.float32)
        st = np.zeros(5, np.float64)
        ok = self._ck(self.L.loam_b200_pipeline_sweep_device(self.h, C.c_void_p(device_ptr), _ip(rs), rs.shape[0],
                                                             _fp(odom), _fp(aft), st.ctypes.data_as(_D)),
                      "pipeline_sweep_device")
        return bool(ok), odom, aft, st


    def sync(self):
        """Everything enqueued by the three stages (helper threads included) has finished on the GPU."""
        self._ck(self.L.loam_b200_pipeline_sync(self.h), "pipeline_sync")

    def stage_seconds(self, reset=False):
        """Streaming mode: seconds per stage thread (registration, odometry, mapping) spent working / waiting / in hand-offs."""
        out = np.zeros(9, np.float64)
        self.L.loam_b200_pipeline_stage_seconds(self.h, out.ctypes.data_as(_D), 1 if reset else 0)
        return {"busy": out[0:3].copy(), "idle": out[3:6].copy(), "handoff": out[6:9].copy()}

    # ---- streaming form: the three stages run concurrently on consecutive sweeps (include/loam_b200_host.h)
    def submit(self, pts=None, ring_sizes=None, device_ptr=None):
        """Queue one sweep (host array `pts` or `device_ptr`); the buffer must stay alive until collected."""
        rs = np.ascontiguousarray(ring_sizes, dtype=np.int32)
        if device_ptr is None:
            pts = _pts(pts)
            self._keep = getattr(self, "_keep", [])
            self._keep.append(pts)
            rc = self.L.loam_b200_pipeline_submit(self.h, _fp(pts), None, _ip(rs), rs.shape[0])
        else:
            rc = self.L.loam_b200_pipeline_submit(self.h, None, C.c_void_p(device_ptr), _ip(rs), rs.shape[0])
        self._ck(rc, "pipeline_submit")

    def collect(self, wait=True):
        """Oldest finished sweep -> (ok, odom_sum6, map_aft6), or None when nothing is pending / ready."""
        odom = np.empty(6, np.float32)
        aft = np.empty(6, np.float32)
        ok = C.c_int(0)
        rc = self._ck(self.L.loam_b200_pipeline_collect(self.h, 1 if wait else 0, _fp(odom), _fp(aft), C.byref(ok)),
                      "pipeline_collect")
        if rc == 0:
            return None
        if getattr(self, "_keep", None):
            self._keep.pop(0)
        return bool(ok.value), odom, aft

    def run_stream(self, sweeps=None, device_ptrs=None):
        """Submit every sweep, collect every result (in order) -> list of (ok, odom_sum6, map_aft6)."""
        res = []
        n = len(sweeps)
        for i in range(n):
            if device_ptrs is not None:
                self.submit(ring_sizes=sweeps[i][1], device_ptr=device_ptrs[i])
            else:
                self.submit(sweeps[i][0], sweeps[i][1])
            r = self.collect(wait=False)
            if r is not None:
                res.append(r)
        while len(res) < n:
            r = self.collect(wait=True)
            if r is None:
                break
            res.append(r)
        self.sync()  # the last sweep's asynchronous map update has been issued and has finished
        return res


def transform_maintenance(sum6, bef6, aft6):
    """loam::BasicTransformMaintenance: updateOdometry + updateMappingTransform + transformAssociateToMap -> mapped pose."""
    a = [np.ascontiguousarray(v, dtype=np.float32).reshape(6) for v in (sum6, bef6, aft6)]
    out = np.zeros(6, np.float32)
    if lib().loam_b200_transform_maintenance(_fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(out)) != 0:
        raise LoamB200Error("transform_maintenance: invalid arguments")
    return out


def gn_solve(AtA, AtB, first_iteration=True, eigen_threshold=10.0):
    """The library's 6 x 6 Gauss-Newton step (host arithmetic, same source as the device loop) -> (x, degenerate)."""
    L = lib()
    a = np.ascontiguousarray(AtA, dtype=np.float32).reshape(36)
    b = np.ascontiguousarray(AtB, dtype=np.float32).reshape(6)
    x = np.zeros(6, np.float32)
    deg = C.c_int(0)
    if L.loam_b200_host_gn_solve(_fp(a), _fp(b), 1 if first_iteration else 0, eigen_threshold, _fp(x), C.byref(deg)) != 0:
        raise LoamB200Error("gn_solve: invalid arguments")
    return x, bool(deg.value)


def nccl_unique_id() -> bytes:
    """128-byte NCCL id (call on rank 0, broadcast to the others)."""
    buf = (C.c_ubyte * 128)()
    if lib().loam_b200_host_nccl_unique_id(buf) != 0:
        raise LoamB200Error(lib().loam_b200_host_last_error().decode())
    return bytes(buf)


def shard_slice(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n queries evaluated by `rank` (mirror of map_iterate_impl in csrc/loam_b200.cu)."""
    return n * rank // world, n * (rank + 1) // world


def set_device(device: int):
    lib().loam_b200_host_set_device(device)
