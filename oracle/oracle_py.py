"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (CPU restatement, parity unpinned).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from mola_lidar_odometry_b200.capi import (DecimateParams, Filter1Params, IcpParams, IcpResult, MapParams)

HERE = Path(__file__).resolve().parent
_LIB = None
_vp, _u32, _u64, _f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float


def lib():
    global _LIB
    if _LIB is None:
        so = HERE / "liboracle.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
        L = C.CDLL(str(so))
        L.orc_voxel_index_map.restype = C.c_int32
        L.orc_voxel_index_map.argtypes = [_f, _f]
        L.orc_voxel_index_filter.restype = C.c_int32
        L.orc_voxel_index_filter.argtypes = [_f, _f]
        L.orc_geman_mcclure.restype = C.c_double
        L.orc_geman_mcclure.argtypes = [C.c_double, C.c_double]
        L.orc_pool_create.restype = _vp
        L.orc_pool_create.argtypes = [C.c_int]
        L.orc_pool_destroy.argtypes = [_vp]
        L.orc_map_create.restype = _vp
        L.orc_map_create.argtypes = [C.POINTER(MapParams)]
        L.orc_map_destroy.argtypes = [_vp]
        L.orc_map_clear.argtypes = [_vp]
        L.orc_map_insert.argtypes = [_vp, _vp, _u32, _u64, _vp]
        L.orc_map_cull.argtypes = [_vp, _vp, _f]
        L.orc_map_stats.argtypes = [_vp, C.POINTER(_u64), C.POINTER(_u64)]
        L.orc_map_nn_single.argtypes = [_vp, _vp, _u32, _u64, _vp, _vp, _vp, C.POINTER(_u64)]
        L.orc_map_nn_plane.argtypes = [_vp, _vp, _u32, _u64, _vp, _vp, _vp, _vp]
        L.orc_map_export.restype = C.c_int
        L.orc_map_export.argtypes = [_vp, _vp, _vp, _vp, _u64, _u64, C.POINTER(_u64), C.POINTER(_u64)]
        L.orc_voxel_decimate_first.argtypes = [_vp, _u32, _u64, C.POINTER(DecimateParams), _vp, C.POINTER(_u64)]
        L.orc_filter_1st_pass.argtypes = [_vp, _u32, _u64, C.POINTER(Filter1Params), _vp, C.POINTER(_u64), _vp,
                                          C.POINTER(_u64)]
        L.orc_icp_align.argtypes = [_vp, _vp, _u32, _u64, _vp, C.POINTER(IcpParams), C.POINTER(IcpResult), _vp, _vp,
                                    _vp, _u32]
        L.orc_se3_exp.argtypes = [_vp, _vp]
        L.orc_se3_log.argtypes = [_vp, _vp]
        L.orc_cov_tangent_to_ypr.argtypes = [_vp, _vp, _vp]
        L.orc_pose_minus.argtypes = [_vp, _vp, _vp]
        L.orc_pose_compose.argtypes = [_vp, _vp, _vp]
        L.orc_horn.restype = C.c_int
        L.orc_horn.argtypes = [_vp, _vp, _u64, _vp]
        L.orc_scan_register.argtypes = [_vp, _vp, _u32, _u64, C.POINTER(Filter1Params), _vp, C.POINTER(IcpParams),
                                        C.c_int, _f, _vp, C.POINTER(IcpResult), _vp]
        _LIB = L
    return _LIB


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] in (3, 4)
    return a


def _pose(p):
    return np.ascontiguousarray(np.asarray(p, dtype=np.float64)[:3, :4])


def se3_exp(xi):
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    out = np.empty((3, 4))
    lib().orc_se3_exp(xi.ctypes.data, out.ctypes.data)
    return out


def se3_log(pose):
    p = _pose(pose)
    out = np.empty(6)
    lib().orc_se3_log(p.ctypes.data, out.ctypes.data)
    return out


def set_conventions(index_floor=0, gm_form=0, cull_metric=0):
    """[VERIFY] conventions of the oracle (process-wide): mirror of mlo_set_option('convention_*') on the product side."""
    lib().orc_set_conventions(int(index_floor), int(gm_form), int(cull_metric))


def cov_tangent_to_ypr(pose, cov):
    p, c = _pose(pose), np.ascontiguousarray(cov, dtype=np.float64).reshape(6, 6)
    out = np.empty((6, 6))
    lib().orc_cov_tangent_to_ypr(p.ctypes.data, c.ctypes.data, out.ctypes.data)
    return out


def pose_minus(a, b):
    """b^-1 * a  (MRPT 'a - b')."""
    a, b = _pose(a), _pose(b)
    out = np.empty((3, 4))
    lib().orc_pose_minus(a.ctypes.data, b.ctypes.data, out.ctypes.data)
    return out


def pose_error(a, b):
    """(translation error [m], rotation error [deg]) between two 3x4 poses."""
    xi = se3_log(pose_minus(a, b))
    d = pose_minus(a, b)
    return float(np.linalg.norm(d[:, 3])), float(np.rad2deg(np.linalg.norm(xi[3:])))


def horn(g, l):
    g, l = np.ascontiguousarray(g, np.float32), np.ascontiguousarray(l, np.float32)
    out = np.empty((3, 4))
    rc = lib().orc_horn(g.ctypes.data, l.ctypes.data, len(g), out.ctypes.data)
    if rc != 0:
        raise RuntimeError("horn failed")
    return out


class Pool:
    def __init__(self, n: int):
        self.n = n
        self.h = lib().orc_pool_create(n)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_pool_destroy(self.h)
            self.h = None


class OracleMap:
    def __init__(self, voxel_size=1.0, max_points_per_voxel=20, min_distance_between_points=0.0, kind=0,
                 max_eigen_ratio_for_planes=0.05, min_points_for_plane=5):
        p = MapParams(kind, voxel_size, max_points_per_voxel, min_distance_between_points,
                      max_eigen_ratio_for_planes, min_points_for_plane, 0)
        self.h = lib().orc_map_create(C.byref(p))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_map_destroy(self.h)
            self.h = None

    def clear(self):
        lib().orc_map_clear(self.h)

    def insert(self, pts, pose):
        pts, pose = _f32(pts), _pose(pose)
        lib().orc_map_insert(self.h, pts.ctypes.data, pts.shape[1], len(pts), pose.ctypes.data)

    def cull(self, sensor_xyz, dist):
        s = np.ascontiguousarray(sensor_xyz, dtype=np.float64)
        lib().orc_map_cull(self.h, s.ctypes.data, dist)

    def stats(self):
        a, b = _u64(), _u64()
        lib().orc_map_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def nn_single(self, q):
        q = _f32(q)
        n = len(q)
        xyz, d2, f = np.empty((n, 3), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8)
        nc = _u64()
        lib().orc_map_nn_single(self.h, q.ctypes.data, q.shape[1], n, xyz.ctypes.data, d2.ctypes.data, f.ctypes.data,
                                C.byref(nc))
        return xyz, d2, f.astype(bool), nc.value

    def nn_plane(self, q):
        q = _f32(q)
        n = len(q)
        mean, nrm = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        d, f = np.empty(n, np.float32), np.empty(n, np.uint8)
        lib().orc_map_nn_plane(self.h, q.ctypes.data, q.shape[1], n, mean.ctypes.data, nrm.ctypes.data, d.ctypes.data,
                               f.ctypes.data)
        return mean, nrm, d, f.astype(bool)

    def export(self):
        nv, np_ = _u64(), _u64()
        lib().orc_map_export(self.h, None, None, None, 0, 0, C.byref(nv), C.byref(np_))
        keys, cnt = np.empty((nv.value, 3), np.int32), np.empty(nv.value, np.uint32)
        xyz = np.empty((np_.value, 3), np.float32)
        lib().orc_map_export(self.h, keys.ctypes.data, cnt.ctypes.data, xyz.ctypes.data, nv.value, np_.value,
                             C.byref(nv), C.byref(np_))
        return keys, cnt, xyz


def decimate_first(pts, params: DecimateParams):
    pts = _f32(pts)
    idx = np.empty(len(pts), np.uint32)
    n = _u64()
    lib().orc_voxel_decimate_first(pts.ctypes.data, pts.shape[1], len(pts), C.byref(params), idx.ctypes.data,
                                   C.byref(n))
    return idx[:n.value].copy()


def filter_1st_pass(pts, fp: Filter1Params):
    pts = _f32(pts)
    a, b = np.empty((len(pts), 3), np.float32), np.empty((len(pts), 3), np.float32)
    na, nb = _u64(), _u64()
    lib().orc_filter_1st_pass(pts.ctypes.data, pts.shape[1], len(pts), C.byref(fp), a.ctypes.data, C.byref(na),
                              b.ctypes.data, C.byref(nb))
    return a[:na.value].copy(), b[:nb.value].copy()


def filter_1st_pass_xyzt(pts, t, fp: Filter1Params):
    pts = _f32(pts)
    t = None if t is None else np.ascontiguousarray(t, dtype=np.float32)
    a, b = np.empty((len(pts), 4), np.float32), np.empty((len(pts), 4), np.float32)
    na, nb = _u64(), _u64()
    L = lib()
    L.orc_filter_1st_pass_xyzt.argtypes = [_vp, _u32, _vp, _u64, C.POINTER(Filter1Params), _vp, C.POINTER(_u64), _vp,
                                           C.POINTER(_u64)]
    L.orc_filter_1st_pass_xyzt(pts.ctypes.data, pts.shape[1], None if t is None else t.ctypes.data, len(pts), C.byref(fp),
                               a.ctypes.data, C.byref(na), b.ctypes.data, C.byref(nb))
    return a[:na.value].copy(), b[:nb.value].copy()


def deskew(xyzt, twist):
    xyzt = np.ascontiguousarray(xyzt, dtype=np.float32)
    tw = np.ascontiguousarray(twist, dtype=np.float64)
    out = np.empty((len(xyzt), 3), np.float32)
    L = lib()
    L.orc_deskew.argtypes = [_vp, _u64, _vp, _vp]
    L.orc_deskew(xyzt.ctypes.data, len(xyzt), tw.ctypes.data, out.ctypes.data)
    return out


def icp_align(omap: OracleMap, local, init_pose, params: IcpParams, pool: Pool | None = None, trace: bool = False):
    local, init_pose = _f32(local), _pose(init_pose)
    res = IcpResult()
    cap = int(params.max_iterations) + 1
    tp = np.zeros((cap, 3, 4)) if trace else None
    tn = np.zeros(cap, np.uint32) if trace else None
    lib().orc_icp_align(omap.h, local.ctypes.data, local.shape[1], len(local), init_pose.ctypes.data, C.byref(params),
                        C.byref(res), pool.h if pool else None, tp.ctypes.data if trace else None,
                        tn.ctypes.data if trace else None, cap)
    if trace:
        return res, tp, tn
    return res


def scan_register(omap: OracleMap, raw, fp: Filter1Params, init_pose, params: IcpParams, insert=False, cull_dist=0.0,
                  pool: Pool | None = None):
    raw, init_pose = _f32(raw), _pose(init_pose)
    res = IcpResult()
    ms = np.zeros(3)
    lib().orc_scan_register(omap.h, raw.ctypes.data, raw.shape[1], len(raw), C.byref(fp), init_pose.ctypes.data,
                            C.byref(params), int(insert), cull_dist, pool.h if pool else None, C.byref(res),
                            ms.ctypes.data)
    return res, ms


class OracleLidarOdometry:
    """The product's C++ host orchestrator instantiated over the CPU oracle (oracle_lo_capi.cpp) — for
    trajectory-level parity tests and the sequence CPU baseline."""

    def __init__(self, yaml_path_or_text, is_text: bool = False):
        from mola_lidar_odometry_b200.host_api import ScanOutput
        L = lib()
        L.orc_lo_create.restype = _vp
        L.orc_lo_create.argtypes = [C.c_char_p, C.c_int]
        L.orc_lo_destroy.argtypes = [_vp]
        L.orc_lo_on_lidar_t.restype = C.c_int
        L.orc_lo_on_lidar_t.argtypes = [_vp, _vp, _u32, _vp, _u64, C.c_double, C.POINTER(ScanOutput)]
        self._out_t = ScanOutput
        self.h = L.orc_lo_create(str(yaml_path_or_text).encode(), int(is_text))
        if not self.h:
            raise RuntimeError("orc_lo_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_lo_destroy(self.h)
            self.h = None

    def on_lidar(self, pts, stamp: float, t=None):
        pts = _f32(pts)
        out = self._out_t()
        if t is not None:
            t = np.ascontiguousarray(t, dtype=np.float32)
        rc = lib().orc_lo_on_lidar_t(self.h, pts.ctypes.data, pts.shape[1], None if t is None else t.ctypes.data, len(pts),
                                     stamp, C.byref(out))
        if rc != 0:
            raise RuntimeError("orc_lo_on_lidar failed")
        return out


class OracleLidarOdometryFleet:
    """The product's fleet orchestrator (LidarOdometryFleetT) over the CPU oracle: checks on the CPU that lock-step
    batching does not change any per-sequence result."""

    def __init__(self, yaml_path_or_text, n_sequences: int, is_text: bool = False):
        from mola_lidar_odometry_b200.host_api import ScanOutput
        L = lib()
        L.orc_fleet_create.restype = _vp
        L.orc_fleet_create.argtypes = [C.c_char_p, C.c_int, _u32]
        L.orc_fleet_destroy.argtypes = [_vp]
        L.orc_fleet_on_lidar.restype = C.c_int
        L.orc_fleet_on_lidar.argtypes = [_vp, _vp, _u32, _vp, _vp, _vp, C.POINTER(ScanOutput)]
        self._out_t, self.n = ScanOutput, n_sequences
        self.h = L.orc_fleet_create(str(yaml_path_or_text).encode(), int(is_text), n_sequences)
        if not self.h:
            raise RuntimeError("orc_fleet_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fleet_destroy(self.h)
            self.h = None

    def on_lidar(self, clouds, stamps, ts=None, as_arrays: bool = False):
        from mola_lidar_odometry_b200.host_api import _fleet_args, _fleet_outputs
        keep, stride, pts, n, st, tp = _fleet_args(clouds, stamps, ts, _f32)
        out = (self._out_t * self.n)()
        if lib().orc_fleet_on_lidar(self.h, pts, stride, n, st, tp, out) != 0:
            raise RuntimeError("orc_fleet_on_lidar failed")
        del keep
        return _fleet_outputs(out, self.n, as_arrays)
