"""ctypes view of include/mlo_b200_host.h: the C++ host layer (pipeline YAML + LidarOdometry caller contract)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import capi
from .api import Context, MloError, _pts

PIPELINES = Path(__file__).resolve().parent.parent / "pipelines"


class ScanOutput(C.Structure):
    _fields_ = [("processed", C.c_int32), ("icp_ran", C.c_int32), ("icp_good", C.c_int32), ("map_updated", C.c_int32),
                ("pose_3x4", C.c_double * 12), ("quality", C.c_double), ("sigma", C.c_double),
                ("est_max_range", C.c_double), ("icp_iterations", C.c_uint32), ("icp_runs", C.c_uint32),
                ("termination", C.c_int32), ("n_map_layer", C.c_uint64), ("n_icp_layer", C.c_uint64),
                ("icp_had_prior", C.c_int32), ("has_motion_model", C.c_int32), ("prior_info_trace", C.c_double)]

    @property
    def pose(self):
        return np.array(self.pose_3x4[:]).reshape(3, 4)


_vp, _u32, _u64 = C.c_void_p, C.c_uint32, C.c_uint64
HOST_SIGNATURES = {
    "mlo_lo_create": (C.c_int, [_vp, C.c_char_p, C.c_int, C.POINTER(_vp)]),
    "mlo_lo_destroy": (None, [_vp]),
    "mlo_lo_last_error": (C.c_char_p, [_vp]),
    "mlo_lo_on_lidar": (C.c_int, [_vp, _vp, _u32, _u64, C.c_double, C.POINTER(ScanOutput)]),
    "mlo_lo_on_lidar_t": (C.c_int, [_vp, _vp, _u32, _vp, _u64, C.c_double, C.POINTER(ScanOutput)]),
    "mlo_lo_trajectory": (C.c_int, [_vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "mlo_lo_reset": (C.c_int, [_vp]),
    "mlo_fleet_create": (C.c_int, [_vp, C.c_char_p, C.c_int, _u32, C.POINTER(_vp)]),
    "mlo_fleet_destroy": (None, [_vp]),
    "mlo_fleet_last_error": (C.c_char_p, [_vp]),
    "mlo_fleet_on_lidar": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp, C.POINTER(ScanOutput)]),
    "mlo_fleet_prefetch": (C.c_int, [_vp, _vp, _u32, _vp]),
    "mlo_fleet_phase_times": (C.c_int, [_vp, _vp, C.c_int]),
    "mlo_fleet_trajectory": (C.c_int, [_vp, _u32, _vp, _vp, _u64, C.POINTER(_u64)]),
    "mlo_host_last_error": (C.c_char_p, []),
    "mlo_host_icp_tables": (C.c_int, [C.c_char_p, C.c_double, _u32, _vp, _vp, _vp, C.POINTER(capi.IcpParams)]),
    "mlo_host_filter1": (C.c_int, [C.c_char_p, C.c_double, C.c_double, C.POINTER(capi.Filter1Params)]),
    "mlo_host_mapdef": (C.c_int, [C.c_char_p, C.c_double, C.POINTER(capi.MapParams), C.POINTER(C.c_float)]),
    "mlo_host_eval_formula": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), _vp, _u32, C.POINTER(C.c_double)]),
}
_BOUND = False


def lib():
    global _BOUND
    L = capi.load()
    if not _BOUND:
        for name, (res, args) in HOST_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _BOUND = True
    return L


def _host_check(rc):
    if rc != 0:
        raise MloError(rc, lib().mlo_host_last_error().decode())


def icp_tables(yaml_text: str, sigma: float, n_it: int):
    t1, t2, t3 = np.zeros(n_it), np.zeros(n_it), np.zeros(n_it)
    sc = capi.IcpParams()
    _host_check(lib().mlo_host_icp_tables(yaml_text.encode(), sigma, n_it, t1.ctypes.data, t2.ctypes.data, t3.ctypes.data,
                                          C.byref(sc)))
    return t1, t2, t3, sc


def filter1(yaml_text: str, est: float, inst: float) -> capi.Filter1Params:
    f = capi.Filter1Params()
    _host_check(lib().mlo_host_filter1(yaml_text.encode(), est, inst, C.byref(f)))
    return f


def mapdef(yaml_text: str, est: float):
    m, cull = capi.MapParams(), C.c_float()
    _host_check(lib().mlo_host_mapdef(yaml_text.encode(), est, C.byref(m), C.byref(cull)))
    return m, cull.value


def eval_formula(expr: str, **variables) -> float:
    names = (C.c_char_p * len(variables))(*[k.encode() for k in variables])
    vals = np.array(list(variables.values()), dtype=np.float64)
    out = C.c_double()
    _host_check(lib().mlo_host_eval_formula(expr.encode(), names, vals.ctypes.data, len(variables), C.byref(out)))
    return out.value


class LidarOdometry:
    """mola::LidarOdometry (hot-path subset) on one GPU context: initialize(yaml) then on_lidar(cloud, stamp)."""

    def __init__(self, ctx: Context, yaml_path_or_text, is_text: bool = False):
        self.ctx = ctx
        h = _vp()
        rc = lib().mlo_lo_create(ctx.h, str(yaml_path_or_text).encode(), int(is_text), C.byref(h))
        if rc != 0:
            raise MloError(rc, lib().mlo_lo_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().mlo_lo_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def on_lidar(self, pts, stamp: float, t=None) -> ScanOutput:
        pts = _pts(pts)
        out = ScanOutput()
        if t is not None:
            t = np.ascontiguousarray(t, dtype=np.float32)
            assert len(t) == len(pts)
        rc = lib().mlo_lo_on_lidar_t(self.h, pts.ctypes.data, pts.shape[1], None if t is None else t.ctypes.data, len(pts),
                                     stamp, C.byref(out))
        if rc != 0:
            raise MloError(rc, lib().mlo_lo_last_error(self.h).decode())
        return out

    def trajectory(self):
        n = _u64()
        lib().mlo_lo_trajectory(self.h, None, None, 0, C.byref(n))
        st, ps = np.zeros(n.value), np.zeros((n.value, 3, 4))
        lib().mlo_lo_trajectory(self.h, st.ctypes.data, ps.ctypes.data, n.value, C.byref(n))
        return st, ps


# numpy view of an array of ScanOutput (same layout as the C struct): lets a driver read S results without S x 13
# ctypes attribute accesses
SCAN_OUTPUT_DTYPE = np.dtype({"names": [f[0] for f in ScanOutput._fields_],
                              "formats": [np.int32, np.int32, np.int32, np.int32, (np.float64, (3, 4)), np.float64, np.float64,
                                          np.float64, np.uint32, np.uint32, np.int32, np.uint64, np.uint64, np.int32, np.int32, np.float64],
                              "offsets": [getattr(ScanOutput, f[0]).offset for f in ScanOutput._fields_],
                              "itemsize": C.sizeof(ScanOutput)})


def _addr(a: np.ndarray) -> int:
    return a.__array_interface__["data"][0]


def _as_cloud(c, as_pts):
    if type(c) is np.ndarray and c.dtype == np.float32 and c.ndim == 2 and c.shape[1] in (3, 4) and c.flags.c_contiguous:
        return c
    return as_pts(c)


def _fleet_args(clouds, stamps, ts, as_pts):
    """ctypes argument arrays of one lock step; a cloud of None leaves that sequence idle."""
    S = len(clouds)
    keep = [None if c is None else _as_cloud(c, as_pts) for c in clouds]
    stride = next((c.shape[1] for c in keep if c is not None), 3)
    assert all(c is None or c.shape[1] == stride for c in keep)
    pts = (_vp * S)(*[None if c is None else _addr(c) for c in keep])
    n = (_u64 * S)(*[0 if c is None else c.shape[0] for c in keep])
    st = (C.c_double * S)(*stamps)
    tp = None
    if ts is not None:
        tk = [None if t is None else np.ascontiguousarray(t, dtype=np.float32) for t in ts]
        keep.append(tk)
        tp = (_vp * S)(*[None if t is None else _addr(t) for t in tk])
    return keep, stride, pts, n, st, tp


def _fleet_outputs(out, n, as_arrays):
    """The ScanOutput array of one lock step as a list of structs, or (as_arrays) as ONE numpy structured array
    (fields as in ScanOutput; `pose_3x4` is [n, 3, 4])."""
    if not as_arrays:
        return list(out)
    return np.frombuffer(out, dtype=SCAN_OUTPUT_DTYPE, count=n).copy()


class LidarOdometryFleet:
    """n independent mola::LidarOdometry instances advanced in lock step on one GPU context (mlo_fleet_*): one filter
    pass, one align pass over per-sequence local maps and one insert pass per step."""

    def __init__(self, ctx: Context, yaml_path_or_text, n_sequences: int, is_text: bool = False):
        self.ctx, self.n = ctx, n_sequences
        self._arg_cache = {}
        h = _vp()
        rc = lib().mlo_fleet_create(ctx.h, str(yaml_path_or_text).encode(), int(is_text), n_sequences, C.byref(h))
        if rc != 0:
            raise MloError(rc, lib().mlo_fleet_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().mlo_fleet_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def on_lidar(self, clouds, stamps, ts=None, as_arrays: bool = False):
        """clouds[i] -> sequence i (None = idle).  Returns a list of ScanOutput, or with as_arrays one numpy structured
        array of dtype SCAN_OUTPUT_DTYPE."""
        assert len(clouds) == self.n and len(stamps) == self.n
        cached = self._arg_cache.pop(tuple(map(id, clouds)), None) if ts is None else None
        if cached is not None:
            keep, stride, pts, n = cached             # the argument arrays built by prefetch() for these very objects
            st, tp = (C.c_double * self.n)(*stamps), None
        else:
            keep, stride, pts, n, st, tp = _fleet_args(clouds, stamps, ts, _pts)
        out = (ScanOutput * self.n)()
        rc = lib().mlo_fleet_on_lidar(self.h, pts, stride, n, st, tp, out)
        if rc != 0:
            raise MloError(rc, lib().mlo_fleet_last_error(self.h).decode())
        del keep
        return _fleet_outputs(out, self.n, as_arrays)

    def prefetch(self, clouds):
        """Announce the clouds of the NEXT on_lidar call (the same array objects must be passed then)."""
        keep, stride, pts, n, _, _ = _fleet_args(clouds, (0.0,) * self.n, None, _pts)
        # Reusable by on_lidar only when the caller's own arrays are the ones pointed to (no conversion copy was made):
        # the cache then holds references to exactly those objects, so their ids cannot be recycled meanwhile.
        same = all(k is c for k, c in zip(keep, clouds))
        self._prefetched = keep                       # (the announced buffers must stay alive until they are consumed)
        if same:
            if len(self._arg_cache) >= 2:             # the announcement for step k+1 is made before step k runs: keep two
                self._arg_cache.pop(next(iter(self._arg_cache)))
            self._arg_cache[tuple(map(id, clouds))] = (keep, stride, pts, n)
        rc = lib().mlo_fleet_prefetch(self.h, pts, stride, n)
        if rc != 0:
            raise MloError(rc, lib().mlo_fleet_last_error(self.h).decode())

    def phase_times(self, reset: bool = True) -> dict:
        a = np.zeros(8)
        lib().mlo_fleet_phase_times(self.h, a.ctypes.data, int(reset))
        steps = max(a[6], 1.0)
        names = ["host_before_filter", "filter", "deskew", "align", "host_after_icp", "insert"]
        return {k + "_ms_per_step": float(a[i] / steps) for i, k in enumerate(names)}

    def trajectory(self, sequence: int):
        n = _u64()
        lib().mlo_fleet_trajectory(self.h, sequence, None, None, 0, C.byref(n))
        st, ps = np.zeros(n.value), np.zeros((n.value, 3, 4))
        lib().mlo_fleet_trajectory(self.h, sequence, st.ctypes.data, ps.ctypes.data, n.value, C.byref(n))
        return st, ps
