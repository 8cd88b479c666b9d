"""Python host layer over the C ABI (libmlo_b200.so).  Names follow the reference's domain:
mola::HashedVoxelPointCloud / mola::NDT (local map), mp2p_icp_filters::FilterDecimateVoxels,
mp2p_icp::ICP::align.  Everything here calls the CUDA library; nothing computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import capi
from .capi import (DecimateParams, Filter1Params, IcpParams, IcpResult, MapParams, Profile)


class MloError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mlo error {code}: {msg}")
        self.code = code


def _pts(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError("point cloud must be [n,3] or [n,4] float32")
    return a


def _pose(p) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(p, dtype=np.float64)[..., :3, :4])


class Context:
    """One CUDA device + one stream + one caller thread (mirrors one LidarOdometry worker, LidarOdometry.h:546-549)."""

    def __init__(self, device: int = 0):
        self.lib = capi.load()
        h = C.c_void_p()
        rc = self.lib.mlo_create(device, C.byref(h))
        if rc != 0:
            raise MloError(rc, "mlo_create failed (no sm_100 device?) — there is no CPU fallback")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.mlo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            raise MloError(rc, self.lib.mlo_last_error(self.h).decode())

    @property
    def stream(self) -> int:
        return int(self.lib.mlo_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.mlo_launch_count(self.h))

    def device_info(self):
        name = C.create_string_buffer(128)
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        self.check(self.lib.mlo_device_info(self.h, name, 128, C.byref(sm), C.byref(ma), C.byref(mi)))
        return name.value.decode(), sm.value, (ma.value, mi.value)

    def set_option(self, name: str, value: int):
        """Launch-policy knob (include/mlo_b200.h mlo_set_option): e.g. align_path = 1 forces the large-batch launch sequence."""
        self.check(self.lib.mlo_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int64()
        self.check(self.lib.mlo_get_option(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def icp_log_enable(self, max_records: int):
        """Keep per-iteration records of every following align call (mp2p_icp's ICP log, default.yaml:177-182); 0 = off."""
        self.check(self.lib.mlo_icp_log_enable(self.h, int(max_records)))

    def icp_log(self, problem: int = 0):
        n = C.c_uint32()
        self.check(self.lib.mlo_icp_log_read(self.h, problem, None, 0, C.byref(n)))
        out = (capi.IcpIterationRecord * max(1, n.value))()
        self.check(self.lib.mlo_icp_log_read(self.h, problem, out, n.value, C.byref(n)))
        return list(out)[:n.value]

    def profile_enable(self, on: bool = True):
        self.check(self.lib.mlo_profile_enable(self.h, int(on)))

    def profile_get(self, reset: bool = True) -> Profile:
        p = Profile()
        self.check(self.lib.mlo_profile_get(self.h, C.byref(p), int(reset)))
        return p

    # ---- filters
    def voxel_decimate_first(self, pts, params: DecimateParams) -> np.ndarray:
        pts = _pts(pts)
        idx = np.empty(max(len(pts), 1), np.uint32)
        n = C.c_uint64()
        self.check(self.lib.mlo_voxel_decimate_first(self.h, pts.ctypes.data, pts.shape[1], len(pts), C.byref(params),
                                                     idx.ctypes.data, C.byref(n)))
        return idx[:n.value].copy()

    def filter_1st_pass(self, pts, fp: Filter1Params):
        pts = _pts(pts)
        a = np.empty((max(len(pts), 1), 3), np.float32)
        b = np.empty((max(len(pts), 1), 3), np.float32)
        na, nb = C.c_uint64(), C.c_uint64()
        self.check(self.lib.mlo_filter_1st_pass(self.h, pts.ctypes.data, pts.shape[1], len(pts), C.byref(fp),
                                                a.ctypes.data, C.byref(na), b.ctypes.data, C.byref(nb)))
        return a[:na.value].copy(), b[:nb.value].copy()

    def filter_1st_pass_xyzt(self, pts, t, fp: Filter1Params):
        """1st-pass filter carrying per-point timestamps: returns the two '_skewed' layers as [n,4] x,y,z,t."""
        pts = _pts(pts)
        t = None if t is None else np.ascontiguousarray(t, dtype=np.float32)
        a = np.empty((max(len(pts), 1), 4), np.float32)
        b = np.empty((max(len(pts), 1), 4), np.float32)
        na, nb = C.c_uint64(), C.c_uint64()
        self.check(self.lib.mlo_filter_1st_pass_xyzt(self.h, pts.ctypes.data, pts.shape[1], None if t is None else t.ctypes.data,
                                                     len(pts), C.byref(fp), a.ctypes.data, C.byref(na), b.ctypes.data,
                                                     C.byref(nb)))
        return a[:na.value].copy(), b[:nb.value].copy()

    def deskew(self, xyzt, twist) -> np.ndarray:
        xyzt = np.ascontiguousarray(xyzt, dtype=np.float32)
        assert xyzt.ndim == 2 and xyzt.shape[1] == 4
        tw = np.ascontiguousarray(twist, dtype=np.float64)
        out = np.empty((max(len(xyzt), 1), 3), np.float32)
        self.check(self.lib.mlo_deskew(self.h, xyzt.ctypes.data, len(xyzt), tw.ctypes.data, out.ctypes.data))
        return out[:len(xyzt)]

    # ---- ICP
    def icp_align(self, local, lmap: "LocalMap", init_pose, params: IcpParams) -> IcpResult:
        local, init_pose = _pts(local), _pose(init_pose)
        res = IcpResult()
        self.check(self.lib.mlo_icp_align(self.h, local.ctypes.data, local.shape[1], len(local), lmap.h,
                                          init_pose.ctypes.data, C.byref(params), C.byref(res)))
        return res

    def icp_align_soa(self, x, y, z, lmap: "LocalMap", init_pose, params: IcpParams) -> IcpResult:
        x, y, z = (np.ascontiguousarray(v, np.float32) for v in (x, y, z))
        init_pose = _pose(init_pose)
        res = IcpResult()
        self.check(self.lib.mlo_icp_align_soa(self.h, x.ctypes.data, y.ctypes.data, z.ctypes.data, len(x), lmap.h,
                                              init_pose.ctypes.data, C.byref(params), C.byref(res)))
        return res

    @staticmethod
    def _concat(clouds: Sequence[np.ndarray]):
        clouds = [_pts(c) for c in clouds]
        stride = clouds[0].shape[1]
        assert all(c.shape[1] == stride for c in clouds)
        offs = np.zeros(len(clouds) + 1, np.uint64)
        offs[1:] = np.cumsum([len(c) for c in clouds])
        return np.ascontiguousarray(np.concatenate(clouds, 0)), offs, stride

    @staticmethod
    def _params_array(params: Sequence[IcpParams]):
        arr = (IcpParams * len(params))()
        for i, p in enumerate(params):
            C.memmove(C.byref(arr, i * C.sizeof(IcpParams)), C.byref(p), C.sizeof(IcpParams))
        return arr

    def icp_align_batch(self, locals_: Sequence[np.ndarray], lmap: "LocalMap", init_poses, params: Sequence[IcpParams]):
        flat, offs, stride = self._concat(locals_)
        B = len(locals_)
        init_poses = _pose(init_poses).reshape(B, 12)
        parr = self._params_array(params)
        out = (IcpResult * B)()
        self.check(self.lib.mlo_icp_align_batch(self.h, B, flat.ctypes.data, stride, offs.ctypes.data, lmap.h,
                                                init_poses.ctypes.data, parr, out))
        return list(out)

    def scan_register(self, lmap: "LocalMap", raw, fp: Filter1Params, init_pose, params: IcpParams, insert: bool = False,
                      cull_dist: float = 0.0) -> IcpResult:
        raw, init_pose = _pts(raw), _pose(init_pose)
        res = IcpResult()
        self.check(self.lib.mlo_scan_register(self.h, lmap.h, raw.ctypes.data, raw.shape[1], len(raw), C.byref(fp),
                                              init_pose.ctypes.data, C.byref(params), int(insert), cull_dist,
                                              C.byref(res)))
        return res

    def scan_register_batch(self, lmap: "LocalMap", raws: Sequence[np.ndarray], fps: Sequence[Filter1Params], init_poses,
                            params: Sequence[IcpParams]):
        flat, offs, stride = self._concat(raws)
        return self.scan_register_batch_flat(lmap, flat, offs, stride, fps, init_poses, params)

    def scan_register_batch_flat(self, lmap, flat, offs, stride, fps, init_poses, params):
        B = len(offs) - 1
        init_poses = _pose(init_poses).reshape(B, 12)
        farr = (Filter1Params * B)(*fps)
        parr = self._params_array(params)
        out = (IcpResult * B)()
        self.check(self.lib.mlo_scan_register_batch(self.h, lmap.h, B, flat.ctypes.data, stride, offs.ctypes.data, farr,
                                                    init_poses.ctypes.data, parr, out))
        return list(out)

    def stage_upload_async(self, slot: int, flat: np.ndarray, offs: np.ndarray, stride: int):
        """Enqueue the H2D of one batch (flat float32 array, pinned memory recommended) into staging slot 0/1."""
        self.check(self.lib.mlo_stage_upload_async(self.h, slot, flat.ctypes.data, stride, len(offs) - 1, offs.ctypes.data))

    def scan_register_batch_staged(self, lmap, slot: int, fps, init_poses, params):
        B = len(fps)
        init_poses = _pose(init_poses).reshape(B, 12)
        farr = (Filter1Params * B)(*fps)
        parr = self._params_array(params)
        out = (IcpResult * B)()
        self.check(self.lib.mlo_scan_register_batch_staged(self.h, lmap.h, slot, farr, init_poses.ctypes.data, parr, out))
        return list(out)

    def upload_batch(self, clouds: Sequence[np.ndarray]) -> "DeviceClouds":
        flat, offs, stride = self._concat(clouds)
        h = C.c_void_p()
        self.check(self.lib.mlo_dcloud_upload_batch(self.h, flat.ctypes.data, stride, len(clouds), offs.ctypes.data,
                                                    C.byref(h)))
        return DeviceClouds(self, h, len(clouds))

    def scan_register_batch_resident(self, lmap, dclouds: "DeviceClouds", fps, init_poses, params):
        B = dclouds.n_clouds
        init_poses = _pose(init_poses).reshape(B, 12)
        farr = (Filter1Params * B)(*fps)
        parr = self._params_array(params)
        out = (IcpResult * B)()
        self.check(self.lib.mlo_scan_register_batch_resident(self.h, lmap.h, dclouds.h, farr, init_poses.ctypes.data,
                                                             parr, out))
        return list(out)

    def icp_align_batch_resident(self, dclouds: "DeviceClouds", lmap, init_poses, params):
        B = dclouds.n_clouds
        init_poses = _pose(init_poses).reshape(B, 12)
        parr = self._params_array(params)
        out = (IcpResult * B)()
        self.check(self.lib.mlo_icp_align_batch_resident(self.h, dclouds.h, lmap.h, init_poses.ctypes.data, parr, out))
        return list(out)


class DeviceClouds:
    def __init__(self, ctx: Context, h, n_clouds: int):
        self.ctx, self.h, self.n_clouds = ctx, h, n_clouds

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.mlo_dcloud_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ScanSet:
    """Device-resident layers of n_slots scans (mlo_scanset_*): filter -> (deskew) -> align -> insert, each ONE device
    pass over all listed scans, every scan against its own local map (fleets of independent sequences)."""

    def __init__(self, ctx: Context, n_slots: int):
        self.ctx, self.n_slots = ctx, n_slots
        h = C.c_void_p()
        ctx.check(ctx.lib.mlo_scanset_create(ctx.h, n_slots, C.byref(h)))
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.mlo_scanset_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def filter(self, slots: Sequence[int], clouds: Sequence[np.ndarray], fps: Sequence[Filter1Params], ts=None):
        clouds = [_pts(c) for c in clouds]
        stride = clouds[0].shape[1] if clouds else 3
        assert all(c.shape[1] == stride for c in clouds)
        ts = [None] * len(clouds) if ts is None else [None if t is None else np.ascontiguousarray(t, dtype=np.float32) for t in ts]
        jobs = (capi.ScanJob * max(1, len(clouds)))()
        for j, (sl, c, fp, t) in enumerate(zip(slots, clouds, fps, ts)):
            jobs[j].slot, jobs[j].pts, jobs[j].n, jobs[j].fp = sl, c.ctypes.data, len(c), fp
            jobs[j].t = None if t is None else t.ctypes.data
        info = (capi.ScanInfo * max(1, len(clouds)))()
        self.ctx.check(self.ctx.lib.mlo_scanset_filter(self.h, len(clouds), jobs, stride, info))
        return list(info)[:len(clouds)]

    def deskew(self, slots: Sequence[int], twists):
        sl = np.ascontiguousarray(slots, dtype=np.uint32)
        tw = np.ascontiguousarray(twists, dtype=np.float64).reshape(len(sl), 6)
        info = (capi.ScanInfo * max(1, len(sl)))()
        self.ctx.check(self.ctx.lib.mlo_scanset_deskew(self.h, len(sl), sl.ctypes.data, tw.ctypes.data, info))
        return list(info)[:len(sl)]

    def align(self, slots: Sequence[int], maps: Sequence["LocalMap"], init_poses, params: Sequence[IcpParams]):
        n = len(slots)
        jobs = (capi.AlignJob * max(1, n))()
        init = _pose(init_poses).reshape(n, 12)
        for j in range(n):
            jobs[j].slot, jobs[j].map, jobs[j].params = slots[j], maps[j].h, params[j]
            jobs[j].init_pose_3x4[:] = init[j].tolist()
        out = (IcpResult * max(1, n))()
        self.ctx.check(self.ctx.lib.mlo_scanset_align(self.h, n, jobs, out))
        return list(out)[:n]

    def insert(self, slots: Sequence[int], maps: Sequence["LocalMap"], poses, cull=None):
        n = len(slots)
        jobs = (capi.InsertJob * max(1, n))()
        ps = _pose(poses).reshape(n, 12)
        for j in range(n):
            jobs[j].slot, jobs[j].map = slots[j], maps[j].h
            jobs[j].pose_3x4[:] = ps[j].tolist()
            jobs[j].cull_farther_than = 0.0 if cull is None else float(cull[j])
        out = (capi.MapCounts * max(1, n))()
        self.ctx.check(self.ctx.lib.mlo_scanset_insert(self.h, n, jobs, out))
        return [(o.n_voxels, o.n_points) for o in list(out)[:n]]

    def download(self, slot: int, layer: int) -> np.ndarray:
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.mlo_scanset_download(self.h, slot, layer, None, 0, C.byref(n)))
        out = np.zeros((n.value, 3), dtype=np.float32)
        if n.value:
            self.ctx.check(self.ctx.lib.mlo_scanset_download(self.h, slot, layer, out.ctypes.data, n.value, C.byref(n)))
        return out


class LocalMap:
    """Hash-voxel local map in HBM: mola::HashedVoxelPointCloud (kind 0) or mola::NDT (kind 1)."""

    def __init__(self, ctx: Context, voxel_size: float = 1.0, max_points_per_voxel: int = 20,
                 min_distance_between_points: float = 0.0, capacity_voxels: int = 1 << 18, kind: int = capi.MAP_POINTS,
                 max_eigen_ratio_for_planes: float = 0.05, min_points_for_plane: int = 5):
        self.ctx = ctx
        p = MapParams(kind, voxel_size, max_points_per_voxel, min_distance_between_points, max_eigen_ratio_for_planes,
                      min_points_for_plane, capacity_voxels)
        h = C.c_void_p()
        ctx.check(ctx.lib.mlo_map_create(ctx.h, C.byref(p), C.byref(h)))
        self.h = h
        self.params = p

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.mlo_map_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        self.ctx.check(self.ctx.lib.mlo_map_clear(self.h))

    def insert(self, pts, pose):
        pts, pose = _pts(pts), _pose(pose)
        self.ctx.check(self.ctx.lib.mlo_map_insert(self.h, pts.ctypes.data, pts.shape[1], len(pts), pose.ctypes.data))

    def insert_soa(self, x, y, z, pose):
        x, y, z = (np.ascontiguousarray(v, np.float32) for v in (x, y, z))
        pose = _pose(pose)
        self.ctx.check(self.ctx.lib.mlo_map_insert_soa(self.h, x.ctypes.data, y.ctypes.data, z.ctypes.data, len(x),
                                                       pose.ctypes.data))

    def cull(self, sensor_xyz, dist: float):
        s = np.ascontiguousarray(sensor_xyz, dtype=np.float64)
        self.ctx.check(self.ctx.lib.mlo_map_cull(self.h, s.ctypes.data, dist))

    def stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.ctx.check(self.ctx.lib.mlo_map_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def nn_single(self, q):
        q = _pts(q)
        n = len(q)
        xyz, d2, f = np.empty((max(n, 1), 3), np.float32), np.empty(max(n, 1), np.float32), np.empty(max(n, 1), np.uint8)
        self.ctx.check(self.ctx.lib.mlo_map_nn_single(self.h, q.ctypes.data, q.shape[1], n, xyz.ctypes.data,
                                                      d2.ctypes.data, f.ctypes.data))
        return xyz[:n], d2[:n], f[:n].astype(bool)

    def nn_plane(self, q):
        q = _pts(q)
        n = len(q)
        mean, nrm = np.empty((max(n, 1), 3), np.float32), np.empty((max(n, 1), 3), np.float32)
        d, f = np.empty(max(n, 1), np.float32), np.empty(max(n, 1), np.uint8)
        self.ctx.check(self.ctx.lib.mlo_map_nn_plane(self.h, q.ctypes.data, q.shape[1], n, mean.ctypes.data,
                                                     nrm.ctypes.data, d.ctypes.data, f.ctypes.data))
        return mean[:n], nrm[:n], d[:n], f[:n].astype(bool)

    def export(self):
        nv, npts = C.c_uint64(), C.c_uint64()
        self.ctx.check(self.ctx.lib.mlo_map_export(self.h, None, None, None, 0, 0, C.byref(nv), C.byref(npts)))
        keys, cnt = np.empty((max(nv.value, 1), 3), np.int32), np.empty(max(nv.value, 1), np.uint32)
        xyz = np.empty((max(npts.value, 1), 3), np.float32)
        self.ctx.check(self.ctx.lib.mlo_map_export(self.h, keys.ctypes.data, cnt.ctypes.data, xyz.ctypes.data, nv.value,
                                                   npts.value, C.byref(nv), C.byref(npts)))
        return keys[:nv.value], cnt[:nv.value], xyz[:npts.value]
