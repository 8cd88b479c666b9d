"""ctypes view of include/mlo_b200.h (the C ABI of libmlo_b200.so).

The structs here are byte-compatible with the header; `load()` fails loudly when the CUDA library
is missing or unloadable — there is no CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["MLO_B200_LIB"]) if os.environ.get("MLO_B200_LIB") else PKG / "libmlo_b200.so"  # (override: scratch builds)

MLO_OK = 0
TERM_NAMES = {0: "Undefined", 1: "NoPairings", 2: "SolverError", 3: "MaxIterations", 4: "Stalled", 5: "HookRequest"}
MATCHER_PT2PT, MATCHER_PT2PL = 1, 2
SOLVER_GN, SOLVER_HORN = 0, 1
KERNEL_NONE, KERNEL_GM, KERNEL_CAUCHY = 0, 1, 2
MAP_POINTS, MAP_NDT = 0, 1

c_dp = C.POINTER(C.c_double)


class MapParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("voxel_size", C.c_float), ("max_points_per_voxel", C.c_uint32),
                ("min_distance_between_points", C.c_float), ("max_eigen_ratio_for_planes", C.c_float),
                ("min_points_for_plane", C.c_uint32), ("capacity_voxels", C.c_uint64)]


class DecimateParams(C.Structure):
    _fields_ = [("voxel_filter_resolution", C.c_float), ("minimum_input_points_to_filter", C.c_uint32),
                ("use_range", C.c_int32), ("range_min", C.c_float), ("range_max", C.c_float),
                ("use_bbox_outside", C.c_int32), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3)]


class Filter1Params(C.Structure):
    _fields_ = [("for_map", DecimateParams), ("for_icp", DecimateParams)]


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_uint32), ("min_abs_step_trans", C.c_double), ("min_abs_step_rot", C.c_double),
                ("solver", C.c_int32), ("gn_max_iterations", C.c_uint32), ("gn_min_delta", C.c_double),
                ("robust_kernel", C.c_int32), ("matcher_mask", C.c_uint32), ("table_len", C.c_uint32),
                ("pt2pt_threshold_by_iter", c_dp), ("pt2pl_threshold_by_iter", c_dp), ("kernel_param_by_iter", c_dp),
                ("threshold_angular_deg", C.c_double), ("pt2pt_weight", C.c_double), ("pt2pl_weight", C.c_double),
                ("has_prior", C.c_int32), ("prior_pose_3x4", C.c_double * 12), ("prior_info_6x6", C.c_double * 36),
                ("hook_enabled", C.c_int32), ("hook_min_trans", C.c_double), ("hook_min_rot_rad", C.c_double),
                ("hook_checkpoint_pose_3x4", C.c_double * 12)]


class IcpResult(C.Structure):
    _fields_ = [("pose_3x4", C.c_double * 12), ("cov_6x6", C.c_double * 36), ("quality", C.c_double),
                ("n_iterations", C.c_uint32), ("termination", C.c_int32), ("n_pairings", C.c_uint64),
                ("n_potential_pairings", C.c_uint64), ("n_query_iterations", C.c_uint64),
                ("n_candidate_points", C.c_uint64)]

    @property
    def pose(self) -> np.ndarray:
        return np.array(self.pose_3x4[:], dtype=np.float64).reshape(3, 4)

    @property
    def cov(self) -> np.ndarray:
        return np.array(self.cov_6x6[:], dtype=np.float64).reshape(6, 6)


class IcpIterationRecord(C.Structure):
    _fields_ = [("iteration", C.c_uint32), ("n_pairings", C.c_uint32), ("pose_3x4", C.c_double * 12),
                ("threshold_pt2pt", C.c_double), ("threshold_pt2pl", C.c_double), ("kernel_param", C.c_double),
                ("step_trans", C.c_double), ("step_rot", C.c_double), ("termination", C.c_int32), ("pad", C.c_int32)]

    @property
    def pose(self) -> np.ndarray:
        return np.array(self.pose_3x4[:], dtype=np.float64).reshape(3, 4)


class Profile(C.Structure):
    _fields_ = [("filter_1st_ms", C.c_double), ("run_icp_ms", C.c_double), ("update_local_map_ms", C.c_double),
                ("nn_kernel_ms", C.c_double), ("nn_kernel_launches", C.c_uint64), ("nn_query_iterations", C.c_uint64),
                ("nn_candidate_points", C.c_uint64), ("nn_blocks", C.c_uint64)]


class ScanJob(C.Structure):
    _fields_ = [("slot", C.c_uint32), ("pts", C.c_void_p), ("t", C.c_void_p), ("n", C.c_uint64), ("fp", Filter1Params)]


class ScanInfo(C.Structure):
    _fields_ = [("n_map", C.c_uint64), ("n_icp", C.c_uint64), ("icp_min", C.c_float * 3), ("icp_max", C.c_float * 3)]


class AlignJob(C.Structure):
    _fields_ = [("slot", C.c_uint32), ("map", C.c_void_p), ("init_pose_3x4", C.c_double * 12), ("params", IcpParams)]


class InsertJob(C.Structure):
    _fields_ = [("slot", C.c_uint32), ("map", C.c_void_p), ("pose_3x4", C.c_double * 12), ("cull_farther_than", C.c_float)]


class MapCounts(C.Structure):
    _fields_ = [("n_voxels", C.c_uint64), ("n_points", C.c_uint64)]


def decimate_params(resolution: float, min_points: int = 2000, range_minmax=None, bbox_outside=None) -> DecimateParams:
    p = DecimateParams()
    p.voxel_filter_resolution = resolution
    p.minimum_input_points_to_filter = min_points
    if range_minmax is not None:
        p.use_range = 1
        p.range_min, p.range_max = range_minmax
    if bbox_outside is not None:
        p.use_bbox_outside = 1
        p.bbox_min[:] = list(bbox_outside[0])
        p.bbox_max[:] = list(bbox_outside[1])
    return p


def filter1_default(est_max_range: float, inst_max_range: float | None = None) -> Filter1Params:
    """observations_filter_1st_pass with the formulas of pipelines/lidar3d-default.yaml:285-319."""
    R = float(est_max_range)
    Ri = float(inst_max_range if inst_max_range is not None else est_max_range)
    f = Filter1Params()
    f.for_map = decimate_params(max(0.20, 0.55e-2 * R))
    f.for_icp = decimate_params(max(0.60, 1.6e-2 * R), 2000, (max(1.0, 0.03 * R), 1.2 * R),
                                ((-0.20 * Ri, -0.20 * Ri, 0.01 * Ri), (0.20 * Ri, 0.20 * Ri, 0.10 * Ri)))
    return f


class IcpParamsOwner:
    """An IcpParams plus the numpy tables it points to (keeps them alive)."""

    def __init__(self, sigma: float = 2.0, max_iterations: int = 300, pipeline: str = "default"):
        it = np.arange(max(1, min(max_iterations, 64)), dtype=np.float64)
        # default.yaml:190,198 — formulas over ADAPTIVE_THRESHOLD_SIGMA and ICP_ITERATION (constant after it=30)
        base = np.maximum(sigma, 2.0 * sigma - (2.0 * sigma - 0.5 * sigma) * it / 30.0)
        self.thr_pt2pt = np.ascontiguousarray(2.0 * base)
        self.kparam = np.ascontiguousarray(0.5 * base)
        self.thr_pt2pl = np.ascontiguousarray(np.full_like(it, 1.0 * sigma))  # ndt.yaml:197
        p = IcpParams()
        p.max_iterations = max_iterations
        p.min_abs_step_trans, p.min_abs_step_rot = (1e-4, 5e-5) if pipeline == "default" else (5e-4, 5e-4)
        p.solver = SOLVER_GN
        p.gn_max_iterations = 2 if pipeline == "default" else 1
        p.gn_min_delta = 1e-7
        p.robust_kernel = KERNEL_GM
        p.matcher_mask = MATCHER_PT2PT if pipeline == "default" else (MATCHER_PT2PL | MATCHER_PT2PT)
        p.pt2pt_weight = p.pt2pl_weight = 1.0
        self.p = p
        self.refresh()

    def refresh(self):
        p = self.p
        p.table_len = len(self.thr_pt2pt)
        p.pt2pt_threshold_by_iter = self.thr_pt2pt.ctypes.data_as(c_dp)
        p.pt2pl_threshold_by_iter = self.thr_pt2pl.ctypes.data_as(c_dp)
        p.kernel_param_by_iter = self.kparam.ctypes.data_as(c_dp)

    def set_tables(self, thr_pt2pt, kparam, thr_pt2pl=None):
        self.thr_pt2pt = np.ascontiguousarray(thr_pt2pt, dtype=np.float64)
        self.kparam = np.ascontiguousarray(kparam, dtype=np.float64)
        self.thr_pt2pl = np.ascontiguousarray(thr_pt2pl if thr_pt2pl is not None else np.zeros_like(self.kparam),
                                              dtype=np.float64)
        assert len(self.thr_pt2pt) == len(self.kparam) == len(self.thr_pt2pl)
        self.refresh()

    def set_prior(self, pose34, info66):
        self.p.has_prior = 1
        self.p.prior_pose_3x4[:] = list(np.asarray(pose34, dtype=np.float64).reshape(-1))
        self.p.prior_info_6x6[:] = list(np.asarray(info66, dtype=np.float64).reshape(-1))

    def set_hook(self, checkpoint34, min_trans=0.15, min_rot_deg=0.75):
        self.p.hook_enabled = 1
        self.p.hook_min_trans = min_trans
        self.p.hook_min_rot_rad = np.deg2rad(min_rot_deg)
        self.p.hook_checkpoint_pose_3x4[:] = list(np.asarray(checkpoint34, dtype=np.float64).reshape(-1))


_vp, _u32, _u64, _f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float
_SIGNATURES = {
    "mlo_abi_version": (C.c_int, []),
    "mlo_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "mlo_destroy": (None, [_vp]),
    "mlo_last_error": (C.c_char_p, [_vp]),
    "mlo_device_info": (C.c_int, [_vp, C.c_char_p, _u32, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mlo_stream": (_vp, [_vp]),
    "mlo_launch_count": (_u64, [_vp]),
    "mlo_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "mlo_get_option": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int64)]),
    "mlo_map_create": (C.c_int, [_vp, C.POINTER(MapParams), C.POINTER(_vp)]),
    "mlo_map_destroy": (None, [_vp]),
    "mlo_map_clear": (C.c_int, [_vp]),
    "mlo_map_insert": (C.c_int, [_vp, _vp, _u32, _u64, _vp]),
    "mlo_map_insert_soa": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _vp]),
    "mlo_map_cull": (C.c_int, [_vp, _vp, _f]),
    "mlo_map_nn_single": (C.c_int, [_vp, _vp, _u32, _u64, _vp, _vp, _vp]),
    "mlo_map_nn_plane": (C.c_int, [_vp, _vp, _u32, _u64, _vp, _vp, _vp, _vp]),
    "mlo_map_stats": (C.c_int, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "mlo_map_export": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "mlo_voxel_index": (C.c_int32, [_f, _f]),
    "mlo_se3_exp": (None, [_vp, _vp]),
    "mlo_se3_log": (None, [_vp, _vp]),
    "mlo_se3_right_jacobian_inv": (None, [_vp, _vp]),
    "mlo_cov_tangent_to_ypr": (None, [_vp, _vp, _vp]),
    "mlo_voxel_decimate_first": (C.c_int, [_vp, _vp, _u32, _u64, C.POINTER(DecimateParams), _vp, C.POINTER(_u64)]),
    "mlo_filter_1st_pass": (C.c_int, [_vp, _vp, _u32, _u64, C.POINTER(Filter1Params), _vp, C.POINTER(_u64), _vp,
                                      C.POINTER(_u64)]),
    "mlo_filter_1st_pass_xyzt": (C.c_int, [_vp, _vp, _u32, _vp, _u64, C.POINTER(Filter1Params), _vp, C.POINTER(_u64), _vp,
                                           C.POINTER(_u64)]),
    "mlo_deskew": (C.c_int, [_vp, _vp, _u64, _vp, _vp]),
    "mlo_icp_params_default": (None, [C.POINTER(IcpParams)]),
    "mlo_icp_log_enable": (C.c_int, [_vp, _u32]),
    "mlo_icp_log_read": (C.c_int, [_vp, _u32, _vp, _u32, C.POINTER(_u32)]),
    "mlo_icp_align": (C.c_int, [_vp, _vp, _u32, _u64, _vp, _vp, C.POINTER(IcpParams), C.POINTER(IcpResult)]),
    "mlo_icp_align_soa": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, C.POINTER(IcpParams), C.POINTER(IcpResult)]),
    "mlo_icp_align_batch": (C.c_int, [_vp, _u32, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "mlo_scan_register": (C.c_int, [_vp, _vp, _vp, _u32, _u64, C.POINTER(Filter1Params), _vp, C.POINTER(IcpParams),
                                    C.c_int, _f, C.POINTER(IcpResult)]),
    "mlo_scan_register_batch": (C.c_int, [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "mlo_stage_upload_async": (C.c_int, [_vp, C.c_int, _vp, _u32, _u32, _vp]),
    "mlo_scan_register_batch_staged": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "mlo_dcloud_upload": (C.c_int, [_vp, _vp, _u32, _u64, C.POINTER(_vp)]),
    "mlo_dcloud_upload_batch": (C.c_int, [_vp, _vp, _u32, _u32, _vp, C.POINTER(_vp)]),
    "mlo_dcloud_destroy": (None, [_vp]),
    "mlo_dcloud_size": (_u64, [_vp]),
    "mlo_scan_register_batch_resident": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mlo_icp_align_batch_resident": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mlo_scanset_create": (C.c_int, [_vp, C.c_uint32, C.POINTER(_vp)]),
    "mlo_scanset_destroy": (None, [_vp]),
    "mlo_scanset_filter": (C.c_int, [_vp, C.c_uint32, C.POINTER(ScanJob), C.c_uint32, C.POINTER(ScanInfo)]),
    "mlo_scanset_prefetch": (C.c_int, [_vp, C.c_uint32, _vp, _vp, C.c_uint32]),
    "mlo_scanset_deskew": (C.c_int, [_vp, C.c_uint32, _vp, _vp, C.POINTER(ScanInfo)]),
    "mlo_scanset_align": (C.c_int, [_vp, C.c_uint32, C.POINTER(AlignJob), C.POINTER(IcpResult)]),
    "mlo_scanset_insert": (C.c_int, [_vp, C.c_uint32, C.POINTER(InsertJob), C.POINTER(MapCounts)]),
    "mlo_scanset_download": (C.c_int, [_vp, C.c_uint32, C.c_int, _vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mlo_profile_enable": (C.c_int, [_vp, C.c_int]),
    "mlo_profile_get": (C.c_int, [_vp, C.POINTER(Profile), C.c_int]),
}

_LIB = None


def declared_symbols():
    return sorted(_SIGNATURES)


def load() -> C.CDLL:
    """dlopen libmlo_b200.so and bind every symbol of include/mlo_b200.h.  Raises when anything is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.mlo_abi_version() != 1:
        raise RuntimeError("libmlo_b200.so ABI version mismatch")
    _LIB = lib
    return lib
