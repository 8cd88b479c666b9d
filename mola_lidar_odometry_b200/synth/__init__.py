"""Seeded synthetic KITTI-shaped inputs: scene, sensors, trajectory.

Input generator only — nothing here is on the registration path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        p = Path(__file__).resolve().parent / "libmlo_synth.so"
        if not p.exists():
            from .._build import build_synth
            build_synth()
        L = C.CDLL(str(p))
        L.synth_scene_create.restype = C.c_void_p
        L.synth_scene_create.argtypes = [C.c_uint64, C.c_float, C.c_int, C.c_int]
        L.synth_scene_destroy.argtypes = [C.c_void_p]
        L.synth_scan.restype = C.c_uint64
        L.synth_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                 C.c_double, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
        L.synth_scan_skewed.restype = C.c_uint64
        L.synth_scan_skewed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_double,
                                        C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
        _LIB = L
    return _LIB


@dataclass(frozen=True)
class Sensor:
    name: str
    n_beams: int
    n_az: int
    el_top_deg: float
    el_bot_deg: float
    max_range: float = 120.0
    noise_sigma: float = 0.02


K64 = Sensor("K64", 64, 2048, 2.0, -24.8)        # KITTI HDL-64E shaped: 131 072 rays
O128 = Sensor("O128", 128, 1800, 22.5, -22.5)    # 230 400 rays
SENSOR_HEIGHT = 1.73


class Scene:
    """Scene S(seed): street grid with buildings, poles and parked cars."""

    def __init__(self, seed: int = 42, extent_m: float = 1500.0, n_poles: int = 8000, n_cars: int = 6000):
        self._h = _lib().synth_scene_create(seed, extent_m, n_poles, n_cars)
        self.extent = extent_m

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().synth_scene_destroy(self._h)
            self._h = None

    def scan(self, pose_world: np.ndarray, sensor: Sensor = K64, scan_seed: int = 1000, with_time: bool = False):
        """pose_world: 3x4 (or 4x4) sensor->world.  Returns float32 [n,4] x,y,z,intensity in the sensor frame."""
        pose = np.ascontiguousarray(np.asarray(pose_world, dtype=np.float64)[:3, :4])
        cap = sensor.n_beams * sensor.n_az
        out = np.empty((cap, 4), dtype=np.float32)
        t = np.empty(cap, dtype=np.float32) if with_time else None
        n = _lib().synth_scan(self._h, pose.ctypes.data, sensor.n_beams, sensor.n_az, sensor.el_top_deg,
                              sensor.el_bot_deg, sensor.max_range, sensor.noise_sigma, scan_seed,
                              out.ctypes.data, t.ctypes.data if with_time else None, cap)
        if with_time:
            return out[:n].copy(), t[:n].copy()
        return out[:n].copy()


def scan_skewed(scene: "Scene", pose_world, twist_body, sensor: Sensor = None, scan_seed: int = 1000, sweep_s: float = 0.1):
    """A sweep taken while the sensor moves with body-frame twist (vx vy vz wx wy wz); pose_world is the pose at the
    middle of the sweep.  Returns (float32 [n,4] x,y,z,intensity in the instantaneous sensor frames, float32 [n] t)."""
    sensor = sensor or K64
    pose = np.ascontiguousarray(np.asarray(pose_world, dtype=np.float64)[:3, :4])
    tw = np.ascontiguousarray(twist_body, dtype=np.float64)
    cap = sensor.n_beams * sensor.n_az
    out = np.empty((cap, 4), dtype=np.float32)
    t = np.empty(cap, dtype=np.float32)
    n = _lib().synth_scan_skewed(scene._h, pose.ctypes.data, tw.ctypes.data, sweep_s, sensor.n_beams, sensor.n_az,
                                 sensor.el_top_deg, sensor.el_bot_deg, sensor.max_range, sensor.noise_sigma, scan_seed,
                                 out.ctypes.data, t.ctypes.data, cap)
    return out[:n].copy(), t[:n].copy()


def body_twists(traj: np.ndarray, dt: float = 0.1) -> np.ndarray:
    """Body-frame twist at every pose of a trajectory (central differences of log(T_k^-1 T_k+1))."""
    from scipy.spatial.transform import Rotation as Rot
    n = len(traj)
    tw = np.zeros((n, 6))
    for k in range(n):
        a, b = traj[max(k - 1, 0)], traj[min(k + 1, n - 1)]
        span = (min(k + 1, n - 1) - max(k - 1, 0)) * dt
        rel = relative(a, b)
        tw[k, :3] = rel[:, 3] / span                      # small-step approximation of the SE(3) log
        tw[k, 3:] = Rot.from_matrix(rel[:, :3]).as_rotvec() / span
    return tw


def rot_zyx(yaw: float, pitch: float = 0.0, roll: float = 0.0) -> np.ndarray:
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def pose34(x, y, z, yaw, pitch=0.0, roll=0.0) -> np.ndarray:
    T = np.zeros((3, 4))
    T[:, :3] = rot_zyx(yaw, pitch, roll)
    T[:, 3] = (x, y, z)
    return T


def to44(T34: np.ndarray) -> np.ndarray:
    T = np.eye(4)
    T[:3, :4] = np.asarray(T34)[:3, :4]
    return T


def relative(T_ref: np.ndarray, T: np.ndarray) -> np.ndarray:
    """T_ref^-1 * T as 3x4."""
    return (np.linalg.inv(to44(T_ref)) @ to44(T))[:3, :4]


def compose(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    return (to44(A) @ to44(B))[:3, :4]


def perturb(T: np.ndarray, rng: np.random.Generator, trans: float = 0.3, rot_deg: float = 1.0) -> np.ndarray:
    """T * small random motion: translation U[-trans,trans]^3, yaw/pitch/roll U[-rot,rot] (SURVEY §8d config[1])."""
    d = pose34(*rng.uniform(-trans, trans, 3), *np.deg2rad(rng.uniform(-rot_deg, rot_deg, 3)))
    return compose(T, d)


_DIRS = ((1, 0), (0, 1), (-1, 0), (0, -1))


def _street_path(rng: np.random.Generator, length_m: float, extent_m: float, ds: float) -> np.ndarray:
    """Polyline (x, y, yaw) sampled every `ds` along street centre lines with 6 m-radius 90-degree arcs."""
    grid, rad = 40.0, 6.0
    nb = int(extent_m // grid)
    ix, iy, heading = nb // 2, nb // 2, 0
    x, y = ix * grid, iy * grid          # at an intersection centre, about to leave along `heading`
    skip = 0.0                           # distance already covered past the intersection centre (after an arc)
    chunks, total = [], 0.0
    while total < length_m:
        dx, dy = _DIRS[heading]
        nx, ny = ix + dx, iy + dy        # next intersection
        options = []
        for turn, pr in ((0, 0.6), (1, 0.2), (-1, 0.2)):
            h2 = (heading + turn) % 4
            jx, jy = nx + _DIRS[h2][0], ny + _DIRS[h2][1]
            if 1 <= jx <= nb - 1 and 1 <= jy <= nb - 1:
                options.append((turn, pr))
        if not options:
            options = [(1, 1.0)]
        pr = np.array([o[1] for o in options])
        turn = options[rng.choice(len(options), p=pr / pr.sum())][0]
        yaw0 = heading * np.pi / 2
        straight = grid - skip - (rad if turn else 0.0)
        n_s = max(1, int(round(straight / ds)))
        s = np.arange(n_s) * (straight / n_s)
        x0, y0 = x + dx * skip, y + dy * skip
        chunks.append(np.stack([x0 + dx * s, y0 + dy * s, np.full(n_s, yaw0)], 1))
        total += straight
        x, y = nx * grid, ny * grid
        ix, iy = nx, ny
        skip = 0.0
        if turn:
            # arc from (x - d*rad) to (x + d2*rad) around the corner centre
            h2 = (heading + turn) % 4
            dx2, dy2 = _DIRS[h2]
            cx, cy = x - dx * rad + dx2 * rad, y - dy * rad + dy2 * rad
            a0 = np.arctan2((y - dy * rad) - cy, (x - dx * rad) - cx)
            n_a = max(1, int(round((np.pi / 2 * rad) / ds)))
            a = np.arange(n_a) / n_a * (np.pi / 2)
            chunks.append(np.stack([cx + rad * np.cos(a0 + turn * a), cy + rad * np.sin(a0 + turn * a),
                                    yaw0 + turn * a], 1))
            total += np.pi / 2 * rad
            heading, skip = h2, rad
    path = np.concatenate(chunks, 0)
    path[:, 2] = np.unwrap(path[:, 2])
    # clothoid-like corners: smooth the heading over ~6 m of arc length and re-integrate the position, so the yaw
    # rate ramps up and down instead of stepping (a vehicle cannot change curvature instantly)
    k = np.exp(-0.5 * (np.arange(-int(9.0 / ds), int(9.0 / ds) + 1) * ds / 3.0) ** 2)
    k /= k.sum()
    pad = len(k) // 2
    yaw = np.convolve(np.pad(path[:, 2], pad, mode="edge"), k, mode="valid")
    x = path[0, 0] + np.concatenate([[0.0], np.cumsum(np.cos(yaw[:-1]) * ds)])
    y = path[0, 1] + np.concatenate([[0.0], np.cumsum(np.sin(yaw[:-1]) * ds)])
    return np.stack([x, y, yaw], 1)


def trajectory_T00(n_poses: int = 4541, seed: int = 7, extent_m: float = 1500.0, dt: float = 0.1, v0: float = 8.0) -> np.ndarray:
    """KITTI-00-shaped drive: straights and 90-degree turns on a street grid, an Ornstein-Uhlenbeck speed
    profile in [0, 15] m/s, smoothed 0.2-degree pitch/roll noise.  Returns [n,3,4] sensor->world poses
    (world: ground z = 0, sensor at 1.73 m)."""
    rng = np.random.default_rng(seed)
    ds = 0.02
    path = _street_path(rng, n_poses * dt * 15.0 + 100.0, extent_m, ds)
    # speed: Ornstein-Uhlenbeck around 12.5 m/s on straights, braking to ~3 m/s for the 6 m-radius corners
    # (yaw rate <= ~30 deg/s like KITTI; 8 m/s through such a corner would be 76 deg/s)
    turning = np.abs(np.gradient(path[:, 2])) > 1e-9
    look = int(9.0 / ds)
    csum = np.concatenate([[0], np.cumsum(turning)])
    s_now, vv = 0.0, float(v0)
    xyyaw = np.empty((n_poses, 3))
    for k in range(n_poses):
        f = s_now / ds
        i0 = min(int(f), len(path) - 2)
        w = f - i0
        xyyaw[k] = path[i0] * (1 - w) + path[i0 + 1] * w
        near_turn = csum[min(i0 + look, len(path))] - csum[max(i0 - int(2.0 / ds), 0)] > 0
        target = 3.0 if near_turn else 12.5
        vv += (0.15 if near_turn else 0.06) * (target - vv) + 0.35 * rng.standard_normal() * (0.3 if near_turn else 1.0)
        vv = min(15.0, max(0.5 if k else vv, vv))
        s_now += max(vv, 0.0) * dt
    pr = rng.standard_normal((n_poses, 2))
    k = np.exp(-0.5 * (np.arange(-20, 21) / 6.0) ** 2)
    k /= np.sqrt((k ** 2).sum())
    pr = np.stack([np.convolve(pr[:, i], k, mode="same") for i in range(2)], 1) * np.deg2rad(0.2)
    out = np.empty((n_poses, 3, 4))
    for i in range(n_poses):
        out[i] = pose34(xyyaw[i, 0], xyyaw[i, 1], SENSOR_HEIGHT, xyyaw[i, 2], pr[i, 0], pr[i, 1])
    return out
