"""Scratch: timeline of problem 0 inside k_icp_block (needs the -DMLO_TRACE build, see trace_build.sh; run with
MLO_B200_LIB=scratch/libmlo_b200_trace.so)."""
import os, sys, ctypes as C, collections
import numpy as np
os.environ.setdefault("MOLA_OPTIMIZE_TWIST", "false"); os.environ.setdefault("MOLA_INITIAL_VX", "8.0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mola_lidar_odometry_b200 import capi, synth
from mola_lidar_odometry_b200.api import Context
from mola_lidar_odometry_b200.host_api import LidarOdometryFleet, PIPELINES
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pipe = sys.argv[2] if len(sys.argv) > 2 else "lidar3d-default.yaml"
scene = synth.Scene(42)
trajs = [synth.trajectory_T00(40, seed=7 + s) for s in range(S)]
ctx = Context(0)
fleet = LidarOdometryFleet(ctx, str(PIPELINES / pipe), S)
lib = capi.load()
buf = (C.c_ulonglong * 16384)(); n = C.c_uint()
names = {11: "iter_start", 12: "match_done", 13: "reduced", 14: "solved", 15: "accumulated", 16: "solve_ret", 20: "transformed", 21: "probed", 22: "own_drained", 23: "nb_selected", 24: "nb_drained"}
for k in range(30):
    outs = fleet.on_lidar([scene.scan(trajs[s][k], scan_seed=(7 + s) * 1000 + k) for s in range(S)], [0.1 * k] * S)
    lib.mlo_debug_trace_read(buf, 16384, C.byref(n))
ev = sorted(((buf[i] >> 8, buf[i] & 0xFF) for i in range(n.value)))
print("events", n.value, "iterations", outs[0].icp_iterations, "n_icp", outs[0].n_icp_layer)
t0 = ev[0][0]; prev = t0; seg = collections.defaultdict(list); last_code = None
for t, c in ev:
    seg[(last_code, c)].append(t - prev); prev = t; last_code = c
for (a, b), v in sorted(seg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{names.get(a, a)!s:>12} -> {names.get(b, b)!s:<12} n={len(v):4d} mean={np.mean(v)/1e3:7.2f} us  total={sum(v)/1e3:8.1f} us")
print("total span us", (ev[-1][0] - t0) / 1e3, "per ICP iteration us", (ev[-1][0] - t0) / 1e3 / max(1, sum(1 for _, c in ev if c == 11)))
