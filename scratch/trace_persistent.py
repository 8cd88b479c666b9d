"""Scratch: timeline of one problem inside the persistent ICP kernel (needs the -DMLO_TRACE build, see trace_build.sh)."""
import os, sys, ctypes as C, collections
import numpy as np
os.environ.setdefault("MOLA_OPTIMIZE_TWIST", "false"); os.environ.setdefault("MOLA_INITIAL_VX", "8.0")
sys.path.insert(0, "/root/repo")
from mola_lidar_odometry_b200 import capi, synth
from mola_lidar_odometry_b200.api import Context
from mola_lidar_odometry_b200.host_api import LidarOdometryFleet
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
scene = synth.Scene(42)
trajs = [synth.trajectory_T00(40, seed=7 + s) for s in range(S)]
ctx = Context(0)
YAML = sys.argv[2] if len(sys.argv) > 2 else "lidar3d-default.yaml"
fleet = LidarOdometryFleet(ctx, "/root/repo/pipelines/" + YAML, S)
lib = capi.load()
buf = (C.c_ulonglong * 16384)(); n = C.c_uint()
names = {1: "pop_match0", 2: "pop_acc0", 3: "chunk0_done", 4: "last_chunk_done", 5: "partials_summed", 6: "solve1_done", 7: "fused_done", 8: "next_published", 40: "sys_laid_out", 41: "prior_added", 42: "ldlt_done", 43: "retracted", 44: "measured", 45: "staged", 46: "core_ret", 47: "prior_e", 48: "prior_J"}
for k in range(30):
    outs = fleet.on_lidar([scene.scan(trajs[s][k], scan_seed=(7 + s) * 1000 + k) for s in range(S)], [0.1 * k] * S)
    lib.mlo_debug_trace_read(buf, 16384, C.byref(n))
ev = sorted(((buf[i] >> 8, buf[i] & 0xFF) for i in range(n.value)))
print("events", n.value, "iterations", outs[0].icp_iterations, "n_icp", outs[0].n_icp_layer)
# per-iteration deltas: split at code 8 (next match published)
t0 = ev[0][0]; prev = t0; seg = collections.defaultdict(list); last_code = None
for t, c in ev:
    seg[(last_code, c)].append(t - prev); prev = t; last_code = c
for (a, b), v in sorted(seg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{names.get(a, a)!s:>16} -> {names.get(b, b)!s:<16} n={len(v):4d} mean={np.mean(v)/1e3:7.2f} us  total={sum(v)/1e3:8.1f} us")
print("total span us", (ev[-1][0] - t0) / 1e3)
