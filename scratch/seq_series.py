"""Scratch: per-10-step align wall time and ICP layer size along a fleet run."""
import os, sys, time
import numpy as np
os.environ.setdefault("MOLA_OPTIMIZE_TWIST", "false"); os.environ.setdefault("MOLA_INITIAL_VX", "8.0")
sys.path.insert(0, "/root/repo")
from concurrent.futures import ThreadPoolExecutor
from mola_lidar_odometry_b200 import synth
from mola_lidar_odometry_b200.api import Context
from mola_lidar_odometry_b200.host_api import LidarOdometryFleet
S, N = int(sys.argv[1]), int(sys.argv[2])
scene = synth.Scene(42)
trajs = [synth.trajectory_T00(N + 5, seed=7 + s) for s in range(S)]
with ThreadPoolExecutor(16) as ex:
    scans = [list(ex.map(lambda k, tr=tr, sd=7 + s: scene.scan(tr[k], scan_seed=sd * 100000 + k), range(N))) for s, tr in enumerate(trajs)]
ctx = Context(0)
for rep in range(2):
    fleet = LidarOdometryFleet(ctx, "/root/repo/pipelines/lidar3d-default.yaml", S)
    fleet.phase_times()
    for k in range(N):
        outs = fleet.on_lidar([scans[s][k] for s in range(S)], [0.1 * k] * S)
        if k % 10 == 9 and rep == 1:
            p = fleet.phase_times()
            print(f"k={k:3d} align {p['align_ms_per_step']:.3f} filter {p['filter_ms_per_step']:.3f} insert {p['insert_ms_per_step']:.3f} n_icp mean {np.mean([o.n_icp_layer for o in outs]):.0f} max {max(o.n_icp_layer for o in outs)} its mean {np.mean([o.icp_iterations for o in outs]):.1f} max {max(o.icp_iterations for o in outs)} launches {ctx.launch_count}", flush=True)
    fleet.close()
