"""Scratch: per-step wall time of a large fleet, to find the occasional slow lock step (S=128: 3.6 ms usual, 8-26 ms seen)."""
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
    rows = []
    for k in range(N):
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        outs = fleet.on_lidar([scans[s][k] for s in range(S)], [0.1 * k] * S)
        dt = 1e3 * (time.perf_counter() - t0)
        p = fleet.phase_times()
        rows.append((k, dt, p['align_ms_per_step'], p['filter_ms_per_step'], p['insert_ms_per_step'], ctx.get_option("last_align_path"), ctx.get_option("last_tail_handover"),
                     int(sum(o.n_icp_layer for o in outs)), max(o.icp_iterations for o in outs), max(o.icp_runs for o in outs), ctx.launch_count - l0))
    if rep == 1:
        med = np.median([r[1] for r in rows])
        print("median step ms", round(med, 3))
        for r in rows:
            flag = "  <-- slow" if r[1] > 2 * med else ""
            print("k=%3d step %.2f ms align %.2f filter %.2f insert %.2f path %d tail %d queries %d max_it %d max_runs %d launches %d%s" % (r + (flag,)))
    fleet.close()
