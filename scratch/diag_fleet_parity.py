"""Scratch: where does a GPU fleet trajectory first leave the oracle's?  (GPU ray caster feeds both.)"""
import os, sys
os.environ.setdefault("MOLA_OPTIMIZE_TWIST", "false"); os.environ.setdefault("MOLA_INITIAL_VX", "8.0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mola_lidar_odometry_b200 import synth
from mola_lidar_odometry_b200.api import Context
from mola_lidar_odometry_b200.host_api import LidarOdometryFleet, PIPELINES
from mola_lidar_odometry_b200.synth.gpu import GpuSynth
from oracle import oracle_py as O
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
yaml = str(PIPELINES / "lidar3d-default.yaml")
scene = synth.Scene(42); gs = GpuSynth(scene, "cuda:0")
seeds = [7 + i for i in range(S)]
trajs = [synth.trajectory_T00(N + 5, seed=sd) for sd in seeds]
ctx = Context(0); fleet = LidarOdometryFleet(ctx, yaml, S)
solo = [O.OracleLidarOdometry(yaml) for _ in range(S)]
first = {}
F = ("icp_ran", "icp_good", "map_updated", "icp_iterations", "icp_runs", "termination", "n_map_layer", "n_icp_layer", "icp_had_prior", "has_motion_model")
for k in range(N):
    poses = np.stack([trajs[s][k] for s in range(S)]); sds = np.array([seeds[s] * 100000 + k for s in range(S)], dtype=np.uint64)
    flat, offs = gs.scan_batch(poses, sds, synth.K64); fl = flat[:, :3].contiguous().cpu().numpy(); of = offs.cpu().numpy()
    clouds = [fl[of[s]:of[s + 1]] for s in range(S)]
    outs = fleet.on_lidar(clouds, [0.1 * k] * S)
    for s in range(S):
        b = solo[s].on_lidar(clouds[s], 0.1 * k)
        a = outs[s]
        et, er = O.pose_error(a.pose, b.pose)
        if s not in first and (et > 1e-6 or any(getattr(a, f) != getattr(b, f) for f in F)):
            first[s] = k
            print(f"seq {s} (seed {seeds[s]}) first differs at scan {k}: dpose=({et:.3e} m, {er:.3e} deg)")
            print("   gpu   ", {f: getattr(a, f) for f in F}, "q=%.6f sigma=%.6f trace=%.6f" % (a.quality, a.sigma, a.prior_info_trace))
            print("   oracle", {f: getattr(b, f) for f in F}, "q=%.6f sigma=%.6f trace=%.6f" % (b.quality, b.sigma, b.prior_info_trace))
worst = 0.0
print("sequences that differ:", sorted(first.items()))
