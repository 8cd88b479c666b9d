"""Generates tests/golden/icp_case_*.npz: small, self-contained input/output vectors for the hot path.

The reference's own tests hold no vectors for this path and its dependencies are absent, so the expected outputs
here are produced by the CPU oracle (oracle/, parity unpinned) — they pin the oracle against silent drift and give
the GPU tests a committed fixture.  Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from mola_lidar_odometry_b200 import capi, synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

OUT = Path(__file__).resolve().parent


def crop(raw, rmax, step):
    r = np.linalg.norm(raw[:, :3], axis=1)
    return np.ascontiguousarray(raw[r < rmax][::step])


def make_case(name, seed, voxel, cap, sigma, n_map=4):
    scene = synth.Scene(42)
    traj = synth.trajectory_T00(40, seed=seed)
    fp = capi.Filter1Params()
    fp.for_map = capi.decimate_params(0.6, 500)
    fp.for_icp = capi.decimate_params(1.5, 300, (1.5, 60.0), ((-4.0, -4.0, 0.2), (4.0, 4.0, 3.0)))
    omap = O.OracleMap(voxel, cap)
    layers, poses = [], []
    for k in range(n_map):
        raw = crop(scene.scan(traj[2 * k], scan_seed=77 + k), 45.0, 3)
        a, _ = O.filter_1st_pass(raw, fp)
        T = synth.relative(traj[0], traj[2 * k])
        omap.insert(a, T)
        layers.append(a)
        poses.append(T)
    raw_q = crop(scene.scan(traj[2 * n_map + 1], scan_seed=99), 45.0, 3)
    gt = synth.relative(traj[0], traj[2 * n_map + 1])
    init = synth.perturb(gt, np.random.default_rng(seed), 0.25, 0.8)
    ip = capi.IcpParamsOwner(sigma=sigma)
    map_q, icp_q = O.filter_1st_pass(raw_q, fp)
    res = O.icp_align(omap, icp_q, init, ip.p)
    keys, cnt, xyz = omap.export()
    Rq, tq = res.pose[:, :3], res.pose[:, 3]
    g = (icp_q.astype(np.float64) @ Rq.T + tq).astype(np.float32)
    nn_xyz, nn_d2, nn_f, _ = omap.nn_single(g)
    dec_idx = O.decimate_first(raw_q, fp.for_map)
    np.savez_compressed(
        OUT / f"icp_case_{name}.npz", voxel=voxel, cap=cap, sigma=sigma,
        map_layers=np.concatenate(layers), map_layer_sizes=np.array([len(a) for a in layers]), map_poses=np.stack(poses),
        raw_query=raw_q, init_pose=init, gt_pose=gt,
        f_for_map=np.array([0.6, 500]), f_for_icp=np.array([1.5, 300, 1.5, 60.0, -4.0, -4.0, 0.2, 4.0, 4.0, 3.0]),
        exp_map_keys=keys, exp_map_counts=cnt, exp_map_xyz=xyz, exp_map_layer=map_q, exp_icp_layer=icp_q,
        exp_decimate_idx=dec_idx, exp_pose=res.pose, exp_iterations=res.n_iterations, exp_termination=res.termination,
        exp_pairings=res.n_pairings, exp_quality=res.quality, exp_candidates=res.n_candidate_points,
        exp_nn_xyz=nn_xyz, exp_nn_d2=nn_d2, exp_nn_found=nn_f)
    print(name, "map", omap.stats(), "icp pts", len(icp_q), "iters", res.n_iterations, capi.TERM_NAMES[res.termination],
          "err vs gt", O.pose_error(res.pose, gt))


def refresh_expectations(path):
    """Recompute the oracle-made expectations of an existing fixture from its COMMITTED inputs (used when an oracle routine
    is corrected: round 2 replaced the small-angle closed forms of oracle/se3.hpp by series, which moved the expected
    poses by <= 7e-15; counts, indices and map contents were unchanged and are asserted to be)."""
    g = dict(np.load(path))
    off = np.concatenate([[0], np.cumsum(g["map_layer_sizes"])])
    m = O.OracleMap(float(g["voxel"]), int(g["cap"]))
    for i, T in enumerate(g["map_poses"]):
        m.insert(g["map_layers"][off[i]:off[i + 1]], T)
    a, b = g["f_for_map"], g["f_for_icp"]
    fp = capi.Filter1Params()
    fp.for_map = capi.decimate_params(float(a[0]), int(a[1]))
    fp.for_icp = capi.decimate_params(float(b[0]), int(b[1]), (float(b[2]), float(b[3])), (tuple(b[4:7]), tuple(b[7:10])))
    _, icp_q = O.filter_1st_pass(g["raw_query"], fp)
    r = O.icp_align(m, icp_q, g["init_pose"], capi.IcpParamsOwner(sigma=float(g["sigma"])).p)
    assert (r.n_iterations, r.termination, r.n_pairings, r.n_candidate_points) == \
        (int(g["exp_iterations"]), int(g["exp_termination"]), int(g["exp_pairings"]), int(g["exp_candidates"]))
    print(path, "pose moved by", np.abs(r.pose - g["exp_pose"]).max())
    g["exp_pose"] = r.pose
    gq = (icp_q.astype(np.float64) @ r.pose[:, :3].T + r.pose[:, 3]).astype(np.float32)
    g["exp_nn_xyz"], g["exp_nn_d2"], g["exp_nn_found"], _ = m.nn_single(gq)
    np.savez_compressed(path, **g)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "refresh":
        for f in sorted(OUT.glob("icp_case_*.npz")):
            refresh_expectations(f)
    else:
        # (the synthetic scene has been enriched since the committed fixtures were made: running this regenerates the
        # INPUTS too; use `refresh` to keep them)
        make_case("v1p0_cap20", seed=7, voxel=1.0, cap=20, sigma=2.0)
        make_case("v0p5_cap8", seed=8, voxel=0.5, cap=8, sigma=1.0)
