"""GPU parity of EVERY device path of ICP::align (module/src/LidarOdometry.cpp:961-962) against the CPU oracle.

The library picks one of three device paths by batch size (include/mlo_b200.h mlo_set_option "align_path"):
  1  one kernel per phase (k_match_accumulate_wl4 -> k_solve [-> k_accumulate -> k_solve]) over stream groups, with a
     hand-over of the stragglers to a single-launch kernel: the LARGE-batch path the headline and the roofline are
     quoted on (>= 148 x 1024 total queries),
  2  the queue-driven persistent kernel (k_icp_persistent),
  3  one thread block per problem (k_icp_block): small batches, single sequences, fleets.
Each is forced here and compared with the oracle for EVERY problem of the batch: pose (1 mm / 0.01 deg, north_star),
termination reason, iteration count, pairings, potential pairings, candidate counts, quality.
"""
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
import pytest

from mola_lidar_odometry_b200 import capi, synth
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu

TOL_TRANS_M = 1e-3   # BASELINE.json north_star: <= 1 mm / 0.01 deg per scan
TOL_ROT_DEG = 1e-2


def _check(gr, orr, it_slack=0):
    et, er = O.pose_error(gr.pose, orr.pose)
    assert et <= TOL_TRANS_M and er <= TOL_ROT_DEG, f"pose differs: {et} m, {er} deg"
    assert gr.termination == orr.termination
    assert abs(int(gr.n_iterations) - int(orr.n_iterations)) <= it_slack
    if int(gr.n_iterations) == int(orr.n_iterations):
        assert gr.n_pairings == orr.n_pairings
        assert gr.n_potential_pairings == orr.n_potential_pairings
        assert gr.n_candidate_points == orr.n_candidate_points
        assert gr.n_query_iterations == orr.n_query_iterations
        assert abs(gr.quality - orr.quality) < 1e-12


class _Options:
    """Set launch-policy knobs on the (session-scoped) context and put them back afterwards."""

    def __init__(self, ctx, **kw):
        self.ctx, self.kw, self.old = ctx, kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = self.ctx.get_option(k)
            self.ctx.set_option(k, v)
        return self

    def __exit__(self, *a):
        for k, v in self.old.items():
            self.ctx.set_option(k, v)


# ------------------------------------------------------------------ small batch, every path, every solver feature
@pytest.fixture(scope="module")
def small_case(ctx, world):
    from mola_lidar_odometry_b200.api import LocalMap
    g, o = LocalMap(ctx, 1.0, 20, 0.0, 1 << 16), O.OracleMap(1.0, 20, 0.0)
    for fr in world["frames"][:12]:
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
    rng = np.random.default_rng(21)
    locals_, inits, owners = [], [], []
    for j, k in enumerate((12, 13, 14, 15, 16, 17, 18, 19, 20, 21)):
        fr = world["frames"][k]
        init = synth.perturb(fr["gt"], rng, 0.3, 1.0)
        ip = capi.IcpParamsOwner(sigma=2.0 if j % 2 == 0 else 1.0)
        if j == 1:
            ip.p.robust_kernel = capi.KERNEL_CAUCHY
        if j == 2:
            ip.p.robust_kernel = capi.KERNEL_NONE
        if j == 3:
            info = np.diag([50.0, 50.0, 50.0, 2000.0, 2000.0, 2000.0])
            info[0, 1] = info[1, 0] = 5.0
            ip.set_prior(init, info)
        if j == 4:
            init = synth.compose(fr["gt"], synth.pose34(0.6, 0.1, 0, 0.01))
            ip.set_hook(init, 0.15, 0.75)
        if j == 5:
            ip = capi.IcpParamsOwner(sigma=2.0, max_iterations=3)
        if j == 6:
            init = synth.compose(fr["gt"], synth.pose34(5000, 0, 0, 0))   # NoPairings
        if j == 7:
            ip.p.gn_max_iterations = 1
        locals_.append(fr["icp_layer"])
        inits.append(init)
        owners.append(ip)
    refs = [O.icp_align(o, l, i, ip.p) for l, i, ip in zip(locals_, inits, owners)]
    return dict(g=g, locals=locals_, inits=np.stack(inits), owners=owners, refs=refs)


@pytest.mark.parametrize("path", [1, 2, 3])
@pytest.mark.parametrize("fuse", [1, 0])
def test_every_align_path_matches_the_oracle(ctx, small_case, path, fuse):
    c = small_case
    with _Options(ctx, align_path=path, fuse_inner=fuse):
        res = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [o.p for o in c["owners"]])
        assert ctx.get_option("last_align_path") == path
    terms = set()
    for gr, orr in zip(res, c["refs"]):
        _check(gr, orr)
        terms.add(int(gr.termination))
    assert {1, 3, 4, 5} <= terms   # NoPairings, MaxIterations, Stalled, HookRequest all occur in this batch


@pytest.mark.parametrize("path", [1, 2])
@pytest.mark.parametrize("fuse", [1, 0])
def test_prior_prepared_ahead_changes_nothing_but_summation_order(ctx, small_case, path, fuse):
    """mlo_set_option "prior_ahead": a second warp linearises the prior term for the next solve while the first finishes
    the current one (and, in the fused loop, while three warps instead of four re-linearise: the only arithmetic
    difference is the grouping of that sum).  With the prior on EVERY problem of the batch: same iteration counts and
    terminations, poses equal to 1e-9 m, and both settings match the oracle."""
    c = small_case
    owners = []
    for j, init in enumerate(c["inits"]):
        ip = capi.IcpParamsOwner(sigma=2.0 if j % 2 == 0 else 1.0)
        info = np.diag([50.0, 50.0, 50.0, 2000.0, 2000.0, 2000.0]) * (1.0 + 0.1 * j)
        info[0, 4] = info[4, 0] = 3.0
        ip.set_prior(init, info)
        owners.append(ip)
    o = O.OracleMap(1.0, 20, 0.0)
    gk, gc, gp = c["g"].export()
    o.insert(gp, np.eye(4)[:3])          # (export order re-inserted: the same map, tests/test_gpu_full_size.py)
    refs = [O.icp_align(o, l, i, ip.p) for l, i, ip in zip(c["locals"], c["inits"], owners)]
    res = {}
    for ahead in (0, 1):
        with _Options(ctx, align_path=path, fuse_inner=fuse, prior_ahead=ahead):
            res[ahead] = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [w.p for w in owners])
    for a, b, r in zip(res[0], res[1], refs):
        assert np.allclose(a.pose, b.pose, rtol=0, atol=1e-9) and np.allclose(a.cov, b.cov, rtol=1e-9, atol=0)
        assert int(a.n_iterations) == int(b.n_iterations) and int(a.termination) == int(b.termination)
        _check(a, r)
        _check(b, r)


@pytest.mark.parametrize("threads", [256, 512])
@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_block_kernel_geometries(ctx, small_case, threads, cluster):
    """k_icp_block with every cluster size (blocks per problem, reductions through distributed shared memory) and both
    block sizes; with 256 threads and one block the ~1.5 k queries of a problem take several passes."""
    c = small_case
    with _Options(ctx, align_path=3, block_threads=threads, block_cluster=cluster):
        res = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [o.p for o in c["owners"]])
        assert ctx.get_option("last_block_cluster") == cluster and ctx.get_option("last_block_threads") == threads
    for gr, orr in zip(res, c["refs"]):
        _check(gr, orr)


def test_block_kernel_single_problem_and_horn(ctx, small_case, world):
    c = small_case
    for j in (0, 3, 4):
        with _Options(ctx, align_path=3):
            gr = ctx.icp_align(c["locals"][j], c["g"], c["inits"][j], c["owners"][j].p)
        _check(gr, c["refs"][j])
    ip = capi.IcpParamsOwner(sigma=1.0, max_iterations=40)
    ip.p.solver = capi.SOLVER_HORN
    o = O.OracleMap(1.0, 20, 0.0)
    for fr in world["frames"][:12]:
        o.insert(fr["map_layer"], fr["gt"])
    orr = O.icp_align(o, c["locals"][0], c["inits"][0], ip.p)
    for path in (1, 2, 3):
        with _Options(ctx, align_path=path):
            gr = ctx.icp_align(c["locals"][0], c["g"], c["inits"][0], ip.p)
        _check(gr, orr, it_slack=1)


def test_block_kernel_ndt_point_to_plane(ctx, world):
    from mola_lidar_odometry_b200.api import LocalMap
    g = LocalMap(ctx, 1.0, 32, 0.2, 1 << 16, kind=capi.MAP_NDT)
    o = O.OracleMap(1.0, 32, 0.2, kind=capi.MAP_NDT)
    for fr in world["frames"][:12]:
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
    rng = np.random.default_rng(3)
    locals_, inits, owners = [], [], []
    for k in (12, 14, 16, 18):
        fr = world["frames"][k]
        locals_.append(fr["icp_layer"])
        inits.append(synth.perturb(fr["gt"], rng, 0.2, 0.5))
        owners.append(capi.IcpParamsOwner(sigma=1.0, pipeline="ndt"))
    refs = [O.icp_align(o, l, i, ip.p) for l, i, ip in zip(locals_, inits, owners)]
    for path in (1, 2, 3):
        with _Options(ctx, align_path=path):
            res = ctx.icp_align_batch(locals_, g, np.stack(inits), [ip.p for ip in owners])
        for gr, orr in zip(res, refs):
            _check(gr, orr)
            assert gr.n_pairings > 0


# ------------------------------------------------------------------ the LARGE-batch launch sequence (config[1]-shaped)
N_LARGE = 128


@pytest.fixture(scope="module")
def large_case(ctx, scene):
    """128 K64 scans against a 0.5 m / cap-20 map of >= 2^18 voxels (BASELINE.json configs[1] shape):
    >= 148 x 1024 total queries, so the library's own policy takes the one-kernel-per-phase launch sequence."""
    from mola_lidar_odometry_b200.api import LocalMap
    threads = os.cpu_count() or 4
    traj = synth.trajectory_T00(2000, seed=7)
    T0 = traj[0]
    fp = capi.filter1_default(100.0)
    g, o = LocalMap(ctx, 0.5, 20, 0.0, 1 << 19), O.OracleMap(0.5, 20, 0.0)
    ks = list(range(0, 2000, 5))
    with ThreadPoolExecutor(threads) as ex:
        for c0 in range(0, len(ks), 32):
            part = ks[c0:c0 + 32]
            layers = list(ex.map(lambda k: O.filter_1st_pass(scene.scan(traj[k], scan_seed=1000 + k), fp)[0], part))
            for k, a in zip(part, layers):
                T = synth.relative(T0, traj[k])
                g.insert(a, T)
                o.insert(a, T)
            if o.stats()[0] >= (1 << 18):
                last = part[-1]
                break
        else:
            raise RuntimeError("trajectory exhausted before 2^18 voxels")
        assert g.stats() == o.stats() and g.stats()[0] >= (1 << 18)
        rng = np.random.default_rng(5)
        poses, inits = [], []
        for i, k in enumerate(rng.integers(0, last, N_LARGE)):
            Tw = synth.compose(traj[k], synth.pose34(0, 0, 0, np.deg2rad(3.0 * i)))
            poses.append(Tw)
            inits.append(synth.perturb(synth.relative(T0, Tw), rng, 0.3, 1.0))
        locals_ = list(ex.map(lambda a: O.filter_1st_pass(scene.scan(a[1], scan_seed=700000 + a[0]), fp)[1], enumerate(poses)))
        owners = [capi.IcpParamsOwner(sigma=2.0) for _ in range(N_LARGE)]
        refs = list(ex.map(lambda a: O.icp_align(o, a[0], a[1], a[2].p), zip(locals_, inits, owners)))
    total_q = sum(len(l) for l in locals_)
    assert total_q >= 148 * 1024, total_q
    return dict(g=g, locals=locals_, inits=np.stack(inits), owners=owners, refs=refs, total_q=total_q)


@pytest.mark.parametrize("groups,tail,tail_path,fuse", [(3, 1, 3, 1), (1, 0, 3, 1), (3, 1, 2, 0), (2, 1, 3, 0), (1, 1, 3, 1)])
def test_large_batch_launch_sequence_matches_the_oracle(ctx, large_case, groups, tail, tail_path, fuse):
    c = large_case
    l0 = ctx.launch_count
    with _Options(ctx, stream_groups=groups, tail_handover=tail, tail_path=tail_path, fuse_inner=fuse):
        res = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [o.p for o in c["owners"]])
        assert ctx.get_option("align_path") == 0           # the library's own policy ...
        assert ctx.get_option("last_align_path") == 1      # ... took the launch sequence
        assert ctx.get_option("last_stream_groups") == groups
        assert ctx.get_option("last_tail_handover") == (tail_path if tail else 0)
    assert ctx.launch_count - l0 >= 2 * min(int(r.n_iterations) for r in res)   # (several launches per ICP iteration)
    worst = (0.0, 0.0)
    for gr, orr in zip(res, c["refs"]):
        _check(gr, orr)
        worst = max(worst, O.pose_error(gr.pose, orr.pose))
    print(f"large batch ({c['total_q']} queries): worst GPU-vs-oracle pose delta {worst}")


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19])
def test_large_batch_drain_variants_are_identical(ctx, large_case, variant):
    """The drain loop of the work-list kernel exists in several forms (include/mlo_b200.h "wl_variant": segment-wise
    merge, software-pipelined, cp.async.bulk staging, contiguous ranges with a register-resident best).  All of them take
    the minimum of the same 64-bit (distance bits, visiting order) keys, so poses, iteration counts and pairing counts
    must be IDENTICAL between variants, not merely within tolerance."""
    c = large_case
    with _Options(ctx, wl_variant=3):
        base = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [o.p for o in c["owners"]])
    with _Options(ctx, wl_variant=variant):
        res = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [o.p for o in c["owners"]])
        assert ctx.get_option("last_align_path") == 1
    for a, b, orr in zip(res, base, c["refs"]):
        if variant in (10, 11):   # one partial per warp: the same pairings, summed in a different grouping
            assert np.allclose(np.asarray(a.pose), np.asarray(b.pose), rtol=0, atol=1e-9)
        else:
            assert np.array_equal(np.asarray(a.pose), np.asarray(b.pose))
        assert int(a.n_iterations) == int(b.n_iterations) and int(a.n_pairings) == int(b.n_pairings)
        assert int(a.n_candidate_points) == int(b.n_candidate_points)
        _check(a, orr)


@pytest.mark.parametrize("per_sm,every", [(2048, 2), (64, 8), (512, 1)])
def test_large_batch_hand_over_point_does_not_change_results(ctx, large_case, per_sm, every):
    """Where the launch sequence hands its stragglers to the queue-driven kernel (mlo_set_option "tail_queries_per_sm",
    "check_every") is a scheduling decision: every problem still matches the oracle."""
    c = large_case
    with _Options(ctx, tail_queries_per_sm=per_sm, check_every=every):
        res = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [o.p for o in c["owners"]])
        assert ctx.get_option("last_align_path") == 1
    for gr, orr in zip(res, c["refs"]):
        _check(gr, orr)


def test_large_batch_other_paths_agree(ctx, large_case):
    """The same 128 problems through the queue-driven kernel and the block kernel (two blocks per SM at this size)."""
    c = large_case
    for path in (2, 3):
        with _Options(ctx, align_path=path):
            res = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [o.p for o in c["owners"]])
        for gr, orr in zip(res, c["refs"]):
            _check(gr, orr)


# ------------------------------------------------------------------ the ICP log (mp2p_icp generateDebugFiles / saveIterationDetails)
@pytest.mark.parametrize("path", [1, 2, 3])
def test_icp_iteration_log(ctx, small_case, path):
    """One record per executed ICP iteration (default.yaml:177-182), written by the device as the loop runs: iteration
    index, pose after the iteration, realised thresholds, pairings, step measure, termination verdict."""
    c = small_case
    ctx.icp_log_enable(64)
    try:
        with _Options(ctx, align_path=path):
            res = ctx.icp_align_batch(c["locals"], c["g"], c["inits"], [o.p for o in c["owners"]])
        for b, (gr, own) in enumerate(zip(res, c["owners"])):
            log = ctx.icp_log(b)
            if gr.termination == 1:      # NoPairings: the loop broke before any solver step
                assert len(log) == 0
                continue
            # Stalled / HookRequest end INSIDE an iteration that nIterations does not count
            assert len(log) == gr.n_iterations + (1 if gr.termination in (4, 5) else 0)
            assert [r.iteration for r in log] == list(range(len(log)))
            assert all(r.termination == 0 for r in log[:-1]) and log[-1].termination == gr.termination
            assert np.array_equal(log[-1].pose, gr.pose) and log[-1].n_pairings == gr.n_pairings
            thr = np.ctypeslib.as_array(own.p.pt2pt_threshold_by_iter, shape=(own.p.table_len,))
            assert all(r.threshold_pt2pt == thr[min(r.iteration, len(thr) - 1)] for r in log)
            if gr.termination == 4:
                assert log[-1].step_trans < own.p.min_abs_step_trans and log[-1].step_rot < own.p.min_abs_step_rot
    finally:
        ctx.icp_log_enable(0)
