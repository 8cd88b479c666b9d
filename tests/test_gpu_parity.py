"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact for voxel indices, decimation indices, map contents and NN results; SE(3) within
1 mm / 0.01 deg per scan for the ICP result (BASELINE.json north_star), tolerance written below.
"""
import numpy as np
import pytest

from mola_lidar_odometry_b200 import capi, synth
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu

TOL_TRANS_M = 1e-3
TOL_ROT_DEG = 1e-2


def _mk_maps(ctx, voxel=1.0, cap=20, min_dist=0.0, capacity=1 << 16, kind=0):
    from mola_lidar_odometry_b200.api import LocalMap
    g = LocalMap(ctx, voxel, cap, min_dist, capacity, kind=kind)
    o = O.OracleMap(voxel, cap, min_dist, kind=kind)
    return g, o


def _assert_maps_equal(g, o):
    gk, gc, gp = g.export()
    ok, oc, op = o.export()
    assert gk.shape == ok.shape and np.array_equal(gk, ok), "voxel keys differ"
    assert np.array_equal(gc, oc), "per-voxel counts differ"
    assert np.array_equal(gp.view(np.uint32), op.view(np.uint32)), "stored points differ (bitwise)"


def test_device_is_blackwell(ctx):
    name, sms, cc = ctx.device_info()
    assert cc[0] == 10 and sms > 0


@pytest.mark.parametrize("voxel,cap,min_dist", [(1.0, 20, 0.0), (0.5, 20, 0.0), (1.0, 32, 0.2), (2.0, 4, 0.0)])
def test_map_insert_bit_exact(ctx, world, voxel, cap, min_dist):
    g, o = _mk_maps(ctx, voxel, cap, min_dist)
    for fr in world["frames"][:8]:
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
        assert g.stats() == o.stats()
    _assert_maps_equal(g, o)


def test_map_grows_on_demand_bit_exact(ctx, world):
    """capacity_voxels is only the initial size (upstream's HashedVoxelPointCloud is unbounded): inserts that would not
    fit re-hash the map into larger buffers; contents stay bit-identical to the oracle's, culling and NN included."""
    g, o = _mk_maps(ctx, 1.0, 20, 0.0, capacity=256)
    for k, fr in enumerate(world["frames"][:10]):
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
        assert g.stats() == o.stats()
        if k == 5:
            s = fr["gt"][:, 3]
            g.cull(s, 40.0)
            o.cull(s, 40.0)
    assert g.stats()[0] > 10 * 256
    _assert_maps_equal(g, o)
    q = world["frames"][11]["icp_layer"]
    gx, gd, gf = g.nn_single(q)
    ox, od, of, _ = o.nn_single(q)
    assert np.array_equal(gf, of) and np.array_equal(gd.view(np.uint32)[gf], od.view(np.uint32)[of])


def test_map_insert_soa_and_stride4(ctx, world):
    g, o = _mk_maps(ctx)
    fr = world["frames"][0]
    raw = fr["raw"][:30000]
    g.insert(raw, fr["gt"])                      # stride 4 (KITTI layout)
    o.insert(raw, fr["gt"])
    p = world["frames"][1]["map_layer"]
    g.insert_soa(p[:, 0], p[:, 1], p[:, 2], world["frames"][1]["gt"])
    o.insert(p, world["frames"][1]["gt"])
    _assert_maps_equal(g, o)


def test_map_cull_bit_exact(ctx, world):
    g, o = _mk_maps(ctx)
    for fr in world["frames"][:6]:
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
    s = world["frames"][5]["gt"][:, 3]
    g.cull(s, 30.0)
    o.cull(s, 30.0)
    assert g.stats() == o.stats() and g.stats()[0] > 0
    _assert_maps_equal(g, o)
    # inserting after a cull keeps working (rebuild swapped buffers)
    fr = world["frames"][6]
    g.insert(fr["map_layer"], fr["gt"])
    o.insert(fr["map_layer"], fr["gt"])
    _assert_maps_equal(g, o)


def test_map_cull_in_place_reuses_rows_bit_exact(ctx, world):
    """Repeated insert + cull along a drive: culled voxel ids go through the free stack and are reused; the map content
    stays bit-identical to the oracle's erase-voxels loop, with a capacity that is only enough when rows are reused."""
    frames = world["frames"]
    o = O.OracleMap(1.0, 20, 0.0)
    peak = created = prev = 0
    for fr in frames:
        o.insert(fr["map_layer"], fr["gt"])
        nv = o.stats()[0]
        peak, created = max(peak, nv), created + nv - prev
        o.cull(fr["gt"][:, 3], 25.0)
        prev = o.stats()[0]
    from mola_lidar_odometry_b200.api import LocalMap
    cap = int(peak * 1.05)
    assert created > 3 * cap, "the drive must create more voxels than the capacity holds unless culled rows are reused"
    g = LocalMap(ctx, 1.0, 20, 0.0, cap)
    o = O.OracleMap(1.0, 20, 0.0)
    for i, fr in enumerate(frames):
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
        g.cull(fr["gt"][:, 3], 25.0)
        o.cull(fr["gt"][:, 3], 25.0)
        assert g.stats() == o.stats(), i
        if i % 6 == 5:
            _assert_maps_equal(g, o)
    _assert_maps_equal(g, o)
    # NN over the culled + refilled map (dead column buckets stay on the probe chains)
    q = frames[-1]["icp_layer"]
    qg = (frames[-1]["gt"][:, :3] @ q.T).T + frames[-1]["gt"][:, 3]
    gx, gd, gf = g.nn_single(qg.astype(np.float32))
    ox, od, of, _ = o.nn_single(qg.astype(np.float32))
    assert np.array_equal(gf, of) and np.array_equal(gd.view(np.uint32), od.view(np.uint32)) and gf.mean() > 0.3


def test_map_clear_and_empty_inputs(ctx, world):
    g, o = _mk_maps(ctx)
    g.insert(np.zeros((0, 3), np.float32), np.eye(4)[:3])
    assert g.stats() == (0, 0)
    xyz, d2, f = g.nn_single(np.array([[1.0, 2.0, 3.0]], np.float32))
    assert not f[0] and np.isinf(d2[0])
    g.insert(world["frames"][0]["map_layer"], np.eye(4)[:3])
    assert g.stats()[0] > 0
    g.clear()
    assert g.stats() == (0, 0)


def test_map_capacity_is_only_the_initial_size(ctx, world):
    """A tiny initial capacity is not an error any more (upstream's map is unbounded): the first insert re-hashes the
    map into larger buffers and the same context keeps working; only sizes above the hard limit are refused."""
    from mola_lidar_odometry_b200.api import LocalMap, MloError
    g = LocalMap(ctx, 1.0, 20, 0.0, 64)
    layer = world["frames"][0]["map_layer"]
    g.insert(layer, np.eye(4)[:3])
    o = O.OracleMap(1.0, 20, 0.0)
    o.insert(layer, np.eye(4)[:3])
    assert g.stats() == o.stats() and g.stats()[0] > 64
    with pytest.raises(MloError):
        LocalMap(ctx, 1.0, 20, 0.0, (1 << 26) + 1)


def test_nn_single_bit_exact(ctx, world):
    g, o = _mk_maps(ctx)
    for fr in world["frames"][:10]:
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
    rng = np.random.default_rng(3)
    fr = world["frames"][10]
    R, t = fr["gt"][:, :3], fr["gt"][:, 3]
    q = (fr["icp_layer"].astype(np.float64) @ R.T + t).astype(np.float32)
    q = np.concatenate([q, q + rng.normal(0, 0.4, q.shape).astype(np.float32),
                        rng.uniform(-300, 300, (500, 3)).astype(np.float32)])
    gx, gd, gf = g.nn_single(q)
    ox, od, of, _ = o.nn_single(q)
    assert np.array_equal(gf, of)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(gx[gf].view(np.uint32), ox[of].view(np.uint32))
    assert gf.mean() > 0.5


def test_nn_ties_and_diagonal_cell(ctx):
    """Toy map: exact ties resolve to the first cell in (cx,cy,cz) order; best neighbour in a diagonal cell."""
    g, o = _mk_maps(ctx)
    pts = np.array([[0.5, 0.5, 0.5], [2.5, 0.5, 0.5], [1.9, 1.9, 1.9], [-0.5, 0.5, 0.5]], np.float32)
    I = np.eye(4)[:3]
    g.insert(pts, I)
    o.insert(pts, I)
    q = np.array([[1.5, 0.5, 0.5],      # equidistant to (0.5,..) and (2.5,..): tie
                  [1.1, 1.1, 1.1],      # nearest lives in the (+1,+1,+1)... same cell actually
                  [0.99, 0.99, 0.99],   # cell 0, nearest is in cell (1,1,1) diagonal
                  [0.0, 0.5, 0.5]], np.float32)
    gx, gd, gf = g.nn_single(q)
    ox, od, of, _ = o.nn_single(q)
    assert np.array_equal(gf, of) and np.array_equal(gx.view(np.uint32), ox.view(np.uint32))
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    assert tuple(gx[0]) == (0.5, 0.5, 0.5)  # first in cell order wins the tie


@pytest.mark.parametrize("res,min_pts", [(0.55, 2000), (1.6, 2000), (0.2, 10), (1.0, 10 ** 9)])
def test_decimate_first_bit_exact(ctx, world, res, min_pts):
    raw = world["frames"][3]["raw"]
    p = capi.decimate_params(res, min_pts)
    gi = ctx.voxel_decimate_first(raw, p)
    oi = O.decimate_first(raw, p)
    assert np.array_equal(gi, oi)
    if min_pts > len(raw):
        assert len(gi) == len(raw)


def test_decimate_permuted_and_small(ctx, world):
    raw = world["frames"][2]["raw"][:, :3].copy()
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(raw))
    p = capi.decimate_params(0.8, 100)
    a = ctx.voxel_decimate_first(raw[perm], p)
    b = O.decimate_first(raw[perm], p)
    assert np.array_equal(a, b)
    small = raw[:50]
    assert np.array_equal(ctx.voxel_decimate_first(small, p), np.arange(50, dtype=np.uint32))
    assert len(ctx.voxel_decimate_first(np.zeros((0, 3), np.float32), p)) == 0
    # predicates fused in front
    p2 = capi.decimate_params(0.8, 100, (3.0, 40.0), ((-5, -5, -1), (5, 5, 1)))
    assert np.array_equal(ctx.voxel_decimate_first(raw, p2), O.decimate_first(raw, p2))


def test_filter_1st_pass_bit_exact(ctx, world):
    for k in (0, 7, 13):
        raw = world["frames"][k]["raw"]
        for R in (100.0, 60.0):
            fp = capi.filter1_default(R)
            ga, gb = ctx.filter_1st_pass(raw, fp)
            oa, ob = O.filter_1st_pass(raw, fp)
            assert np.array_equal(ga.view(np.uint32), oa.view(np.uint32))
            assert np.array_equal(gb.view(np.uint32), ob.view(np.uint32))
            assert len(gb) > 500


def _build_pair(ctx, world, n_frames=12, **kw):
    g, o = _mk_maps(ctx, **kw)
    for fr in world["frames"][:n_frames]:
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
    return g, o


def _check_result(gr, orr, it_slack=0):
    et, er = O.pose_error(gr.pose, orr.pose)
    assert et <= TOL_TRANS_M and er <= TOL_ROT_DEG, f"pose differs: {et} m, {er} deg"
    assert gr.termination == orr.termination
    assert abs(int(gr.n_iterations) - int(orr.n_iterations)) <= it_slack
    if it_slack == 0:
        assert gr.n_pairings == orr.n_pairings
        assert gr.n_potential_pairings == orr.n_potential_pairings
        assert gr.n_candidate_points == orr.n_candidate_points
        assert gr.n_query_iterations == orr.n_query_iterations
        assert abs(gr.quality - orr.quality) < 1e-12


def test_icp_align_parity(ctx, world):
    g, o = _build_pair(ctx, world)
    rng = np.random.default_rng(11)
    worst = (0.0, 0.0)
    for k in (12, 14, 17, 20):
        fr = world["frames"][k]
        init = synth.perturb(fr["gt"], rng, 0.3, 1.0)
        ip = capi.IcpParamsOwner(sigma=2.0)
        gr = ctx.icp_align(fr["icp_layer"], g, init, ip.p)
        orr = O.icp_align(o, fr["icp_layer"], init, ip.p)
        _check_result(gr, orr)
        worst = max(worst, O.pose_error(gr.pose, orr.pose))
        # converged near ground truth (sanity of the workload, not a parity claim)
        et, er = O.pose_error(gr.pose, fr["gt"])
        assert et < 0.3 and er < 0.5
        assert np.allclose(gr.cov, orr.cov, rtol=1e-6, atol=1e-12)
    print("worst GPU-vs-oracle pose delta (m, deg):", worst)


def test_icp_termination_reasons(ctx, world):
    g, o = _build_pair(ctx, world, 6)
    fr = world["frames"][6]
    # MaxIterations
    ip = capi.IcpParamsOwner(sigma=2.0, max_iterations=3)
    gr, orr = ctx.icp_align(fr["icp_layer"], g, fr["gt"], ip.p), O.icp_align(o, fr["icp_layer"], fr["gt"], ip.p)
    assert gr.termination == orr.termination == 3 and gr.n_iterations == orr.n_iterations == 3
    _check_result(gr, orr)
    # exhausted budget: maxIterations == 0 returns the initial pose
    ip0 = capi.IcpParamsOwner(sigma=2.0, max_iterations=0)
    ip0.p.max_iterations = 0
    gr = ctx.icp_align(fr["icp_layer"], g, fr["gt"], ip0.p)
    assert gr.termination == 3 and gr.n_iterations == 0 and np.array_equal(gr.pose, fr["gt"])
    # NoPairings: far away from the map
    far = synth.compose(fr["gt"], synth.pose34(5000, 0, 0, 0))
    ip = capi.IcpParamsOwner(sigma=2.0)
    gr, orr = ctx.icp_align(fr["icp_layer"], g, far, ip.p), O.icp_align(o, fr["icp_layer"], far, ip.p)
    assert gr.termination == orr.termination == 1 and gr.quality == 0.0
    # hook as data: large initial error triggers HookRequest at the first iteration that moves > 0.15 m
    init = synth.compose(fr["gt"], synth.pose34(0.6, 0.1, 0, 0.01))
    ip = capi.IcpParamsOwner(sigma=2.0)
    ip.set_hook(init, 0.15, 0.75)
    gr, orr = ctx.icp_align(fr["icp_layer"], g, init, ip.p), O.icp_align(o, fr["icp_layer"], init, ip.p)
    assert gr.termination == orr.termination == 5
    _check_result(gr, orr)


def test_icp_prior_and_kernels(ctx, world):
    g, o = _build_pair(ctx, world, 8)
    fr = world["frames"][8]
    rng = np.random.default_rng(5)
    init = synth.perturb(fr["gt"], rng, 0.2, 0.5)
    for kernel in (capi.KERNEL_NONE, capi.KERNEL_GM, capi.KERNEL_CAUCHY):
        ip = capi.IcpParamsOwner(sigma=1.0)
        ip.p.robust_kernel = kernel
        _check_result(ctx.icp_align(fr["icp_layer"], g, init, ip.p), O.icp_align(o, fr["icp_layer"], init, ip.p))
    ip = capi.IcpParamsOwner(sigma=1.0)
    info = np.diag([50.0, 50.0, 50.0, 2000.0, 2000.0, 2000.0])
    info[0, 1] = info[1, 0] = 5.0
    ip.set_prior(init, info)
    gr, orr = ctx.icp_align(fr["icp_layer"], g, init, ip.p), O.icp_align(o, fr["icp_layer"], init, ip.p)
    _check_result(gr, orr)
    # the prior pulls the solution towards the (wrong) initial guess
    free = ctx.icp_align(fr["icp_layer"], g, init, capi.IcpParamsOwner(sigma=1.0).p)
    assert O.pose_error(gr.pose, init)[0] < O.pose_error(free.pose, init)[0]


def test_icp_horn_solver(ctx, world):
    g, o = _build_pair(ctx, world, 8)
    fr = world["frames"][9]
    init = synth.perturb(fr["gt"], np.random.default_rng(2), 0.2, 0.5)
    ip = capi.IcpParamsOwner(sigma=1.0, max_iterations=40)
    ip.p.solver = capi.SOLVER_HORN
    gr, orr = ctx.icp_align(fr["icp_layer"], g, init, ip.p), O.icp_align(o, fr["icp_layer"], init, ip.p)
    _check_result(gr, orr, it_slack=1)


def test_icp_batch_matches_single(ctx, world):
    g, o = _build_pair(ctx, world)
    rng = np.random.default_rng(21)
    locals_, inits, owners = [], [], []
    for k in (12, 13, 15, 16, 18, 19, 21):
        fr = world["frames"][k]
        locals_.append(fr["icp_layer"])
        inits.append(synth.perturb(fr["gt"], rng, 0.3, 1.0))
        owners.append(capi.IcpParamsOwner(sigma=float(rng.uniform(1.0, 2.5))))
    locals_.append(np.zeros((0, 3), np.float32))      # ragged: an empty problem in the batch
    inits.append(np.eye(4)[:3])
    owners.append(capi.IcpParamsOwner(sigma=2.0))
    batch = ctx.icp_align_batch(locals_, g, np.stack(inits), [w.p for w in owners])
    for i, (l, T, w) in enumerate(zip(locals_, inits, owners)):
        single = ctx.icp_align(l, g, T, w.p)
        # the warp decomposition adapts to the batch size, so sums may differ in the last bits only
        assert np.allclose(batch[i].pose, single.pose, rtol=0, atol=1e-9)
        assert batch[i].n_iterations == single.n_iterations and batch[i].termination == single.termination
        _check_result(batch[i], O.icp_align(o, l, T, w.p))
    assert batch[-1].termination == 1  # NoPairings for the empty cloud


def test_scan_register_sequence_parity(ctx, world):
    """filter -> align -> insert (+cull) over consecutive scans, both sides starting from the same map."""
    from mola_lidar_odometry_b200.api import LocalMap
    g, o = _mk_maps(ctx)
    fr0 = world["frames"][0]
    g.insert(fr0["map_layer"], fr0["gt"])
    o.insert(fr0["map_layer"], fr0["gt"])
    gp = op = fr0["gt"]
    fp = world["fp"]
    for k in range(1, 10):
        raw = world["frames"][k]["raw"]
        ip = capi.IcpParamsOwner(sigma=2.0)
        gr = ctx.scan_register(g, raw, fp, gp, ip.p, insert=True, cull_dist=120.0)
        orr, _ = O.scan_register(o, raw, fp, op, ip.p, insert=True, cull_dist=120.0)
        _check_result(gr, orr, it_slack=1)
        gp, op = gr.pose, orr.pose
        gv, ov = g.stats(), o.stats()
        assert abs(gv[0] - ov[0]) <= max(2, ov[0] // 1000)


def test_scan_register_batch_and_resident(ctx, world):
    g, o = _build_pair(ctx, world)
    rng = np.random.default_rng(8)
    ks = (12, 14, 16, 18)
    raws = [world["frames"][k]["raw"] for k in ks]
    inits = np.stack([synth.perturb(world["frames"][k]["gt"], rng, 0.3, 1.0) for k in ks])
    owners = [capi.IcpParamsOwner(sigma=2.0) for _ in ks]
    fps = [world["fp"]] * len(ks)
    res = ctx.scan_register_batch(g, raws, fps, inits, [w.p for w in owners])
    d = ctx.upload_batch(raws)
    res2 = ctx.scan_register_batch_resident(g, d, fps, inits, [w.p for w in owners])
    for i, k in enumerate(ks):
        orr, _ = O.scan_register(o, raws[i], world["fp"], inits[i], owners[i].p)
        _check_result(res[i], orr)
        assert np.array_equal(res[i].pose, res2[i].pose)  # same decomposition: bit-identical
    assert ctx.launch_count > 0


# ------------------------------------------------------------------------------------------------ NDT / point-to-plane
def test_ndt_map_and_plane_query_bit_exact(ctx, world):
    """mola::NDT (ndt.yaml:234-254): insert with min_distance_between_points 0.2 and the hard point limit, per-voxel
    mean / normal / planarity, nearest-plane query — all bit-exact against the oracle."""
    g, o = _mk_maps(ctx, 1.0, 0, 0.2, 1 << 16, kind=capi.MAP_NDT)
    for fr in world["frames"][:10]:
        g.insert(fr["map_layer"], fr["gt"])
        o.insert(fr["map_layer"], fr["gt"])
    assert g.stats() == o.stats()
    _assert_maps_equal(g, o)
    fr = world["frames"][10]
    R, t = fr["gt"][:, :3], fr["gt"][:, 3]
    q = (fr["icp_layer"].astype(np.float64) @ R.T + t).astype(np.float32)
    gm, gn, gd, gf = g.nn_plane(q)
    om, on, od, of = o.nn_plane(q)
    assert np.array_equal(gf, of) and gf.mean() > 0.3
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(gm[gf].view(np.uint32), om[of].view(np.uint32))
    assert np.array_equal(gn[gf].view(np.uint32), on[of].view(np.uint32))
    # culling keeps the statistics of the surviving voxels
    s = fr["gt"][:, 3]
    g.cull(s, 25.0)
    o.cull(s, 25.0)
    _assert_maps_equal(g, o)
    gm, gn, gd, gf = g.nn_plane(q)
    om, on, od, of = o.nn_plane(q)
    assert np.array_equal(gf, of) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_icp_ndt_pipeline_parity(ctx, world):
    """lidar3d-ndt.yaml ICP: Matcher_Point2Plane first, Matcher_Points_DistanceThreshold for the rest, GN x1."""
    g, o = _build_pair(ctx, world, 12, voxel=1.0, cap=0, min_dist=0.2, kind=capi.MAP_NDT)
    rng = np.random.default_rng(31)
    for k in (12, 15, 19):
        fr = world["frames"][k]
        init = synth.perturb(fr["gt"], rng, 0.3, 1.0)
        ip = capi.IcpParamsOwner(sigma=2.0, pipeline="ndt")
        gr = ctx.icp_align(fr["icp_layer"], g, init, ip.p)
        orr = O.icp_align(o, fr["icp_layer"], init, ip.p)
        _check_result(gr, orr)
        assert gr.n_potential_pairings == 2 * len(fr["icp_layer"])
    # a point-to-plane matcher against a plain point map is rejected, not silently ignored
    g2, _ = _build_pair(ctx, world, 2)
    from mola_lidar_odometry_b200.api import MloError
    with pytest.raises(MloError):
        ctx.icp_align(world["frames"][3]["icp_layer"], g2, np.eye(4)[:3], capi.IcpParamsOwner(sigma=2.0, pipeline="ndt").p)


# ------------------------------------------------------------------------------------------------ FilterDeskew (row f1)
def test_deskew_and_xyzt_filter_bit_exact(ctx, world):
    raw = world["frames"][5]["raw"]
    rng = np.random.default_rng(12)
    t = rng.uniform(-0.05, 0.05, len(raw)).astype(np.float32)
    ga, gb = ctx.filter_1st_pass_xyzt(raw, t, world["fp"])
    oa, ob = O.filter_1st_pass_xyzt(raw, t, world["fp"])
    assert np.array_equal(ga.view(np.uint32), oa.view(np.uint32)) and np.array_equal(gb.view(np.uint32), ob.view(np.uint32))
    for twist in ([8.0, 0.2, -0.1, 0.01, -0.02, 0.5], [0, 0, 0, 0, 0, 0], [15.0, 0, 0, 0, 0, -0.9]):
        gd, od = ctx.deskew(ga, twist), O.deskew(oa, twist)
        assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))       # small-angle series: bit-exact
    # large angles take the libm branch: tolerance test (floating point, 1e-5 m)
    big = ga.copy()
    big[:, 3] = 1.0
    assert np.allclose(ctx.deskew(big, [1, 2, 3, 0.3, -0.4, 1.2]), O.deskew(big, [1, 2, 3, 0.3, -0.4, 1.2]), atol=1e-5)
    assert len(ctx.deskew(np.zeros((0, 4), np.float32), [0] * 6)) == 0


# ------------------------------------------------------------------------------------------------ scan sets
def test_scanset_layers_align_insert_parity(ctx, world):
    """mlo_scanset_*: three scans filtered in one pass, layers bit-exact vs the oracle; one align pass against THREE
    different local maps; one insert pass (+ cull); a ragged job list (empty cloud)."""
    from mola_lidar_odometry_b200.api import LocalMap, ScanSet
    frames, fp = world["frames"], world["fp"]
    S = 3
    gm = [LocalMap(ctx, 1.0, 20, 0.0, 1 << 16) for _ in range(S)]
    om = [O.OracleMap(1.0, 20, 0.0) for _ in range(S)]
    for s in range(S):                      # map s holds frames s .. s+5
        for fr in frames[s:s + 6]:
            gm[s].insert(fr["map_layer"], fr["gt"])
            om[s].insert(fr["map_layer"], fr["gt"])
    ks = [7, 9, 11]
    raws = [frames[k]["raw"] for k in ks]
    sset = ScanSet(ctx, S + 1)
    info = sset.filter([0, 1, 2, 3], raws + [np.zeros((0, raws[0].shape[1]), np.float32)], [fp] * 4)
    for s in range(S):
        assert (info[s].n_map, info[s].n_icp) == (len(frames[ks[s]]["map_layer"]), len(frames[ks[s]]["icp_layer"]))
        assert np.array_equal(sset.download(s, 0), frames[ks[s]]["map_layer"])
        assert np.array_equal(sset.download(s, 1), frames[ks[s]]["icp_layer"])
        assert np.array_equal(np.array(info[s].icp_min[:], np.float32), frames[ks[s]]["icp_layer"].min(axis=0))
        assert np.array_equal(np.array(info[s].icp_max[:], np.float32), frames[ks[s]]["icp_layer"].max(axis=0))
    assert (info[3].n_map, info[3].n_icp) == (0, 0)
    rng = np.random.default_rng(5)
    inits = np.stack([synth.perturb(frames[k]["gt"], rng, 0.3, 1.0) for k in ks] + [np.eye(4)[:3]])
    owners = [capi.IcpParamsOwner(sigma=float(rng.uniform(1.5, 2.5))) for _ in range(S + 1)]
    res = sset.align([0, 1, 2, 3], gm + [gm[0]], inits, [w.p for w in owners])
    for s in range(S):
        _check_result(res[s], O.icp_align(om[s], frames[ks[s]]["icp_layer"], inits[s], owners[s].p))
        single = ctx.icp_align(frames[ks[s]]["icp_layer"], gm[s], inits[s], owners[s].p)
        assert np.allclose(res[s].pose, single.pose, rtol=0, atol=1e-9) and res[s].n_iterations == single.n_iterations
    assert res[3].termination == 1           # NoPairings for the empty scan
    counts = sset.insert([0, 1, 2], gm, np.stack([r.pose for r in res[:S]]), cull=[0.0, 40.0, 40.0])
    for s in range(S):
        om[s].insert(frames[ks[s]]["map_layer"], np.array(res[s].pose))
        if s > 0:
            om[s].cull(np.array(res[s].pose)[:, 3], 40.0)
        assert counts[s] == gm[s].stats()
    # poses agree to ~1e-13 only, so a point within that distance of a voxel face may land next door: compare sizes
    for s in range(S):
        gv, ov = gm[s].stats(), om[s].stats()
        assert abs(gv[0] - ov[0]) <= max(2, ov[0] // 1000) and abs(gv[1] - ov[1]) <= max(4, ov[1] // 1000)
    sset.close()


def test_scanset_deskew_bit_exact(ctx, world, scene, traj):
    from mola_lidar_odometry_b200.api import ScanSet
    tw = synth.body_twists(traj)
    fp = world["fp"]
    raws, ts = zip(*[synth.scan_skewed(scene, traj[k], tw[k], scan_seed=1000 + k) for k in (3, 4)])
    ts = [t - 0.5 * (t.min() + t.max()) for t in ts]
    sset = ScanSet(ctx, 2)
    info = sset.filter([1, 0], list(raws), [fp, fp], ts=list(ts))
    with pytest.raises(Exception):
        sset.download(0, 1)                   # skewed layers: deskew first
    twists = np.stack([tw[3], tw[4]])
    info2 = sset.deskew([1, 0], twists)
    for j, slot in enumerate([1, 0]):
        a, b = O.filter_1st_pass_xyzt(raws[j], ts[j], fp)
        assert (info[j].n_map, info[j].n_icp) == (len(a), len(b)) == (info2[j].n_map, info2[j].n_icp)
        assert np.array_equal(sset.download(slot, 0).view(np.uint32), O.deskew(a, twists[j]).view(np.uint32))
        assert np.array_equal(sset.download(slot, 1).view(np.uint32), O.deskew(b, twists[j]).view(np.uint32))
        assert np.array_equal(np.array(info2[j].icp_max[:], np.float32), O.deskew(b, twists[j]).max(axis=0))
    sset.close()


def test_scanset_prefetch_is_transparent(ctx, world):
    """mlo_scanset_prefetch only moves the upload of the NEXT filter call onto the copy stream: announced, ignored
    (other clouds) and matched calls all give the layers and ICP results of the plain path."""
    from mola_lidar_odometry_b200.api import ScanSet
    import ctypes as C
    frames, fp = world["frames"], world["fp"]
    g, o = _build_pair(ctx, world)
    raws = [np.ascontiguousarray(frames[k]["raw"], dtype=np.float32) for k in (12, 13, 14, 15)]

    def announce(sset, clouds):
        pts = (C.c_void_p * len(clouds))(*[c.ctypes.data for c in clouds])
        n = (C.c_uint64 * len(clouds))(*[len(c) for c in clouds])
        ctx.check(ctx.lib.mlo_scanset_prefetch(sset.h, len(clouds), pts, n, clouds[0].shape[1]))

    sset = ScanSet(ctx, 2)
    rng = np.random.default_rng(3)
    ip = capi.IcpParamsOwner(sigma=2.0)
    announce(sset, raws[2:4])                         # for the call AFTER the next one
    info = sset.filter([0, 1], raws[0:2], [fp, fp])   # not the announced clouds: plain upload, announcement kept
    assert [i.n_icp for i in info] == [len(frames[12]["icp_layer"]), len(frames[13]["icp_layer"])]
    inits = np.stack([synth.perturb(frames[k]["gt"], rng, 0.3, 1.0) for k in (12, 13)])
    r0 = sset.align([0, 1], [g, g], inits, [ip.p, ip.p])   # the announced transfer is enqueued inside this call
    info = sset.filter([0, 1], raws[2:4], [fp, fp])   # matched: consumes the prefetched buffer
    for s, k in enumerate((14, 15)):
        assert np.array_equal(sset.download(s, 0), frames[k]["map_layer"])
        assert np.array_equal(sset.download(s, 1), frames[k]["icp_layer"])
    inits2 = np.stack([synth.perturb(frames[k]["gt"], rng, 0.3, 1.0) for k in (14, 15)])
    r1 = sset.align([0, 1], [g, g], inits2, [ip.p, ip.p])
    for s, k in enumerate((14, 15)):
        _check_result(r1[s], O.icp_align(o, frames[k]["icp_layer"], inits2[s], ip.p))
    announce(sset, raws[0:2])                         # announced and consumed with no compute call in between
    info = sset.filter([1, 0], raws[0:2], [fp, fp])
    assert np.array_equal(sset.download(1, 1), frames[12]["icp_layer"]) and np.array_equal(sset.download(0, 1), frames[13]["icp_layer"])
    sset.close()


def test_scanset_argument_errors(ctx, world):
    """Bad jobs are refused with MLO_ERR_INVALID_ARG (the adapter turns that into the exception the reference's worker
    latches, LidarOdometry.cpp:614-619) and leave the set usable."""
    from mola_lidar_odometry_b200.api import LocalMap, MloError, ScanSet
    frames, fp = world["frames"], world["fp"]
    raw = frames[3]["raw"]
    sset = ScanSet(ctx, 2)
    with pytest.raises(MloError):
        sset.filter([2], [raw], [fp])                       # slot out of range
    with pytest.raises(MloError):
        sset.filter([0, 0], [raw, raw], [fp, fp])           # two jobs for one slot
    g = LocalMap(ctx, 1.0, 20, 0.0, 1 << 14)
    ip = capi.IcpParamsOwner(sigma=2.0)
    sset.filter([1], [raw], [fp])
    with pytest.raises(MloError):
        sset.align([0], [g], np.eye(4)[None, :3], [ip.p])   # slot 0 holds no scan
    with pytest.raises(MloError):
        sset.insert([1, 1], [g, g], np.stack([np.eye(4)[:3]] * 2))   # two insert jobs for one map
    with pytest.raises(MloError):
        sset.deskew([1], np.zeros((1, 6)))                  # no skewed layers: the filter ran without timestamps
    res = sset.align([1], [g], np.eye(4)[None, :3], [ip.p])  # empty map: the matcher pairs nothing
    assert res[0].termination == 1 and res[0].n_pairings == 0
    assert sset.insert([1], [g], np.eye(4)[None, :3])[0][1] == len(frames[3]["map_layer"])
    sset.close()
    g.close()
