"""The [VERIFY] conventions (voxel-index rounding, Geman-McClure weight form, cull metric) are
switchable on BOTH sides - oracle: O.set_conventions(...), product: ctx.set_option('convention_*') - so that whoever
checks upstream's source flips a flag instead of editing kernels.  Every setting keeps GPU == oracle."""
import itertools

import numpy as np
import pytest

from mola_lidar_odometry_b200 import capi, synth
from oracle import oracle_py as O


@pytest.fixture
def conventions():
    """Set the same conventions on the oracle and (when given) on the device context; restore the defaults afterwards."""
    touched = []

    def setter(ctx=None, index_floor=0, gm_form=0, cull_metric=0):
        O.set_conventions(index_floor, gm_form, cull_metric)
        if ctx is not None:
            ctx.set_option("convention_index_floor", index_floor)
            ctx.set_option("convention_gm_form", gm_form)
            ctx.set_option("convention_cull_metric", cull_metric)
            touched.append(ctx)
    yield setter
    O.set_conventions(0, 0, 0)
    for c in touched:
        for k in ("convention_index_floor", "convention_gm_form", "convention_cull_metric"):
            c.set_option(k, 0)


def test_oracle_conventions_known_answers(built, conventions):
    pts = np.array([[0.3, 0.3, 0.3], [-0.3, 0.3, 0.3], [0.4, 0.2, 0.1], [-1.2, 0.1, 0.1]], np.float32)
    p = capi.decimate_params(1.0, 0)
    assert list(O.decimate_first(pts, p)) == [0, 3]                 # truncation: -0.3 shares cell 0 with +0.3
    conventions(index_floor=1)
    assert list(O.decimate_first(pts, p)) == [0, 1, 3]              # floor: -0.3 -> cell -1, -1.2 -> cell -2
    m = O.OracleMap(1.0, 20, 0.0)
    m.insert(np.array([[-0.5, 0.5, 0.5], [0.5, 0.5, 0.5]], np.float32), np.eye(4)[:3])
    assert m.stats()[0] == 2                                        # floor: two voxels (one under truncation)
    conventions(index_floor=0)
    m = O.OracleMap(1.0, 20, 0.0)
    m.insert(np.array([[-0.5, 0.5, 0.5], [0.5, 0.5, 0.5]], np.float32), np.eye(4)[:3])
    assert m.stats()[0] == 1
    # cull metric: voxel at cell (3, 3, 0), sensor at the origin, distance 4 cells
    for metric, kept in ((0, True), (1, False), (2, False)):        # max-norm 3 <= 4; L1 6 > 4; Euclid 18 > 16
        conventions(cull_metric=metric)
        m = O.OracleMap(1.0, 20, 0.0)
        m.insert(np.array([[3.5, 3.5, 0.5]], np.float32), np.eye(4)[:3])
        m.cull(np.zeros(3), 4.0)
        assert (m.stats()[0] == 1) == kept, metric


@pytest.mark.gpu
@pytest.mark.parametrize("index_floor,gm_form,cull_metric", [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 2), (1, 1, 2)])
def test_gpu_matches_oracle_under_every_convention(ctx, world, conventions, index_floor, gm_form, cull_metric):
    from mola_lidar_odometry_b200.api import LocalMap
    conventions(ctx, index_floor, gm_form, cull_metric)
    fp = world["fp"]
    # decimation: bit-exact kept indices
    raw = world["frames"][3]["raw"]
    p = capi.decimate_params(0.55, 2000)
    assert np.array_equal(ctx.voxel_decimate_first(raw, p), O.decimate_first(raw, p))
    a_g, b_g = ctx.filter_1st_pass(raw, fp)
    a_o, b_o = O.filter_1st_pass(raw, fp)
    assert np.array_equal(a_g.view(np.uint32), a_o.view(np.uint32)) and np.array_equal(b_g.view(np.uint32), b_o.view(np.uint32))
    # map: insert + cull bit-exact, NN bit-exact (the pruning bounds follow the cell geometry of the index convention)
    g, o = LocalMap(ctx, 1.0, 20, 0.0, 1 << 12), O.OracleMap(1.0, 20, 0.0)
    frames = [dict(f) for f in world["frames"]]
    for k, fr in enumerate(frames[:10]):
        layer = O.filter_1st_pass(fr["raw"], fp)[0]     # (the layers of the fixture were made under the default conventions)
        g.insert(layer, fr["gt"])
        o.insert(layer, fr["gt"])
        if k == 6:
            g.cull(fr["gt"][:, 3], 45.0)
            o.cull(fr["gt"][:, 3], 45.0)
        assert g.stats() == o.stats()
    gk, gc, gp = g.export()
    ok, oc, op = o.export()
    assert np.array_equal(gk, ok) and np.array_equal(gc, oc) and np.array_equal(gp.view(np.uint32), op.view(np.uint32))
    q = O.filter_1st_pass(frames[11]["raw"], fp)[1]
    gx, gd, gf = g.nn_single(q)
    ox, od, of, _ = o.nn_single(q)
    assert np.array_equal(gf, of) and np.array_equal(gd.view(np.uint32)[gf], od.view(np.uint32)[of])
    assert np.array_equal(gx.view(np.uint32)[gf], ox.view(np.uint32)[of])
    # ICP on every device path
    init = synth.perturb(frames[11]["gt"], np.random.default_rng(4), 0.3, 1.0)
    ip = capi.IcpParamsOwner(sigma=2.0)
    orr = O.icp_align(o, q, init, ip.p)
    for path in (1, 2, 3):
        ctx.set_option("align_path", path)
        try:
            gr = ctx.icp_align(q, g, init, ip.p)
        finally:
            ctx.set_option("align_path", 0)
        et, er = O.pose_error(gr.pose, orr.pose)
        assert et <= 1e-3 and er <= 1e-2
        assert (gr.n_iterations, gr.termination, gr.n_pairings, gr.n_candidate_points) == \
               (orr.n_iterations, orr.termination, orr.n_pairings, orr.n_candidate_points)


@pytest.mark.gpu
def test_conventions_change_results(ctx, world, conventions):
    """The switches are live: each alternative setting gives a different (self-consistent) answer than the default."""
    raw = world["frames"][3]["raw"]
    p = capi.decimate_params(0.55, 2000)
    base = ctx.voxel_decimate_first(raw, p)
    conventions(ctx, index_floor=1)
    assert len(ctx.voxel_decimate_first(raw, p)) != len(base)
