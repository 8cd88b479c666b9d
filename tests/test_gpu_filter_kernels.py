"""The two device forms of the FirstPoint decimation chain (include/mlo_b200.h mlo_set_option "filter_kernel"):

  1  k_decim_claim / k_decim_finalize: global scratch tables, the blocks of a cloud spread over the device (small batches)
  2  k_decim_cta: one thread block per cloud, table and winner bitmap in shared memory (large batches)

Both must reproduce the oracle's indices / layers BIT FOR BIT (mp2p_icp_filters::FilterDecimateVoxels FirstPoint with
FilterByRange / FilterBoundingBox, pipelines/lidar3d-default.yaml:285-319), including the cases where form 2 cannot hold
a cloud (voxel indices outside its 32-bit key box, more voxels than its table) and the library repeats the batch with
form 1 on its own.
"""
import numpy as np
import pytest

from mola_lidar_odometry_b200 import capi
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


class _Options:
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


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("res,min_pts", [(0.55, 2000), (1.6, 2000), (0.2, 10), (1.0, 10 ** 9)])
def test_decimate_first_both_kernels(ctx, world, kernel, res, min_pts):
    raw = world["frames"][3]["raw"]
    p = capi.decimate_params(res, min_pts)
    oi = O.decimate_first(raw, p)
    with _Options(ctx, filter_kernel=kernel):
        gi = ctx.voxel_decimate_first(raw, p)
        # (a 0.2 m grid keeps more voxels of a 130 k-point sweep than the shared-memory table holds: form 2 hands over to form 1)
        assert ctx.get_option("last_filter_kernel") == (1 if len(oi) > 20_000 and min_pts <= len(raw) else kernel)
    assert np.array_equal(gi, oi)


@pytest.mark.parametrize("kernel", [1, 2])
def test_predicates_pass_through_and_empty(ctx, world, kernel):
    raw = world["frames"][2]["raw"][:, :3].copy()
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(raw))
    p = capi.decimate_params(0.8, 100)
    with _Options(ctx, filter_kernel=kernel):
        assert np.array_equal(ctx.voxel_decimate_first(raw[perm], p), O.decimate_first(raw[perm], p))
        # fewer inputs than minimum_input_points_to_filter: everything passes
        assert np.array_equal(ctx.voxel_decimate_first(raw[:50], p), np.arange(50, dtype=np.uint32))
        assert len(ctx.voxel_decimate_first(np.zeros((0, 3), np.float32), p)) == 0
        # predicates in front of the decimation
        p2 = capi.decimate_params(0.8, 100, (3.0, 40.0), ((-5, -5, -1), (5, 5, 1)))
        assert np.array_equal(ctx.voxel_decimate_first(raw, p2), O.decimate_first(raw, p2))
        # ... that leave fewer survivors than minimum_input_points_to_filter: the survivors pass undecimated
        p3 = capi.decimate_params(0.8, 100_000, (3.0, 6.0))
        assert len(raw) > 100_000
        a, b = ctx.voxel_decimate_first(raw, p3), O.decimate_first(raw, p3)
        assert 0 < len(b) < 100_000 and np.array_equal(a, b)
        r2 = (raw.astype(np.float32) ** 2).sum(axis=1)
        assert len(b) == int(((r2 >= np.float32(9.0)) & (r2 <= np.float32(36.0))).sum())   # every survivor of the range test


@pytest.mark.parametrize("kernel", [1, 2])
def test_filter_chain_both_kernels(ctx, world, kernel):
    with _Options(ctx, filter_kernel=kernel):
        for k in (0, 7):
            raw = world["frames"][k]["raw"]
            for R in (100.0, 60.0):
                fp = capi.filter1_default(R)
                ga, gb = ctx.filter_1st_pass(raw, fp)
                oa, ob = O.filter_1st_pass(raw, fp)
                assert np.array_equal(ga.view(np.uint32), oa.view(np.uint32))
                assert np.array_equal(gb.view(np.uint32), ob.view(np.uint32))


def test_ragged_batch_and_policy(ctx, world):
    """40 clouds of different sizes (one empty, one tiny) in one filter pass: by default a batch of this size takes the
    block-per-cloud kernel; layers bit-exact either way."""
    from mola_lidar_odometry_b200.api import ScanSet
    frames, fp = world["frames"], world["fp"]
    rng = np.random.default_rng(3)
    NB = 40
    clouds = []
    for s in range(NB):
        raw = frames[s % len(frames)]["raw"]
        n = int(rng.integers(3000, len(raw)))
        clouds.append(np.ascontiguousarray(raw[:n]))
    clouds[5] = np.zeros((0, clouds[0].shape[1]), np.float32)
    clouds[9] = np.ascontiguousarray(clouds[9][:37])
    want = [O.filter_1st_pass(c, fp) if len(c) else (np.zeros((0, 3), np.float32),) * 2 for c in clouds]
    sset = ScanSet(ctx, NB)
    for kernel, expect in ((0, 2), (1, 1), (2, 2)):
        with _Options(ctx, filter_kernel=kernel):
            info = sset.filter(list(range(NB)), clouds, [fp] * NB)
            assert ctx.get_option("last_filter_kernel") == expect
        for s in range(NB):
            assert (info[s].n_map, info[s].n_icp) == (len(want[s][0]), len(want[s][1])), (kernel, s)
            if len(clouds[s]):
                assert np.array_equal(sset.download(s, 0), want[s][0])
                assert np.array_equal(sset.download(s, 1), want[s][1])
    # a small batch stays with the spread-out kernels
    with _Options(ctx, filter_kernel=0):
        sset.filter([0, 1], clouds[:2], [fp] * 2)
        assert ctx.get_option("last_filter_kernel") == 1
    sset.close()


def test_block_per_cloud_kernel_falls_back_on_its_own(ctx, world):
    """(a) a grid so fine that points sit more than 1024 cells from the origin, (b) more voxels than the shared-memory
    table holds: the library repeats the batch with the global-table kernels; the caller sees exact results and no error."""
    raw = world["frames"][4]["raw"]
    far = np.abs(raw[:, :3]).max()
    with _Options(ctx, filter_kernel=2):
        res = float(far) / 1500.0                       # -> indices up to ~1500 cells
        p = capi.decimate_params(res, 10)
        gi = ctx.voxel_decimate_first(raw, p)
        assert ctx.get_option("last_filter_kernel") == 1
        assert np.array_equal(gi, O.decimate_first(raw, p))
        rng = np.random.default_rng(1)                   # 200 k points, nearly all in voxels of their own, inside the key box
        cloud = rng.uniform(-60.0, 60.0, (200_000, 3)).astype(np.float32)
        p = capi.decimate_params(0.25, 10)
        gi = ctx.voxel_decimate_first(cloud, p)
        assert ctx.get_option("last_filter_kernel") == 1
        oi = O.decimate_first(cloud, p)
        assert np.array_equal(gi, oi) and len(oi) > 150_000
        # and the next well-behaved cloud goes through the block-per-cloud kernel again (explicit choice: no back-off)
        p = capi.decimate_params(0.8, 100)
        assert np.array_equal(ctx.voxel_decimate_first(raw, p), O.decimate_first(raw, p))
        assert ctx.get_option("last_filter_kernel") == 2


def test_timestamps_ride_through_both_kernels(ctx, world):
    raw = world["frames"][6]["raw"][:, :3].copy()
    t = np.linspace(-0.05, 0.05, len(raw)).astype(np.float32)
    fp = capi.filter1_default(100.0)
    outs = []
    for kernel in (1, 2):
        with _Options(ctx, filter_kernel=kernel):
            outs.append(ctx.filter_1st_pass_xyzt(raw, t, fp))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))
