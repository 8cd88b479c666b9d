"""Independent checks of the oracle's nearest-neighbour search (SURVEY.md §4 "Implication for the build"):
a brute-force numpy search restricted to the 27-cell neighbourhood must agree with the hash-voxel oracle, on
random clouds (hypothesis) and on a synthetic LiDAR frame."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import oracle_py as O

I34 = np.eye(4)[:3]


def _brute(pts, q, voxel, cap):
    """Reference semantics: per-voxel first-`cap` points in input order; NN over the 27 cells around key(q)."""
    inv = np.float32(1.0) / np.float32(voxel)
    keys = (pts * inv).astype(np.int32)          # truncation toward zero, float32 arithmetic
    kept = np.zeros(len(pts), bool)
    counts = {}
    for i, k in enumerate(map(tuple, keys)):
        c = counts.get(k, 0)
        if c < cap:
            kept[i] = True
            counts[k] = c + 1
    P, K = pts[kept], keys[kept]
    out_d2 = np.full(len(q), np.inf, np.float32)
    out_found = np.zeros(len(q), bool)
    for j, p in enumerate(q):
        kq = (p * inv).astype(np.int32)
        m = np.all(np.abs(K - kq) <= 1, axis=1)
        if m.any():
            d = P[m] - p
            d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
            out_d2[j] = d2.min()
            out_found[j] = True
    return out_d2, out_found


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), voxel=st.sampled_from([0.5, 1.0, 2.0]), cap=st.sampled_from([1, 4, 20]))
def test_nn_matches_brute_force_random(seed, voxel, cap):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-6, 6, (400, 3)).astype(np.float32)
    q = rng.uniform(-8, 8, (150, 3)).astype(np.float32)
    m = O.OracleMap(voxel, cap)
    m.insert(pts, I34)
    _, d2, f, _ = m.nn_single(q)
    bd2, bf = _brute(pts, q, voxel, cap)
    assert np.array_equal(f, bf)
    assert np.array_equal(d2[f].view(np.uint32), bd2[bf].view(np.uint32))


def test_nn_matches_brute_force_lidar_frame(world):
    fr = world["frames"][0]
    pts = fr["map_layer"]
    m = O.OracleMap(1.0, 20)
    m.insert(pts, I34)
    q = world["frames"][1]["icp_layer"][:400]
    _, d2, f, ncand = m.nn_single(q)
    bd2, bf = _brute(pts, q, 1.0, 20)
    assert np.array_equal(f, bf) and np.array_equal(d2[f].view(np.uint32), bd2[bf].view(np.uint32))
    assert ncand > 0
