import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build the native pieces once per session (nvcc cross-compiles without a GPU)."""
    from mola_lidar_odometry_b200 import _build
    _build.build_all()
    return True


@pytest.fixture(scope="session")
def scene(built):
    from mola_lidar_odometry_b200 import synth
    return synth.Scene(42)


@pytest.fixture(scope="session")
def traj(built):
    from mola_lidar_odometry_b200 import synth
    return synth.trajectory_T00(200, seed=7)


@pytest.fixture(scope="session")
def ctx(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mola_lidar_odometry_b200.api import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def world(scene, traj):
    """A small keyframe map workload: decimated scans + ground-truth relative poses."""
    from mola_lidar_odometry_b200 import capi, synth
    from oracle import oracle_py as O
    fp = capi.filter1_default(100.0)
    T0 = traj[0]
    frames = []
    for k in range(0, 24):
        raw = scene.scan(traj[k], scan_seed=1000 + k)
        a, b = O.filter_1st_pass(raw, fp)
        frames.append(dict(raw=raw, map_layer=a, icp_layer=b, gt=synth.relative(T0, traj[k])))
    return dict(frames=frames, fp=fp)
