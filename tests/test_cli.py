"""apps/mlo-lidar-odometry-cli: flag surface of the reference CLI (apps/mola-lidar-odometry-cli.cpp:84-161) for the
hot-path subset; on a GPU box it must reproduce the trajectory of the library-driven odometry in TUM format."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CLI = ROOT / "apps" / "mlo-lidar-odometry-cli"
YAML = ROOT / "pipelines" / "lidar3d-default.yaml"


def test_cli_usage_and_errors(built, tmp_path):
    r = subprocess.run([str(CLI), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--input-kitti-seq" in r.stderr and "--output-tum-path" in r.stderr
    assert subprocess.run([str(CLI)], capture_output=True).returncode == 2                       # -c is required
    assert subprocess.run([str(CLI), "-c", str(YAML), "--bogus"], capture_output=True).returncode == 2
    r = subprocess.run([str(CLI), "-c", str(YAML), "--input-bin-dir", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 2 and "no *.bin clouds" in r.stderr


@pytest.mark.gpu
def test_cli_matches_library_trajectory(ctx, scene, traj, tmp_path, monkeypatch):
    from mola_lidar_odometry_b200.host_api import LidarOdometry
    from oracle import oracle_py as O
    import os
    env = dict(os.environ, MOLA_OPTIMIZE_TWIST="false", MOLA_INITIAL_VX="8.0")
    monkeypatch.setenv("MOLA_OPTIMIZE_TWIST", "false")
    monkeypatch.setenv("MOLA_INITIAL_VX", "8.0")
    n = 12
    lo = LidarOdometry(ctx, YAML)
    for k in range(n):
        raw = scene.scan(traj[k], scan_seed=1000 + k)
        raw.tofile(tmp_path / f"{k:06d}.bin")                       # KITTI velodyne layout: x y z intensity float32
        lo.on_lidar(raw, k / 10.0)
    st, ps = lo.trajectory()
    out = tmp_path / "traj.tum"
    r = subprocess.run([str(CLI), "-c", str(YAML), "--input-bin-dir", str(tmp_path), "--output-tum-path", str(out)],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    tum = np.loadtxt(out)
    assert tum.shape == (len(st), 8)
    assert np.allclose(tum[:, 0], st, atol=1e-6)
    assert np.allclose(tum[:, 1:4], ps[:, :, 3], atol=1e-5)          # same library, same inputs: same trajectory
    assert np.allclose(np.linalg.norm(tum[:, 4:8], axis=1), 1.0, atol=1e-5)
    # --only-first-n / --skip-first-n
    r = subprocess.run([str(CLI), "-c", str(YAML), "--input-bin-dir", str(tmp_path), "--output-tum-path", str(out),
                        "--only-first-n", "5"], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and len(np.loadtxt(out)) == 5
    lo.close()


@pytest.mark.gpu
def test_cli_fleet_of_two_inputs(ctx, scene, tmp_path):
    """Comma-separated inputs run as a lock-step fleet on one GPU (the counterpart of `parallel -j` in eval/cli_kitti.sh):
    each trajectory equals the one of a stand-alone run; a shorter input simply ends earlier."""
    import os
    from mola_lidar_odometry_b200 import synth
    env = dict(os.environ, MOLA_OPTIMIZE_TWIST="false", MOLA_INITIAL_VX="8.0")
    lens = {"a": 8, "b": 6}
    for name, seed in (("a", 7), ("b", 8)):
        tr = synth.trajectory_T00(12, seed=seed)
        (tmp_path / name).mkdir()
        for k in range(lens[name]):
            scene.scan(tr[k], scan_seed=seed * 1000 + k).tofile(tmp_path / name / f"{k:06d}.bin")
    out = tmp_path / "fleet.tum"
    r = subprocess.run([str(CLI), "-c", str(YAML), "--input-bin-dir", f"{tmp_path / 'a'},{tmp_path / 'b'}", "--output-tum-path",
                        str(out)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    for name in ("a", "b"):
        solo = tmp_path / f"solo_{name}.tum"
        r = subprocess.run([str(CLI), "-c", str(YAML), "--input-bin-dir", str(tmp_path / name), "--output-tum-path", str(solo)],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        f, s1 = np.loadtxt(tmp_path / f"fleet_{name}.tum"), np.loadtxt(solo)
        assert f.shape == s1.shape == (lens[name], 8)
        assert np.allclose(f, s1, atol=2e-6)
