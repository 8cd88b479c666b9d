"""world_size-2 gloo test of the multi-GPU plumbing (sharding + result gather + max-over-ranks clock)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _R:
    def __init__(self, u):
        self.pose_3x4 = [float(u + k) for k in range(12)]
        self.quality = 0.5 + u
        self.n_iterations = 10 + u
        self.termination = 4


def _worker(rank, world, port, n_units, q):
    import torch.distributed as dist
    from mola_lidar_odometry_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = D.shard_units(n_units, rank, world)
    block = D.pack_results(mine, [_R(u) for u in mine])
    full = D.gather_results(block, n_units)
    t = D.max_over_ranks(1.0 + rank)
    dist.barrier()
    q.put((rank, mine, full, t))
    dist.destroy_process_group()


def test_shard_and_gather_two_ranks():
    from mola_lidar_odometry_b200 import dist as D
    n_units, world = 7, 2
    assert D.shard_units(n_units, 0, 2) == [0, 2, 4, 6] and D.shard_units(n_units, 1, 2) == [1, 3, 5]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in ps:
        p.start()
    outs = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mine, full, t in outs:
        assert t == 2.0                                   # max over ranks
        assert full.shape == (n_units, D.RESULT_BLOCK)
        for u in range(n_units):                          # every rank sees every unit's block, ordered by unit id
            assert full[u, 15] == u and full[u, 13] == 10 + u and full[u, 0] == float(u)


def test_single_process_gather_is_identity():
    from mola_lidar_odometry_b200 import dist as D
    block = D.pack_results([2, 0, 1], [_R(2), _R(0), _R(1)])
    full = D.gather_results(block, 3)
    assert np.array_equal(full[:, 15], [0, 1, 2])


# ---------------------------------------------------------------------------------------------- fleets sharded over ranks
class _Out:
    def __init__(self, o):
        self.pose_3x4, self.quality, self.n_iterations, self.termination = o.pose_3x4, o.quality, o.icp_iterations, o.termination


def _fleet_steps(n_seq, n_scans):
    from mola_lidar_odometry_b200 import synth
    scene = synth.Scene(42)
    trajs = [synth.trajectory_T00(n_scans + 2, seed=7 + s) for s in range(n_seq)]
    return [[scene.scan(trajs[s][k], scan_seed=(7 + s) * 1000 + k) for s in range(n_seq)] for k in range(n_scans)]


def _run_fleet(units, steps, yaml_path):
    from oracle import oracle_py as O
    fleet = O.OracleLidarOdometryFleet(yaml_path, len(units))
    outs = None
    for k, clouds in enumerate(steps):
        outs = fleet.on_lidar([clouds[u] for u in units], [0.1 * k] * len(units))
    return [_Out(o) for o in outs]


def _fleet_worker(rank, world, port, n_seq, n_scans, yaml_path, q):
    import torch.distributed as dist
    from mola_lidar_odometry_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MOLA_OPTIMIZE_TWIST"] = "false"
    os.environ["MOLA_INITIAL_VX"] = "8.0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = D.shard_units(n_seq, rank, world)          # rank r owns sequences {s : s mod world == r}, one fleet per rank
    outs = _run_fleet(mine, _fleet_steps(n_seq, n_scans), yaml_path)
    full = D.gather_results(D.pack_results(mine, outs), n_seq)
    dist.barrier()
    q.put((rank, full))
    dist.destroy_process_group()


def test_fleet_sharded_over_two_ranks_equals_one_fleet(built, monkeypatch):
    """bench.py --gpus N --workload sequence on CPU: every rank runs a lock-step fleet over its shard of the sequences
    (here over the oracle backend); the gathered final poses are those of ONE fleet holding all sequences."""
    from pathlib import Path
    from mola_lidar_odometry_b200 import dist as D
    monkeypatch.setenv("MOLA_OPTIMIZE_TWIST", "false")
    monkeypatch.setenv("MOLA_INITIAL_VX", "8.0")
    yaml_path = str(Path(__file__).resolve().parent.parent / "pipelines" / "lidar3d-default.yaml")
    n_seq, n_scans, world = 3, 4, 2
    ref = D.pack_results(list(range(n_seq)), _run_fleet(list(range(n_seq)), _fleet_steps(n_seq, n_scans), yaml_path))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_fleet_worker, args=(r, world, port, n_seq, n_scans, yaml_path, q)) for r in range(world)]
    for p in ps:
        p.start()
    outs = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full in outs:
        assert np.array_equal(full, ref)              # bit-identical: sharding changes nothing per sequence
