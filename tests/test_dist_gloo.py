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
