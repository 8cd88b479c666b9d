"""Multi-GPU plumbing: one process per GPU, whole units (scans / sequences / Monte-Carlo poses) sharded across
ranks, no data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only for the
barrier around the timed region, the max-over-ranks clock, and one all_gather of fixed-size result blocks
(the reference's analogue is `parallel -j` over sequences, eval/cli_kitti.sh:23)."""
from __future__ import annotations

import numpy as np

RESULT_BLOCK = 16  # doubles per unit: pose 3x4 (12), quality, iterations, termination, unit id


def shard_units(n_units: int, rank: int, world: int) -> list[int]:
    """Rank r owns units {u : u mod world == r}."""
    return [u for u in range(n_units) if u % world == rank]


def pack_results(unit_ids, results) -> np.ndarray:
    out = np.zeros((len(unit_ids), RESULT_BLOCK), dtype=np.float64)
    for i, (u, r) in enumerate(zip(unit_ids, results)):
        out[i, :12] = np.asarray(r.pose_3x4[:], dtype=np.float64)
        out[i, 12] = r.quality
        out[i, 13] = r.n_iterations
        out[i, 14] = r.termination
        out[i, 15] = u
    return out


def gather_results(local_block: np.ndarray, n_units: int, device=None) -> np.ndarray:
    """all_gather of per-rank result blocks -> [n_units, RESULT_BLOCK] ordered by unit id on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        full = np.zeros((n_units, RESULT_BLOCK))
        full[local_block[:, 15].astype(int)] = local_block
        return full
    world = dist.get_world_size()
    per = (n_units + world - 1) // world
    buf = torch.full((per, RESULT_BLOCK), -1.0, dtype=torch.float64, device=device)
    if len(local_block):
        buf[:len(local_block)] = torch.from_numpy(local_block).to(buf.device)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    full = np.zeros((n_units, RESULT_BLOCK))
    for o in outs:
        a = o.cpu().numpy()
        a = a[a[:, 15] >= 0]
        full[a[:, 15].astype(int)] = a
    return full


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
