"""Multi-GPU plumbing: independent tracking units sharded over ranks, one final gather.

The reference tracks on a single device (`device = cuda:0`, pixtrack/localization/pixloc_pose_refiners.py:35-39)
and a sequence is temporally serial (frame t starts from the pose of t-1, pixloc_tracker_r9.py:228-237), so the
only parallel axis across GPUs is independent units: sequences, objects, or independent frames
(BASELINE.json configs 4-5).  One process per GPU, every rank holds replicas of the
weights, NO collective on the data path; the per-unit results (pose 12 floats, failed flag, iterations) are
gathered once at the end.  Backend-agnostic (`nccl` on the GPUs, `gloo` in the CPU tests).
"""
from typing import List, Sequence

import torch
import torch.distributed as dist

RESULT_WIDTH = 16     # pose (12) + failed + n_iters + unit id + pad


def units_of_rank(n_units: int, world: int, rank: int) -> List[int]:
    """unit i -> rank i mod world (round robin keeps long and short sequences mixed)."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    return list(range(rank, n_units, world))


def pack_results(unit_ids: Sequence[int], T: torch.Tensor, failed: torch.Tensor, n_iters: torch.Tensor) -> torch.Tensor:
    """[n_local, RESULT_WIDTH] float32 on T's device."""
    n = len(unit_ids)
    out = torch.zeros((n, RESULT_WIDTH), dtype=torch.float32, device=T.device)
    if n:
        out[:, :12] = T.reshape(n, 12).float()
        out[:, 12] = failed.reshape(n).float()
        out[:, 13] = n_iters.reshape(n).float()
        out[:, 14] = torch.as_tensor(list(unit_ids), dtype=torch.float32, device=T.device)
    return out


def gather_results(local: torch.Tensor, n_units: int) -> torch.Tensor:
    """All ranks get the [n_units, RESULT_WIDTH] table ordered by unit id.  Ranks may hold different
    numbers of units; rows are padded to the maximum for the fixed-size all_gather."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        table = local
    else:
        per = (n_units + world - 1) // world
        buf = torch.full((per, RESULT_WIDTH), -1.0, dtype=torch.float32, device=local.device)
        buf[:local.shape[0]] = local
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        table = torch.cat(parts, 0)
        table = table[table[:, 14] >= 0]
    order = torch.argsort(table[:, 14])
    table = table[order]
    if table.shape[0] != n_units or not torch.equal(table[:, 14].long().cpu(), torch.arange(n_units)):
        raise RuntimeError('gathered result table does not cover every unit exactly once')
    return table


def max_over_ranks(value: float, device) -> float:
    """Timing reduction: the slowest rank defines the step time."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
