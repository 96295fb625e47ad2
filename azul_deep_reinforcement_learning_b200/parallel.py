"""Multi-GPU plumbing: one process per GPU, games sharded by global id, no data-path collective.

Games are independent (`azul.py:17-61` holds no shared state), so rank r simply owns the global game
ids ``[r * games_per_rank, (r + 1) * games_per_rank)``.  The Philox schedule is keyed by global id,
which makes every result independent of the number of ranks.  The only collectives are the
16-counter statistic reduction (C2) and the max-over-ranks timing; both go through
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard(rank, games_per_rank):
    """First global game id owned by ``rank`` (the ``game_id_base`` of its engine handle)."""
    return rank * games_per_rank


def init(backend="nccl", local_rank=0):
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)


def reduce_counters(counters):
    """Sum the rollout counters over all ranks (in place); no-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (device-timed durations are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
