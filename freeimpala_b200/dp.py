"""Data-parallel host logic: one process per GPU, the trajectory batch sharded
across ranks, one sum-all-reduce of the flat gradient arena per step (inside fi_learner_step, NCCL).

The only host-side pieces are the shard arithmetic and the transport of the NCCL ids created by rank 0;
torch.distributed is the transport (the reference's MPI mains would use MPI_Bcast). They are backend-agnostic,
so the CPU test-suite runs them under gloo with world_size 2.
"""
from __future__ import annotations


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Trajectories [lo, hi) of a global batch owned by `rank`: contiguous, sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} not in [0, {world})")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_ids(num_players: int, rank: int, world: int, create_ids) -> bytes:
    """Rank 0 calls create_ids(num_players) -> bytes (FI_DP_ID_BYTES per player); every rank returns them."""
    import torch.distributed as dist
    box = [create_ids(num_players) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def init_learner_dp(learner, rank: int, world: int) -> None:
    """Join the learner (all players) to the data-parallel group of the current torch.distributed job."""
    if world <= 1:
        return
    ids = exchange_ids(learner.num_players, rank, world, type(learner).dp_create_ids)
    learner.dp_init(ids, rank, world)
