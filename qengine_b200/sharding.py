"""Shard geometry of the multi-GPU path (mirrors qengine_b200/csrc/transfers.cu; one process per GPU).

* pass 1 (visibility): tasks are dealt in granules of TASK_GRANULE (512), granule g belongs to rank g % world; every word
  of the bit matrix is therefore written by exactly one rank and the exchange is one sum all-reduce.
* passes 2/3 and the bounces: whole 32-slot slices in `world` equal contiguous runs; exchange buffers are padded to
  world * shard so that all-gathers use equal shards.
"""
from __future__ import annotations

TASK_GRANULE = 512


def task_owner(task: int, world: int) -> int:
    return (task // TASK_GRANULE) % world


def local_to_global_task(ltask: int, rank: int, world: int) -> int:
    return ((ltask // TASK_GRANULE) * world + rank) * TASK_GRANULE + ltask % TASK_GRANULE


def num_local_tasks(num_tasks: int, rank: int, world: int) -> int:
    granules = (num_tasks + TASK_GRANULE - 1) // TASK_GRANULE
    mine = granules // world + (1 if granules % world > rank else 0)
    return mine * TASK_GRANULE


def slice_shard(num_slices: int, rank: int, world: int):
    """-> (slice_lo, slice_hi, shard_slots, padded_slots)"""
    per = (num_slices + world - 1) // world
    lo, hi = min(num_slices, per * rank), min(num_slices, per * (rank + 1))
    return lo, hi, per * 32, per * 32 * world
