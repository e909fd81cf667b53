"""The path's only exchange step when the environment batch is sharded over the GPUs of one
box: the best cut, its argmax and the winner's spins.  One NCCL all-reduce(MAX)
of an int64 key (cut << 32 | ~global_env_id: ties go to the lowest global env id) and one
broadcast of the owner's row.  Latency-bound (16 B + N bytes); NVLink bandwidth is irrelevant.
"""
from __future__ import annotations

from typing import Tuple

import torch as th
import torch.distributed as dist

TEN = th.Tensor
_LOW = 0xFFFFFFFF


def local_best_key(vs: TEN, rank: int, envs_per_rank: int) -> TEN:
    """int64 [1]: max over local envs of (cut << 32) | (0xFFFFFFFF - global_env_id)."""
    gid = th.arange(vs.shape[0], device=vs.device, dtype=th.int64) + rank * envs_per_rank
    return ((vs.to(th.int64) << 32) | (_LOW - gid)).max().reshape(1)


def decode_key(key: int) -> Tuple[int, int]:
    return key >> 32, _LOW - (key & _LOW)


def best_allreduce(vs: TEN, xs: TEN, rank: int, world: int, envs_per_rank: int, group=None):
    """Returns (best_cut, global_env_id, best_x) -- identical on every rank."""
    key = local_best_key(vs, rank, envs_per_rank)
    if world > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
    best_cut, gid = decode_key(int(key.item()))
    owner, local_id = gid // envs_per_rank, gid % envs_per_rank
    row = xs[local_id].clone() if rank == owner else th.empty_like(xs[0])
    if world > 1:
        dist.broadcast(row, src=owner, group=group)
    return best_cut, gid, row
