"""The path's only exchange step when the environment batch is sharded over the GPUs of one
box: the best cut, its argmax and the winner's spins.

ONE collective and no host synchronisation: every rank contributes a record of 8 + N bytes -- the
int64 key `(cut << 32) | (0xFFFFFFFF - global_env_id)` of its local best row (ties go to the lowest
global env id) followed by that row -- to an all-gather over NCCL/NVLink; every rank then picks
the record with the largest key locally.  Latency-bound (world * (8 + N) bytes); the results stay
device tensors, so the step loop never waits on the host.
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
    """Returns (best_cut int64 [], global_env_id int64 [], best_x bool [N]) as tensors on the
    input's device -- identical on every rank.  No `.item()`: nothing here blocks the host."""
    gid = th.arange(vs.shape[0], device=vs.device, dtype=th.int64) + rank * envs_per_rank
    keys = (vs.to(th.int64) << 32) | (_LOW - gid)
    local = keys.argmax()                         # keys are distinct (they embed the env id)
    record = th.cat([keys[local].reshape(1).view(th.uint8), xs[local].view(th.uint8)])
    if world > 1:
        gathered = th.empty((world, record.numel()), dtype=th.uint8, device=vs.device)
        dist.all_gather_into_tensor(gathered, record.unsqueeze(0), group=group)
    else:
        gathered = record.unsqueeze(0)
    all_keys = gathered[:, :8].contiguous().view(th.int64).reshape(-1)
    win = all_keys.argmax()
    key = all_keys[win]
    return key >> 32, _LOW - (key & _LOW), gathered[win, 8:].view(th.bool)
