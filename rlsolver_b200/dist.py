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


def _local_record(vs: TEN, xs: TEN, rank: int, envs_per_rank: int) -> TEN:
    """uint8 [8 + N]: key of the best local row, then the row."""
    n = xs.shape[1]
    if vs.is_cuda and vs.dtype == th.int64 and xs.dtype == th.bool and vs.is_contiguous() and xs.is_contiguous():
        from . import _lib                      # one kernel instead of a dozen tiny torch ops
        record = th.empty((8 + n,), dtype=th.uint8, device=vs.device)
        _lib.check(_lib.lib().rlsb_best_record(vs.data_ptr(), xs.data_ptr(), vs.shape[0], n, rank * envs_per_rank,
                                               record.data_ptr(), th.cuda.current_stream(vs.device).cuda_stream),
                   "best_record")
        return record
    gid = th.arange(vs.shape[0], device=vs.device, dtype=th.int64) + rank * envs_per_rank
    keys = (vs.to(th.int64) << 32) | (_LOW - gid)
    local = keys.argmax()                         # keys are distinct (they embed the env id)
    return th.cat([keys[local].reshape(1).view(th.uint8), xs[local].view(th.uint8)])


def best_allreduce(vs: TEN, xs: TEN, rank: int, world: int, envs_per_rank: int, group=None):
    """Returns (best_cut int64 [], global_env_id int64 [], best_x bool [N]) as tensors on the
    input's device -- identical on every rank.  No `.item()`: nothing here blocks the host."""
    record = _local_record(vs, xs, rank, envs_per_rank)
    if world > 1:
        gathered = th.empty((world, record.numel()), dtype=th.uint8, device=vs.device)
        dist.all_gather_into_tensor(gathered, record.unsqueeze(0), group=group)
    else:
        gathered = record.unsqueeze(0)
    if gathered.is_cuda:
        from . import _lib
        out2 = th.empty((2,), dtype=th.int64, device=vs.device)
        row = th.empty((xs.shape[1],), dtype=th.bool, device=vs.device)
        _lib.check(_lib.lib().rlsb_best_pick(gathered.data_ptr(), gathered.shape[0], xs.shape[1], out2.data_ptr(),
                                             row.data_ptr(), th.cuda.current_stream(vs.device).cuda_stream), "best_pick")
        return out2[0], out2[1], row
    all_keys = gathered[:, :8].contiguous().view(th.int64).reshape(-1)
    win = all_keys.argmax()
    key = all_keys[win]
    return key >> 32, _LOW - (key & _LOW), gathered[win, 8:].view(th.bool)
