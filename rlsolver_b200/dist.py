"""The path's only exchange step when the environment batch is sharded over the GPUs of one
box: the best cut, its argmax and the winner's spins.

ONE collective and no host synchronisation: every rank contributes a record of 8 + N bytes -- the
64-bit key `((cut + 2^31) << 32) | (0xFFFFFFFF - global_env_id)` of its local best row (compared
unsigned: negative values order below positive ones, ties go to the lowest global env id; values are
saturated to the int32 range) followed by that row -- to an all-gather over NCCL/NVLink; every rank then picks
the record with the largest key locally.  Latency-bound (world * (8 + N) bytes); the results stay
device tensors, so the step loop never waits on the host.
"""
from __future__ import annotations

from typing import Tuple

import torch as th
import torch.distributed as dist

from .graph_store import on_device

TEN = th.Tensor
_LOW = 0xFFFFFFFF


_BIAS = 1 << 31


def _keys(vs: TEN, rank: int, envs_per_rank: int) -> TEN:
    """int64 [E] holding the 64-bit keys MINUS 2^63 (so signed int64 order == unsigned key order)."""
    gid = th.arange(vs.shape[0], device=vs.device, dtype=th.int64) + rank * envs_per_rank
    v = vs.to(th.int64).clamp(-_BIAS, _BIAS - 1)
    return (v << 32) | (_LOW - gid)          # ((v + 2^31) << 32 | low) - 2^63 == (v << 32) | low in two's complement


def local_best_key(vs: TEN, rank: int, envs_per_rank: int) -> TEN:
    """int64 [1]: the largest local key (in the signed form of _keys)."""
    return _keys(vs, rank, envs_per_rank).max().reshape(1)


def decode_key(key: int) -> Tuple[int, int]:
    """(value, global env id) of a signed-form key (see _keys)."""
    return key >> 32, _LOW - (key & _LOW)


def _to_wire(signed_keys: TEN) -> TEN:
    """signed form -> the unsigned wire form the kernels write (adds 2^63 = flips the top bit)."""
    return signed_keys ^ th.iinfo(th.int64).min


def _local_record(vs: TEN, xs: TEN, rank: int, envs_per_rank: int) -> TEN:
    """uint8 [8 + N]: key of the best local row, then the row."""
    n = xs.shape[1]
    if vs.is_cuda and vs.dtype == th.int64 and xs.dtype == th.bool and vs.is_contiguous() and xs.is_contiguous():
        from . import _lib                      # one kernel instead of a dozen tiny torch ops
        record = th.empty((8 + n,), dtype=th.uint8, device=vs.device)
        with on_device(vs.device):
            _lib.check(_lib.lib().rlsb_best_record(vs.data_ptr(), xs.data_ptr(), vs.shape[0], n, rank * envs_per_rank,
                                                   record.data_ptr(), th.cuda.current_stream(vs.device).cuda_stream),
                       "best_record")
        return record
    keys = _keys(vs, rank, envs_per_rank)
    local = keys.argmax()                         # keys are distinct (they embed the env id)
    return th.cat([_to_wire(keys[local].reshape(1)).view(th.uint8), xs[local].view(th.uint8)])


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
        with on_device(vs.device):
            _lib.check(_lib.lib().rlsb_best_pick(gathered.data_ptr(), gathered.shape[0], xs.shape[1], out2.data_ptr(),
                                                 row.data_ptr(), th.cuda.current_stream(vs.device).cuda_stream), "best_pick")
        return out2[0], out2[1], row
    all_keys = _to_wire(gathered[:, :8].reshape(-1).clone().view(th.int64))     # wire -> signed form
    win = all_keys.argmax()
    key = all_keys[win]
    return key >> 32, _LOW - (key & _LOW), gathered[win, 8:].view(th.bool)



class BestExchange:
    """The exchange with every buffer allocated once, so that record kernel -> ncclAllGather -> pick kernel is a
    fixed launch sequence: it can be captured into the CUDA graph of the step (NCCL collectives are capturable)
    and replays with no host work at all.  `ex(vs, xs)` takes bool rows [E, N]; `ex.packed(vs, packed, store)`
    takes packed tiles.  Returns (best_cut int64 [], global_env_id int64 [], best_x bool [N]) -- views of
    buffers owned by this object, overwritten by the next call."""

    def __init__(self, num_nodes: int, rank: int, world: int, envs_per_rank: int, device, group=None):
        self.n, self.rank, self.world, self.envs, self.group = int(num_nodes), rank, world, envs_per_rank, group
        self.device = th.device(device)
        stride = (8 + self.n + 7) // 8 * 8                 # records stay 8-byte aligned inside the gathered buffer
        self.stride = stride
        self.record = th.zeros((stride,), dtype=th.uint8, device=self.device)
        self.gathered = th.zeros((world, stride), dtype=th.uint8, device=self.device)
        self.out2 = th.zeros((2,), dtype=th.int64, device=self.device)
        self.row = th.zeros((self.n,), dtype=th.bool, device=self.device)

    def _finish(self):
        from . import _lib
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.record.unsqueeze(0), group=self.group)
            src = self.gathered
        else:
            src = self.record.unsqueeze(0)
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_best_pick_strided(src.data_ptr(), src.shape[0], self.n, self.stride,
                                                         self.out2.data_ptr(), self.row.data_ptr(),
                                                         th.cuda.current_stream(self.device).cuda_stream), "best_pick")
        return self.out2[0], self.out2[1], self.row

    def post(self, vs: TEN, xs: TEN) -> None:
        """First half: the record kernel (reads vs / xs).  Once it has run, vs / xs may be overwritten: a caller that
        pipelines the exchange behind its next step only has to order that step after this launch."""
        from . import _lib
        assert vs.dtype == th.int64 and xs.dtype == th.bool and vs.is_contiguous() and xs.is_contiguous()
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_best_record(vs.data_ptr(), xs.data_ptr(), vs.shape[0], self.n,
                                                   self.rank * self.envs, self.record.data_ptr(),
                                                   th.cuda.current_stream(self.device).cuda_stream), "best_record")

    def finish(self):
        """Second half: all-gather of the records + pick kernel (touches only this object's buffers)."""
        return self._finish()

    def __call__(self, vs: TEN, xs: TEN):
        self.post(vs, xs)
        return self._finish()

    def packed(self, vs: TEN, packed: TEN, store):
        from . import _lib
        assert vs.dtype == th.int64 and packed.dtype == th.int32 and packed.is_contiguous()
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_best_record_packed(vs.data_ptr(), packed.data_ptr(), vs.shape[0], self.n,
                                                          store.padded_nodes, self.rank * self.envs,
                                                          self.record.data_ptr(),
                                                          th.cuda.current_stream(self.device).cuda_stream),
                       "best_record_packed")
        return self._finish()


class PeerBestExchange:
    """The same exchange as ONE kernel over NVLink peer memory (csrc/peer_exchange.cu): every rank stores its record
    straight into a mailbox in each peer's HBM (mapped through CUDA IPC), raises an arrival word and polls its own --
    no NCCL call, one launch, and because the launch carries no per-call arguments it can be captured into the CUDA
    graph of the step.  Same interface and the same results as BestExchange.  One process per GPU of one box; the
    constructor is collective (the IPC handles travel through one all-gather of `group`)."""

    def __init__(self, num_nodes: int, rank: int, world: int, envs_per_rank: int, device, group=None):
        import ctypes as C
        from . import _lib
        self.n, self.rank, self.world, self.envs, self.group = int(num_nodes), rank, world, envs_per_rank, group
        self.device = th.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PeerBestExchange needs CUDA devices (use BestExchange / best_allreduce elsewhere)")
        lib = _lib.lib()
        hb = int(lib.rlsb_peer_exchange_handle_bytes())
        handle = (C.c_uint8 * hb)()
        self._h = C.c_void_p()
        # Every rank walks through the same collectives whatever fails locally (a rank that raised early would leave
        # the others waiting in the all-gather); the outcome is agreed on at the end and raised everywhere.
        error = None
        try:
            with on_device(self.device):
                _lib.check(lib.rlsb_peer_exchange_create(rank, world, self.n, C.byref(self._h), handle),
                           "peer_exchange_create")
        except Exception as exc:        # noqa: BLE001 - reported after the collectives
            error, self._h = exc, None
        mine = th.frombuffer(bytearray(handle), dtype=th.uint8).to(self.device)
        if world > 1:
            every = th.empty((world, hb), dtype=th.uint8, device=self.device)
            dist.all_gather_into_tensor(every, mine.unsqueeze(0), group=group)
        else:
            every = mine.unsqueeze(0)
        if error is None:
            try:
                with on_device(self.device):
                    _lib.check(lib.rlsb_peer_exchange_connect(self._h, every.cpu().contiguous().numpy().tobytes()),
                               "peer_exchange_connect")
            except Exception as exc:    # noqa: BLE001
                error = exc
        if world > 1:
            # also the barrier: nobody stores into a mailbox that is not mapped everywhere yet
            ok = th.tensor([0 if error is not None else 1], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0 and error is None:
                error = RuntimeError("another rank could not set up its mailbox")
        if error is not None:
            self.close()
            raise RuntimeError(f"PeerBestExchange: CUDA IPC / peer access between the ranks is not available "
                               f"({error}); use BestExchange (NCCL all-gather)") from error
        self.out2 = th.zeros((2,), dtype=th.int64, device=self.device)
        self.row = th.zeros((self.n,), dtype=th.bool, device=self.device)

    def __call__(self, vs: TEN, xs: TEN):
        from . import _lib
        assert vs.dtype == th.int64 and xs.dtype == th.bool and vs.is_contiguous() and xs.is_contiguous()
        assert xs.shape[1] == self.n
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_peer_exchange_best(self._h, vs.data_ptr(), xs.data_ptr(), vs.shape[0],
                                                          self.rank * self.envs, self.out2.data_ptr(), self.row.data_ptr(),
                                                          th.cuda.current_stream(self.device).cuda_stream),
                       "peer_exchange_best")
        return self.out2[0], self.out2[1], self.row

    def packed(self, vs: TEN, packed: TEN, store):
        from . import _lib
        assert vs.dtype == th.int64 and packed.dtype == th.int32 and packed.is_contiguous()
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_peer_exchange_best_packed(self._h, vs.data_ptr(), packed.data_ptr(), vs.shape[0],
                                                                 store.padded_nodes, self.rank * self.envs,
                                                                 self.out2.data_ptr(), self.row.data_ptr(),
                                                                 th.cuda.current_stream(self.device).cuda_stream),
                       "peer_exchange_best_packed")
        return self.out2[0], self.out2[1], self.row

    def status(self) -> Tuple[int, int]:
        """(calls completed, polls that timed out) -- synchronises the current stream.  A non-zero second number
        means a peer did not make the matching call within the bound and the results since then are not valid."""
        import ctypes as C
        from . import _lib
        out = (C.c_uint32 * 2)()
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_peer_exchange_status(self._h, out, th.cuda.current_stream(self.device).cuda_stream),
                       "peer_exchange_status")
        return int(out[0]), int(out[1])

    def close(self) -> None:
        from . import _lib
        if getattr(self, "_h", None):
            with on_device(self.device):
                _lib.lib().rlsb_peer_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001 - interpreter shutdown
            pass
