"""Bookkeeping for consuming torch's CUDA Philox stream inside fused kernels (csrc/philox.cuh).

`torch_call_geometry(device, numel)` reproduces ATen's calc_execution_policy for one distribution call
(torch.rand / rand_like / randint over `numel` elements): T = 256 * grid threads and the number of
Philox counter increments the call consumes.  `take(device, numel, calls)` returns the (seed, offset,
T, iters) a kernel needs to regenerate `calls` consecutive calls and advances the generator exactly
as those calls would have.
"""
from __future__ import annotations

from typing import Tuple

import torch as th


_MAX_GRID = {}      # device index -> SMs * (max threads per SM // 256): ATen's cap on the grid of a distribution kernel


def _max_grid(device: th.device) -> int:
    key = device.index if device.index is not None else th.cuda.current_device()
    cap = _MAX_GRID.get(key)
    if cap is None:
        props = th.cuda.get_device_properties(key)
        cap = _MAX_GRID[key] = props.multi_processor_count * (props.max_threads_per_multi_processor // 256)
    return cap


def _geometry(device: th.device, numel: int) -> Tuple[int, int]:
    grid = min(_max_grid(device), (numel + 255) // 256)
    threads = 256 * max(grid, 1)
    iters = (max(numel, 1) - 1) // (threads * 4) + 1
    return threads, iters


_CHECKED = set()     # device indices whose torch build was verified against this module's model of it


def self_check(device: th.device) -> None:
    """First-use check, once per device and process: `_geometry` restates ATen's private calc_execution_policy
    and csrc/philox.cuh restates its element -> (counter, subsequence) map.  A torch build that changes either
    would silently break "same seed, same flip sequence", so regenerate torch.randn / torch.rand draws with the
    library (rlsb_torch_randn, the same device code the fused kernels use) for a small, a one-round and a
    multi-round call and compare bit for bit, including the generator offset after the calls.  Raises
    RuntimeError on any difference.  The caller's generator state is left untouched."""
    idx = device.index if device.index is not None else th.cuda.current_device()
    if idx in _CHECKED or th.cuda.is_current_stream_capturing():
        return
    from . import _lib
    from .graph_store import on_device
    dev = th.device("cuda", idx)
    gen = generator(dev)
    saved = gen.get_state()
    try:
        cap = _max_grid(dev) * 256
        for numel in (1000, cap * 4, cap * 4 + 77, cap * 9 + 5):
            gen.manual_seed(0x5EED + numel)
            gen.set_offset(8)
            seed, off0 = int(gen.initial_seed()), int(gen.get_offset())
            want = th.randn((2, numel), dtype=th.float32, device=dev)       # ONE call of 2 * numel elements
            threads, iters = _geometry(dev, 2 * numel)
            if int(gen.get_offset()) != off0 + 4 * iters:
                raise RuntimeError(
                    f"rlsolver_b200.rng: torch.randn({2 * numel}) advanced the CUDA generator by "
                    f"{int(gen.get_offset()) - off0}, this module expects {4 * iters} "
                    f"(torch {th.__version__}: calc_execution_policy changed?)")
            got = th.empty((2 * numel,), dtype=th.float32, device=dev)
            with on_device(dev):
                _lib.check(_lib.lib().rlsb_torch_randn(got.data_ptr(), 2 * numel, seed, off0, None, threads, iters, 1,
                                                       th.cuda.current_stream(dev).cuda_stream), "torch_randn")
            if not th.equal(got, want.reshape(-1)):
                bad = int((got != want.reshape(-1)).sum())
                raise RuntimeError(
                    f"rlsolver_b200.rng: the in-kernel restatement of torch's CUDA Philox stream disagrees with "
                    f"torch.randn on {bad} of {2 * numel} elements (torch {th.__version__}); the fused-RNG paths "
                    f"would not reproduce the reference's flip sequence on this build")
    finally:
        gen.set_state(saved)
    _CHECKED.add(idx)


def torch_call_geometry(device: th.device, numel: int) -> Tuple[int, int]:
    self_check(device)
    return _geometry(device, numel)


def generator(device: th.device) -> th.Generator:
    idx = device.index if device.index is not None else th.cuda.current_device()
    return th.cuda.default_generators[idx]


def peek(device: th.device, numel: int) -> Tuple[int, int, int, int]:
    """(seed, offset, threads, iters) for calls over `numel` elements, without consuming anything."""
    gen = generator(device)
    threads, iters = torch_call_geometry(device, numel)
    return int(gen.initial_seed()) & 0xFFFFFFFFFFFFFFFF, int(gen.get_offset()), threads, iters


def advance(device: th.device, numel: int, calls: int) -> None:
    """Move the generator past `calls` consecutive distribution calls of `numel` elements each."""
    if calls <= 0 or numel <= 0:
        return
    gen = generator(device)
    _, iters = torch_call_geometry(device, numel)
    gen.set_offset(int(gen.get_offset()) + 4 * iters * int(calls))
