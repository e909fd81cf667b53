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


def torch_call_geometry(device: th.device, numel: int) -> Tuple[int, int]:
    grid = min(_max_grid(device), (numel + 255) // 256)
    threads = 256 * max(grid, 1)
    iters = (max(numel, 1) - 1) // (threads * 4) + 1
    return threads, iters


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
