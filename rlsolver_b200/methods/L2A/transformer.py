"""`sub_set_sampling` of rlsolver/methods/L2A/transformer.py:335-353, the dREINFORCE sampler: the
N - top_k most certain bits keep `start_xs`; each of the top_k least certain columns is
resampled on the num_repeats-times repeated batch with `rand < |p - 0.5|`.  The reference loops
over the top_k columns in Python (rand_like + lt + scatter per column); here the loop is one
kernel that regenerates torch's Philox stream for those rand_like calls (csrc/samplers.cu)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch as th

from ... import _lib, rng
from ...graph_store import _ptr, _stream_ptr, on_device, require_cuda

TEN = th.Tensor


def sub_set_sampling(probs: TEN, start_xs: TEN, num_repeats: int, top_k: int,
                     _explicit_u: Optional[TEN] = None) -> Tuple[TEN, TEN]:
    device = require_cuda(probs.device)
    determinism = th.abs(probs - 0.5)
    max_k = probs.shape[1]
    top_values, top_ids = th.topk(determinism, k=max(max_k - top_k, 0), largest=True, dim=1)
    probs = probs.scatter(dim=1, index=top_ids, src=probs.gather(dim=1, index=top_ids).lt(0.5).float())

    xs = start_xs.repeat(num_repeats, 1)
    top_values, top_ids = th.topk(determinism, k=min(top_k, max_k), largest=False, dim=1)
    k = top_values.shape[1]
    rows, num_sims = xs.shape[0], start_xs.shape[0]
    if k > 0 and rows > 0:
        if xs.dtype != th.bool:
            raise TypeError("sub_set_sampling: start_xs must be a bool tensor")
        seed, offset, threads, iters = rng.peek(device, rows)
        u = None
        if _explicit_u is not None:
            u = _explicit_u.to(device=device, dtype=th.float32).contiguous()
            assert u.shape == (k, rows)
        with on_device(device):
            _lib.check(_lib.lib().rlsb_subset_sampling(_ptr(xs), rows, max_k, num_sims, k, _ptr(top_ids.contiguous()),
                                                       _ptr(top_values.float().contiguous()), _ptr(u), seed, offset,
                                                       threads, iters, _stream_ptr(device)), "subset_sampling")
        if u is None:
            rng.advance(device, rows, k)              # one rand_like(_prob) per resampled column
    return xs, probs
