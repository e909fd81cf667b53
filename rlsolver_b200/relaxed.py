"""Relaxed (probabilistic) max-cut objective with its gradient, as one differentiable op.

`relaxed_cut(store, probs)` = `-(p0 + p1 - 2 p0 p1).sum(1)` over the graph's original edge list
(rlsolver/envs/env_k_spin.py:191-193; PIGNN `hamiltonian_maxcut`, rlsolver/methods/PIGNN/util.py:4-8).  Forward and
backward are the CUDA kernels of csrc/relaxed.cu; autograd sees a normal function.
"""
from __future__ import annotations

import torch as th

from . import _lib
from .graph_store import GraphStore, _ptr, _stream_ptr

TEN = th.Tensor


def _check(store: GraphStore, probs: TEN) -> TEN:
    if probs.device != store.device:
        raise RuntimeError(f"probs must live on {store.device} (rlsolver_b200 has no CPU path)")
    if probs.dim() != 2 or probs.shape[1] != store.num_nodes:
        raise IndexError(f"probs must be [num_envs, {store.num_nodes}], got {tuple(probs.shape)}")
    if probs.dtype != th.float32:
        raise TypeError("probs must be float32")
    return probs.contiguous()


class _RelaxedCut(th.autograd.Function):
    @staticmethod
    def forward(ctx, probs: TEN, store: GraphStore) -> TEN:
        p = _check(store, probs.detach())
        out = th.empty((p.shape[0],), dtype=th.float32, device=p.device)
        with store._op("relaxed_cut"):
            _lib.check(store._lib.rlsb_relaxed_cut(store._h, _ptr(p), p.shape[0], _ptr(out), _stream_ptr(store.device)),
                       "relaxed_cut")
        ctx.store = store
        ctx.save_for_backward(p)
        return out

    @staticmethod
    def backward(ctx, grad_out: TEN):
        (p,) = ctx.saved_tensors
        store = ctx.store
        go = grad_out.to(th.float32).contiguous()
        grad = th.empty_like(p)
        with store._op("relaxed_cut_grad"):
            _lib.check(store._lib.rlsb_relaxed_cut_grad(store._h, _ptr(p), _ptr(go), p.shape[0], _ptr(grad),
                                                        _stream_ptr(store.device)), "relaxed_cut_grad")
        return grad, None


def relaxed_cut(store: GraphStore, probs: TEN) -> TEN:
    """float32 [E]: minus the expected cut of independent Bernoulli(probs) spins; differentiable in `probs`."""
    return _RelaxedCut.apply(probs, store)


__all__ = ["relaxed_cut"]
