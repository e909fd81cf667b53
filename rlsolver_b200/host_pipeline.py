"""Batches that live in HOST memory, streamed through `EnvMaxcut.local_search_inplace` / `local_search_packed`.

A caller that keeps its spins on the host (the reference's solvers hold `xs` as torch tensors and move them with
`.to(device)`, rlsolver/methods/L2A/maxcut_simulator... -> envs/env_L2A.py:87) pays PCIe for every step: 8 MB in and
8 MB out per G22 x 4096 batch as bool rows, 1 MB as packed tiles.  This class hides that behind the step:

    H2D(i+1)  |  step(i)  |  D2H(i-1)

on three CUDA streams with `depth` (default 3) device buffers, so that the copy engines of both directions and the SMs
work at the same time; every buffer has its own captured CUDA graph of the step (the search runs in place on the
buffer, no device-to-device staging), so the host side of a step is one replay plus two async copies.  Random numbers
come from the simulator's device-resident generator cursor exactly as in any captured local search: batch i sees the
draws eager calls number i would have seen.

    pipe = HostPipeline(sim, num_envs, layout="bool")          # or "packed"
    for h_in, h_out, h_vs in batches:                           # pinned host tensors
        pipe.submit(h_in, h_out, h_vs)
    pipe.drain()                                                # h_out / h_vs of every submitted batch are complete
"""
from __future__ import annotations

from typing import Callable, Optional

import torch as th

from .graph_store import on_device

TEN = th.Tensor


class HostPipeline:
    def __init__(self, sim, num_envs: int, layout: str = "bool", num_iters: int = 8, num_spin: int = 8,
                 noise_std: float = 0.3, depth: int = 3, pre_step: Optional[Callable[[], None]] = None,
                 in_graph_post: Optional[Callable] = None, eager_post: Optional[Callable] = None):
        """layout "bool": host batches are bool [E, N] rows; "packed": int32 [ceil(E/32), Np] tiles (store.pack).
        pre_step(): captured in front of the search (a benchmark's L2-evicting write); in_graph_post(xs_or_packed, vs):
        captured behind it (a kernel-only exchange such as PeerBestExchange); eager_post(xs_or_packed, vs): issued
        eagerly behind every replay (calls that cannot be captured, e.g. an NCCL collective)."""
        if layout not in ("bool", "packed"):
            raise ValueError("layout must be 'bool' or 'packed'")
        if depth < 2:
            raise ValueError("depth must be at least 2")
        self.sim, self.envs, self.layout, self.depth = sim, int(num_envs), layout, int(depth)
        self.device = sim.device
        with on_device(self.device):         # streams, graphs and events belong to the simulator's device
            self._build(num_iters, num_spin, noise_std, pre_step, in_graph_post, eager_post)

    def _build(self, num_iters, num_spin, noise_std, pre_step, in_graph_post, eager_post) -> None:
        sim, layout, depth, dev = self.sim, self.layout, self.depth, self.device
        st = sim.store
        self.compute = th.cuda.current_stream(dev)
        self.s_in, self.s_out = th.cuda.Stream(device=dev), th.cuda.Stream(device=dev)
        if layout == "bool":
            shape, dtype = (self.envs, sim.num_nodes), th.bool
        else:
            shape, dtype = (st.tiles(self.envs), st.padded_nodes), th.int32
        self.d_in = [th.zeros(shape, dtype=dtype, device=dev) for _ in range(depth)]
        self.d_out = self.d_in if layout == "bool" else [th.zeros(shape, dtype=dtype, device=dev) for _ in range(depth)]
        self.eager_post = eager_post
        sentinel = th.empty(())

        def body(b):
            if pre_step is not None:
                pre_step()
            if layout == "bool":
                xs, vs = sim.local_search_inplace(self.d_in[b], sentinel, num_iters, num_spin, noise_std)
                res = xs
            else:
                pk, vs = sim.local_search_packed(self.d_in[b], num_iters, num_spin, noise_std, num_sims=self.envs)
                self.d_out[b].copy_(pk)              # the result sits in the shared workspace: keep a copy per buffer
                res = self.d_out[b]
            if in_graph_post is not None:
                in_graph_post(res, vs)
            return res, vs

        # one eager pass (allocations, lazily loaded kernels), then one graph per buffer
        side = th.cuda.Stream(device=dev)
        side.wait_stream(self.compute)
        with th.cuda.stream(side):
            body(0)
        self.compute.wait_stream(side)
        th.cuda.synchronize(dev)
        st.rng_cursor_sync()
        self.graphs, self.outs = [], []
        for b in range(depth):
            g = th.cuda.CUDAGraph()
            # one memory pool for all of them: they only ever replay one after the other, in capture order
            with th.cuda.graph(g, pool=self.graphs[0].pool() if self.graphs else None):
                out = body(b)
            self.graphs.append(g)
            self.outs.append(out)
        th.cuda.synchronize(dev)
        self.in_ready = [th.cuda.Event() for _ in range(depth)]
        self.done = [th.cuda.Event() for _ in range(depth)]
        self.buf_free = [th.cuda.Event() for _ in range(depth)]
        for e in self.buf_free:
            e.record(self.compute)
        self.count = 0

    def submit(self, h_in: TEN, h_out: TEN, h_vs: TEN) -> None:
        """Enqueues one batch: h_in -> device, the step, results -> h_out (same layout as h_in) and h_vs (int64 [E]).
        Returns at once; the host tensors must stay alive (and pinned, for the copies to be asynchronous) until
        `drain()`."""
        with on_device(self.device):
            self._submit(h_in, h_out, h_vs)

    def _submit(self, h_in: TEN, h_out: TEN, h_vs: TEN) -> None:
        b = self.count % self.depth
        self.count += 1
        with th.cuda.stream(self.s_in):
            self.s_in.wait_event(self.buf_free[b])           # the D2H of the batch that used this buffer last
            self.d_in[b].copy_(h_in, non_blocking=True)
            self.in_ready[b].record(self.s_in)
        with th.cuda.stream(self.compute):                  # whatever stream the caller is on right now
            self.compute.wait_event(self.in_ready[b])
            self.graphs[b].replay()
            res, vs = self.outs[b]
            if self.eager_post is not None:
                self.eager_post(res, vs)
            self.done[b].record(self.compute)
        with th.cuda.stream(self.s_out):
            self.s_out.wait_event(self.done[b])
            h_out.copy_(res, non_blocking=True)
            h_vs.copy_(vs, non_blocking=True)
            self.buf_free[b].record(self.s_out)

    def drain(self) -> None:
        """Blocks until every submitted batch has landed in its host tensors."""
        self.compute.wait_stream(self.s_out)
        self.compute.wait_stream(self.s_in)
        self.compute.synchronize()

    def sync_rng(self) -> None:
        """device cursor <- torch's CUDA generator (call after seeding, before the next submit)."""
        self.sim.store.rng_cursor_sync()

    def commit_rng(self) -> None:
        """torch's CUDA generator <- device cursor (after drain: later eager calls continue the stream)."""
        self.sim.store.rng_cursor_commit()

    def join(self) -> None:
        """Orders the compute stream behind every submitted batch without blocking the host (for event timing)."""
        self.compute.wait_stream(self.s_out)
        self.compute.wait_stream(self.s_in)


__all__ = ["HostPipeline"]
