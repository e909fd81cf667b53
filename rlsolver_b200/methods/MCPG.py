"""Drop-in for the sampling half of `rlsolver/methods/MCPG.py` (single-file MCPG for max-cut):
`metro_sampling` (88-117), `sampler_func` (120-166) and the parts of `maxcut_dataloader` (187-232)
those two read.  Layout as in the reference: node-major float32 `[N, C]`, C = total_mcmc_num *
repeat_times chains.

The reference runs `sampler_func` as num_ls * N Python iterations (neighbour gather + sum +
torch.rand(C) + compare per node) and `metro_sampling` as up to 5 * max_transfer_time iterations
with a host sync each.  Here each is one or two kernel launches (csrc/samplers.cu).  The random
numbers are NOT drawn by separate torch calls: the kernels regenerate torch's own Philox stream
from the CUDA generator's (seed, offset) and the generator is advanced by what the reference's
call sequence would have consumed, so a given seed yields the reference's samples bit for bit.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch as th

from .. import _lib, rng
from ..graph_store import GraphStore, _ptr, _stream_ptr, on_device, require_cuda
from .config import MyGraph

TEN = th.Tensor

# scratch of the split metro_sampling (8 bytes per chain and iteration); beyond this the two-pass kernel is used
_METRO_SPLIT_MAX_BYTES = 8 << 30


class McpgData:
    """What maxcut_dataloader returns, reduced to the fields the samplers read
    (MCPG.py:187-232): num_nodes, num_edges, edge_index [2, M], single_degree, weighted_degree,
    sorted_degree_nodes (descending degree, torch CPU argsort as in the reference)."""

    def __init__(self, mygraph: MyGraph, num_nodes: int, device):
        self.device = require_cuda(device)
        arr = np.asarray([(a, b) for a, b, _ in mygraph], dtype=np.int64).reshape(-1, 2)
        self.num_nodes = int(num_nodes)
        self.edge_index = th.from_numpy(arr.T.copy()).to(self.device)
        self.num_edges = int(arr.shape[0])
        deg = np.bincount(arr.reshape(-1), minlength=self.num_nodes)
        self.single_degree = deg.tolist()
        self.weighted_degree = [float(d) for d in deg]          # append_neighbors hard-codes weight 1 (MCPG.py:245)
        self.sorted_degree_nodes = th.argsort(th.tensor(self.weighted_degree), descending=True)
        self.store = GraphStore(mygraph, True, device=self.device, num_nodes=self.num_nodes)
        order = np.ascontiguousarray(self.sorted_degree_nodes.numpy(), dtype=np.int32)
        handle = C.c_void_p()
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_mcpg_plan_create(self.store.handle, order.ctypes.data, C.byref(handle)),
                       "mcpg_plan_create")
        self._plan = handle
        self.num_levels = int(_lib.lib().rlsb_mcpg_plan_num_levels(handle))

    def __del__(self):
        h, self._plan = getattr(self, "_plan", None), None
        if h:
            try:
                _lib.lib().rlsb_mcpg_plan_destroy(h)
            except Exception:
                pass


def maxcut_dataloader(path: str, device=None):
    """`N M` header + 1-based `u v [w]` rows (weights ignored, MCPG.py:245) -> (data, num_nodes)."""
    with open(path) as fh:
        first = fh.readline().split()
        num_nodes, num_edges = int(first[0]), int(first[1])
        edges = []
        for line in fh:
            tok = line.split()
            if len(tok) >= 2:
                edges.append((int(tok[0]) - 1, int(tok[1]) - 1, 1))
    assert len(edges) == num_edges, f"{path}: header says {num_edges} edges, file has {len(edges)}"
    device = th.device("cuda", th.cuda.current_device()) if device is None else device
    return McpgData(edges, num_nodes, device), num_nodes


def _as_chains(t: TEN, device) -> TEN:
    return t.to(device=device, dtype=th.float32).contiguous()


def metro_sampling(probs: TEN, start_status: TEN, max_transfer_time: int, device=None,
                   _explicit=None) -> TEN:
    """MCPG.py:88-117.  Returns float32 [N, C].  `_explicit = (index_rows int64 [T, C], rands float32
    [T, C])` replays recorded draws instead of the generator (tests)."""
    device = require_cuda(start_status.device if device is None else device)
    lib = _lib.lib()
    num_node, num_chain = len(probs), start_status.shape[1]
    if start_status.dtype == th.float32 and start_status.device == device and start_status.is_contiguous():
        start = start_status            # the kernels read `value != 0` (== start_status.bool()): no conversion pass
    else:
        start = _as_chains(start_status.bool(), device)
    p = probs.detach().to(device=device, dtype=th.float32).contiguous()
    tmax = int(max_transfer_time) * 5
    out = th.empty_like(start)
    if tmax <= 0 or num_chain == 0:
        return _as_chains(start_status.bool(), device)
    seed, offset, threads, iters = rng.peek(device, num_chain)
    idx_ptr = u_ptr = C.c_void_p(0)
    if _explicit is not None:
        e_idx, e_u = (t.to(device).contiguous() for t in _explicit)
        assert e_idx.dtype == th.int64 and e_u.dtype == th.float32 and e_idx.shape[0] >= 1
        tmax = min(tmax, e_idx.shape[0])
        idx_ptr, u_ptr = _ptr(e_idx), _ptr(e_u)
    st = _stream_ptr(device)
    thresh = num_chain * int(max_transfer_time)
    if _explicit is None:
        need = int(lib.rlsb_metro_workspace_bytes(num_node, num_chain, tmax))
        if 0 <= need <= _METRO_SPLIT_MAX_BYTES:
            # split form: draws in parallel, one pass of the chain, stop rule on the device, surplus moves undone
            ws = th.empty((need,), dtype=th.uint8, device=device)
            num_iters = th.empty((1,), dtype=th.int32, device=device)
            with on_device(device):
                _lib.check(lib.rlsb_metro_sampling_split(num_node, _ptr(p), _ptr(start), _ptr(out), num_chain, tmax, thresh,
                                                         seed, offset, threads, iters, _ptr(num_iters), _ptr(ws), st),
                           "metro_sampling_split")
            rng.advance(device, num_chain, 2 * int(num_iters.item()))   # one randint + one rand per executed iteration
            return out
    acc = th.zeros((tmax,), dtype=th.int32, device=device)
    with on_device(device):
        _lib.check(lib.rlsb_metro_sampling(num_node, _ptr(p), _ptr(start), None, num_chain, tmax, None, idx_ptr, u_ptr,
                                           seed, offset, threads, iters, _ptr(acc), 1, st), "metro_sampling(count)")
    # the reference checks `count >= num_chain * max_transfer_time` BEFORE every iteration (MCPG.py:101-103)
    before = th.cat([acc.new_zeros(1), acc.cumsum(0)[:-1]])
    num_iters = (before < thresh).sum().to(th.int32).reshape(1)
    with on_device(device):
        _lib.check(lib.rlsb_metro_sampling(num_node, _ptr(p), _ptr(start), _ptr(out), num_chain, tmax, _ptr(num_iters),
                                           idx_ptr, u_ptr, seed, offset, threads, iters, None, 0, st),
                   "metro_sampling(apply)")
    if _explicit is None:
        rng.advance(device, num_chain, 2 * int(num_iters.item()))      # one randint + one rand per executed iteration
    return out


def sampler_func(data: McpgData, xs_sample: TEN, num_ls: int, total_mcmc_num: int, repeat_times: int,
                 device=None, _explicit_u: Optional[TEN] = None):
    """MCPG.py:120-166.  Returns (vs_good [T], xs_good [N, T], value [C])."""
    device = data.device
    lib = _lib.lib()
    num_edges = data.num_edges
    num_chain = total_mcmc_num * repeat_times
    xs_loc = _as_chains(xs_sample, device).clone()
    assert xs_loc.shape == (data.num_nodes, num_chain)
    if num_ls <= 0:
        # no sweep: the state keeps the {-0.5, 1.5} encoding (MCPG.py:132-133); pure layout, torch ops
        xs_loc = xs_loc * 2 - 0.5
        n0, n1 = data.edge_index[0], data.edge_index[1]
        expected = ((2 * xs_loc[n0] - 1) * (2 * xs_loc[n1] - 1)).sum(dim=0)
    else:
        expected = th.empty((num_chain,), dtype=th.float32, device=device)
        seed, offset, threads, iters = rng.peek(device, num_chain)
        u = None
        if _explicit_u is not None:
            u = _explicit_u.to(device=device, dtype=th.float32).contiguous()
            assert u.shape == (num_ls * data.num_nodes, num_chain)
        with on_device(device):
            _lib.check(lib.rlsb_mcpg_sweeps(data.store.handle, data._plan, _ptr(xs_loc), num_chain, int(num_ls), _ptr(u),
                                            seed, offset, threads, iters, _ptr(expected), _stream_ptr(device)),
                       "mcpg_sweeps")
        if u is None:
            rng.advance(device, num_chain, int(num_ls) * data.num_nodes)   # one torch.rand(C) per node visit
    expected_reshape = expected.reshape((-1, total_mcmc_num))
    index = th.argmin(expected_reshape, dim=0)
    index = th.arange(total_mcmc_num, device=device) + index * total_mcmc_num
    max_cut = expected[index]
    vs_good = (num_edges - max_cut) / 2
    xs_good = xs_loc[:, index]
    value = expected.float()
    value = value - value.mean()
    return vs_good, xs_good, value


# ----------------------------------------------------------------------------- weighted sampler (MCPG/sampling.py)
class WeightedMcpgData:
    """What `maxcut_dataloader` of rlsolver/methods/MCPG/dataloader.py:53-103 returns, reduced to the fields
    `mcpg_sampling_maxcut` reads: num_nodes, edge_index [2, M], edge_attr [M, 1] float32, edge_weight_sum,
    neighbors / neighbor_edges (as CSR, edge order, both directions), weighted_degree (float(sum of the float32
    row)), sorted_degree_nodes (descending |weighted| degree, torch CPU argsort as in the reference)."""

    def __init__(self, edges, weights, num_nodes: int, device):
        self.device = require_cuda(device)
        e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        w = np.asarray(weights, dtype=np.float32).reshape(-1)
        self.num_nodes, self.num_edges = int(num_nodes), int(e.shape[0])
        self.edge_index = th.from_numpy(e.T.copy()).to(self.device)
        self.edge_attr = th.from_numpy(w.copy()).reshape(-1, 1).to(self.device)
        self.edge_weight_sum = float(th.sum(th.from_numpy(w.copy())))
        src = np.concatenate([e[:, 0], e[:, 1]])
        dst = np.concatenate([e[:, 1], e[:, 0]])
        pos = np.concatenate([2 * np.arange(e.shape[0]), 2 * np.arange(e.shape[0]) + 1])     # edge order, (row, col) first
        order = np.lexsort((pos, src))
        ptr = np.zeros(self.num_nodes + 1, np.int64)
        np.add.at(ptr, src + 1, 1)
        ptr = np.cumsum(ptr)
        col, nw = dst[order], np.concatenate([w, w])[order]
        self.single_degree = np.diff(ptr).tolist()
        rows = [th.from_numpy(nw[ptr[i]:ptr[i + 1]].copy()) for i in range(self.num_nodes)]
        self.weighted_degree = [float(th.sum(r)) for r in rows]
        abs_deg = th.tensor([float(th.sum(th.abs(r))) for r in rows])
        self.sorted_degree_nodes = th.argsort(abs_deg, descending=True)
        thr = np.asarray([np.float32(d / 2 + 0.125) for d in self.weighted_degree], dtype=np.float32)
        to = lambda a, dt: th.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)   # noqa: E731
        self._order = to(self.sorted_degree_nodes.numpy(), np.int32)
        self._ptr, self._col, self._w, self._thr = to(ptr, np.int32), to(col, np.int32), to(nw, np.float32), to(thr, np.float32)
        self._eu, self._ev, self._ew = to(e[:, 0], np.int32), to(e[:, 1], np.int32), to(w, np.float32)


def weighted_maxcut_dataloader(path: str, device=None):
    """`N M` header + 1-based `u v w` rows with float weights -> (data, num_nodes) (MCPG/dataloader.py:53-103)."""
    with open(path) as fh:
        first = fh.readline().split()
        num_nodes, num_edges = int(first[0]), int(first[1])
        rows = [ln.split() for ln in fh if ln.strip()]
    assert len(rows) == num_edges, f"{path}: header says {num_edges} edges, file has {len(rows)}"
    device = th.device("cuda", th.cuda.current_device()) if device is None else device
    edges = [(int(r[0]) - 1, int(r[1]) - 1) for r in rows]
    return WeightedMcpgData(edges, [float(r[2]) for r in rows], num_nodes, device), num_nodes


def mcpg_sampling_maxcut(data: WeightedMcpgData, start_result: TEN, probs: TEN, num_ls: int, change_times: int,
                         total_mcmc_num: int, device=None, _explicit=None, _explicit_u: Optional[TEN] = None):
    """rlsolver/methods/MCPG/sampling.py:89-127: metro_sampling, then the weighted local-search sweeps and the
    expected cut in one kernel (csrc/mcpg_weighted.cu).  Returns ((edge_weight_sum - cut) / 2 [T], xs [N, T],
    the metro output [N, C], advantage [C])."""
    device = data.device
    lib = _lib.lib()
    graph_probs = metro_sampling(probs, start_result.clone(), change_times, device, _explicit=_explicit)
    start = graph_probs.clone()
    num_chain = graph_probs.shape[1]
    sweeps = max(int(num_ls), 1)                      # the reference's `while True` runs once for num_ls <= 1
    expected = th.empty((num_chain,), dtype=th.float32, device=device)
    seed, offset, threads, iters = rng.peek(device, num_chain)
    u = None
    if _explicit_u is not None:
        u = _explicit_u.to(device=device, dtype=th.float32).contiguous()
        assert u.shape == (sweeps * data.num_nodes, num_chain)
    with on_device(device):
        _lib.check(lib.rlsb_mcpg_weighted_sweeps(data.num_nodes, num_chain, _ptr(data._order), _ptr(data._ptr),
                                                 _ptr(data._col), _ptr(data._w), _ptr(data._thr), data.num_edges,
                                                 _ptr(data._eu), _ptr(data._ev), _ptr(data._ew), _ptr(graph_probs),
                                                 sweeps, _ptr(u), seed, offset, threads, iters, _ptr(expected),
                                                 _stream_ptr(device)), "mcpg_weighted_sweeps")
    if u is None:
        rng.advance(device, num_chain, sweeps * data.num_nodes)      # one torch.rand(C) per node visit
    index = th.argmin(expected.reshape((-1, total_mcmc_num)), dim=0)
    index = th.arange(total_mcmc_num, device=device) + index * total_mcmc_num
    max_cut = expected[index]
    return (data.edge_weight_sum - max_cut) / 2, graph_probs[:, index], start, expected - th.mean(expected)
