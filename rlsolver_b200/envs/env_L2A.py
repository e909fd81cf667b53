"""Drop-in for `rlsolver.envs.env_L2A.EnvMaxcut` (rlsolver/envs/env_L2A.py:24-116; the same
class is `rlsolver.envs.env_MCPG.EnvMaxcut`, env_MCPG.py:24-116).

Same constructor, attributes, method names, argument meaning, in-place mutation and RNG call
sequence as the reference; the tensor work runs in the sm_100a kernels behind the C ABI
(include/rlsolver_b200.h).  CUDA only -- constructing it on a CPU device raises.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch as th

from ..graph_store import GraphStore, require_cuda
from ..methods.config import MyGraph
from ..methods.util import build_adjacency_bool
from ..methods.util_read_data import load_mygraph2, update_xs_by_vs

TEN = th.Tensor

# noise tensors handed to one fused ls_noisy_iters launch are capped to this many bytes so a
# huge E*N batch does not hold every draw of a call alive at once
_NOISE_BYTES_PER_LAUNCH = 4 << 30


class EnvMaxcut:
    # True: the noisy iterations consume torch's CUDA generator in place (csrc/noise_masks.cu: the numbers
    # the reference's randn calls would return are recomputed in registers, only flip bits leave the
    # kernel; the generator ends where the reference's calls would leave it).  False: every draw is a
    # torch.randn tensor streamed through rlsb_ls_run -- the form to use when the noise itself must be
    # supplied (tests replaying recorded draws) and the one taken when the mask path does not apply.
    fused_rng = True
    # True: under CUDA-graph capture the threshold draw is issued on a second stream next to ls_begin (it does not
    # depend on the state).  Same results, but measured on B200 it buys nothing (235.6 us in line vs 237.8 / 239.6 us per G22 x
    # 4096 step, profiles/r02_threshold_draw_second_stream.log): the draw's 8288 short blocks are dispatched first and
    # fill every SM, the begin kernel's 128 CTAs of 512 threads x 104 registers only find room when the draw is
    # nearly over -- a launch priority on the begin kernel does not change that -- so the default keeps one stream.
    overlap_threshold_draw = False

    def __init__(self, sim_name: str = 'max_cut', mygraph: MyGraph = (),
                 device=th.device('cpu'), if_bidirectional: bool = False):
        self.device = require_cuda(device)
        self.sim_name = sim_name
        self.int_type = th.long
        self.if_maximize = True
        self.if_bidirectional = if_bidirectional

        mygraph = mygraph if mygraph else load_mygraph2(graph_name=sim_name)
        self._mygraph = mygraph
        self.store = GraphStore(mygraph, if_bidirectional, device=self.device)
        self.num_nodes = self.store.num_nodes
        self.num_edges = self.store.num_edges

        arrs = self.store.export()
        ptr, col = arrs["listed_ptr"].astype(np.int64), arrs["listed_col"].astype(np.int64)
        self._listed_ptr, self._listed_col = ptr, col
        n0 = np.repeat(np.arange(self.num_nodes, dtype=np.int64), np.diff(ptr))
        self.n0_ids = th.from_numpy(n0).to(self.device)[None, :]
        self.n1_ids = th.from_numpy(col.copy()).to(self.device)[None, :]
        self.sim_ids = th.zeros((1, self.store.num_listed), dtype=th.long, device=self.device)
        self.n0_num_n1 = th.from_numpy(np.diff(ptr)).to(self.device)[None, :]
        self._adjacency_bool: Optional[TEN] = None
        self._adjacency_indies: Optional[List[TEN]] = None

    # the two O(N^2) / O(N)-tensor attributes are only read by a few callers: built on first use
    @property
    def adjacency_bool(self) -> TEN:
        if self._adjacency_bool is None:
            self._adjacency_bool = build_adjacency_bool(self._mygraph, if_bidirectional=True).to(self.device)
        return self._adjacency_bool

    @property
    def adjacency_indies(self) -> List[TEN]:
        if self._adjacency_indies is None:
            flat = th.from_numpy(self._listed_col.copy()).to(self.device)
            self._adjacency_indies = list(th.split(flat, np.diff(self._listed_ptr).tolist()))
        return self._adjacency_indies

    # ------------------------------------------------------------------ objective
    def calculate_obj_values(self, xs: TEN, if_sum: bool = True) -> TEN:
        if if_sum:
            return self.store.cut_eval(xs)
        values = self.store.cut_edges(xs)
        if self.if_bidirectional:     # reference: bool // 2 -> int64 zeros
            values = values.to(th.long) // 2
        return values

    def calculate_obj_values_for_loop(self, xs: TEN, if_sum: bool = True) -> TEN:
        num_sims = xs.shape[0]
        packed = self.store.pack(xs)
        cross, _, _ = self.store.cross_counts(packed, num_sims, want_minmax=False)
        values = cross[:, :self.num_nodes].to(th.long) & 0xFFFF       # uint16 payload
        if if_sum:
            values = values.sum(dim=1)
        if self.if_bidirectional:
            values = values.float() / 2
        return values

    def generate_xs_randomly(self, num_sims):
        xs = th.randint(0, 2, size=(num_sims, self.num_nodes), dtype=th.bool, device=self.device)
        xs[:, 0] = 0
        return xs

    # ------------------------------------------------------------------ local search
    def local_search_inplace(self, good_xs: TEN, good_vs: TEN,
                             num_iters: int = 8, num_spin: int = 8, noise_std: float = 0.3):
        """env_L2A.py:87-116.  RNG: 1 + num_iters draws of randn [E, N] float32 on self.device,
        in the reference's order, so the flip sequence matches the reference's for a given seed."""
        st = self.store
        num_sims = good_xs.shape[0]
        if not good_xs.is_contiguous():
            raise RuntimeError("local_search_inplace mutates good_xs in place: it must be contiguous")
        if good_vs.shape == ():
            vs_in = None
        else:
            vs_in = good_vs.long()
            if not vs_in.is_contiguous():
                vs_in = vs_in.contiguous()
        ws = st.ls_workspace(num_sims)
        fused = self.fused_rng and num_sims > 0 and st.ls_mask_words(num_sims) >= 0
        if fused and self.overlap_threshold_draw:
            st.ls_prefetch_threshold_draw(num_sims)      # graph capture only: the draw runs next to ls_begin
        good_vs = st.ls_begin(good_xs, vs_in, 1, noise_std, ws)
        if fused:
            st.ls_fused(good_vs, 1, num_spin, num_iters, False, good_xs, ws)
            return good_xs, good_vs
        shape = (num_sims, self.num_nodes)
        noise0 = th.randn(shape, dtype=th.float32, device=self.device)
        per_launch = max(1, min(16, _NOISE_BYTES_PER_LAUNCH // max(1, 4 * num_sims * self.num_nodes)))
        done = 0
        while done < num_iters or noise0 is not None:
            now = min(per_launch, num_iters - done)
            noises = [th.randn(shape, dtype=th.float32, device=self.device) for _ in range(now)]
            done += now
            # the first launch also derives the threshold from noise0 (env_L2A.py:94-96)
            st.ls_run(good_vs, 1, noise0, num_spin, noises, done == num_iters, good_xs, ws)
            noise0 = None
        return good_xs, good_vs


    def local_search_packed(self, packed: TEN, num_iters: int = 8, num_spin: int = 8, noise_std: float = 0.3,
                            num_sims: Optional[int] = None, good_vs: Optional[TEN] = None):
        """`local_search_inplace` for a batch kept as packed tiles: `packed` int32 [ceil(E/32), Np] with bit b of
        word [t][i] = node i of env 32t + b (`store.pack(xs)` / `rlsb_pack_spins` produce it, `store.unpack`
        reads it).  Same algorithm, same RNG use; nothing is expanded to one byte per spin on the way in or out,
        so a host that ships spins across PCIe moves N/8 bytes per env.  Returns `(packed_out, vs)`; `packed_out`
        is a view of the simulator's workspace, valid until its next local-search call."""
        st = self.store
        num_sims = packed.shape[0] * 32 if num_sims is None else int(num_sims)
        if not (self.fused_rng and num_sims > 0 and st.ls_mask_words(num_sims) >= 0):
            raise NotImplementedError("local_search_packed needs the fused generator path (degrees <= 255, "
                                      "fewer than 2^31 elements per draw)")
        ws = st.ls_workspace(num_sims)
        vs = st.ls_begin_packed(packed, num_sims, None if good_vs is None else good_vs.long().contiguous(), 1,
                                noise_std, ws)
        st.ls_fused(vs, 1, num_spin, num_iters, False, None, ws)
        return st.ls_section(ws, num_sims, "packed").view(st.tiles(num_sims), st.padded_nodes), vs


def metropolis_hastings_sampling_TNCO(probs: TEN, start_xs: TEN, num_repeats: int, num_iters: int = -1,
                                      accept_rate: float = 0.25) -> TEN:
    """env_L2A.py:233-276: row-major independent-site Metropolis toward product Bernoulli(probs) -- `num_repeats`
    copies of every start row, up to 4 rounds over the columns in a random order, stopping after the column at which
    the number of accepted flips reaches rows * num_iters (default: dim * accept_rate per row).
    RNG == the reference's: one `th.randperm(dim)` per round (torch's own call) and one `th.rand(rows)` per visited
    column, recomputed in the kernels from the generator state, which is advanced past exactly the calls the reference
    would have made -- same seed, same samples, same generator state afterwards.  One host sync per round (the
    reference has one per column)."""
    from .. import _lib, rng
    from ..graph_store import on_device
    dev = require_cuda(start_xs.device)
    if probs.shape != start_xs.shape or start_xs.dim() != 2 or start_xs.dtype != th.bool:
        raise TypeError("probs float32 [S, dim] and start_xs bool [S, dim] are required")
    xs = start_xs.repeat(num_repeats, 1).contiguous()
    ps = probs.to(device=dev, dtype=th.float32).contiguous()
    num, dim = xs.shape
    num_sims = start_xs.shape[0]
    num_iters = int(dim * accept_rate) if num_iters == -1 else num_iters
    target = num * num_iters
    if num == 0:
        return xs
    state = th.zeros((2,), dtype=th.int64, device=dev)
    counts = th.empty((dim,), dtype=th.int32, device=dev)
    lib = _lib.lib()
    for _ in range(4):
        ids = th.randperm(dim, device=dev)
        seed, offset, threads, iters = rng.peek(dev, num)
        with on_device(dev):
            _lib.check(lib.rlsb_mh_rows_round(ps.data_ptr(), xs.data_ptr(), ids.data_ptr(), num, num_sims, dim, target, seed,
                                              offset, threads, iters, state.data_ptr(), counts.data_ptr(),
                                              th.cuda.current_stream(dev).cuda_stream), "mh_rows_round")
        count, visited = (int(t) for t in state.tolist())
        rng.advance(dev, num, visited)
        if count >= target:
            break
    return xs


__all__ = ["EnvMaxcut", "update_xs_by_vs", "metropolis_hastings_sampling_TNCO"]
