"""Python handle of the native graph store (csrc/graph_store.cu) plus thin op wrappers.

Everything here is plumbing: torch owns device memory and streams, the C-ABI library does
the work.  No op has a CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch as th

from . import _lib, rng

TEN = th.Tensor


def _stream_ptr(device: th.device) -> C.c_void_p:
    return C.c_void_p(th.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[TEN]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def require_cuda(device) -> th.device:
    device = th.device(device)
    if device.type != "cuda":
        raise RuntimeError(
            f"rlsolver_b200 runs on CUDA (sm_100a) only; got device '{device}'. There is no CPU fallback.")
    if device.index is None:
        device = th.device("cuda", th.cuda.current_device())
    return device


class OpTimer:
    """Optional per-op CUDA-event timing (bench.py's roofline pass).  Events are recorded on
    the stream the kernels are launched on (torch's current stream)."""

    def __init__(self):
        self.spans = {}
        self.launches = 0

    def begin(self, name):
        ev = th.cuda.Event(enable_timing=True)
        ev.record()
        return name, ev

    def end(self, tok, launches=1):
        name, start = tok
        stop = th.cuda.Event(enable_timing=True)
        stop.record()
        self.spans.setdefault(name, []).append((start, stop))
        self.launches += launches

    def summary(self):
        th.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.spans.items()}


class on_device:
    """Makes `device` the current CUDA device for the duration of a library call (and restores the previous
    one).  The C ABI launches on the stream it is handed and calls cudaFuncSetAttribute / cudaMalloc on the
    CURRENT device, so a simulator built for cuda:1 must not launch while cuda:0 is current."""
    __slots__ = ("index", "prev")

    def __init__(self, device: Optional[th.device]):
        self.index = None if device is None else device.index

    def __enter__(self):
        self.prev = None
        if self.index is not None:
            cur = th.cuda.current_device()
            if cur != self.index:
                self.prev = cur
                th.cuda.set_device(self.index)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            th.cuda.set_device(self.prev)
        return False


class _Span(on_device):
    """One library call of a GraphStore: device guard + optional per-op timing."""
    __slots__ = ("timer", "name", "launches", "tok")

    def __init__(self, timer, name, launches, device=None):
        on_device.__init__(self, device)
        self.timer, self.name, self.launches = timer, name, launches

    def __enter__(self):
        on_device.__enter__(self)
        if self.timer is not None:
            self.tok = self.timer.begin(self.name)

    def __exit__(self, *exc):
        if self.timer is not None:
            self.timer.end(self.tok, self.launches)
        return on_device.__exit__(self, *exc)


class GraphStore:
    """CSR + edge list + sweep levels, built natively from a reference-style MyGraph
    (list of (n0, n1, weight), 0-based).  `device=None` keeps it host-only (no CUDA needed)."""

    def __init__(self, mygraph: Sequence[Tuple[int, int, int]], if_bidirectional: bool = False,
                 device: Optional[th.device] = None, num_nodes: int = 0):
        self._lib = _lib.lib()
        arr = np.asarray(list(mygraph), dtype=np.int64).reshape(-1, 3) if len(mygraph) else np.zeros((0, 3), np.int64)
        if arr.size and (arr[:, :2].max() >= 2 ** 31 or arr[:, :2].min() < 0):
            raise IndexError("node ids must be in [0, 2^31)")
        n0 = np.ascontiguousarray(arr[:, 0], dtype=np.int32)
        n1 = np.ascontiguousarray(arr[:, 1], dtype=np.int32)
        w = np.ascontiguousarray(arr[:, 2], dtype=np.int32)
        self.device = None if device is None else require_cuda(device)
        handle = C.c_void_p()
        st = self._lib.rlsb_graph_create(int(num_nodes), int(arr.shape[0]), n0.ctypes.data, n1.ctypes.data,
                                         w.ctypes.data, int(bool(if_bidirectional)),
                                         -1 if self.device is None else self.device.index, C.byref(handle))
        if st == 1 and "IndexError" in self._lib.rlsb_last_error().decode():
            raise IndexError(self._lib.rlsb_last_error().decode())
        _lib.check(st, "graph_create")
        self._h = handle
        L = self._lib
        self.num_nodes = int(L.rlsb_graph_num_nodes(handle))
        self.padded_nodes = int(L.rlsb_graph_padded_nodes(handle))
        self.num_edges = int(L.rlsb_graph_num_edges(handle))
        self.num_listed = int(L.rlsb_graph_num_listed(handle))
        self.num_full = int(L.rlsb_graph_num_full(handle))
        self.num_levels = int(L.rlsb_graph_num_levels(handle))
        self.max_listed_degree = int(L.rlsb_graph_max_listed_degree(handle))
        self.max_full_degree = int(L.rlsb_graph_max_full_degree(handle))
        self.if_bidirectional = bool(if_bidirectional)
        # weights other than 1: the weighted entry points apply (the unweighted ones keep ignoring weights,
        # like the reference's EnvMaxcut)
        self.weighted = bool(arr.shape[0]) and bool((arr[:, 2] != 1).any())
        loops = arr[:, 0] == arr[:, 1]
        self.weight_sum = int(arr[~loops, 2].sum()) if arr.shape[0] else 0
        self.timer: Optional[OpTimer] = None      # set by bench.py for the per-kernel pass
        self.launch_count = 0                     # kernels of this library launched through this store

    def _op(self, name: str, launches: int = 1) -> _Span:
        self.launch_count += launches
        return _Span(self.timer, name, launches, self.device)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.rlsb_graph_destroy(h)
            except Exception:
                pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def export(self) -> Dict[str, np.ndarray]:
        """Host copies of the built arrays (listed CSR == the reference's n0_ids/n1_ids order)."""
        n = self.num_nodes
        out = {
            "listed_ptr": np.zeros(n + 1, np.int32), "listed_col": np.zeros(self.num_listed, np.int32),
            "full_ptr": np.zeros(n + 1, np.int32), "full_col": np.zeros(self.num_full, np.int32),
            "level_ptr": np.zeros(self.num_levels + 1, np.int32), "level_nodes": np.zeros(n, np.int32),
        }
        _lib.check(self._lib.rlsb_graph_export(self._h, *[out[k].ctypes.data for k in (
            "listed_ptr", "listed_col", "full_ptr", "full_col", "level_ptr", "level_nodes")]), "graph_export")
        return out

    def listed_degree_numpy(self) -> np.ndarray:
        """Listed degree per node (== n0_num_n1 of the reference), int64 [N]."""
        return np.diff(self.export()["listed_ptr"].astype(np.int64))

    def export_sell(self, which: int) -> Dict[str, np.ndarray]:
        """Host copy of a SELL-32 structure (0 = listed neighbours, 1 = sweep order)."""
        sizes = np.zeros(3, np.int64)
        _lib.check(self._lib.rlsb_graph_sell_sizes(self._h, which, sizes.ctypes.data), "graph_sell_sizes")
        out = {"off": np.zeros(sizes[0] + 1, np.int32), "node": np.zeros(sizes[0] * 32, np.uint16),
               "half": np.zeros(sizes[0] * 32, np.uint16), "col": np.zeros(sizes[1], np.uint16),
               "level_slice": np.zeros(sizes[2], np.int32)}
        _lib.check(self._lib.rlsb_graph_sell_export(self._h, which, *[out[k].ctypes.data for k in (
            "off", "node", "half", "col", "level_slice")]), "graph_sell_export")
        return out

    # ------------------------------------------------------------------ layout helpers
    def tiles(self, num_envs: int) -> int:
        return (num_envs + 31) // 32

    def new_packed(self, num_envs: int) -> TEN:
        return th.empty((self.tiles(num_envs), self.padded_nodes), dtype=th.int32, device=self.device)

    def _check_xs(self, xs: TEN) -> TEN:
        if xs.dtype != th.bool:
            raise TypeError(f"xs must be torch.bool, got {xs.dtype}")
        if xs.dim() != 2 or xs.shape[1] != self.num_nodes:
            raise IndexError(f"xs must be [num_envs, {self.num_nodes}], got {tuple(xs.shape)}")
        if xs.device != self.device:
            raise RuntimeError(f"xs is on {xs.device}, the simulator is on {self.device}")
        return xs if xs.is_contiguous() else xs.contiguous()

    # ------------------------------------------------------------------ ops (device tensors in/out)
    def pack(self, xs: TEN, out: Optional[TEN] = None) -> TEN:
        xs = self._check_xs(xs)
        e = xs.shape[0]
        out = self.new_packed(e) if out is None else out
        with self._op("pack_spins"):
            _lib.check(self._lib.rlsb_pack_spins(_ptr(xs), e, self.num_nodes, self.padded_nodes, _ptr(out),
                                                 _stream_ptr(self.device)), "pack_spins")
        return out

    def unpack(self, packed: TEN, num_envs: int, out: Optional[TEN] = None) -> TEN:
        if out is None:
            out = th.empty((num_envs, self.num_nodes), dtype=th.bool, device=self.device)
        elif not out.is_contiguous():
            raise RuntimeError("unpack target must be contiguous")
        with self._op("unpack_spins"):
            _lib.check(self._lib.rlsb_unpack_spins(_ptr(packed), num_envs, self.num_nodes, self.padded_nodes,
                                                   _ptr(out), _stream_ptr(self.device)), "unpack_spins")
        return out

    def cut_eval(self, xs: TEN) -> TEN:
        xs = self._check_xs(xs)
        vs = th.empty((xs.shape[0],), dtype=th.int64, device=self.device)
        with self._op("cut_eval"):
            _lib.check(self._lib.rlsb_cut_eval(self._h, _ptr(xs), xs.shape[0], _ptr(vs), _stream_ptr(self.device)),
                       "cut_eval")
        return vs

    def cut_eval_packed(self, packed: TEN, num_envs: int, out: Optional[TEN] = None) -> TEN:
        vs = th.empty((num_envs,), dtype=th.int64, device=self.device) if out is None else out
        with self._op("cut_eval_packed"):
            _lib.check(self._lib.rlsb_cut_eval_packed(self._h, _ptr(packed), num_envs, _ptr(vs),
                                                      _stream_ptr(self.device)), "cut_eval_packed")
        return vs

    def cut_eval_weighted(self, xs: Optional[TEN] = None, packed: Optional[TEN] = None,
                          num_envs: Optional[int] = None) -> TEN:
        """int64 [E]: sum of w over the cut edges, from bool rows `xs` or from packed tiles."""
        if (xs is None) == (packed is None):
            raise ValueError("cut_eval_weighted takes bool rows OR packed tiles")
        if xs is not None:
            xs = self._check_xs(xs)
            num_envs = xs.shape[0]
        vs = th.empty((num_envs,), dtype=th.int64, device=self.device)
        with self._op("cut_eval_weighted"):
            _lib.check(self._lib.rlsb_cut_eval_weighted(self._h, _ptr(xs), _ptr(packed), num_envs, _ptr(vs),
                                                        _stream_ptr(self.device)), "cut_eval_weighted")
        return vs

    def node_fields_weighted(self, packed: TEN, num_envs: int) -> TEN:
        """int32 [E, Np]: per node the weight of its incident edges that are cut."""
        out = th.empty((num_envs, self.padded_nodes), dtype=th.int32, device=self.device)
        with self._op("node_fields_weighted"):
            _lib.check(self._lib.rlsb_node_fields_weighted(self._h, _ptr(packed), num_envs, _ptr(out),
                                                           _stream_ptr(self.device)), "node_fields_weighted")
        return out

    def cut_edges(self, xs: TEN) -> TEN:
        xs = self._check_xs(xs)
        out = th.empty((xs.shape[0], self.num_listed), dtype=th.bool, device=self.device)
        for lo in range(0, xs.shape[0], 32768):
            part = xs[lo:lo + 32768]
            with self._op("cut_edges"):
                _lib.check(self._lib.rlsb_cut_edges(self._h, _ptr(part), part.shape[0], _ptr(out[lo:lo + 32768]),
                                                    _stream_ptr(self.device)), "cut_edges")
        return out

    def cross_counts(self, packed: TEN, num_envs: int, want_minmax: bool = True):
        """uint16 counts as an int16 tensor [E, Np] (+ int32 [N] min / max over envs)."""
        cross = th.empty((num_envs, self.padded_nodes), dtype=th.int16, device=self.device)
        cmin = cmax = None
        if want_minmax:
            cmin = th.empty((self.num_nodes,), dtype=th.int32, device=self.device)
            cmax = th.empty((self.num_nodes,), dtype=th.int32, device=self.device)
        with self._op("node_cross_counts", 3 if want_minmax else 1):
            _lib.check(self._lib.rlsb_node_cross_counts(self._h, _ptr(packed), num_envs, _ptr(cross), _ptr(cmin),
                                                        _ptr(cmax), _stream_ptr(self.device)), "node_cross_counts")
        return cross, cmin, cmax

    # ---- local search (env_L2A.py:87-116 / LocalSearch.py:53-86): begin -> thresh -> search
    def ls_workspace(self, num_envs: int) -> TEN:
        need = int(self._lib.rlsb_ls_workspace_bytes(self._h, num_envs))
        if need < 0:
            _lib.check(3, "ls_workspace_bytes")
        ws = getattr(self, "_ls_ws", None)
        if ws is None or ws.numel() < need:
            ws = self._ls_ws = th.empty((need,), dtype=th.uint8, device=self.device)
        return ws

    def ls_section(self, workspace: TEN, num_envs: int, section: str) -> TEN:
        """View of one workspace section (tests): 'packed', 'col_min', 'col_max', 'rd_std', 'thresh'."""
        idx, dtype, count = {"packed": (0, th.int32, self.tiles(num_envs) * self.padded_nodes),
                             "col_min": (2, th.int32, self.padded_nodes), "col_max": (3, th.int32, self.padded_nodes),
                             "rd_std": (5, th.float32, self.padded_nodes), "thresh": (6, th.float32, num_envs)}[section]
        off = int(self._lib.rlsb_ls_workspace_offset(self._h, num_envs, idx))
        return workspace[off:off + 4 * count].view(dtype)

    def ls_begin(self, xs: TEN, vs: Optional[TEN], ws_mult: int, noise_std: float, workspace: TEN) -> TEN:
        """Packs xs, computes the cross counts / their spread over the batch (and the cut values when
        `vs` is None).  Returns vs (int64 [E])."""
        xs = self._check_xs(xs)
        e = xs.shape[0]
        compute = vs is None
        if compute:
            vs = th.empty((e,), dtype=th.int64, device=self.device)
        with self._op("ls_begin", 3):
            _lib.check(self._lib.rlsb_ls_begin(self._h, _ptr(xs), e, _ptr(vs), int(compute), int(ws_mult),
                                               float(noise_std), _ptr(workspace), _stream_ptr(self.device)),
                       "ls_begin")
        return vs

    def ls_thresh(self, num_envs: int, ws_mult: int, noise: TEN, num_spin: int, workspace: TEN) -> None:
        self._check_noise(noise, num_envs)
        with self._op("ls_thresh"):
            _lib.check(self._lib.rlsb_ls_thresh(self._h, num_envs, int(ws_mult), _ptr(noise), int(num_spin),
                                                _ptr(workspace), _stream_ptr(self.device)), "ls_thresh")

    def ls_search(self, vs: TEN, ws_mult: int, noises: Sequence[TEN], finish: bool, xs_out: Optional[TEN],
                  workspace: TEN) -> None:
        e = vs.shape[0]
        for t in noises:
            self._check_noise(t, e)
        ptrs = (C.c_void_p * max(1, len(noises)))(*[t.data_ptr() for t in noises])
        with self._op("ls_search", max(1, (len(noises) + 15) // 16)):
            _lib.check(self._lib.rlsb_ls_search(self._h, e, _ptr(vs), int(ws_mult), ptrs, len(noises), int(finish),
                                                _ptr(xs_out), _ptr(workspace), _stream_ptr(self.device)),
                       "ls_search")

    def ls_run(self, vs: TEN, ws_mult: int, thresh_noise: Optional[TEN], num_spin: int, noises: Sequence[TEN],
               finish: bool, xs_out: Optional[TEN], workspace: TEN) -> None:
        """Threshold from `thresh_noise` (if given) + one noisy iteration per tensor of `noises` + (finish)
        the single-flip pass, one launch per 16 tensors."""
        e = vs.shape[0]
        for t in noises:
            self._check_noise(t, e)
        if thresh_noise is not None:
            self._check_noise(thresh_noise, e)
        ptrs = (C.c_void_p * max(1, len(noises)))(*[t.data_ptr() for t in noises])
        with self._op("ls_search", max(1, (len(noises) + 15) // 16)):
            _lib.check(self._lib.rlsb_ls_run(self._h, e, _ptr(vs), int(ws_mult), _ptr(thresh_noise), int(num_spin),
                                             ptrs, len(noises), int(finish), _ptr(xs_out), _ptr(workspace),
                                             _stream_ptr(self.device)), "ls_run")

    # ---- noisy iterations without noise tensors (csrc/noise_masks.cu): the draws of torch's CUDA generator
    # are recomputed in place and only the flip bits leave the kernel
    def ls_mask_words(self, num_envs: int) -> int:
        """uint32 words per draw of the flip-mask arrays; -1 when this path is not available for the
        graph / batch (counts wider than 8 bits, 2^31 and more elements per draw)."""
        return int(self._lib.rlsb_ls_mask_words(self._h, num_envs))

    def rng_cursor(self) -> TEN:
        """Device-resident generator state {seed, offset} used while a CUDA graph is being captured /
        replayed (the kernels read it, rng_cursor_advance moves it on the device)."""
        cur = getattr(self, "_rng_cursor", None)
        if cur is None:
            cur = self._rng_cursor = th.zeros((2,), dtype=th.int64, device=self.device)
        return cur

    def rng_cursor_sync(self) -> None:
        """cursor <- torch's CUDA generator (seed, offset).  Call before capturing / replaying a graph
        that contains local-search calls."""
        gen = rng.generator(self.device)
        seed = int(gen.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        seed = seed - (1 << 64) if seed >= (1 << 63) else seed
        self.rng_cursor().copy_(th.tensor([seed, int(gen.get_offset())], dtype=th.int64))

    def rng_cursor_commit(self) -> None:
        """torch's CUDA generator offset <- cursor (after the replays; synchronises)."""
        rng.generator(self.device).set_offset(int(self.rng_cursor()[1].item()))

    def rng_cursor_advance(self, delta: int) -> None:
        with self._op("rng_cursor_advance"):
            _lib.check(self._lib.rlsb_rng_cursor_advance(_ptr(self.rng_cursor()), int(delta),
                                                         _stream_ptr(self.device)), "rng_cursor_advance")

    def torch_randn(self, numel: int, num_draws: int, seed: int, offset: int, threads: int, iters: int,
                    cursor: Optional[TEN] = None, out: Optional[TEN] = None) -> TEN:
        """float32 [num_draws, numel]: the values `num_draws` consecutive torch.randn(numel) calls would
        return from generator state (seed, offset) -- or from the device cursor + offset."""
        if out is None:
            out = th.empty((num_draws, numel), dtype=th.float32, device=self.device)
        with self._op("torch_randn"):
            _lib.check(self._lib.rlsb_torch_randn(_ptr(out), int(numel), int(seed), int(offset), _ptr(cursor),
                                                  int(threads), int(iters), int(num_draws),
                                                  _stream_ptr(self.device)), "torch_randn")
        return out

    def ls_noise_masks(self, num_envs: int, ws_mult: int, num_draws: int, seed: int, offset: int, threads: int,
                       iters: int, workspace: TEN, cursor: Optional[TEN] = None, out: Optional[TEN] = None,
                       reuse_bound: bool = False) -> TEN:
        """Flip masks (bit e*N + n of row k) of `num_draws` consecutive randn [E, N] draws starting at
        generator state (seed, offset), against the thresholds in the workspace."""
        words = self.ls_mask_words(num_envs)
        if words < 0:
            _lib.check(3, "ls_mask_words")
        masks = out if out is not None else th.empty((max(num_draws, 1), words), dtype=th.int32, device=self.device)
        with self._op("ls_noise_masks", 2 if reuse_bound else 3):
            _lib.check(self._lib.rlsb_ls_noise_masks(self._h, num_envs, int(ws_mult), int(seed), int(offset),
                                                     _ptr(cursor), int(threads), int(iters), int(num_draws),
                                                     int(reuse_bound), _ptr(masks), _ptr(workspace),
                                                     _stream_ptr(self.device)), "ls_noise_masks")
        return masks

    def ls_run_masks(self, vs: TEN, masks: Optional[TEN], num_iters: int, finish: bool, xs_out: Optional[TEN],
                     workspace: TEN) -> None:
        with self._op("ls_run_masks"):
            _lib.check(self._lib.rlsb_ls_run_masks(self._h, vs.shape[0], _ptr(vs), _ptr(masks), int(num_iters),
                                                   int(finish), _ptr(xs_out), _ptr(workspace),
                                                   _stream_ptr(self.device)), "ls_run_masks")

    # True: the mask generator runs next to the tile kernel (rlsb_ls_fused_search); False (default): one after the
    # other (rlsb_ls_noise_masks, then rlsb_ls_run_masks).  Same results.  Measured on B200 (profiles/r02_fused_*):
    # beside a tile CTA the generator gets a quarter of the register file (8 warps per SM) and runs > 2x slower than
    # alone, which costs more than the overlap hides -- 395 vs 372 us per eager G22 x 4096 step, 4.67 vs 3.76 ms
    # at G70 x 16384 -- so the sequential form stays the default; DESIGN.md section 5 has the analysis.
    overlap_generator = False

    def _mask_scratch(self, num_draws: int, words: int) -> TEN:
        m = getattr(self, "_masks", None)
        if m is None or m.shape[0] < num_draws or m.shape[1] != words:
            m = self._masks = th.empty((max(num_draws, 1), words), dtype=th.int32, device=self.device)
        return m

    def ls_fused_search(self, vs: TEN, ws_mult: int, num_iters: int, seed: int, offset: int, threads: int, iters: int,
                        finish: bool, xs_out: Optional[TEN], workspace: TEN, cursor: Optional[TEN] = None) -> None:
        """Noisy iterations (draws recomputed from (seed, offset) or the device cursor) + single-flip pass, the
        generator streaming next to the tile kernel."""
        e = vs.shape[0]
        words = self.ls_mask_words(e)
        if words < 0:
            _lib.check(3, "ls_mask_words")
        masks = self._mask_scratch(num_iters, words)
        with self._op("ls_fused_search", 4):
            _lib.check(self._lib.rlsb_ls_fused_search(self._h, e, _ptr(vs), int(ws_mult), int(seed), int(offset),
                                                      _ptr(cursor), int(threads), int(iters), int(num_iters),
                                                      int(finish), _ptr(xs_out), _ptr(masks), _ptr(workspace),
                                                      _stream_ptr(self.device)), "ls_fused_search")

    def ls_fused_status(self, num_envs: int, workspace: TEN):
        """(generator blocks started, stalled tile CTAs, units of group 0 finished) of the last fused search."""
        out = (C.c_uint32 * 3)()
        _lib.check(self._lib.rlsb_ls_fused_status(self._h, num_envs, _ptr(workspace), out, _stream_ptr(self.device)),
                   "ls_fused_status")
        return int(out[0]), int(out[1]), int(out[2])

    def ls_begin_packed(self, packed: TEN, num_envs: int, vs: Optional[TEN], ws_mult: int, noise_std: float,
                        workspace: TEN) -> TEN:
        """ls_begin for a state given as packed tiles (uint32 [ceil(E/32), Np]).  Returns vs (int64 [E])."""
        if packed.dtype != th.int32 or tuple(packed.shape) != (self.tiles(num_envs), self.padded_nodes) \
                or not packed.is_contiguous() or packed.device != self.device:
            raise TypeError(f"packed must be a contiguous int32 [{self.tiles(num_envs)}, {self.padded_nodes}] tensor "
                            f"on {self.device}")
        compute = vs is None
        if compute:
            vs = th.empty((num_envs,), dtype=th.int64, device=self.device)
        with self._op("ls_begin", 3):
            _lib.check(self._lib.rlsb_ls_begin_packed(self._h, _ptr(packed), num_envs, _ptr(vs), int(compute),
                                                      int(ws_mult), float(noise_std), _ptr(workspace),
                                                      _stream_ptr(self.device)), "ls_begin_packed")
        return vs

    def ls_prefetch_threshold_draw(self, num_envs: int) -> None:
        """Under CUDA-graph capture: issue the threshold draw of the coming ls_fused call NOW, on a second stream, so
        that it runs next to rlsb_ls_begin (the draw does not depend on the state).  The next ls_fused joins the
        stream and uses the tensor.  Outside capture this does nothing (torch's own randn call is kept there)."""
        if not th.cuda.is_current_stream_capturing():
            return
        n = self.num_nodes
        numel = num_envs * n
        threads, iters = rng.torch_call_geometry(self.device, numel)
        buf = getattr(self, "_noise0", None)
        if buf is None or buf.numel() != numel:
            raise RuntimeError("ls_prefetch_threshold_draw: run one eager local search of this batch size before capturing "
                               "(the noise buffer and the side stream are created there)")
        cur = th.cuda.current_stream(self.device)
        side = self._side_stream
        side.wait_stream(cur)
        with th.cuda.stream(side):
            self.torch_randn(numel, 1, 0, 0, threads, iters, cursor=self.rng_cursor(), out=buf.view(1, numel))
        self._noise0_pending = True

    def ls_fused(self, vs: TEN, ws_mult: int, num_spin: int, num_iters: int, first_draw_is_iter: bool,
                 xs_out: Optional[TEN], workspace: TEN) -> None:
        """Threshold + noisy iterations + single-flip pass with the generator consumed in place.
        RNG use == the reference's: one randn [E, N] for the threshold -- which is also the first
        iteration's noise when `first_draw_is_iter` (LocalSearch.py:66-68) -- then one per iteration.
        Eager: torch draws the threshold noise, the rest is recomputed from (seed, offset) and torch's
        generator is advanced past it.  While a CUDA graph is captured every draw comes from the device
        cursor (rng_cursor_sync before, rng_cursor_commit after the replays).  xs_out None: the result
        stays in the workspace's packed tiles."""
        e, n = vs.shape[0], self.num_nodes
        numel = e * n
        threads, iters = rng.torch_call_geometry(self.device, numel)
        draws = num_iters + (0 if first_draw_is_iter else 1)      # randn calls of the reference
        if draws <= 0:
            self.ls_run_masks(vs, None, 0, True, xs_out, workspace)
            return
        first = 0 if first_draw_is_iter else 1
        capturing = th.cuda.is_current_stream_capturing()
        if capturing:
            cur, seed, base = self.rng_cursor(), 0, 0
            if getattr(self, "_noise0_pending", False):       # drawn on the side stream next to ls_begin
                th.cuda.current_stream(self.device).wait_stream(self._side_stream)
                noise0 = self._noise0.view(e, n)
                self._noise0_pending = False
            else:
                noise0 = self.torch_randn(numel, 1, 0, 0, threads, iters, cursor=cur).view(e, n)
        else:
            cur = None
            seed, base, _, _ = rng.peek(self.device, numel)
            noise0 = th.randn((e, n), dtype=th.float32, device=self.device)
            if getattr(self, "_noise0", None) is None or self._noise0.numel() != numel:
                # what a later captured call of this batch size needs: a persistent buffer and a second stream
                self._noise0 = th.empty((numel,), dtype=th.float32, device=self.device)
                self._side_stream = th.cuda.Stream(device=self.device)
        self.ls_run(vs, ws_mult, noise0, num_spin, [], False, None, workspace)
        done = 0
        chunk = 1024 if self.overlap_generator else 16384           # kLsMaxFusedDraws per fused launch
        while True:
            now = min(chunk, num_iters - done)
            last = done + now == num_iters
            off = base + 4 * iters * (first + done)
            if self.overlap_generator:
                self.ls_fused_search(vs, ws_mult, now, seed, off, threads, iters, last, xs_out if last else None,
                                     workspace, cursor=cur)
            else:
                masks = self.ls_noise_masks(e, ws_mult, now, seed, off, threads, iters, workspace, cursor=cur,
                                            out=self._mask_scratch(now, self.ls_mask_words(e)))
                self.ls_run_masks(vs, masks, now, last, xs_out if last else None, workspace)
            done += now
            if last:
                break
        if capturing:
            self.rng_cursor_advance(4 * iters * draws)
        else:
            rng.advance(self.device, numel, draws - 1)

    def _check_noise(self, t: TEN, num_envs: int) -> None:
        if t.dtype != th.float32 or tuple(t.shape) != (num_envs, self.num_nodes) or not t.is_contiguous() \
                or t.device != self.device:
            raise TypeError(f"noise must be a contiguous float32 [{num_envs}, {self.num_nodes}] tensor on {self.device}")

    def step_flip(self, xs: TEN, action: TEN, reward: TEN, cut: TEN, bad: TEN) -> None:
        with self._op("step_flip"):
            _lib.check(self._lib.rlsb_step_flip(self._h, _ptr(xs), _ptr(action), xs.shape[0], _ptr(reward), _ptr(cut),
                                                _ptr(bad), _stream_ptr(self.device)), "step_flip")

    def greedy_best_flip(self, xs: TEN, max_flips: int, strict: bool = True):
        """In place on xs (bool [E,N]).  Returns (vs int64 [E], flips int32 [E])."""
        xs_c = self._check_xs(xs)
        if xs_c.data_ptr() != xs.data_ptr():
            raise RuntimeError("greedy_best_flip mutates xs in place: it must be contiguous")
        e = xs.shape[0]
        vs = th.empty((e,), dtype=th.int64, device=self.device)
        flips = th.zeros((e,), dtype=th.int32, device=self.device)
        with self._op("greedy_best_flip"):
            _lib.check(self._lib.rlsb_greedy_best_flip(self._h, _ptr(xs), e, _ptr(vs), _ptr(flips), int(max_flips),
                                                       int(bool(strict)), _stream_ptr(self.device)),
                       "greedy_best_flip")
        return vs, flips

    def flip_sweep(self, packed: TEN, vs: TEN) -> None:
        with self._op("flip_sweep"):
            _lib.check(self._lib.rlsb_flip_sweep(self._h, _ptr(packed), _ptr(vs), vs.shape[0],
                                                 _stream_ptr(self.device)), "flip_sweep")


def select_rows(xs0: TEN, vs0: TEN, xs1: TEN, vs1: TEN, if_maximize: bool = True) -> None:
    """In-place update_xs_by_vs on device tensors (bool [E,N] contiguous, int64 [E])."""
    lib = _lib.lib()
    dev = require_cuda(xs0.device)
    if not (xs0.is_contiguous() and vs0.is_contiguous()):
        raise RuntimeError("select_rows mutates xs0 / vs0 in place: they must be contiguous")
    if xs0.dtype != th.bool or xs1.dtype != th.bool or xs0.shape != xs1.shape:
        raise TypeError("select_rows: xs0/xs1 must be bool tensors of one shape")
    if vs0.dtype != th.int64:
        raise TypeError("select_rows: vs0 must be int64")
    vs1 = vs1.to(th.int64)
    with on_device(dev):
        _lib.check(lib.rlsb_select_rows(_ptr(xs0), _ptr(vs0), _ptr(xs1.contiguous()), _ptr(vs1.contiguous()),
                                        xs0.shape[0], xs0.shape[1], int(bool(if_maximize)), _stream_ptr(dev)),
                   "select_rows")


def pick_best(xs: TEN, vs: TEN, num_repeats: int, if_maximize: bool = True) -> Tuple[TEN, TEN]:
    lib = _lib.lib()
    dev = require_cuda(xs.device)
    if xs.dtype != th.bool:
        raise TypeError("pick_best: xs must be bool")
    n = xs.shape[1]
    sims = xs.shape[0] // num_repeats
    xs = xs.contiguous()
    vs64 = vs.to(th.int64).contiguous()
    out_xs = th.empty((sims, n), dtype=th.bool, device=dev)
    out_vs = th.empty((sims,), dtype=th.int64, device=dev)
    with on_device(dev):
        _lib.check(lib.rlsb_pick_best(_ptr(xs), _ptr(vs64), int(num_repeats), sims, n, int(bool(if_maximize)),
                                      _ptr(out_xs), _ptr(out_vs), _stream_ptr(dev)), "pick_best")
    return out_xs, out_vs.to(vs.dtype)
