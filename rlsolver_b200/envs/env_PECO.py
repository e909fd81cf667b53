"""Drop-in for the PECO pattern-I environment of the reference: `SpinSystemFactory` /
`SpinSystemUnbiased` (rlsolver/methods/ECO_S2V/src/envs/spinsystem_PECO.py:16-50, 67-193,
306-497, 587-662), the enums and observable lists of `src/envs/util_envs.py:11-60`, and the
on-device graph generators of `src/envs/util_envs_PECO.py:15-113`.  (`rlsolver/envs/env_PECO.py` is
an empty placeholder in the reference; this module is what it was meant to hold.)

One graph PER environment (`matrix [E, N, N]`), spins +-1, `state [E, num_obs, N]` float32 with the
reference's observables, `reset() -> obs`, `step(action) -> (obs, rew, done)`,
`get_observation() -> [E, num_obs + N, N]`, `get_best_cut()`, `best_spins`, `score`, `matrix`.
The reference recomputes all local fields with a batched matmul and clones the state every step.

Two resident layouts behind the same attributes:
  * COMPACT (csrc/peco_compact.cu) whenever every weight is -1, 0 or +1 -- all the reference's generators: adjacency
    and sign bit rows, packed spins, int8 / int16 fields, a 64-bit hashed visited set.  `state`, `matrix`, `best_spins` and
    the observation are materialised from it on request, with the reference's float32 values.  3.3 KB per env
    instead of 43 KB at N = 100; the step touches ~0.5 KB.
  * DENSE (csrc/peco.cu) for arbitrary float weights handed in through SetMatrixGenerator: the reference's tensors,
    updated in place.
The ER / BA generators write the compact layout straight from torch's Philox stream (same seed, same graphs as the
reference's torch ops; `get()` still returns the dense tensor for callers that want it).
Supported configuration = what PECO trains with (train_PECO.py:34-44): unbiased graphs, ExtraAction.NONE, infinite
memory, reversible spins -- plus the S2V-DQN pattern of the reference's single-env NumPy environment
(ECO_S2V/src/envs/spinsystem.py:242-247, 476-480; train_S2V.py:37-47): `reversible_spins=False` starts every env at all
+1 and ends an env's episode when no +1 spin is left; with `S2V_OBSERVABLES`, `RewardSignal.DENSE` and `norm_rewards` a
step is flip -> reward = delta cut / N.  (The batched reference's own irreversible / finite-memory / ExtraAction branches
index the [E, obs, N] state as if it were one env's and are not reproduced.)
"""
from __future__ import annotations

import ctypes as C
from enum import Enum
from typing import List, Optional

import numpy as np
import torch as th

from .. import _lib, rng
from ..graph_store import _ptr, _stream_ptr, on_device, require_cuda

TEN = th.Tensor


class EdgeType(Enum):
    UNIFORM = 1
    DISCRETE = 2
    RANDOM = 3


class RewardSignal(Enum):
    DENSE = 1
    BLS = 2
    SINGLE = 3
    CUSTOM_BLS = 4


class ExtraAction(Enum):
    PASS = 1
    RANDOMISE = 2
    NONE = 3


class OptimisationTarget(Enum):
    CUT = 1
    ENERGY = 2


class SpinBasis(Enum):
    SIGNED = 1
    BINARY = 2


class Observable(Enum):
    SPIN_STATE = 1
    IMMEDIATE_REWARD_AVAILABLE = 2
    TIME_SINCE_FLIP = 3
    EPISODE_TIME = 4
    TERMINATION_IMMANENCY = 5
    NUMBER_OF_GREEDY_ACTIONS_AVAILABLE = 6
    DISTANCE_FROM_BEST_SCORE = 7
    DISTANCE_FROM_BEST_STATE = 8


ECO_PECO_OBSERVABLES = [Observable.SPIN_STATE,
                        Observable.IMMEDIATE_REWARD_AVAILABLE,
                        Observable.TIME_SINCE_FLIP,
                        Observable.DISTANCE_FROM_BEST_SCORE,
                        Observable.DISTANCE_FROM_BEST_STATE,
                        Observable.NUMBER_OF_GREEDY_ACTIONS_AVAILABLE,
                        Observable.TERMINATION_IMMANENCY]
S2V_OBSERVABLES = [Observable.SPIN_STATE]

_STEP_ROWS = [Observable.IMMEDIATE_REWARD_AVAILABLE, Observable.TIME_SINCE_FLIP, Observable.EPISODE_TIME,
              Observable.TERMINATION_IMMANENCY, Observable.NUMBER_OF_GREEDY_ACTIONS_AVAILABLE,
              Observable.DISTANCE_FROM_BEST_SCORE, Observable.DISTANCE_FROM_BEST_STATE]


# ----------------------------------------------------------------------------- compact graphs
def _words(n: int) -> int:
    return (n + 31) // 32


def pack_rows(bits: TEN) -> TEN:
    """bool [..., n] -> int32 [..., ceil(n/32)] (bit j of word k = element 32k + j).  Small tensors only (sign
    masks, spins): plain torch ops."""
    n = bits.shape[-1]
    w = _words(n)
    padded = th.zeros((*bits.shape[:-1], w * 32), dtype=th.int64, device=bits.device)
    padded[..., :n] = bits.to(th.int64)
    weights = th.ones(32, dtype=th.int64, device=bits.device) << th.arange(32, device=bits.device)
    words = (padded.reshape(*bits.shape[:-1], w, 32) * weights).sum(dim=-1)
    return th.where(words >= 2 ** 31, words - 2 ** 32, words).to(th.int32)


def unpack_rows(words: TEN, n: int) -> TEN:
    """int32 [..., W] -> bool [..., n]."""
    shifts = th.arange(32, device=words.device)
    bits = ((words.to(th.int64).unsqueeze(-1) >> shifts) & 1).bool()
    return bits.reshape(*words.shape[:-1], -1)[..., :n]


class CompactGraphs:
    """adjacency bit rows int32 [E, N, W] + sign bit rows: per env [E, N, W], shared [N, W] or None (all +1)."""

    def __init__(self, adj: TEN, sgn: Optional[TEN], n_spins: int):
        self.adj, self.sgn, self.n = adj, sgn, n_spins
        self.num_envs = adj.shape[0]
        self.device = adj.device

    @property
    def sgn_stride(self) -> int:
        return 0 if self.sgn is None or self.sgn.dim() == 2 else self.n * _words(self.n)

    def dense(self, out: Optional[TEN] = None, env_stride: Optional[int] = None) -> TEN:
        """float32 [E, N, N] (or written into `out` with `env_stride` floats between envs)."""
        n = self.n
        if out is None:
            out = th.empty((self.num_envs, n, n), dtype=th.float32, device=self.device)
            env_stride = n * n
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_peco_compact_expand_matrix(_ptr(self.adj), _ptr(self.sgn), self.sgn_stride,
                                                                  self.num_envs, n, out.data_ptr(), int(env_stride),
                                                                  _stream_ptr(self.device)), "peco_compact_expand_matrix")
        return out

    @staticmethod
    def from_dense(matrix: TEN) -> Optional["CompactGraphs"]:
        """None when an entry is not in {-1, 0, +1} (the dense layout serves those)."""
        e, n, _ = matrix.shape
        dev = matrix.device
        m = matrix.to(th.float32).contiguous()
        adj = th.empty((e, n, _words(n)), dtype=th.int32, device=dev)
        sgn = th.empty_like(adj)
        bad = th.zeros((1,), dtype=th.int32, device=dev)
        with on_device(dev):
            _lib.check(_lib.lib().rlsb_peco_compact_from_dense(_ptr(m), e, n, _ptr(adj), _ptr(sgn), _ptr(bad),
                                                               _stream_ptr(dev)), "peco_compact_from_dense")
        if int(bad.item()) != 0:
            return None
        return CompactGraphs(adj, sgn if bool(sgn.any()) else None, n)


# ----------------------------------------------------------------------------- graph generators
class GraphGenerator:
    def __init__(self, n_spins, edge_type, biased=False, num_envs=None):
        self.n_spins, self.edge_type, self.biased, self.num_envs = n_spins, edge_type, biased, num_envs

    def _sign_rows(self) -> Optional[TEN]:
        """Sign bit rows of the edge-weight mask the reference multiplies the adjacency with
        (util_envs_PECO.py:21-38 / 67-84), consuming the generator like its randint call does: UNIFORM draws
        nothing; DISCRETE draws ONE [n, n] matrix of {0, 1} shared by all envs (0 -> weight -1); RANDOM one per env.
        The mask is made symmetric from its lower triangle: entry (i, j) takes the draw at (max(i, j), min(i, j))."""
        n, dev = self.n_spins, self.device
        if self.edge_type == EdgeType.UNIFORM:
            return None
        if self.edge_type == EdgeType.DISCRETE:
            neg = th.randint(0, 2, (n, n), device=dev) == 0
        elif self.edge_type == EdgeType.RANDOM:
            neg = th.randint(0, 2, (self.num_envs, n, n), dtype=th.float32, device=dev) == 0
        else:
            raise NotImplementedError()
        lower = th.ones((n, n), dtype=th.bool, device=dev).tril()
        return pack_rows(th.where(lower, neg, neg.transpose(-1, -2)))

    def get_compact(self) -> CompactGraphs:
        raise NotImplementedError

    def get(self, with_padding=False):
        """The reference's return value: float32 [E, N, N]."""
        return self.get_compact().dense()

    def _chunks(self, elems_per_env: int):
        """Env ranges whose RNG call stays below 2^31 elements (torch itself splits larger calls differently: for
        such batches the graphs equal the reference's run chunk by chunk)."""
        per = max(1, ((1 << 31) - 1) // max(elems_per_env, 1))
        return [(lo, min(lo + per, self.num_envs)) for lo in range(0, self.num_envs, per)]


class RandomERGraphGenerator(GraphGenerator):
    """util_envs_PECO.py:15-57: adj[i < j] = rand(E, n, n) < p, mirrored; csrc/peco_compact.cu peco_gen_er_kernel."""

    def __init__(self, n_spins=20, p_connection=0.2, edge_type=EdgeType.DISCRETE, num_envs=8, device="cuda"):
        super().__init__(n_spins, edge_type, False, num_envs)
        self.p_connection, self.device = p_connection, require_cuda(device)

    def get_compact(self) -> CompactGraphs:
        e, n, dev = self.num_envs, self.n_spins, self.device
        adj = th.empty((e, n, _words(n)), dtype=th.int32, device=dev)
        lib = _lib.lib()
        for lo, hi in self._chunks(n * n):
            numel = (hi - lo) * n * n
            seed, offset, threads, iters = rng.peek(dev, numel)
            with on_device(dev):
                _lib.check(lib.rlsb_peco_gen_er(_ptr(adj[lo:hi]), hi - lo, n, float(np.float32(self.p_connection)),
                                                seed, offset, threads, iters, _stream_ptr(dev)), "peco_gen_er")
            rng.advance(dev, numel, 1)
        return CompactGraphs(adj, self._sign_rows(), n)


class RandomBAGraphGenerator(GraphGenerator):
    """util_envs_PECO.py:60-113, including its quirk: the initial clique loop also sets the first m+1 diagonal
    entries (self loops), kept for parity; csrc/peco_compact.cu peco_gen_ba_kernel."""

    def __init__(self, n_spins=20, m_insertion_edges=4, edge_type=EdgeType.DISCRETE, num_envs=8, device="cuda"):
        super().__init__(n_spins, edge_type, False, num_envs)
        self.m_insertion_edges, self.device = m_insertion_edges, require_cuda(device)

    def get_compact(self) -> CompactGraphs:
        e, n, m, dev = self.num_envs, self.n_spins, self.m_insertion_edges, self.device
        adj = th.empty((e, n, _words(n)), dtype=th.int32, device=dev)
        lib = _lib.lib()
        calls = max(0, n - m - 1)                      # one exponential_ over [E, n] per inserted node
        for lo, hi in self._chunks(n):
            numel = (hi - lo) * n
            seed, offset, threads, iters = rng.peek(dev, numel)
            with on_device(dev):
                _lib.check(lib.rlsb_peco_gen_ba(_ptr(adj[lo:hi]), hi - lo, n, m, seed, offset, threads, iters,
                                                _stream_ptr(dev)), "peco_gen_ba")
            rng.advance(dev, numel, calls)
        return CompactGraphs(adj, self._sign_rows(), n)


class RandomPLGraphGenerator(GraphGenerator):
    """Power-law cluster graphs (Holme-Kim), the reference's third graph family (`GraphType.PL`:
    `nx.powerlaw_cluster_graph(n, m=4, p=0.05)`, rlsolver/methods/util_generate.py:75-93), one graph per env on the
    device (csrc/peco_compact.cu peco_gen_pl_kernel).  The reference grows single graphs on the host from Python's
    `random`, so there is no stream to reproduce: the graphs follow the same growth process from torch's CUDA
    generator state (seed, offset; the offset moves by 64 * n_spins per call) and agree with networkx statistically."""

    def __init__(self, n_spins=20, m_insertion_edges=4, p_triangle=0.05, edge_type=EdgeType.DISCRETE, num_envs=8,
                 device="cuda"):
        super().__init__(n_spins, edge_type, False, num_envs)
        self.m_insertion_edges, self.p_triangle, self.device = m_insertion_edges, p_triangle, require_cuda(device)

    def get_compact(self) -> CompactGraphs:
        e, n, dev = self.num_envs, self.n_spins, self.device
        adj = th.empty((e, n, _words(n)), dtype=th.int32, device=dev)
        gen = rng.generator(dev)
        seed, offset = int(gen.initial_seed()) & 0xFFFFFFFFFFFFFFFF, int(gen.get_offset())
        with on_device(dev):
            _lib.check(_lib.lib().rlsb_peco_gen_pl(_ptr(adj), e, n, int(self.m_insertion_edges),
                                                   float(self.p_triangle), seed, offset, _stream_ptr(dev)), "peco_gen_pl")
        gen.set_offset(offset + 64 * n)
        return CompactGraphs(adj, self._sign_rows(), n)


class SetMatrixGenerator(GraphGenerator):
    """Hands out a fixed `[E, N, N]` tensor (replaying recorded graphs)."""

    def __init__(self, matrix: TEN):
        super().__init__(matrix.shape[-1], EdgeType.DISCRETE, False, matrix.shape[0])
        self.matrix, self.device = matrix, matrix.device

    def get(self, with_padding=False):
        return self.matrix.clone()

    def get_compact(self):
        return CompactGraphs.from_dense(self.matrix)


# ----------------------------------------------------------------------------- environment
class SpinSystemFactory(object):
    @staticmethod
    def get(graph_generator=None, max_steps=20, observables=ECO_PECO_OBSERVABLES, reward_signal=RewardSignal.DENSE,
            extra_action=ExtraAction.PASS, optimisation_target=OptimisationTarget.ENERGY, spin_basis=SpinBasis.SIGNED,
            norm_rewards=False, memory_length=None, horizon_length=None, stag_punishment=None, basin_reward=None,
            reversible_spins=True, init_snap=None, seed=None, device=None, num_envs=None):
        if graph_generator.biased:
            raise NotImplementedError("biased spin systems are outside the max-cut hot path")
        return SpinSystemUnbiased(graph_generator, max_steps, observables, reward_signal, extra_action,
                                  optimisation_target, spin_basis, norm_rewards, memory_length, horizon_length,
                                  stag_punishment, basin_reward, reversible_spins, init_snap, seed, device, num_envs)


_ZOBRIST_SEED = 0x9E3779B97F4A7C15


class SpinSystemUnbiased:
    # False forces the dense layout (csrc/peco.cu: exact linear visited-state scan, the reference's tensors resident)
    allow_compact = True

    def __init__(self, graph_generator=None, max_steps=20, observables=ECO_PECO_OBSERVABLES,
                 reward_signal=RewardSignal.DENSE, extra_action=ExtraAction.PASS,
                 optimisation_target=OptimisationTarget.ENERGY, spin_basis=SpinBasis.SIGNED, norm_rewards=False,
                 memory_length=None, horizon_length=None, stag_punishment=None, basin_reward=None,
                 reversible_spins=False, init_snap=None, seed=None, device=None, num_envs=None):
        if seed is not None:
            np.random.seed(seed)
        assert observables[0] == Observable.SPIN_STATE, "First observable must be Observation.SPIN_STATE."
        if extra_action != ExtraAction.NONE:
            raise NotImplementedError("only ExtraAction.NONE (PECO's configuration, train_PECO.py:36)")
        if memory_length is not None or init_snap is not None:
            raise NotImplementedError("infinite memory only (train_PECO.py:40-44; the reference's finite-memory branch "
                                      "indexes the batched state as a single env's)")
        if reward_signal == RewardSignal.SINGLE:
            raise NotImplementedError("RewardSignal.SINGLE")
        if optimisation_target != OptimisationTarget.CUT:
            raise NotImplementedError("OptimisationTarget.CUT only: this is the max-cut path")
        self.device = require_cuda(device)
        self.num_envs = num_envs
        self.observables = list(enumerate(observables))
        self.extra_action = extra_action
        self.gg = graph_generator
        self.n_spins = self.gg.n_spins
        self.max_steps = max_steps
        self.reward_signal = reward_signal
        self.norm_rewards = norm_rewards
        self.n_actions = self.n_spins
        self.current_step = 0
        self.optimisation_target = optimisation_target
        self.spin_basis = spin_basis
        self.memory_length = memory_length
        self.horizon_length = horizon_length if horizon_length is not None else self.max_steps
        self.stag_punishment = stag_punishment
        self.basin_reward = basin_reward
        self.reversible_spins = reversible_spins
        self._lib = _lib.lib()
        self._rows = np.asarray([next((i for i, o in self.observables if o == want), -1) for want in _STEP_ROWS],
                                dtype=np.int32)
        self._bad = th.zeros((1,), dtype=th.int32, device=self.device)
        self._history = None
        # `x / n_spins` with a Python scalar: torch CUDA multiplies by the float32 reciprocal, torch CPU divides
        # (1 ulp apart).  True reproduces the reference running on this GPU; the CPU-generated goldens need False.
        self.scalar_div_as_cuda = True
        # k-fold float32 accumulation of 1 / max_steps: what `state[row] += 1 / max_steps` has produced after k steps
        inv = np.float32(1. / self.max_steps)
        table = np.zeros(self.max_steps + 2, np.float32)
        for k in range(1, table.size):
            table[k] = np.float32(table[k - 1] + inv)
        self._table = th.from_numpy(table).to(self.device)
        zr = np.random.default_rng(_ZOBRIST_SEED & 0xFFFFFFFF)
        self._zobrist = th.from_numpy(zr.integers(1, 2 ** 63 - 1, size=self.n_spins, dtype=np.int64)).to(self.device)
        self._compact: Optional[CompactGraphs] = None
        self._draw_graphs()                  # the reference draws one graph batch in __init__ and another in reset()
        self.reset(return_observation=False)

    # ------------------------------------------------------------------ layout plumbing
    @property
    def state_layout(self) -> str:
        return (f"compact: adjacency / sign bit rows, packed spins, {'int8' if self.n_spins <= 128 else 'int16'} fields, "
                "hashed visited set"
                if self._compact is not None else "dense float32 matrix [E,N,N] + state [E,obs,N]")

    def _draw_graphs(self) -> None:
        """One graph batch from the generator, in the compact layout when its weights allow it."""
        self._compact = None
        self._matrix_cache = None
        if self.allow_compact and hasattr(self.gg, "get_compact"):
            self._compact = self.gg.get_compact()
        if self._compact is None:
            self._matrix_cache = self.gg.get().to(self.device).to(th.float32).contiguous()

    @property
    def matrix(self) -> TEN:
        if self._matrix_cache is None:
            self._matrix_cache = self._compact.dense()
        return self._matrix_cache

    @property
    def matrix_obs(self) -> TEN:
        return self.matrix

    @property
    def _field_dtype(self):
        """(A s)_j is bounded by the degree: one byte up to 128 spins (csrc/peco_compact.cu kPcByteFields)."""
        return th.int8 if self.n_spins <= 128 else th.int16

    def mean_degree(self) -> float:
        if self._compact is not None:
            sample = self._compact.adj[:1024]
            return float(unpack_rows(sample, self.n_spins).float().sum() / sample.shape[0] / self.n_spins)
        return float((self.matrix[:1024] != 0).float().sum() / min(self.num_envs, 1024) / self.n_spins)

    def step_algorithmic_bytes(self) -> int:
        """HBM bytes one step has to move per launch: compact = action + one adjacency and one sign row +
        the fields of the acted node's neighbours (read + write) + every field once (the greedy-actions count needs them
        all) + spin words (r/w) + best spins (r) + last_flip entry + score / best / max_local / reward scalars."""
        n, w, e = self.n_spins, _words(self.n_spins), self.num_envs
        if self._compact is None:
            return e * (4 * n + 8 * n + 2 * 4 * n * 4 + 40)
        deg = self.mean_degree()
        fb = 1 if n <= 128 else 2                    # bytes per field
        return int(e * (8 + 8 * w + 2 * fb * deg + fb * n + 8 * w + 4 * w + 2 + 24))

    # ------------------------------------------------------------------ kernels
    def _fields(self, spins: TEN, want_as: bool = False, want_cut: bool = False):
        """Dense layout: (A s) s, A s and the cut for float spins [E, N]."""
        e, n = spins.shape
        spins = spins.contiguous()
        fields = th.empty((e, n), dtype=th.float32, device=self.device)
        as_ = th.empty((e, n), dtype=th.float32, device=self.device) if want_as else None
        cut = th.empty((e,), dtype=th.float32, device=self.device) if want_cut else None
        with on_device(self.device):
            _lib.check(self._lib.rlsb_peco_fields(_ptr(self.matrix), _ptr(spins), e, n, _ptr(as_), _ptr(fields), _ptr(cut),
                                                  _stream_ptr(self.device)), "peco_fields")
        return fields, as_, cut

    def _compact_fields(self, spin_words: TEN, fields: Optional[TEN], cut: Optional[TEN], max_local: Optional[TEN],
                        empty: Optional[TEN]) -> None:
        c = self._compact
        with on_device(self.device):
            _lib.check(self._lib.rlsb_peco_compact_fields(_ptr(c.adj), _ptr(c.sgn), c.sgn_stride, _ptr(spin_words),
                                                          self.num_envs, self.n_spins, _ptr(fields), _ptr(cut),
                                                          _ptr(max_local), _ptr(empty), _stream_ptr(self.device)),
                       "peco_compact_fields")

    def _get_immeditate_cuts_avaialable(self, spins: TEN, matrix: Optional[TEN] = None) -> TEN:
        if self._compact is None:
            return self._fields(spins)[0]
        words = pack_rows(spins > 0)
        fields = th.empty((self.num_envs, _words(self.n_spins) * 32), dtype=self._field_dtype, device=self.device)
        self._compact_fields(words, fields, None, None, None)
        return fields[:, :self.n_spins].float() * spins

    # ------------------------------------------------------------------ reset (spinsystem_PECO.py:151-193)
    def reset(self, spins=None, return_observation: bool = True):
        """`return_observation=False` skips materialising the [E, obs + N, N] observation (43 GB at 10^6 envs)."""
        self.current_step = 0
        e, n, dev = self.num_envs, self.n_spins, self.device
        w = _words(n)
        while True:
            self._draw_graphs()
            if self._compact is not None:
                ones = th.full((e, w), -1, dtype=th.int32, device=dev)
                self.max_local_reward_available_ = th.empty((e,), dtype=th.float32, device=dev)
                empty = th.zeros((1,), dtype=th.int32, device=dev)
                self._compact_fields(ones, None, None, self.max_local_reward_available_, empty)
                retry = int(empty.item()) != 0
            else:
                local = self._fields(th.ones(e, n, device=dev))[0]
                retry = bool(th.any(th.eq(th.sum(th.abs(local), dim=-1), 0)))
                if not retry:
                    self.max_local_reward_available_ = th.max(local, dim=-1).values
                    retry = bool((self.max_local_reward_available_ == 0).any())
            if not retry:                 # an empty graph was generated: draw again (as the reference does)
                break
        self.max_local_reward_available = self.max_local_reward_available_.unsqueeze(1).expand(-1, n)
        if spins is None:
            if self.reversible_spins:
                spins_f = 2 * th.randint(0, 2, (e, n), device=dev, dtype=th.float) - 1
            else:       # irreversible (S2V-DQN, ECO_S2V/src/envs/spinsystem.py:242-247): every spin may still be flipped
                spins_f = th.ones((e, n), device=dev, dtype=th.float)
        else:
            spins_f = spins.to(dev).to(th.float32)
        self._state_cache = None
        self._hist_len = 0
        self._history = None
        use_hist = self.stag_punishment is not None or self.basin_reward is not None
        if self._compact is not None:
            self._spins = pack_rows(spins_f > 0)
            self._cfields = th.empty((e, w * 32), dtype=self._field_dtype, device=dev)
            self.score = th.empty((e,), dtype=th.float32, device=dev)
            self._compact_fields(self._spins, self._cfields, self.score, None, None)
            self._last_flip = th.zeros((e, w * 32), dtype=th.int16, device=dev)
            self._best_words = self._spins.clone()
            self._hset = self._hkey = None
            if use_hist:
                cap = 32
                while cap < 2 * (self.max_steps + 1):
                    cap *= 2
                self._hcap = cap
                self._hset = th.zeros((e, cap), dtype=th.int64, device=dev)
                self._hkey = th.zeros((e,), dtype=th.int64, device=dev)
        else:
            self._dense_state = self._reset_state(spins_f)
            self.score = self.calculate_score()
            self._dense_best_spins = self._dense_state[:, 0, :n].clone()
            if use_hist:
                self._history = th.empty((self.max_steps + 1, e, w), dtype=th.int32, device=dev)
        self.best_score = self.score.clone()
        self.best_obs_score = self.score.clone()
        self.best_obs_spins = None                      # == best_spins (materialised on request)
        self._seed_best()
        return self.get_observation() if return_observation else None

    def _seed_best(self) -> None:
        """Hook: what best_score / best_spins start from (here: every env's own start, set by reset)."""

    def _reset_state(self, spins_f: TEN):
        """Dense layout: the reference's state tensor after reset (spinsystem_PECO.py:173-193)."""
        state = th.zeros(self.num_envs, len(self.observables), self.n_actions, device=self.device)
        state[:, 0, :] = spins_f
        imm, self._as, _ = self._fields(state[:, 0, :self.n_spins], want_as=True)
        for idx, obs in self.observables:
            if obs == Observable.IMMEDIATE_REWARD_AVAILABLE:
                state[:, idx, :self.n_spins] = imm / self.max_local_reward_available
            elif obs == Observable.NUMBER_OF_GREEDY_ACTIONS_AVAILABLE:
                cnt = th.sum(imm <= 0, dim=-1).float()
                frac = cnt / self.n_spins if self.scalar_div_as_cuda else cnt / th.full_like(cnt, self.n_spins)
                state[:, idx, :self.n_spins] = (1 - frac).unsqueeze(-1)
        return state

    # ------------------------------------------------------------------ the reference's tensors, on request
    def _termination(self) -> float:
        if self.current_step == 0:
            return 0.0                      # reset leaves the row at zero (spinsystem_PECO.py:173-193)
        # max(0, float32((step - max_steps) / horizon) + 1) as the reference computes it
        return float(np.float32(max(np.float32(0.), np.float32((self.current_step - self.max_steps) / self.horizon_length)
                                    + np.float32(1.))))

    def _expand_state(self, out: TEN, env_stride: int, binary: bool) -> None:
        with on_device(self.device):
            _lib.check(self._lib.rlsb_peco_compact_expand_state(
                _ptr(self._spins), _ptr(self._best_words), _ptr(self._cfields), _ptr(self._last_flip), _ptr(self.score),
                _ptr(self.best_score), _ptr(self.max_local_reward_available_), _ptr(self._table), out.data_ptr(),
                int(env_stride), self.num_envs, self.n_spins, len(self.observables), self._rows.ctypes.data,
                self.current_step, int(binary), self._termination(), int(bool(self.scalar_div_as_cuda)),
                int(self.current_step == 0), _stream_ptr(self.device)), "peco_compact_expand_state")

    @property
    def state(self) -> TEN:
        """float32 [E, num_obs, N], spins in the +-1 basis (the reference's `self.state`)."""
        if self._compact is None:
            return self._dense_state
        if self._state_cache is None:
            self._state_cache = th.empty((self.num_envs, len(self.observables), self.n_spins), dtype=th.float32,
                                         device=self.device)
            self._expand_state(self._state_cache, len(self.observables) * self.n_spins, False)
        return self._state_cache

    @property
    def best_spins(self) -> TEN:
        if self._compact is None:
            return self._dense_best_spins
        return unpack_rows(self._best_words, self.n_spins).float() * 2 - 1

    @best_spins.setter
    def best_spins(self, value: TEN) -> None:
        if self._compact is None:
            self._dense_best_spins = value
        else:
            self._best_words = pack_rows(value > 0)

    def calculate_cut(self, spins=None):
        if self._compact is None:
            if spins is None:
                spins = self._dense_state[:, 0, :self.n_spins]
            return self._fields(spins, want_cut=True)[2]
        words = self._spins if spins is None else pack_rows(spins > 0)
        cut = th.empty((self.num_envs,), dtype=th.float32, device=self.device)
        self._compact_fields(words, None, cut, None, None)
        return cut

    def calculate_score(self, spins=None):
        return self.calculate_cut(spins)

    def get_best_cut(self):
        return self.best_score

    # ------------------------------------------------------------------ step (spinsystem_PECO.py:306-486)
    def step(self, action: TEN, return_observation: bool = True):
        self.current_step += 1
        if self.current_step > self.max_steps:
            print("The environment has already returned done. Stop it!")
            raise NotImplementedError
        if action.dtype != th.int64 or action.device != self.device:
            action = action.to(device=self.device, dtype=th.int64)
        action = action.reshape(self.num_envs).contiguous()
        rew = th.empty((self.num_envs,), dtype=th.float32, device=self.device)
        done = th.empty((self.num_envs,), dtype=th.bool, device=self.device)
        use_stag, use_basin = self.stag_punishment is not None, self.basin_reward is not None
        if self._compact is not None:
            c = self._compact
            with on_device(self.device):
                _lib.check(self._lib.rlsb_peco_compact_step(
                    _ptr(c.adj), _ptr(c.sgn), c.sgn_stride, _ptr(self._spins), _ptr(self._cfields), _ptr(self._last_flip),
                    _ptr(self._best_words), _ptr(self.score), _ptr(self.best_score), _ptr(self.max_local_reward_available_),
                    _ptr(rew), _ptr(action), _ptr(self._hset), getattr(self, "_hcap", 0) if self._hset is not None else 0,
                    _ptr(self._hkey), _ptr(self._zobrist), _ptr(self._bad), self.num_envs, self.n_spins,
                    self.current_step, self.reward_signal.value, int(bool(self.norm_rewards)), int(use_stag),
                    float(self.stag_punishment or 0.0), int(use_basin), float(self.basin_reward or 0.0),
                    int(bool(self.scalar_div_as_cuda)), _ptr(done), int(self.current_step == self.max_steps),
                    int(not self.reversible_spins), _stream_ptr(self.device)), "peco_compact_step")
            self._state_cache = None
        else:
            with on_device(self.device):
                _lib.check(self._lib.rlsb_peco_step(
                    _ptr(self.matrix), _ptr(self._dense_state), _ptr(self._as), _ptr(action), _ptr(self.score),
                    _ptr(self.best_score), _ptr(self._dense_best_spins), _ptr(self.max_local_reward_available_), _ptr(rew),
                    _ptr(self._history), self._hist_len, _ptr(self._bad), self.num_envs, self.n_spins,
                    len(self.observables), self._rows.ctypes.data, self.reward_signal.value, int(bool(self.norm_rewards)),
                    float(np.float32(1. / self.max_steps)), self._termination(), int(use_stag),
                    float(self.stag_punishment or 0.0), int(use_basin), float(self.basin_reward or 0.0),
                    int(bool(self.scalar_div_as_cuda)), _stream_ptr(self.device)), "peco_step")
            if self._history is not None:
                self._hist_len += 1
        self.best_obs_score = self.best_score           # infinite memory (spinsystem_PECO.py:427-429)
        self.best_obs_spins = None                      # == best_spins (materialised on request)
        if self._compact is None:           # (the compact kernel wrote `done` itself)
            done.fill_(self.current_step == self.max_steps)
            if not self.reversible_spins:   # "no more spins to flip" (spinsystem.py:476-480), per env of the batch
                done |= ~(self._dense_state[:, 0, :self.n_spins] > 0).any(dim=-1)
        obs = self.get_observation() if return_observation else None
        return obs, rew, done

    def get_observation(self):
        """float32 [E, num_obs + N, N]: the state rows (spins in the configured basis) over the matrix."""
        n, k = self.n_spins, len(self.observables)
        if self._compact is None:
            state = self._dense_state.clone()
            if self.spin_basis == SpinBasis.BINARY:
                state[:, 0, :] = (1 - state[:, 0, :]) / 2
            return th.cat((state, self.matrix), dim=-2)
        obs = th.empty((self.num_envs, k + n, n), dtype=th.float32, device=self.device)
        self._expand_state(obs, (k + n) * n, self.spin_basis == SpinBasis.BINARY)
        self._compact.dense(out=obs[:, k:, :], env_stride=(k + n) * n)
        return obs

    def num_bad_actions(self) -> int:
        return int(self._bad.item())
