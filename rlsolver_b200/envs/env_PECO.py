"""Drop-in for the PECO pattern-I environment of the reference: `SpinSystemFactory` /
`SpinSystemUnbiased` (rlsolver/methods/ECO_S2V/src/envs/spinsystem_PECO.py:16-50, 67-193,
306-497, 587-662), the enums and observable lists of `src/envs/util_envs.py:11-60`, and the
on-device graph generators of `src/envs/util_envs_PECO.py:15-113`.  (`rlsolver/envs/env_PECO.py` is
an empty placeholder in the reference; this module is what it was meant to hold.)

One graph PER environment (`matrix [E, N, N]`), spins +-1, `state [E, num_obs, N]` float32 with the
reference's observables, `reset() -> obs`, `step(action) -> (obs, rew, done)`,
`get_observation() -> [E, num_obs + N, N]`, `get_best_cut()`, `best_spins`, `score`, `matrix`.
The reference recomputes all local fields with a batched matmul and clones the state every step;
here one kernel (csrc/peco.cu) flips, updates the resident fields from one matrix row and rewrites
the observables in place.  Supported configuration = what PECO trains with
(train_PECO.py:34-44): unbiased graphs, ExtraAction.NONE, reversible spins, infinite memory.
"""
from __future__ import annotations

import ctypes as C
from enum import Enum
from typing import List, Optional

import numpy as np
import torch as th

from .. import _lib
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


# ----------------------------------------------------------------------------- graph generators
class GraphGenerator:
    def __init__(self, n_spins, edge_type, biased=False, num_envs=None):
        self.n_spins, self.edge_type, self.biased, self.num_envs = n_spins, edge_type, biased, num_envs

    def _make_mask(self):
        """util_envs_PECO.py:21-38 / 67-84: the per-call edge-weight mask, same RNG calls as the reference."""
        n, dev = self.n_spins, self.device
        if self.edge_type == EdgeType.UNIFORM:
            return th.ones((n, n), device=dev)
        if self.edge_type == EdgeType.DISCRETE:
            mask = 2. * th.randint(0, 2, (n, n), device=dev) - 1.
            return th.tril(mask) + th.triu(mask.T, 1)
        if self.edge_type == EdgeType.RANDOM:
            mask = 2. * th.randint(0, 2, (self.num_envs, n, n), dtype=th.float32, device=dev) - 1
            return th.tril(mask, diagonal=0) + th.triu(mask.transpose(1, 2), diagonal=1)
        raise NotImplementedError()


class RandomERGraphGenerator(GraphGenerator):
    """util_envs_PECO.py:15-57 (torch ops in the reference's order: same seed -> same graphs)."""

    def __init__(self, n_spins=20, p_connection=0.2, edge_type=EdgeType.DISCRETE, num_envs=8, device="cuda"):
        super().__init__(n_spins, edge_type, False, num_envs)
        self.p_connection, self.device = p_connection, device

    def get(self, with_padding=False):
        n = self.n_spins
        adj = (th.rand(self.num_envs, n, n, device=self.device) < self.p_connection).float()
        adj = adj * (1 - th.eye(n, device=self.device).unsqueeze(0))
        adj = th.triu(adj, diagonal=1)
        adj = adj + adj.transpose(1, 2)
        return adj * self._make_mask()


class RandomBAGraphGenerator(GraphGenerator):
    """util_envs_PECO.py:60-113, including its quirk: the initial clique loop also sets the first
    m+1 diagonal entries (self loops), kept for parity."""

    def __init__(self, n_spins=20, m_insertion_edges=4, edge_type=EdgeType.DISCRETE, num_envs=8, device="cuda"):
        super().__init__(n_spins, edge_type, False, num_envs)
        self.m_insertion_edges, self.device = m_insertion_edges, device

    def get(self, with_padding=False):
        e, n, m = self.num_envs, self.n_spins, self.m_insertion_edges
        adj = th.zeros((e, n, n), device=self.device)
        for i in range(m + 1):
            adj[:, i, :i + 1] = 1
            adj[:, :i + 1, i] = 1
        for new_node in range(m + 1, n):
            degree = adj.sum(dim=-1)
            prob = degree / degree.sum(dim=-1, keepdim=True)
            chosen = th.multinomial(prob, num_samples=m, replacement=False)
            batch = th.arange(e, device=self.device).repeat_interleave(m)
            adj[batch, new_node, chosen.view(-1)] = 1
            adj[batch, chosen.view(-1), new_node] = 1
        return adj * self._make_mask()


class SetMatrixGenerator(GraphGenerator):
    """Hands out a fixed `[E, N, N]` tensor (replaying recorded graphs)."""

    def __init__(self, matrix: TEN):
        super().__init__(matrix.shape[-1], EdgeType.DISCRETE, False, matrix.shape[0])
        self.matrix, self.device = matrix, matrix.device

    def get(self, with_padding=False):
        return self.matrix.clone()


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


class SpinSystemUnbiased:
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
        if not reversible_spins or memory_length is not None or init_snap is not None:
            raise NotImplementedError("reversible spins with infinite memory only (train_PECO.py:40-44)")
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
        self.matrix = self.gg.get()          # the reference draws one graph batch in __init__ and another in reset()
        self.reset()
        self.best_score = self.score.clone()
        self.best_spins = self.state[:, 0, :].clone()

    # ------------------------------------------------------------------ kernels
    def _fields(self, spins: TEN, want_as: bool = False, want_cut: bool = False):
        e, n = spins.shape
        spins = spins.contiguous()
        fields = th.empty((e, n), dtype=th.float32, device=self.device)
        as_ = th.empty((e, n), dtype=th.float32, device=self.device) if want_as else None
        cut = th.empty((e,), dtype=th.float32, device=self.device) if want_cut else None
        with on_device(self.device):
            _lib.check(self._lib.rlsb_peco_fields(_ptr(self.matrix), _ptr(spins), e, n, _ptr(as_), _ptr(fields), _ptr(cut),
                                                  _stream_ptr(self.device)), "peco_fields")
        return fields, as_, cut

    def _get_immeditate_cuts_avaialable(self, spins: TEN, matrix: Optional[TEN] = None) -> TEN:
        return self._fields(spins)[0]

    # ------------------------------------------------------------------ reset (spinsystem_PECO.py:151-193)
    def reset(self, spins=None):
        self.current_step = 0
        self.matrix = self.gg.get().to(self.device).to(th.float32).contiguous()
        self.matrix_obs = self.matrix
        spins_one = th.ones(self.num_envs, self.n_spins, device=self.device)
        local_rewards_available = self._get_immeditate_cuts_avaialable(spins_one)
        if th.any(th.eq(th.sum(th.abs(local_rewards_available), dim=-1), 0)):
            self.reset()                      # an empty graph was generated: try again (as the reference does)
        else:
            self.max_local_reward_available_ = th.max(local_rewards_available, dim=-1).values
            if (self.max_local_reward_available_ == 0).any():
                self.reset()
            self.max_local_reward_available = self.max_local_reward_available_.unsqueeze(1).expand(-1, self.n_spins)
        self.state = self._reset_state(spins)
        self.score = self.calculate_score()
        self.best_score = self.score.clone()
        self.best_obs_score = self.score.clone()
        self.best_spins = self.state[:, 0, :self.n_spins].clone()
        self.best_obs_spins = self.state[:, 0, :self.n_spins].clone()
        self._history = None
        self._hist_len = 0
        if self.stag_punishment is not None or self.basin_reward is not None:
            words = (self.n_spins + 31) // 32
            self._history = th.empty((self.max_steps + 1, self.num_envs, words), dtype=th.int32, device=self.device)
        return self.get_observation()

    def _reset_state(self, spins=None):
        state = th.zeros(self.num_envs, len(self.observables), self.n_actions, device=self.device)
        if spins is None:
            state[:, 0, :self.n_spins] = 2 * th.randint(0, 2, (self.num_envs, self.n_spins,), device=self.device,
                                                        dtype=th.float) - 1
        else:
            state[:, 0, :] = spins.to(self.device).to(th.float32)
        imm, self._as, _ = self._fields(state[:, 0, :self.n_spins], want_as=True)
        for idx, obs in self.observables:
            if obs == Observable.IMMEDIATE_REWARD_AVAILABLE:
                state[:, idx, :self.n_spins] = imm / self.max_local_reward_available
            elif obs == Observable.NUMBER_OF_GREEDY_ACTIONS_AVAILABLE:
                cnt = th.sum(imm <= 0, dim=-1).float()
                frac = cnt / self.n_spins if self.scalar_div_as_cuda else cnt / th.full_like(cnt, self.n_spins)
                state[:, idx, :self.n_spins] = (1 - frac).unsqueeze(-1)
        return state

    def calculate_cut(self, spins=None):
        if spins is None:
            spins = self.state[:, 0, :self.n_spins]
        return self._fields(spins, want_cut=True)[2]

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
        # TERMINATION_IMMANENCY: max(0, float32((step - max_steps) / horizon) + 1) as the reference computes it
        term = np.float32(max(np.float32(0.), np.float32((self.current_step - self.max_steps) / self.horizon_length)
                              + np.float32(1.)))
        use_stag, use_basin = self.stag_punishment is not None, self.basin_reward is not None
        with on_device(self.device):
            _lib.check(self._lib.rlsb_peco_step(
                _ptr(self.matrix), _ptr(self.state), _ptr(self._as), _ptr(action), _ptr(self.score), _ptr(self.best_score),
                _ptr(self.best_spins), _ptr(self.max_local_reward_available_), _ptr(rew), _ptr(self._history),
                self._hist_len, _ptr(self._bad), self.num_envs, self.n_spins, len(self.observables),
                self._rows.ctypes.data, self.reward_signal.value, int(bool(self.norm_rewards)),
                float(np.float32(1. / self.max_steps)), float(term), int(use_stag),
                float(self.stag_punishment or 0.0), int(use_basin), float(self.basin_reward or 0.0),
                int(bool(self.scalar_div_as_cuda)), _stream_ptr(self.device)), "peco_step")
        if self._history is not None:
            self._hist_len += 1
        self.best_obs_score = self.best_score           # infinite memory (spinsystem_PECO.py:427-429)
        self.best_obs_spins = self.best_spins
        done = th.full((self.num_envs,), self.current_step == self.max_steps, device=self.device, dtype=th.bool)
        obs = self.get_observation() if return_observation else None
        return obs, rew, done

    def get_observation(self):
        state = self.state.clone()
        if self.spin_basis == SpinBasis.BINARY:
            state[:, 0, :] = (1 - state[:, 0, :]) / 2
        return th.cat((state, self.matrix_obs), dim=-2)

    def num_bad_actions(self) -> int:
        return int(self._bad.item())
