"""Mirror of rlsolver/methods/ECO_S2V/src/envs/inference_network_env.py (SpinSystemFactory / SpinSystemUnbiased): the
PECO pattern-I environment as it is used at INFERENCE time -- ONE graph shared by `num_envs` parallel searches
(`SetGraphGenerator(matrix [N, N])`, train_and_inference/inference_PECO.py:88-101), driven by
`peco_test_network` (ECO_S2V/util.py:20-62).  Differences from the training env (spinsystem_PECO.py), all kept:

* `matrix` is `[N, N]`; the observation stacks it under every env's state (inference_network_env.py:446-456);
* reset seeds `best_score` / `best_spins` of EVERY env from the batch's best start (171-207: `torch.max(score, 0)`,
  the winner's spins expanded) -- `best_score` is 0-dim until the first step turns it into `[E]` (356-360);
* `step` returns `(observation, done)`: no reward is computed (295-444).

The step itself is the same arithmetic, so it runs on the same kernels (csrc/peco_compact.cu for {-1, 0, 1} weights,
csrc/peco.cu otherwise) with the one graph replicated per env in the resident layout (bit rows: N * ceil(N / 32) * 4
bytes per env).

`use_tensor_core=True` (143-145, 212-236: matrix, spins and `state` in float16).  The resident state here is integers
and bits either way; what the mode changes is the rounding of what is MATERIALISED: `state` / the observation come out
as float16 with torch's half arithmetic reproduced (every operation evaluated in float32 and rounded to half once; the
time observables accumulated in half), for graphs with {-1, 0, 1} weights and at most 1024 edges, where every integer the
reference holds in a half (scores, sums of the matrix) stays below 2048 and is exact.  `score` / `best_score` /
`best_spins` stay float32 tensors holding the same values.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch as th

from .env_PECO import (ECO_PECO_OBSERVABLES, CompactGraphs, EdgeType, ExtraAction, GraphGenerator, Observable,  # noqa: F401
                       OptimisationTarget, RewardSignal, SpinBasis, pack_rows)
from .env_PECO import SpinSystemUnbiased as _BatchedSpinSystem

TEN = th.Tensor


class SetGraphGenerator(GraphGenerator):
    """util_envs_PECO.py:139-170 for one unbiased graph: `get()` hands out the `[N, N]` tensor."""

    def __init__(self, matrices: TEN, biases=None, ordered=False, device="cuda"):
        if biases is not None:
            raise NotImplementedError("biased spin systems are outside the max-cut hot path")
        if matrices.dim() != 2 or matrices.shape[0] != matrices.shape[1]:
            raise NotImplementedError("one [N, N] graph (inference_PECO.py:88)")
        m = matrices.to(device)
        if bool(th.isin(m, th.tensor([0, 1], device=m.device)).all()):
            edge_type = EdgeType.UNIFORM
        elif bool(th.isin(m, th.tensor([0, -1, 1], device=m.device)).all()):
            edge_type = EdgeType.DISCRETE
        else:
            edge_type = EdgeType.RANDOM
        super().__init__(m.shape[0], edge_type, False)
        self.graphs, self.device = m, m.device

    def get(self, with_padding=False):
        return self.graphs


class SpinSystemFactory(object):
    @staticmethod
    def get(graph_generator=None, max_steps=20, observables=ECO_PECO_OBSERVABLES, reward_signal=RewardSignal.DENSE,
            extra_action=ExtraAction.PASS, optimisation_target=OptimisationTarget.ENERGY, spin_basis=SpinBasis.SIGNED,
            norm_rewards=False, memory_length=None, horizon_length=None, stag_punishment=None, basin_reward=None,
            reversible_spins=True, init_snap=None, seed=None, device=None, num_envs=None, if_greedy=False,
            use_tensor_core=False):
        if graph_generator.biased:
            raise NotImplementedError("biased spin systems are outside the max-cut hot path")
        return SpinSystemUnbiased(graph_generator, max_steps, observables, reward_signal, extra_action,
                                  optimisation_target, spin_basis, norm_rewards, memory_length, horizon_length,
                                  stag_punishment, basin_reward, reversible_spins, init_snap, seed, device, num_envs,
                                  use_tensor_core)


class SpinSystemUnbiased(_BatchedSpinSystem):
    def __init__(self, graph_generator=None, max_steps=20, observables=ECO_PECO_OBSERVABLES,
                 reward_signal=RewardSignal.DENSE, extra_action=ExtraAction.PASS,
                 optimisation_target=OptimisationTarget.ENERGY, spin_basis=SpinBasis.SIGNED, norm_rewards=False,
                 memory_length=None, horizon_length=None, stag_punishment=None, basin_reward=None,
                 reversible_spins=False, init_snap=None, seed=None, device=None, num_envs=None, use_tensor_core=False):
        self.use_tensor_core = bool(use_tensor_core)
        # the inference step never looks at its reward: SINGLE (rejected by the training mirror) is as good as any
        signal = RewardSignal.DENSE if reward_signal == RewardSignal.SINGLE else reward_signal
        # the reference's inference step has no history buffer: stag / basin settings are accepted and ignored (295-444)
        super().__init__(graph_generator, max_steps, observables, signal, extra_action, optimisation_target, spin_basis,
                         norm_rewards, memory_length, horizon_length, None, None, reversible_spins, init_snap, seed,
                         device, num_envs)
        self.reward_signal = reward_signal
        self.stag_punishment, self.basin_reward = stag_punishment, basin_reward

    # ------------------------------------------------------------------ one graph, replicated in the resident layout
    def _draw_graphs(self) -> None:
        m = self.gg.get().to(self.device).to(th.float32).contiguous()
        self._shared_matrix = m
        self._compact = None
        self._matrix_cache = None
        if self.allow_compact:
            one = CompactGraphs.from_dense(m.unsqueeze(0))
            if one is not None:
                adj = one.adj.expand(self.num_envs, -1, -1).contiguous()
                sgn = None if one.sgn is None else one.sgn[0].contiguous()        # one sign matrix shared by every env
                self._compact = CompactGraphs(adj, sgn, self.n_spins)
        if self._compact is None:
            self._dense_batch = m.unsqueeze(0).expand(self.num_envs, -1, -1).contiguous()
        if self.use_tensor_core and (self._compact is None or int((m != 0).sum()) > 2048):
            raise NotImplementedError("use_tensor_core=True is reproduced for graphs with weights in {-1, 0, 1} and at most "
                                      "1024 edges (beyond that the reference's float16 sums are no longer exact integers)")

    @property
    def matrix(self) -> TEN:
        """The kernels' per-env view [E, N, N] (dense layout only); `matrix_obs` is the reference's [N, N]."""
        return self._dense_batch if self._compact is None else self._shared_matrix

    @property
    def matrix_obs(self) -> TEN:
        return self._shared_matrix

    def mean_degree(self) -> float:
        return float((self._shared_matrix != 0).float().sum() / self.n_spins)

    # ------------------------------------------------------------------ reset: best_* from the batch's best start
    def _seed_best(self) -> None:
        best, idx = th.max(self.score, dim=0)
        self._best_is_scalar = True
        self._best_scalar = best
        self.best_score = best.expand(self.num_envs).contiguous()          # what the kernels read and update
        if self._compact is not None:
            self._best_words = self._spins[idx].unsqueeze(0).expand(self.num_envs, -1).contiguous()
        else:
            self._dense_best_spins = self._dense_state[idx, 0, :self.n_spins].unsqueeze(0) \
                .expand(self.num_envs, -1).contiguous()

    def get_best_cut(self):
        """0-dim before the first step (the global best of the start states), [E] afterwards -- as in the reference."""
        return self._best_scalar if self._best_is_scalar else self.best_score

    def step(self, action: TEN, return_observation: bool = True):
        obs, _, done = super().step(action, return_observation)
        self._best_is_scalar = False
        return obs, done

    # ------------------------------------------------------------------ float16 materialisation (use_tensor_core)
    def _half_table(self) -> TEN:
        """k-fold `state[row] += 1 / max_steps` on a half tensor: float32 add, rounded to half each time.  The Python
        scalar enters the add as float32 on CUDA (the kernel's opmath type) and rounded to half first on the CPU -- like
        `x / n_spins`, `scalar_div_as_cuda` selects whose arithmetic is reproduced."""
        key = bool(self.scalar_div_as_cuda)
        cached = getattr(self, "_table16", None)
        if cached is None or cached[0] != key:
            inv = np.float32(1. / self.max_steps) if key else np.float32(np.float16(1. / self.max_steps))
            tab = np.zeros(self.max_steps + 2, np.float32)
            for k in range(1, tab.size):
                tab[k] = np.float32(np.float16(np.float32(tab[k - 1]) + inv))
            cached = self._table16 = (key, th.from_numpy(tab).to(self.device))
        return cached[1]

    def _expand_state_half(self, out: TEN, env_stride: int, binary: bool) -> None:
        from .. import _lib
        from ..graph_store import on_device
        with on_device(self.device):
            _lib.check(self._lib.rlsb_peco_compact_expand_state_half(
                self._spins.data_ptr(), self._best_words.data_ptr(), self._cfields.data_ptr(), self._last_flip.data_ptr(),
                self.score.data_ptr(), self.best_score.data_ptr(), self.max_local_reward_available_.data_ptr(),
                self._half_table().data_ptr(), out.data_ptr(), int(env_stride), self.num_envs, self.n_spins,
                len(self.observables), self._rows.ctypes.data, self.current_step, int(binary), self._termination(),
                int(bool(self.scalar_div_as_cuda)), int(self.current_step == 0),
                th.cuda.current_stream(self.device).cuda_stream), "peco_compact_expand_state_half")

    @property
    def state(self) -> TEN:
        if not self.use_tensor_core:
            return _BatchedSpinSystem.state.fget(self)
        out = th.empty((self.num_envs, len(self.observables), self.n_spins), dtype=th.float16, device=self.device)
        self._expand_state_half(out, len(self.observables) * self.n_spins, False)
        return out

    def get_observation(self):
        n, k = self.n_spins, len(self.observables)
        if self.use_tensor_core:
            obs = th.empty((self.num_envs, k + n, n), dtype=th.float16, device=self.device)
            self._expand_state_half(obs, (k + n) * n, self.spin_basis == SpinBasis.BINARY)
            obs[:, k:, :] = self._shared_matrix.to(th.float16)
            return obs
        if self._compact is None:
            state = self._dense_state.clone()
            if self.spin_basis == SpinBasis.BINARY:
                state[:, 0, :] = (1 - state[:, 0, :]) / 2
            return th.cat((state, self._shared_matrix.unsqueeze(0).expand(state.shape[0], -1, -1)), dim=-2)
        obs = th.empty((self.num_envs, k + n, n), dtype=th.float32, device=self.device)
        self._expand_state(obs, (k + n) * n, self.spin_basis == SpinBasis.BINARY)
        obs[:, k:, :] = self._shared_matrix
        return obs


__all__ = ["SetGraphGenerator", "SpinSystemFactory", "SpinSystemUnbiased"]
