"""NumPy (float32) restatement of the reference's PECO pattern-I environment.  TEST INFRASTRUCTURE ONLY.
SpinSystemUnbiased of rlsolver/methods/ECO_S2V/src/envs/spinsystem_PECO.py: _reset_state (212-236),
step (306-486), calculate_cut (601-607), _get_immeditate_cuts_avaialable (660-662); HistoryBuffer.update of
util_envs_PECO.py:262-288.  Configuration subset: unbiased, ExtraAction.NONE, infinite memory, reward signals
DENSE / BLS / CUSTOM_BLS.  Observables are named as in util_envs.py:40-51.
`reversible_spins=False` is the S2V-DQN pattern of the single-env NumPy environment
(rlsolver/methods/ECO_S2V/src/envs/spinsystem.py: every spin starts at +1, :242-247; an episode also ends when no +1
spin is left, :476-480; train_S2V.py:37-47 pairs it with the DENSE reward, norm_rewards and the single SPIN_STATE
observable), applied per env of the batch; pinned by tests/golden/s2v_*.npz (tools/make_goldens_s2v.py).
`seed_best_from_batch=True` is the inference twin (inference_network_env.py: one graph for all envs, best_* seeded
from the batch's best start, no reward); pinned by tests/golden/pecoinf_*.npz (tools/make_goldens_inference.py)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

f32 = np.float32
SPIN, IMM, TSF, EPT, TERM, GREEDY, DSCORE, DSTATE = range(1, 9)       # Observable enum values
DENSE, BLS, CUSTOM_BLS = 1, 2, 4


def fields(matrix: np.ndarray, spins: np.ndarray) -> np.ndarray:
    """matmul(matrix, spins) * spins per env (660-662)."""
    return (np.einsum("ejk,ek->ej", matrix.astype(f32), spins.astype(f32)).astype(f32) * spins).astype(f32)


def cut(matrix: np.ndarray, spins: np.ndarray) -> np.ndarray:
    """(1/4) * sum(matmul(A, s) * -s) + (1/4) * sum(A) (601-607)."""
    t = (np.einsum("ejk,ek->ej", matrix.astype(f32), spins.astype(f32)).astype(f32) * -spins).sum(axis=-1, dtype=f32)
    return (f32(0.25) * t + f32(0.25) * matrix.sum(axis=(-1, -2), dtype=f32)).astype(f32)


class SpinSystem:
    def __init__(self, matrix, spins, observables: List[int], max_steps: int, reward_signal: int, norm_rewards: bool,
                 horizon_length: Optional[int] = None, stag_punishment=None, basin_reward=None,
                 scalar_div_as_cuda: bool = False, reversible_spins: bool = True, seed_best_from_batch: bool = False):
        self.reversible = reversible_spins
        self.m = matrix.astype(f32)
        self.e, self.n = spins.shape
        self.obs = list(enumerate(observables))
        self.max_steps, self.reward_signal, self.norm = max_steps, reward_signal, norm_rewards
        self.horizon = horizon_length if horizon_length is not None else max_steps
        self.stag, self.basin = stag_punishment, basin_reward
        # x / n_spins: torch CPU divides; torch CUDA multiplies by float32(1/n) (div_true_kernel_cuda)
        self.div_n = (lambda x: (x * (f32(1) / f32(self.n))).astype(f32)) if scalar_div_as_cuda else \
            (lambda x: (x / f32(self.n)).astype(f32))
        self.current_step = 0
        ones = np.ones((self.e, self.n), f32)
        self.max_local = fields(self.m, ones).max(axis=-1)
        self.state = np.zeros((self.e, len(observables), self.n), f32)
        self.state[:, 0, :] = spins
        imm = fields(self.m, spins)
        for idx, o in self.obs:
            if o == IMM:
                self.state[:, idx, :] = imm / self.max_local[:, None]
            elif o == GREEDY:
                self.state[:, idx, :] = (f32(1) - self.div_n((imm <= 0).sum(axis=-1).astype(f32)))[:, None]
        self.score = cut(self.m, spins)
        self.best_score = self.score.copy()
        self.best_spins = spins.astype(f32).copy()
        if seed_best_from_batch:
            # the inference env (inference_network_env.py:171-207): every env starts from the batch's best start state
            # (torch.max(score, dim=0): the first index among equal maxima on the CPU and the CUDA kernel alike)
            idx = int(np.argmax(self.score))
            self.best_score = np.full(self.e, self.score[idx], f32)
            self.best_spins = np.repeat(spins[idx:idx + 1].astype(f32), self.e, axis=0)
        self.history: List[np.ndarray] = []

    def step(self, action: np.ndarray):
        self.current_step += 1
        rows = np.arange(self.e)
        st = self.state
        st[rows, 0, action] = -st[rows, 0, action]
        spins = st[:, 0, :]
        imm = fields(self.m, spins)
        delta = -imm[rows, action]
        self.score = (self.score + delta).astype(f32)
        improvement = (self.score - self.best_score).astype(f32)          # best_obs_score == best_score (infinite memory)
        if self.reward_signal == BLS:
            rew = np.where(improvement > 0, improvement, f32(0)).astype(f32)
        elif self.reward_signal == CUSTOM_BLS:
            rew = np.where(improvement > 0, improvement / (improvement + f32(0.1)), f32(0)).astype(f32)
        else:
            rew = delta.astype(f32)
        if self.norm:
            rew = self.div_n(rew)
        if self.stag is not None or self.basin is not None:
            key = [tuple(r) for r in (spins > 0)]
            if not self.history:
                fresh = np.ones(self.e, bool)
            else:
                fresh = np.asarray([all(key[e] != h[e] for h in self.history) for e in range(self.e)])
            self.history.append(key)
            if self.stag is not None:
                rew[~fresh] = (rew[~fresh] - f32(self.stag)).astype(f32)
            if self.basin is not None:
                mask = (imm <= 0).all(axis=-1) & fresh
                rew[mask] = (rew[mask] + f32(self.basin)).astype(f32)
        better = self.score > self.best_score
        self.best_score = np.where(better, self.score, self.best_score).astype(f32)
        self.best_spins = np.where(better[:, None], spins, self.best_spins).astype(f32)
        for idx, o in self.obs:
            if o == IMM:
                st[:, idx, :] = imm / self.max_local[:, None]
            elif o == TSF:
                st[:, idx, :] = (st[:, idx, :] + f32(1. / self.max_steps)).astype(f32)
                st[rows, idx, action] = 0
            elif o == EPT:
                st[:, idx, :] = (st[:, idx, :] + f32(1. / self.max_steps)).astype(f32)
            elif o == TERM:
                st[:, idx, :] = max(f32(0), f32((self.current_step - self.max_steps) / self.horizon) + f32(1))
            elif o == GREEDY:
                st[:, idx, :] = (f32(1) - self.div_n((imm <= 0).sum(axis=-1).astype(f32)))[:, None]
            elif o == DSCORE:
                st[:, idx, :] = (np.abs(self.score - self.best_score) / self.max_local)[:, None]
            elif o == DSTATE:
                st[:, idx, :] = np.count_nonzero(self.best_spins - spins, axis=-1).astype(f32)[:, None]
        done = np.full(self.e, self.current_step == self.max_steps)
        if not self.reversible:
            done = done | ~(spins > 0).any(axis=-1)              # no spin left to flip (spinsystem.py:476-480)
        return rew, done

    def observation(self, binary: bool) -> np.ndarray:
        state = self.state.copy()
        if binary:
            state[:, 0, :] = (1 - state[:, 0, :]) / 2
        return np.concatenate([state, self.m], axis=-2)


def half_state(env: "SpinSystem", scalar_as_cuda: bool = False) -> np.ndarray:
    """float16 `state` of the inference env's use_tensor_core mode (inference_network_env.py:143-145, 212-236) derived
    from the float32 restatement: torch evaluates a half operation in float32 and rounds once, so every observable is
    the float32 value rounded to half, except the greedy row (count / n, then 1 - x: two roundings) and the time rows
    (k-fold half accumulation of 1 / max_steps; the Python scalar is rounded to half first on the CPU, kept float32 on
    CUDA).  Valid while every integer involved is below 2048.  Pinned by tests/golden/pecoinfhalf_*.npz."""
    f16 = np.float16
    e, n = env.e, env.n
    spins = env.state[:, 0, :]
    inv = f32(1. / env.max_steps) if scalar_as_cuda else f32(f16(1. / env.max_steps))
    tab = np.zeros(env.max_steps + 2, f32)
    for k in range(1, tab.size):
        tab[k] = f32(f16(f32(tab[k - 1]) + inv))
    imm = fields(env.m, spins)
    cnt = (imm <= 0).sum(axis=-1).astype(f32)
    at_reset = env.current_step == 0
    out = np.zeros((e, len(env.obs), n), f16)
    for idx, o in env.obs:
        if o == SPIN:
            out[:, idx, :] = spins.astype(f16)
        elif o == IMM:
            out[:, idx, :] = (imm / env.max_local[:, None]).astype(f32).astype(f16)
        elif o == TSF:
            out[:, idx, :] = tab[np.rint(env.state[:, idx, :] * env.max_steps).astype(int)].astype(f16)
        elif o == EPT:
            out[:, idx, :] = f16(tab[env.current_step])
        elif o == TERM:
            out[:, idx, :] = f16(env.state[0, idx, 0])
        elif o == GREEDY:
            x = env.div_n(cnt).astype(f16)
            out[:, idx, :] = (f32(1) - x.astype(f32)).astype(f16)[:, None]
        elif o == DSCORE:
            out[:, idx, :] = 0 if at_reset else (np.abs(env.score - env.best_score) / env.max_local).astype(f32).astype(f16)[:, None]
        elif o == DSTATE:
            out[:, idx, :] = 0 if at_reset else np.count_nonzero(env.best_spins - spins, axis=-1).astype(f16)[:, None]
    return out


# --------------------------------------------------------------------------- generators (torch restatement)
def torch_edge_mask(edge_type: str, n: int, num_envs: int, device):
    """util_envs_PECO.py:21-38 / 67-84, op for op: the edge-weight mask drawn per get() call.
    edge_type in {"UNIFORM", "DISCRETE", "RANDOM"}.  TEST INFRASTRUCTURE: run on the same device and seed as the
    generator kernels (csrc/peco_compact.cu) to check "same seed, same graphs"."""
    import torch as th
    if edge_type == "UNIFORM":
        return th.ones((n, n), device=device)
    if edge_type == "DISCRETE":
        mask = 2. * th.randint(0, 2, (n, n), device=device) - 1.
        return th.tril(mask) + th.triu(mask.T, 1)
    mask = 2. * th.randint(0, 2, (num_envs, n, n), dtype=th.float32, device=device) - 1
    return th.tril(mask, diagonal=0) + th.triu(mask.transpose(1, 2), diagonal=1)


def torch_er_graphs(num_envs: int, n: int, p: float, edge_type: str, device):
    """RandomERGraphGenerator.get (util_envs_PECO.py:40-52) with the reference's torch ops."""
    import torch as th
    adj = (th.rand(num_envs, n, n, device=device) < p).float()
    adj = adj * (1 - th.eye(n, device=device).unsqueeze(0))
    adj = th.triu(adj, diagonal=1)
    adj = adj + adj.transpose(1, 2)
    return adj * torch_edge_mask(edge_type, n, num_envs, device)


def torch_ba_graphs(num_envs: int, n: int, m: int, edge_type: str, device):
    """RandomBAGraphGenerator.get (util_envs_PECO.py:87-107) with the reference's torch ops, including the self loops
    its initial clique loop sets."""
    import torch as th
    adj = th.zeros((num_envs, n, n), device=device)
    for i in range(m + 1):
        adj[:, i, :i + 1] = 1
        adj[:, :i + 1, i] = 1
    for new_node in range(m + 1, n):
        degree = adj.sum(dim=-1)
        prob = degree / degree.sum(dim=-1, keepdim=True)
        chosen = th.multinomial(prob, num_samples=m, replacement=False)
        batch = th.arange(num_envs, device=device).repeat_interleave(m)
        adj[batch, new_node, chosen.view(-1)] = 1
        adj[batch, chosen.view(-1), new_node] = 1
    return adj * torch_edge_mask(edge_type, n, num_envs, device)
