"""Drop-in for `rlsolver.envs.env_PPO.EnvMaxcut` (rlsolver/envs/env_PPO.py:63-126): the
pattern-I gym-style max-cut environment used by rlsolver/methods/PPO.py.

Same constructor (`args` with num_nodes / num_envs / num_steps), attributes and return values:
`reset() -> xs float32 [E, N]`; `step(action int64 [E]) -> (xs, reward, next_done, cur_reward)`
where `xs` is the SAME tensor object mutated in place.  The reference flips with a Python loop
over the envs and re-evaluates every edge of every env; here one kernel flips and returns the
single-flip gain in O(degree) (csrc/fields.cu).  CUDA only.
"""
from __future__ import annotations

import torch as th

from ..graph_store import GraphStore, require_cuda
from ..methods.config import MyGraph

TEN = th.Tensor


class EnvMaxcut:
    def __init__(self, args, mygraph: MyGraph = (), device=th.device('cpu'), if_bidirectional: bool = False):
        self.device = require_cuda(device)
        self.int_type = th.long
        self.if_bidirectional = if_bidirectional
        self.num_nodes = args.num_nodes
        self.num_envs = args.num_envs
        self.xs = None
        self.action_count = 0
        self.last_reward = None
        self.num_steps = args.num_steps

        self.store = GraphStore(mygraph, if_bidirectional, device=self.device)
        if self.store.num_nodes != self.num_nodes:
            # the reference indexes a [E, args.num_nodes] tensor with the graph's node ids
            raise IndexError(f"graph has {self.store.num_nodes} distinct nodes, args.num_nodes is {self.num_nodes}")
        self.num_edges = self.store.num_edges
        self._bad = th.zeros((1,), dtype=th.int32, device=self.device)

    def reset(self):
        self.xs = self.generate_xs_randomly(num_sims=self.num_envs).to(th.float)
        self.last_reward = self.calculate_obj_values().to(th.float)
        return self.xs

    def step(self, action: TEN):
        self.action_count += 1
        if action.dtype != th.int64 or action.device != self.device or action.shape != (self.num_envs,):
            action = action.to(device=self.device, dtype=th.int64).reshape(self.num_envs)
        reward = th.empty((self.num_envs,), dtype=th.float, device=self.device)
        cur_reward = self.last_reward.clone()            # the reference rebinds last_reward every step
        self.store.step_flip(self.xs, action.contiguous(), reward, cur_reward, self._bad)
        self.last_reward = cur_reward
        if self.action_count == self.num_steps:
            self.action_count = 0
            next_done = th.ones([self.num_envs], dtype=th.float, device=self.device)
        else:
            next_done = th.zeros([self.num_envs], dtype=th.float, device=self.device)
        return self.xs, reward, next_done, cur_reward

    def num_bad_actions(self) -> int:
        """Actions outside [0, N) seen so far (the reference raises IndexError; syncs the device)."""
        return int(self._bad.item())

    def calculate_obj_values(self, if_sum: bool = True) -> TEN:
        xs = self.xs > 0
        if if_sum:
            return self.store.cut_eval(xs)
        values = self.store.cut_edges(xs)
        return values.to(th.long) // 2 if self.if_bidirectional else values

    def generate_xs_randomly(self, num_sims):
        xs = th.randint(0, 2, size=(num_sims, self.num_nodes), dtype=th.bool, device=self.device)
        xs[:, 0] = 0
        return xs
