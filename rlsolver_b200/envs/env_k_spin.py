"""Drop-in for `rlsolver.envs.env_k_spin.SimulatorMaxcut` (rlsolver/envs/env_k_spin.py:62-216): the relaxed
("probability") max-cut simulator behind the k-spin / gradient methods.

`get_objectives(probs)` (191-193) and `get_objectives_using_for_loop` (164-189, the same sum node by node) run in the
edge-streaming kernel of csrc/relaxed.cu and are differentiable (autograd gets the hand-written backward);
`get_scores(bool)` (195-197) is the integer cut of the packed-spin kernel.  `graph_tuple = (mygraph, num_nodes,
num_edges)` as in the reference; `graph_name` loads through `load_mygraph2`.  CUDA only.
"""
from __future__ import annotations

import numpy as np
import torch as th

from ..graph_store import GraphStore, require_cuda
from ..methods.util_read_data import load_mygraph2
from ..relaxed import relaxed_cut

TEN = th.Tensor


class SimulatorMaxcut:
    def __init__(self, graph_name: str = 'powerlaw_64', gpu_id: int = 0, graph_tuple=None):
        self.device = require_cuda(th.device(f"cuda:{gpu_id}" if gpu_id >= 0 else "cpu"))
        self.int_type = th.int32
        if graph_tuple:
            graph, num_nodes, num_edges = graph_tuple
        else:
            graph = load_mygraph2(graph_name=graph_name)
            num_nodes = len({a for a, _, _ in graph} | {b for _, b, _ in graph})
            num_edges = len(graph)
        self.store = GraphStore(graph, False, device=self.device)
        assert num_nodes == self.store.num_nodes
        assert num_edges == self.store.num_edges
        self.num_nodes, self.num_edges = self.store.num_nodes, self.store.num_edges
        # the reference's index tensors: edges grouped by their first node, each group in the order the graph lists
        # them (env_k_spin.py:133-160)
        arr = np.asarray([(a, b) for a, b, _ in graph], dtype=np.int64).reshape(-1, 2)
        order = np.argsort(arr[:, 0], kind="stable")
        self.n0_ids = th.from_numpy(arr[order, 0].astype(np.int32)).to(self.device).unsqueeze(0)
        self.n1_ids = th.from_numpy(arr[order, 1].astype(np.int32)).to(self.device).unsqueeze(0)
        self.env_is = th.zeros(self.num_edges, dtype=th.int32, device=self.device).unsqueeze(0)

    def get_objectives(self, probs: TEN) -> TEN:
        return relaxed_cut(self.store, probs)

    def get_objectives_using_for_loop(self, probs: TEN) -> TEN:
        assert probs.shape[-1] == self.num_nodes
        return relaxed_cut(self.store, probs)

    def get_scores(self, probs: TEN) -> TEN:
        return self.store.cut_eval(probs)

    def get_rand_probs(self, num_envs: int) -> TEN:
        return th.rand((num_envs, self.num_nodes), dtype=th.float32, device=self.device)

    @staticmethod
    def prob_to_bool(p0s, thresh=0.5):
        return p0s > thresh


__all__ = ["SimulatorMaxcut"]
