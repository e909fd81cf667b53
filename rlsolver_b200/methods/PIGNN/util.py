"""`hamiltonian_maxcut` / `eval_maxcut` of rlsolver/methods/PIGNN/util.py:4-19 on the relaxed-cut kernel.

The reference takes `edge_index [2, M]` and `pred [N]` per call; the graph store for an edge_index tensor is built
once and cached on the tensor's storage address."""
from __future__ import annotations

import math

import torch as th

from ...graph_store import GraphStore
from ...relaxed import relaxed_cut

_STORES = {}


def _store_for(edge_index: th.Tensor, num_nodes: int) -> GraphStore:
    key = (edge_index.data_ptr(), tuple(edge_index.shape), num_nodes, str(edge_index.device))
    st = _STORES.get(key)
    if st is None:
        ij = edge_index.detach().cpu().numpy()
        graph = [(int(a), int(b), 1) for a, b in zip(ij[0], ij[1])]
        st = _STORES[key] = GraphStore(graph, False, device=edge_index.device, num_nodes=num_nodes)
    return st


def hamiltonian_maxcut(edge_index, pred):
    """sum(2 p_i p_j - p_i - p_j) over the edges = the relaxed objective of one environment."""
    flat = pred.reshape(1, -1).float()
    return relaxed_cut(_store_for(edge_index, flat.shape[1]), flat)[0]


def eval_maxcut(edge_index, pred, d, n):
    maxcut_energy = -hamiltonian_maxcut(edge_index, pred)
    P = 0.7632
    cut_ub = (d / 4 + (P * math.sqrt(d / 4))) * n
    return maxcut_energy, maxcut_energy / cut_ub
