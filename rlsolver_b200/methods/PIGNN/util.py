"""`hamiltonian_maxcut` / `eval_maxcut` of rlsolver/methods/PIGNN/util.py:4-19 on the relaxed-cut kernel.

The reference takes `edge_index [2, M]` and `pred [N]` per call.  The graph store of an edge_index is looked up by
CONTENT (a 128-bit digest of the edge list bytes, plus the node count and device): the reference's loop feeds a
fresh `batch.edge_index` tensor of the same shape every step, so neither the storage address nor the shape
identifies a graph.  A small LRU keeps the most recent stores; evicted ones release their device blobs.
A caller that owns a store can skip the lookup with `hamiltonian_maxcut(edge_index, pred, store=...)`."""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch as th

from ...graph_store import GraphStore
from ...relaxed import relaxed_cut

STORE_CACHE_SIZE = 16
_STORES: "OrderedDict[tuple, GraphStore]" = OrderedDict()


def store_for(edge_index: th.Tensor, num_nodes: int) -> GraphStore:
    """Graph store for the edge list `edge_index [2, M]` (content-addressed, LRU of STORE_CACHE_SIZE)."""
    ij = np.ascontiguousarray(edge_index.detach().cpu().numpy().astype(np.int64, copy=False))
    key = (hashlib.blake2b(ij.tobytes(), digest_size=16).digest(), tuple(ij.shape), int(num_nodes),
           str(edge_index.device))
    st = _STORES.get(key)
    if st is not None:
        _STORES.move_to_end(key)
        return st
    graph = np.stack([ij[0], ij[1], np.ones_like(ij[0])], axis=1)
    st = GraphStore(graph, False, device=edge_index.device, num_nodes=num_nodes)
    _STORES[key] = st
    while len(_STORES) > STORE_CACHE_SIZE:
        _STORES.popitem(last=False)          # the store's __del__ frees the device blob
    return st


def hamiltonian_maxcut(edge_index, pred, store: Optional[GraphStore] = None):
    """sum(2 p_i p_j - p_i - p_j) over the edges = the relaxed objective of one environment."""
    flat = pred.reshape(1, -1).float()
    st = store if store is not None else store_for(edge_index, flat.shape[1])
    return relaxed_cut(st, flat)[0]


def eval_maxcut(edge_index, pred, d, n, store: Optional[GraphStore] = None):
    maxcut_energy = -hamiltonian_maxcut(edge_index, pred, store)
    P = 0.7632
    cut_ub = (d / 4 + (P * math.sqrt(d / 4))) * n
    return maxcut_energy, maxcut_energy / cut_ub
