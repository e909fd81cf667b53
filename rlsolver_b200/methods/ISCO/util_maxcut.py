"""`load_data` / `record` of the ISCO max-cut drivers (rlsolver/methods/ISCO/util_maxcut.py:7-62)."""
from __future__ import annotations

from typing import Dict, List

import torch as th

from . import config_maxcut as cfg


def load_data(filename: str, device=None) -> Dict:
    """Graph file -> {'num_nodes', 'num_edges', 'edge_from', 'edge_to', 'adj_matrix'}.

    Same parse as the reference (lines with '//' skipped, 1-based `u v w`), same edge order
    (networkx.Graph.edges(): by first endpoint, neighbours in insertion order, each undirected edge
    once, later duplicates overwrite the weight) and the same fp16 adjacency padded to a multiple
    of 8 -- without networkx."""
    device = th.device(device) if device is not None else cfg.DEVICE
    rows: List[List[int]] = []
    with open(filename, 'r') as f:
        for line in f:
            if '//' not in line and line.strip():
                rows.append([int(t) for t in line.split()])
    num_nodes, num_edges = rows[0]
    adj: List[Dict[int, int]] = [dict() for _ in range(num_nodes)]
    for a, b, w in rows[1:]:
        adj[a - 1][b - 1] = w
        adj[b - 1][a - 1] = w
    edge_from, edge_to = [0] * num_edges, [0] * num_edges
    k = 0
    for u in range(num_nodes):
        for v in adj[u]:
            if v >= u:                      # v < u was reported from v's side already
                edge_from[k], edge_to[k] = u, v
                k += 1
    padded = (num_nodes + 7) // 8 * 8
    A = th.zeros((padded, padded), dtype=th.float16)
    for u in range(num_nodes):
        for v, w in adj[u].items():
            A[u, v] = w
    return {'num_nodes': num_nodes, 'num_edges': num_edges,
            'edge_from': th.tensor(edge_from, dtype=th.int64, device=device),
            'edge_to': th.tensor(edge_to, dtype=th.int64, device=device),
            'adj_matrix': A.to(device)}


def record(pre_obj, pre_sample, new_obj, new_sample):
    """util_maxcut.py:56-62: keep the best chain seen so far."""
    max_value, max_index = th.max(new_obj, dim=0)
    if max_value > pre_obj:
        pre_sample = new_sample[max_index]
        pre_obj = max_value
    return pre_obj, pre_sample
