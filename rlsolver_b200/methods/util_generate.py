"""Mirror of the graph generators of rlsolver/methods/util_generate.py:75-93: `generate_mygraph(graph_type, num_nodes)`
returns one synthetic graph as the reference's edge list `[(node0, node1, 1), ...]` with its node and edge counts.
Host code, as in the reference: the same networkx calls with the same parameters (ER p = 0.15, PL m = 4 / p = 0.05,
BA m = 4), so a seeded Python `random` gives the reference's graph.  The BATCHED device-side generators (one graph per
env, bit rows straight into the pattern-I env) are `rlsolver_b200.envs.env_PECO.Random{ER,BA,PL}GraphGenerator`."""
from __future__ import annotations

from typing import Tuple

from .config import GraphType, MyGraph

GRAPH_TYPES = [GraphType.ER, GraphType.PL, GraphType.BA]


def generate_mygraph(graph_type: GraphType, num_nodes: int) -> Tuple[MyGraph, int, int]:
    import networkx as nx
    assert graph_type in GRAPH_TYPES
    if graph_type == GraphType.ER:
        g = nx.erdos_renyi_graph(n=num_nodes, p=0.15)
    elif graph_type == GraphType.PL:
        g = nx.powerlaw_cluster_graph(n=num_nodes, m=4, p=0.05)
    elif graph_type == GraphType.BA:
        g = nx.barabasi_albert_graph(n=num_nodes, m=4)
    else:
        raise ValueError(f"g_type {graph_type} should in {GRAPH_TYPES}")
    graph = [(node0, node1, 1) for node0, node1 in g.edges]
    return graph, num_nodes, len(graph)


__all__ = ["generate_mygraph", "GRAPH_TYPES"]
