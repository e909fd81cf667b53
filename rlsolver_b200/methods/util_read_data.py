"""Graph ingest and the select operators of rlsolver/methods/util_read_data.py.

read_mygraph (76-82), load_mygraph2 (121-140), build_adjacency_indies (144-187),
update_xs_by_vs (190-202), pick_xs_by_vs (204-216).  Same names, arguments and return
conventions (including update_xs_by_vs returning the batch size); the tensor work runs in
the CUDA kernels of csrc/select.cu and the native graph builder.
"""
from __future__ import annotations

import os
import random
from typing import List, Tuple

import torch as th

from ..graph_store import GraphStore, pick_best, select_rows
from .config import GRAPH_TYPE, GRAPH_TYPES, GraphType, MyGraph

TEN = th.Tensor


def read_mygraph(filename: str) -> MyGraph:
    """`N M` header, then 1-based `u v w` rows -> 0-based triples."""
    with open(filename, "r") as fh:
        rows = [[int(tok) for tok in ln.split()] for ln in fh if ln.strip()]
    return [(a - 1, b - 1, w) for a, b, w in rows[1:]]


def generate_mygraph(graph_type: GraphType, num_nodes: int) -> Tuple[MyGraph, int, int]:
    """rlsolver/methods/util_generate.py:75-93 (networkx generators, unit weights)."""
    import networkx as nx
    assert graph_type in GRAPH_TYPES
    if graph_type == GraphType.ER:
        g = nx.erdos_renyi_graph(n=num_nodes, p=0.15)
    elif graph_type == GraphType.PL:
        g = nx.powerlaw_cluster_graph(n=num_nodes, m=4, p=0.05)
    else:
        g = nx.barabasi_albert_graph(n=num_nodes, m=4)
    graph = [(a, b, 1) for a, b in g.edges]
    return graph, num_nodes, len(graph)


def load_mygraph2(dataDir: str = "./data/syn_" + GRAPH_TYPE.value, graph_name: str = "") -> MyGraph:
    if os.path.exists(f"{dataDir}/{graph_name}.txt"):
        return read_mygraph(f"{dataDir}/{graph_name}.txt")
    if os.path.isfile(graph_name) and os.path.splitext(graph_name)[-1] == ".txt":
        return read_mygraph(graph_name)
    if GRAPH_TYPE and graph_name.find("ID") == -1:
        try:
            num_nodes = int(graph_name.split("_")[-1])
        except ValueError:
            raise ValueError(f"DataDir {dataDir} | graph_name {graph_name} txt_path {dataDir}/{graph_name}.txt")
        return generate_mygraph(num_nodes=num_nodes, graph_type=GRAPH_TYPE)[0]
    if GRAPH_TYPE and graph_name.find("ID") >= 0:
        num_nodes, valid_i = graph_name.split("_")[-2:]
        random.seed(int(valid_i[len("ID"):]))
        mygraph = generate_mygraph(num_nodes=int(num_nodes), graph_type=GRAPH_TYPE)[0]
        random.seed()
        return mygraph
    raise ValueError(f"DataDir {dataDir} | graph_name {graph_name} txt_path {dataDir}/{graph_name}.txt")


def build_adjacency_indies(mygraph: MyGraph, if_bidirectional: bool = False) -> Tuple[List[TEN], List[TEN]]:
    """Per-node sorted listed-neighbour tensors (and their edge weights), built by the native
    CSR builder instead of N Python-level argsorts."""
    store = GraphStore(mygraph, if_bidirectional, device=None)
    arrs = store.export()
    ptr, col = arrs["listed_ptr"], arrs["listed_col"]
    wmap = {}
    for a, b, w in mygraph:
        wmap.setdefault((a, b), []).append(w)
        if if_bidirectional:
            wmap.setdefault((b, a), []).append(w)
    n1s, dts = [], []
    for i in range(store.num_nodes):
        nb = col[ptr[i]:ptr[i + 1]].astype("int64")
        n1s.append(th.from_numpy(nb.copy()))
        seen = {}
        ws = []
        for j in nb.tolist():
            k = seen.get(j, 0)
            ws.append(wmap[(i, j)][k])
            seen[j] = k + 1
        dts.append(th.tensor(ws, dtype=th.int64))
    return n1s, dts


def update_xs_by_vs(xs0: TEN, vs0: TEN, xs1: TEN, vs1: TEN, if_maximize: bool) -> int:
    """Rows of (xs1, vs1) replace rows of (xs0, vs0) where not worse; in place; returns the
    batch size like the reference (`good_is.shape[0]`)."""
    int_types = (th.int64, th.int32, th.int16, th.int8, th.uint8)
    fast = (xs0.is_cuda and xs0.dtype == th.bool and xs1.dtype == th.bool and xs0.shape == xs1.shape
            and xs0.is_contiguous() and vs0.dtype == th.int64 and vs0.is_contiguous() and vs1.dtype in int_types)
    if fast:
        select_rows(xs0, vs0, xs1, vs1, if_maximize)
    else:  # any other dtype combination the reference accepts (float values, non-bool rows): same semantics, torch ops
        good = vs1.ge(vs0) if if_maximize else vs1.le(vs0)
        xs0.copy_(th.where(good[:, None], xs1.to(xs0.dtype), xs0))
        vs0.copy_(th.where(good, vs1.to(vs0.dtype), vs0))
    return xs0.shape[0]


def pick_xs_by_vs(xs: TEN, vs: TEN, num_repeats: int, if_maximize: bool) -> Tuple[TEN, TEN]:
    return pick_best(xs, vs, num_repeats, if_maximize)
