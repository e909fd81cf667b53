"""Greedy best-single-flip ascent for max-cut, batched over environments.

`greedy_maxcut(num_steps, graph)` keeps the call shape and return values of
rlsolver/methods/greedy.py:33-78 (start from all zeros, lowest index among the best flips, accept
only a strictly better cut, at most min(num_steps, N) steps) for one graph given as a MyGraph
edge list or a networkx graph; `greedy_maxcut_batched` runs the same rule on any batch of start
states.  The reference re-evaluates the whole objective N times per step in Python (O(N^3) per
step); here all N gains per env stay resident on chip and a flip costs O(degree) (csrc/fields.cu).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch as th

from ..envs.env_L2A import EnvMaxcut
from .config import MyGraph

TEN = th.Tensor


def _as_mygraph(graph) -> MyGraph:
    if hasattr(graph, "edges") and hasattr(graph, "number_of_nodes"):      # networkx
        return [(int(a), int(b), 1) for a, b in graph.edges]
    return list(graph)


def greedy_maxcut_batched(sim: EnvMaxcut, xs: TEN, num_steps: Optional[int] = None,
                          strict: bool = True) -> Tuple[TEN, TEN, TEN]:
    """In place on `xs` (bool [E, N]).  Returns (xs, vs int64 [E], number of flips int32 [E])."""
    n = sim.num_nodes
    steps = n if num_steps is None else min(int(num_steps), n)
    vs, flips = sim.store.greedy_best_flip(xs, steps, strict)
    return xs, vs, flips


def greedy_maxcut(num_steps: Optional[int], graph, device=None) -> Tuple[int, List[int], List[int]]:
    mygraph = _as_mygraph(graph)
    device = th.device("cuda", th.cuda.current_device()) if device is None else device
    sim = EnvMaxcut(mygraph=mygraph, device=device, if_bidirectional=True)
    xs = th.zeros((1, sim.num_nodes), dtype=th.bool, device=sim.device)
    init = int(sim.calculate_obj_values(xs)[0])
    steps = sim.num_nodes if num_steps is None else min(int(num_steps), sim.num_nodes)
    # the reference also returns the score after every accepted step: replay them one step at a time
    scores: List[int] = []
    cur = init
    for _ in range(steps):
        vs, flips = sim.store.greedy_best_flip(xs, 1, True)
        if int(flips[0]) == 0:
            break
        cur = int(vs[0])
        scores.append(cur)
    return cur, xs[0].to(th.int64).tolist(), scores
