"""NumPy restatement of the reference's MCPG / dREINFORCE samplers.  TEST INFRASTRUCTURE ONLY.

  metro_sampling     rlsolver/methods/MCPG.py:88-117 (== MCPG/sampling.py:67-86)
  sampler_func       rlsolver/methods/MCPG.py:120-166 with the data fields of maxcut_dataloader (187-232)
  sub_set_sampling   rlsolver/methods/L2A/transformer.py:335-353
  weighted_sampler   rlsolver/methods/MCPG/sampling.py:89-127 (mcpg_sampling_maxcut, float `edge_attr`) with the data
                     fields of rlsolver/methods/MCPG/dataloader.py:53-103, 106-124

Random draws are passed in (recorded from the reference, or regenerated from torch's generator by
the tests), so every function is deterministic.  Layout as in the reference: node-major [N, C]."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

f32 = np.float32


def neighbours(num_nodes: int, edges: Sequence[Tuple[int, int, int]]) -> List[np.ndarray]:
    """append_neighbors (MCPG.py:235-289): both directions, in edge order, weight 1."""
    nb: List[List[int]] = [[] for _ in range(num_nodes)]
    for a, b, _ in edges:
        nb[a].append(b)
        nb[b].append(a)
    return [np.asarray(x, dtype=np.int64) for x in nb]


def metro_sampling(probs: np.ndarray, start: np.ndarray, max_transfer_time: int,
                   index_rows: np.ndarray, rands: np.ndarray) -> Tuple[np.ndarray, int]:
    """Returns (samples float32 [N, C], number of iterations executed).  index_rows int64 [T, C] and
    rands float32 [T, C] are the randint / rand draws of iteration t."""
    samples = start.astype(bool).copy()
    probs = probs.astype(f32)
    num_chain = samples.shape[1]
    cols = np.arange(num_chain)
    count, t_done = 0, 0
    for t in range(max_transfer_time * 5):
        if count >= num_chain * max_transfer_time:
            break
        rows = index_rows[t]
        base = probs[rows]
        val = samples[rows, cols]
        chosen = np.where(val, base, f32(1) - base).astype(f32)
        rate = ((f32(1) - chosen) / chosen).astype(f32)
        acc = rands[t].astype(f32) < rate
        samples[rows, cols] = np.where(acc, ~val, val)
        count += int(acc.sum())
        t_done += 1
    return samples.astype(f32), t_done


def sampler_func(num_nodes: int, edges: Sequence[Tuple[int, int, int]], order: np.ndarray, xs_sample: np.ndarray,
                 num_ls: int, total_mcmc_num: int, repeat_times: int, rands: np.ndarray):
    """rands float32 [num_ls * N, C]: the torch.rand(C) of the k-th node visit of sweep s at row s*N + k.
    Returns (vs_good [T], xs_good [N, T], value [C], xs_loc [N, C], expected [C])."""
    nb = neighbours(num_nodes, edges)
    deg = np.asarray([len(x) for x in nb], dtype=np.float64)
    k = 0.25
    xs = xs_sample.astype(f32).copy()
    xs *= f32(2)
    xs -= f32(0.5)
    draw = 0
    for _ in range(num_ls):
        for node in order:
            s = xs[nb[node]].sum(axis=0, dtype=f32) if nb[node].size else np.zeros(xs.shape[1], f32)
            v = (s + (rands[draw].astype(f32) * f32(k)).astype(f32)).astype(f32)
            xs[node] = (v < f32((deg[node] + k) / 2)).astype(f32)
            draw += 1
    e = np.asarray([(a, b) for a, b, _ in edges], dtype=np.int64).reshape(-1, 2)
    expected = ((f32(2) * xs[e[:, 0]] - f32(1)) * (f32(2) * xs[e[:, 1]] - f32(1))).sum(axis=0, dtype=f32)
    index = expected.reshape(-1, total_mcmc_num).argmin(axis=0)
    index = np.arange(total_mcmc_num) + index * total_mcmc_num
    max_cut = expected[index]
    vs_good = (f32(len(edges)) - max_cut) / f32(2)
    value = expected - expected.mean(dtype=f32)
    return vs_good, xs[:, index], value, xs, expected


def weighted_fields(num_nodes: int, edges: np.ndarray, weights: np.ndarray):
    """append_neighbors + the degree fields of MCPG/dataloader.py:75-85, 106-124: per node the neighbour ids and
    edge weights in edge order (both directions), weighted_degree = float(sum of the float32 row)."""
    nb: List[List[int]] = [[] for _ in range(num_nodes)]
    nw: List[List[float]] = [[] for _ in range(num_nodes)]
    for (a, b), w in zip(edges, weights):
        nb[a].append(int(b)), nw[a].append(w)
        nb[b].append(int(a)), nw[b].append(w)
    nb_a = [np.asarray(x, dtype=np.int64) for x in nb]
    nw_a = [np.asarray(x, dtype=f32) for x in nw]
    wdeg = [float(x.sum(dtype=f32)) if x.size else 0.0 for x in nw_a]
    return nb_a, nw_a, wdeg


def weighted_sampler(num_nodes: int, edges: np.ndarray, weights: np.ndarray, order: np.ndarray, metro_out: np.ndarray,
                     num_ls: int, total_mcmc_num: int, rands: np.ndarray):
    """mcpg_sampling_maxcut after its metro_sampling call (sampling.py:101-127).  metro_out float32 [N, C] (0/1);
    rands float32 [max(num_ls, 1) * N, C].  Returns (vs_good [T], xs_good [N, T], value [C], expected [C]).
    The neighbour sum is accumulated in float32 in neighbour order; torch.mm's order is its own, so bit-exactness
    against the reference holds for weights whose partial sums are exact (integers, k/8, ...)."""
    nb, nw, wdeg = weighted_fields(num_nodes, edges, weights)
    xs = metro_out.astype(f32).copy()
    xs = (xs + xs[order[0]].copy()) % f32(2)                  # :102-104 symmetry breaking on the top-degree node
    xs = ((xs - f32(0.5)) * f32(2) + f32(0.5)).astype(f32)    # :105  {0, 1} -> {-0.5, 1.5}
    draw = 0
    cnt = 0
    while True:                                               # :109-121 (at least one sweep)
        cnt += 1
        for node in order:
            s = np.zeros(xs.shape[1], f32)
            for j, w in zip(nb[node], nw[node]):
                s = (s + f32(w) * xs[j]).astype(f32)
            v = (s + (rands[draw].astype(f32) / f32(4)).astype(f32)).astype(f32)
            xs[node] = (v < f32(wdeg[node] / 2 + 0.125)).astype(f32)
            draw += 1
        if cnt >= num_ls:
            break
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    terms = (f32(2) * xs[e[:, 0]] - f32(1)) * (f32(2) * xs[e[:, 1]] - f32(1)) * weights.astype(f32)[:, None]
    expected = terms.sum(axis=0, dtype=f32)
    index = expected.reshape(-1, total_mcmc_num).argmin(axis=0)
    index = np.arange(total_mcmc_num) + index * total_mcmc_num
    wsum = f32(float(weights.astype(f32).sum(dtype=f32)))
    vs_good = (wsum - expected[index]) / f32(2)
    value = expected - expected.mean(dtype=f32)
    return vs_good, xs[:, index], value, expected


def sub_set_sampling(top_ids: np.ndarray, top_values: np.ndarray, start_xs: np.ndarray, num_repeats: int,
                     rands: np.ndarray) -> np.ndarray:
    """The resampling loop only (transformer.py:342-352); the two topk calls are torch's and are
    passed in as (top_ids, top_values) [S, K].  rands float32 [K, R*S]."""
    xs = np.tile(start_xs, (num_repeats, 1))
    rows = np.arange(xs.shape[0])
    for i in range(top_values.shape[1]):
        prob = np.tile(top_values[:, i], num_repeats)
        ids = np.tile(top_ids[:, i], num_repeats)
        xs[rows, ids] = rands[i] < prob
    return xs
