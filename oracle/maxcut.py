"""NumPy restatement of the reference's batched max-cut simulator and local search.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function names the
reference lines it restates; paths are relative to the reference checkout.
Arrays: ``xs`` is ``bool [E, N]`` (E environments, N nodes), ``vs`` is ``[E]``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

Edge = Tuple[int, int, int]


# --------------------------------------------------------------------------- graph

def parse_graph_text(text: str) -> Tuple[int, int, List[Edge]]:
    """`N M` header then `u v w` 1-based rows -> 0-based triples.
    Follows rlsolver/methods/util_read_data.py:76-82 (read_mygraph)."""
    rows = [[int(tok) for tok in ln.split()] for ln in text.splitlines() if ln.strip()]
    n_nodes, n_edges = rows[0][0], rows[0][1]
    edges = [(a - 1, b - 1, w) for a, b, w in rows[1:]]
    return n_nodes, n_edges, edges


def read_graph(path: str) -> List[Edge]:
    with open(path, "r") as fh:
        return parse_graph_text(fh.read())[2]


def count_nodes(edges: Sequence[Edge]) -> int:
    """Number of DISTINCT endpoints (isolated nodes are not counted).
    Follows rlsolver/methods/util.py:35-40 (calc_num_nodes_in_mygraph)."""
    seen = set()
    for a, b, _ in edges:
        seen.add(a)
        seen.add(b)
    return len(seen)


@dataclass
class GraphStore:
    """What EnvMaxcut.__init__ builds (rlsolver/envs/env_L2A.py:25-52)."""
    num_nodes: int
    num_edges: int              # len(mygraph), NOT doubled when bidirectional
    bidirectional: bool
    n0: np.ndarray              # int64 [Md] first endpoint of every listed edge, CSR order
    n1: np.ndarray              # int64 [Md] second endpoint
    row_ptr: np.ndarray         # int64 [N+1] listed-neighbour CSR offsets
    listed_degree: np.ndarray   # int64 [N]  == n0_num_n1[0]


def build_graph_store(edges: Sequence[Edge], bidirectional: bool) -> GraphStore:
    """Per-node listed-neighbour lists (forward only, plus reverse when
    bidirectional), each sorted ascending, concatenated in node order.
    Follows util_read_data.py:144-187 (build_adjacency_indies) and
    env_L2A.py:40-52."""
    n = count_nodes(edges)
    buckets: List[List[int]] = [[] for _ in range(n)]
    for a, b, _ in edges:
        buckets[a].append(b)
        if bidirectional:
            buckets[b].append(a)
    n0, n1, row_ptr = [], [], [0]
    for i, nb in enumerate(buckets):
        nb = sorted(nb)
        n0.extend([i] * len(nb))
        n1.extend(nb)
        row_ptr.append(len(n1))
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    return GraphStore(
        num_nodes=n, num_edges=len(edges), bidirectional=bidirectional,
        n0=np.asarray(n0, dtype=np.int64), n1=np.asarray(n1, dtype=np.int64),
        row_ptr=row_ptr, listed_degree=np.diff(row_ptr))


def adjacency_bool(edges: Sequence[Edge], n: int = 0, bidirectional: bool = False) -> np.ndarray:
    """Follows rlsolver/methods/util.py:343-370 (build_adjacency_bool)."""
    n = n or count_nodes(edges)
    adj = np.zeros((n, n), dtype=bool)
    for a, b, _ in edges:
        adj[a, b] = True
    return adj | adj.T if bidirectional else adj


# --------------------------------------------------------------------------- objective

def cut_values(g: GraphStore, xs: np.ndarray, if_sum: bool = True) -> np.ndarray:
    """calculate_obj_values (env_L2A.py:54-66): XOR of the two endpoint bits over
    every listed edge, summed per env, floor-halved iff bidirectional.
    With if_sum=False returns the bool [E, Md] indicators; combined with
    bidirectional the reference's `bool // 2` yields int64 zeros (torch type
    promotion) -- restated as such, no caller uses it."""
    ind = xs[:, g.n0] ^ xs[:, g.n1]
    if not if_sum:
        return ind.astype(np.int64) // 2 if g.bidirectional else ind
    tot = ind.sum(axis=1, dtype=np.int64)
    return tot // 2 if g.bidirectional else tot


def node_cross_counts_raw(g: GraphStore, xs: np.ndarray) -> np.ndarray:
    """int64 [E, N]: for every node, how many of its LISTED neighbours sit on the
    other side.  The integer core of calculate_obj_values_for_loop
    (env_L2A.py:68-76) before the optional sum / float halving."""
    ind = (xs[:, g.n0] ^ xs[:, g.n1]).astype(np.int64)
    run = np.concatenate([np.zeros((xs.shape[0], 1), np.int64), np.cumsum(ind, axis=1)], axis=1)
    return run[:, g.row_ptr[1:]] - run[:, g.row_ptr[:-1]]


def obj_values_for_loop(g: GraphStore, xs: np.ndarray, if_sum: bool = True) -> np.ndarray:
    """calculate_obj_values_for_loop (env_L2A.py:68-80).  Bidirectional results are
    float32 halves even with if_sum=False (line 78-79)."""
    vals = node_cross_counts_raw(g, xs)
    if if_sum:
        vals = vals.sum(axis=1)
    if g.bidirectional:
        vals = vals.astype(np.float32) / np.float32(2)
    return vals


def random_xs(rng: np.random.Generator, g: GraphStore, num_envs: int) -> np.ndarray:
    """generate_xs_randomly (env_L2A.py:82-85): uniform bits, column 0 cleared.
    (The reference draws from torch's generator; the oracle only restates shape
    and the symmetry-breaking column.)"""
    xs = rng.integers(0, 2, size=(num_envs, g.num_nodes)).astype(bool)
    xs[:, 0] = False
    return xs


# --------------------------------------------------------------------------- select ops

def update_xs_by_vs(xs0, vs0, xs1, vs1, if_maximize: bool = True) -> int:
    """util_read_data.py:190-202: rows of (xs1, vs1) replace rows of (xs0, vs0)
    where not worse (>= for maximise).  In place.  Returns E (reference quirk:
    `good_is.shape[0]`, not the number replaced)."""
    keep = vs1 >= vs0 if if_maximize else vs1 <= vs0
    xs0[keep] = xs1[keep]
    vs0[keep] = vs1[keep]
    return int(keep.shape[0])


def pick_xs_by_vs(xs, vs, num_repeats: int, if_maximize: bool = True):
    """util_read_data.py:204-216: view as [R, S, N]; per sim take the repeat with
    the best value (first index on ties, as torch CPU argmax does)."""
    n = xs.shape[1]
    s = xs.shape[0] // num_repeats
    xv = xs.reshape(num_repeats, s, n)
    vv = vs.reshape(num_repeats, s)
    ids = vv.argmax(axis=0) if if_maximize else vv.argmin(axis=0)
    col = np.arange(s)
    return xv[ids, col], vv[ids, col]


def evolutionary_replacement(xs, vs, low_k: int, if_maximize: bool, perm: np.ndarray) -> None:
    """rlsolver/methods/util.py:87-94.  `perm` is the recorded
    randperm(E - low_k); ascending argsort (ties: caller must avoid or accept
    torch's order -- goldens use distinct values)."""
    ids = np.argsort(vs, kind="stable")
    if if_maximize:
        top_ids, low_ids = ids[:-low_k], ids[-low_k:]
    else:
        top_ids, low_ids = ids[:low_k], ids[low_k:]
    rep = top_ids[perm[:low_k]]
    xs[rep] = xs[low_ids]
    vs[rep] = vs[low_ids]


# --------------------------------------------------------------------------- integer-weighted objective

def cut_values_weighted(edges: Sequence[Edge], xs: np.ndarray) -> np.ndarray:
    """int64 [E]: sum of w over the edges whose ends differ -- obj_maxcut (rlsolver/methods/util_obj.py:31-39,
    `obj += adj[i, j]` for i < j with result[i] != result[j]) for a batch of rows; self loops never count.
    Equivalently PISCO's energy: -1/4 s^T A s = cut_w - sum(A) / 4 (rlsolver/envs/env_ISCO.py:436-444)."""
    arr = np.asarray(list(edges), dtype=np.int64).reshape(-1, 3)
    u, v, w = arr[:, 0], arr[:, 1], arr[:, 2]
    return ((xs[:, u] ^ xs[:, v]).astype(np.int64) * w[None, :]).sum(axis=1)


def node_fields_weighted(edges: Sequence[Edge], num_nodes: int, xs: np.ndarray) -> np.ndarray:
    """int64 [E, N]: per node the total weight of its incident cut edges (each undirected edge seen from both
    ends).  With d = 2x - 1: d_i (A d)_i = wdeg_i - 2 * this, the quantity PISCO's gradient carries
    ((1 - 2x_i) grad_i = (A d)_i d_i / 2 up to the temperature, env_ISCO.py:412-418, 436-444)."""
    arr = np.asarray(list(edges), dtype=np.int64).reshape(-1, 3)
    out = np.zeros((xs.shape[0], num_nodes), np.int64)
    for a, b, w in arr:
        if a == b:
            continue
        cut = (xs[:, a] ^ xs[:, b]).astype(np.int64) * w
        out[:, a] += cut
        out[:, b] += cut
    return out


# --------------------------------------------------------------------------- local search

def kth_smallest(a: np.ndarray, k: int) -> np.ndarray:
    """torch.kthvalue(a, k, dim=1).values for 1-based k."""
    return np.partition(a, k - 1, axis=1)[:, k - 1]


def _spin_rand(ws: np.ndarray, noise: np.ndarray, rd_std: np.ndarray) -> np.ndarray:
    """float32 `ws + noise * rd_std` with one rounding per op (two torch kernels)."""
    prod = (noise.astype(np.float32) * rd_std.astype(np.float32)).astype(np.float32)
    return (ws.astype(np.float32) + prod).astype(np.float32)


def sweep_literal(g: GraphStore, xs: np.ndarray, vs: np.ndarray) -> None:
    """Phase (iii), literally: for every node in index order, flip that column of
    a copy, re-evaluate the full objective, keep rows that are not worse.
    env_L2A.py:110-115 / LocalSearch.py:78-83.  O(N) full evaluations."""
    for i in range(g.num_nodes):
        cand = xs.copy()
        cand[:, i] = ~cand[:, i]
        update_xs_by_vs(xs, vs, cand, cut_values(g, cand).astype(vs.dtype), True)


def full_neighbourhood(g: GraphStore) -> Tuple[np.ndarray, np.ndarray]:
    """Undirected CSR with multiplicity and without self loops (a self loop never
    contributes to the cut).  Derived structure used by the O(degree) forms."""
    if g.bidirectional:
        a, b = g.n0, g.n1
    else:
        a = np.concatenate([g.n0, g.n1])
        b = np.concatenate([g.n1, g.n0])
    keep = a != b
    a, b = a[keep], b[keep]
    order = np.lexsort((b, a))
    a, b = a[order], b[order]
    ptr = np.zeros(g.num_nodes + 1, np.int64)
    np.add.at(ptr, a + 1, 1)
    return np.cumsum(ptr), b


def sweep_delta(g: GraphStore, xs: np.ndarray, vs: np.ndarray) -> None:
    """Same result as sweep_literal via the single-flip gain
    (same-side minus other-side neighbours); checked equal in the tests.
    The update rule is the batched form of S2V_PPO/env.py:197-206."""
    ptr, nb = full_neighbourhood(g)
    for i in range(g.num_nodes):
        js = nb[ptr[i]:ptr[i + 1]]
        cross = (xs[:, js] ^ xs[:, [i]]).sum(axis=1)
        gain = js.size - 2 * cross
        acc = gain >= 0
        xs[acc, i] = ~xs[acc, i]
        vs[acc] += gain[acc].astype(vs.dtype)


def batch_rd_std(g: GraphStore, xs: np.ndarray, mult: int, noise_std: float) -> np.ndarray:
    """float32 [1, N]: `(max_e ws - min_e ws) * noise_std` over the WHOLE batch `xs` (env_L2A.py:92-95 with
    mult = 2 if bi else 1; LocalSearch.py:64-66 with mult = 4 if bi else 2).  The spread couples the envs of a
    batch; everything after it is per env, so a test can replay a SUBSET of the rows of a large batch by
    handing this to local_search_inplace / LocalSearch.random_search as `rd_std`."""
    ws = g.listed_degree[None, :] - mult * obj_values_for_loop(g, xs, if_sum=False)
    ws_std = ws.max(axis=0, keepdims=True) - ws.min(axis=0, keepdims=True)
    return (ws_std.astype(np.float32) * np.float32(noise_std)).astype(np.float32)


def local_search_inplace(g: GraphStore, good_xs: np.ndarray, good_vs, noises: Sequence[np.ndarray],
                         num_iters: int = 8, num_spin: int = 8, noise_std: float = 0.3,
                         literal: bool = True, rd_std: np.ndarray = None):
    """EnvMaxcut.local_search_inplace (env_L2A.py:87-116).
    `noises` holds the 1 + num_iters recorded randn draws (float32 [E, N]).
    `good_vs=None` stands for the reference's `()` sentinel (line 91).
    `rd_std`: see batch_rd_std (rows of a larger batch); None = computed from this batch, as the reference does."""
    vs_raw = obj_values_for_loop(g, good_xs, if_sum=False)
    if good_vs is None:
        good_vs = vs_raw.sum(axis=1).astype(np.int64)
    else:
        good_vs = good_vs.astype(np.int64)
    mult = 2 if g.bidirectional else 1
    ws = g.listed_degree[None, :] - mult * vs_raw           # float32 (bi) / int64 (uni)
    if rd_std is None:
        ws_std = ws.max(axis=0, keepdims=True) - ws.min(axis=0, keepdims=True)
        rd_std = (ws_std.astype(np.float32) * np.float32(noise_std)).astype(np.float32)
    thresh = kth_smallest(_spin_rand(ws, noises[0], rd_std), g.num_nodes - num_spin)[:, None]
    for it in range(num_iters):
        mask = _spin_rand(ws, noises[1 + it], rd_std) > thresh
        cand = good_xs ^ mask
        update_xs_by_vs(good_xs, good_vs, cand, cut_values(g, cand), True)
    if g.num_nodes:
        (sweep_literal if literal else sweep_delta)(g, good_xs, good_vs)
    return good_xs, good_vs


class LocalSearch:
    """rlsolver/methods/LocalSearch.py:27-86 restated on NumPy arrays."""

    def __init__(self, g: GraphStore):
        self.g = g
        self.good_xs = None
        self.good_vs = None
        self.num_sims = 0

    def reset(self, xs: np.ndarray) -> np.ndarray:
        vs = cut_values(self.g, xs)
        self.good_xs, self.good_vs, self.num_sims = xs, vs, xs.shape[0]
        return vs

    def random_search(self, noises: Sequence[np.ndarray], num_iters: int = 8, num_spin: int = 8,
                      noise_std: float = 0.3, literal: bool = True, rd_std: np.ndarray = None):
        g = self.g
        kth = g.num_nodes - num_spin
        prev_xs = self.good_xs.copy()
        raw = obj_values_for_loop(g, prev_xs, if_sum=False)
        prev_vs = raw.sum(axis=1)                             # int64 (uni) / float32 (bi)
        mult = 4 if g.bidirectional else 2
        ws = g.listed_degree[None, :] - mult * raw            # constant across iterations (raw is not refreshed)
        if rd_std is None:                                    # else: rows of a larger batch, see batch_rd_std
            ws_std = ws.max(axis=0, keepdims=True) - ws.min(axis=0, keepdims=True)
            rd_std = (ws_std.astype(np.float32) * np.float32(noise_std)).astype(np.float32)
        thresh = None
        for it in range(num_iters):
            sr = _spin_rand(ws, noises[it], rd_std)
            if thresh is None:
                thresh = kth_smallest(sr, kth)[:, None]
            cand = prev_xs ^ (sr > thresh)
            cv = cut_values(g, cand)
            update_xs_by_vs(prev_xs, prev_vs, cand, cv.astype(prev_vs.dtype), True)
        if literal:
            sweep_literal(g, prev_xs, prev_vs)
        else:
            sweep_delta(g, prev_xs, prev_vs)
        n_upd = update_xs_by_vs(self.good_xs, self.good_vs, prev_xs, prev_vs.astype(self.good_vs.dtype), True)
        return self.good_xs, self.good_vs, n_upd


# --------------------------------------------------------------------------- greedy (semantic reference)

def greedy_best_flip(g: GraphStore, xs: np.ndarray, strict: bool = True, max_flips: int = 1 << 30):
    """Batched best-single-flip ascent with the contract of
    rlsolver/methods/greedy.py:33-78 (lowest index among the best gains; accept
    only a strictly positive gain, else that env stops).  Returns (xs, vs, flips)."""
    ptr, nb = full_neighbourhood(g)
    xs = xs.copy()
    e = xs.shape[0]
    vs = cut_values(g, xs).astype(np.int64)
    flips = np.zeros(e, np.int64)
    src = np.repeat(np.arange(g.num_nodes), np.diff(ptr))
    for env in range(e):
        x = xs[env]
        for _ in range(max_flips):
            cross = np.bincount(src, weights=(x[src] ^ x[nb]).astype(np.float64), minlength=g.num_nodes)
            gain = (np.diff(ptr) - 2 * cross).astype(np.int64)
            i = int(gain.argmax())
            if gain[i] > 0 or (not strict and gain[i] >= 0):
                x[i] = ~x[i]
                vs[env] += gain[i]
                flips[env] += 1
                if not strict and gain[i] == 0:
                    break
            else:
                break
    return xs, vs, flips


# --------------------------------------------------------------------------- pattern-I env (env_PPO)

class PPOEnv:
    """rlsolver/envs/env_PPO.py:63-126 restated on NumPy: float32 {0,1} observations, one flip
    per env per step, reward = cut after - cut before by full re-evaluation, `done` every
    num_steps steps."""

    def __init__(self, g: GraphStore, num_steps: int):
        self.g, self.num_steps, self.action_count = g, num_steps, 0
        self.xs = None
        self.last_reward = None

    def reset(self, xs_bool: np.ndarray) -> np.ndarray:
        self.xs = xs_bool.astype(np.float32)                       # env_PPO.py:85-90 (xs drawn by the caller)
        self.last_reward = cut_values(self.g, self.xs > 0).astype(np.float32)
        return self.xs

    def step(self, action: np.ndarray):
        self.action_count += 1
        rows = np.arange(self.xs.shape[0])
        self.xs[rows, action] = np.logical_not(self.xs[rows, action]).astype(np.float32)   # :94-95
        cur = cut_values(self.g, self.xs > 0).astype(np.float32)
        reward = cur - self.last_reward
        self.last_reward = cur
        done = np.full(self.xs.shape[0], float(self.action_count == self.num_steps), np.float32)
        if self.action_count == self.num_steps:
            self.action_count = 0
        return self.xs, reward, done, cur


def greedy_trace(g: GraphStore, x0: np.ndarray, num_steps=None):
    """greedy_maxcut (rlsolver/methods/greedy.py:33-78) for ONE start state, literally: evaluate all N
    single flips by full recompute, first index of the max, accept iff strictly better.  Returns
    (score, solution, scores after every accepted step)."""
    x = x0.copy()
    n = g.num_nodes
    steps = n if num_steps is None else min(num_steps, n)
    cur = int(cut_values(g, x[None, :])[0])
    scores = []
    for _ in range(steps):
        cand = np.repeat(x[None, :], n, axis=0)
        cand[np.arange(n), np.arange(n)] ^= True
        vals = cut_values(g, cand)
        i = int(vals.argmax())
        if vals[i] > cur:
            cur = int(vals[i])
            x = cand[i]
            scores.append(cur)
        else:
            break
    return cur, x, scores
