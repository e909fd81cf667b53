"""Torch restatement of the reference's batched simulator, op for op (index-gather objective,
full re-evaluation per candidate flip).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two uses:
  * run on the SAME device and seed as the CUDA path it consumes torch's RNG in the same
    order, which is what "same seeds, same flip sequence" means (tests, smoke);
  * timed on the host cores it is the `cpu_baseline` / `--impl reference` arm of bench.py:
    it does the reference's work the reference's way (three int64 [E, Md] index tensors,
    N full evaluations per sweep), multi-threaded by torch.
It is checked against the NumPy oracle and the reference-generated goldens in
tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch as th

from . import maxcut as om

TEN = th.Tensor


class TorchSim:
    """Follows rlsolver/envs/env_L2A.py:24-116."""

    def __init__(self, edges: Sequence[Tuple[int, int, int]], bidirectional: bool, device="cpu"):
        g = om.build_graph_store(edges, bidirectional)
        self.g = g
        self.device = th.device(device)
        self.bidirectional = bidirectional
        self.num_nodes, self.num_edges = g.num_nodes, g.num_edges
        self.n0 = th.from_numpy(g.n0).to(self.device)[None, :]
        self.n1 = th.from_numpy(g.n1).to(self.device)[None, :]
        self.rows = th.zeros_like(self.n0)
        self.degree = th.from_numpy(g.listed_degree).to(self.device)[None, :]
        self.neigh = [th.from_numpy(g.n1[g.row_ptr[i]:g.row_ptr[i + 1]]).to(self.device) for i in range(g.num_nodes)]

    def objective(self, xs: TEN, if_sum: bool = True) -> TEN:        # env_L2A.py:54-66
        e = xs.shape[0]
        if e != self.rows.shape[0]:
            self.n0 = self.n0[0].repeat(e, 1)
            self.n1 = self.n1[0].repeat(e, 1)
            self.rows = th.arange(e, device=self.device)[:, None].repeat(1, self.n0.shape[1])
        vals = xs[self.rows, self.n0] ^ xs[self.rows, self.n1]
        if if_sum:
            vals = vals.sum(1)
        return vals // 2 if self.bidirectional else vals

    def objective_for_loop(self, xs: TEN, if_sum: bool = True) -> TEN:   # env_L2A.py:68-80
        out = th.zeros(xs.shape, dtype=th.long, device=self.device)
        for i, nb in enumerate(self.neigh):
            if nb.shape[0]:
                out[:, i] = (xs[:, i, None] ^ xs[:, nb]).sum(dim=1)
        if if_sum:
            out = out.sum(dim=1)
        return out.float() / 2 if self.bidirectional else out

    def random_xs(self, e: int) -> TEN:                               # env_L2A.py:82-85
        xs = th.randint(0, 2, size=(e, self.num_nodes), dtype=th.bool, device=self.device)
        xs[:, 0] = 0
        return xs

    @staticmethod
    def keep_not_worse(xs0, vs0, xs1, vs1):                            # util_read_data.py:190-202
        m = vs1.ge(vs0)
        xs0[m] = xs1[m]
        vs0[m] = vs1[m]
        return m.shape[0]

    def sweep(self, xs: TEN, vs: TEN, first_nodes: int = -1) -> None:  # env_L2A.py:110-115
        stop = self.num_nodes if first_nodes < 0 else min(first_nodes, self.num_nodes)
        for i in range(stop):
            cand = xs.clone()
            cand[:, i] = th.logical_not(cand[:, i])
            self.keep_not_worse(xs, vs, cand, self.objective(cand))

    def local_search_inplace(self, xs: TEN, vs, num_iters=8, num_spin=8, noise_std=0.3, sweep_nodes=-1):
        raw = self.objective_for_loop(xs, if_sum=False)                # env_L2A.py:87-116
        vs = raw.sum(dim=1).long() if vs is None else vs.long()
        ws = self.degree - (2 if self.bidirectional else 1) * raw
        spread = ws.max(dim=0, keepdim=True)[0] - ws.min(dim=0, keepdim=True)[0]
        rd = spread.float() * noise_std
        noisy = ws + th.randn_like(ws, dtype=th.float32) * rd
        thresh = th.kthvalue(noisy, k=self.num_nodes - num_spin, dim=1)[0][:, None]
        for _ in range(num_iters):
            noisy = ws + th.randn_like(ws, dtype=th.float32) * rd
            cand = xs.clone()
            mask = noisy.gt(thresh)
            cand[mask] = th.logical_not(cand[mask])
            self.keep_not_worse(xs, vs, cand, self.objective(cand))
        self.sweep(xs, vs, sweep_nodes)
        return xs, vs


class TorchLocalSearch:
    """Follows rlsolver/methods/LocalSearch.py:27-86 (unidirectional simulators only -- the
    reference itself raises for bidirectional ones)."""

    def __init__(self, sim: TorchSim):
        self.sim = sim
        self.good_xs = self.good_vs = None

    def reset(self, xs: TEN) -> TEN:
        self.good_xs, self.good_vs = xs, self.sim.objective(xs)
        return self.good_vs

    def random_search(self, num_iters=8, num_spin=8, noise_std=0.3):
        sim = self.sim
        px = self.good_xs.clone()
        raw = sim.objective_for_loop(px, if_sum=False)
        pv = raw.sum(dim=1)
        thresh = None
        for _ in range(num_iters):
            ws = sim.degree - (4 if sim.bidirectional else 2) * raw
            spread = ws.max(dim=0, keepdim=True)[0] - ws.min(dim=0, keepdim=True)[0]
            noisy = ws + th.randn_like(ws, dtype=th.float32) * (spread.float() * noise_std)
            if thresh is None:
                thresh = th.kthvalue(noisy, k=sim.num_nodes - num_spin, dim=1)[0][:, None]
            cand = px.clone()
            mask = noisy.gt(thresh)
            cand[mask] = th.logical_not(cand[mask])
            sim.keep_not_worse(px, pv, cand, sim.objective(cand))
        sim.sweep(px, pv)
        n = sim.keep_not_worse(self.good_xs, self.good_vs, px, pv)
        return self.good_xs, self.good_vs, n


def metropolis_hastings_sampling_tnco(probs, start_xs, num_repeats: int, num_iters: int = -1, accept_rate: float = 0.25):
    """rlsolver/envs/env_L2A.py:233-276 op for op (row-major independent-site Metropolis; same torch calls in the same
    order, so on one device and seed it consumes the generator exactly like the reference).  TEST INFRASTRUCTURE:
    pinned on the CPU by tests/golden/mhrows_*.npz (tools/make_goldens_mhrows.py), compared with the kernels on CUDA."""
    xs = start_xs.repeat(num_repeats, 1)
    ps = probs.repeat(num_repeats, 1)
    num, dim = xs.shape
    device = xs.device
    num_iters = int(dim * accept_rate) if num_iters == -1 else num_iters
    count = 0
    for _ in range(4):
        ids = th.randperm(dim, device=device)
        for i in range(dim):
            idx = ids[i]
            chosen_p0 = ps[:, idx]
            chosen_xs = xs[:, idx]
            chosen_ps = th.where(chosen_xs, chosen_p0, 1 - chosen_p0)
            accept_masks = th.rand(num, device=device).lt((1 - chosen_ps) / chosen_ps)
            xs[:, idx] = th.where(accept_masks, th.logical_not(chosen_xs), chosen_xs)
            count += accept_masks.sum()
            if count >= num * num_iters:
                break
        if count >= num * num_iters:
            break
    return xs
