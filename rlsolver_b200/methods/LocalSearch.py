"""Drop-in for `rlsolver.methods.LocalSearch` (rlsolver/methods/LocalSearch.py:13-86).

`LocalSearch(simulator, num_nodes)` keeps `good_xs` / `good_vs` / `num_sims`; `reset(xs)`
evaluates; `random_search(num_iters, num_spin, noise_std)` runs the noisy multi-flip
iterations and the exhaustive single-flip pass on a copy and merges rows that are not worse.
The simulator must be a rlsolver_b200 EnvMaxcut (it owns the native graph store).
"""
from __future__ import annotations

import torch as th

from .util_read_data import update_xs_by_vs as _update_xs_by_vs

TEN = th.Tensor


def update_xs_by_vs(xs0, vs0, xs1, vs1, if_maximize: bool = True):
    return _update_xs_by_vs(xs0, vs0, xs1, vs1, if_maximize)


def merge_searched(good_xs: TEN, good_vs: TEN, backup_xs: TEN, prev_vs: TEN) -> int:
    """update_xs_by_vs(good, prev) (LocalSearch.py:85) when the search ran IN PLACE on good_xs and
    `backup_xs` holds the rows as they were: a row keeps its old spins exactly where the reference
    keeps them, i.e. where prev_vs < good_vs.  That only happens when the caller's good_vs is larger
    than the cut of its good_xs row (a stale value) -- every accepted move has vs' >= vs -- so for
    consistent inputs no row moves and this is two tiny kernels."""
    bar = prev_vs + 1                      # good_vs >= prev_vs + 1  <=>  prev_vs < good_vs (integers)
    _update_xs_by_vs(good_xs, bar, backup_xs, good_vs, True)     # restores the stale rows
    th.maximum(good_vs, prev_vs, out=good_vs)
    return good_vs.shape[0]


class LocalSearch:
    def __init__(self, simulator, num_nodes: int):
        self.simulator = simulator
        self.num_nodes = num_nodes

        self.num_sims = 0
        self.good_xs = th.tensor([])
        self.good_vs = th.tensor([])

    def reset(self, xs: TEN):
        vs = self.simulator.calculate_obj_values(xs=xs)
        self.good_xs = xs
        self.good_vs = vs
        self.num_sims = xs.shape[0]
        return vs

    def reset_search(self, num_sims):
        sim = self.simulator
        xs = th.empty((num_sims, self.num_nodes), dtype=th.bool, device=sim.device)
        for sim_id in range(num_sims):
            _xs = sim.generate_xs_randomly(num_sims=num_sims)
            _vs = sim.calculate_obj_values(_xs)
            xs[sim_id] = _xs[_vs.argmax()]
        return xs

    def random_search(self, num_iters: int = 8, num_spin: int = 8, noise_std: float = 0.3):
        """LocalSearch.py:53-86.  RNG: num_iters draws of randn [E, N] float32; the threshold
        comes from the first draw, which also drives the first iteration."""
        sim = self.simulator
        if sim.if_bidirectional:
            # the reference fails here too: prev_vs is float32 for a bidirectional simulator and
            # index_put of the int64 candidate values raises (LocalSearch.py:24)
            raise RuntimeError("Index put requires the source and destination dtypes match, "
                               "got Float for the destination and Long for the source. "
                               "(LocalSearch.random_search needs if_bidirectional=False, as in the reference)")
        st = sim.store
        num_sims = self.good_xs.shape[0]
        shape = (num_sims, sim.num_nodes)
        if not (self.good_xs.is_contiguous() and self.good_vs.is_contiguous() and self.good_vs.dtype == th.int64):
            raise RuntimeError("random_search updates good_xs / good_vs in place: contiguous bool / int64 needed")
        # The reference searches on prev_xs = good_xs.clone() and merges with update_xs_by_vs(good, prev)
        # (LocalSearch.py:57, 85).  Here the search runs in place on good_xs and the clone is the backup:
        # rows whose searched value stays below a (stale, larger) good_vs are restored from it.
        backup_xs = self.good_xs.clone()
        ws = st.ls_workspace(num_sims)
        prev_vs = st.ls_begin(self.good_xs, None, 2, noise_std, ws)
        if getattr(sim, "fused_rng", False) and num_sims > 0 and num_iters > 0 and st.ls_mask_words(num_sims) >= 0:
            # the threshold comes from the first draw, which also drives iteration 0 (LocalSearch.py:66-68)
            st.ls_fused(prev_vs, 2, num_spin, num_iters, True, self.good_xs, ws)
            num_update = merge_searched(self.good_xs, self.good_vs, backup_xs, prev_vs)
            return self.good_xs, self.good_vs, num_update
        done = 0
        first = True
        while done < num_iters:
            now = min(16, num_iters - done)
            noises = [th.randn(shape, dtype=th.float32, device=sim.device) for _ in range(now)]
            done += now
            # the threshold comes from the first draw, which also drives iteration 0
            st.ls_run(prev_vs, 2, noises[0] if first else None, num_spin, noises, done == num_iters, self.good_xs, ws)
            first = False
        if num_iters <= 0:
            st.ls_search(prev_vs, 2, [], True, self.good_xs, ws)
        num_update = merge_searched(self.good_xs, self.good_vs, backup_xs, prev_vs)
        return self.good_xs, self.good_vs, num_update
