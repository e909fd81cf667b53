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

        packed = st.pack(self.good_xs)                      # prev_xs = good_xs.clone()
        cross, cmin, cmax = st.cross_counts(packed, num_sims)
        prev_vs = st.cut_eval_packed(packed, num_sims)      # prev_vs_raw.sum(dim=1)
        if num_iters > 0:
            noises = [th.randn(shape, dtype=th.float32, device=sim.device)]
            thresh = st.ls_thresh(cross, cmin, cmax, 2, noise_std, noises[0], num_spin)
            done = 0
            while done < num_iters:
                now = min(16, num_iters - done)
                while len(noises) < now:
                    noises.append(th.randn(shape, dtype=th.float32, device=sim.device))
                st.ls_noisy_iters(packed, prev_vs, cross, cmin, cmax, 2, noise_std, noises, thresh)
                done += now
                noises = []
        st.flip_sweep(packed, prev_vs)
        prev_xs = st.unpack(packed, num_sims)
        num_update = update_xs_by_vs(self.good_xs, self.good_vs, prev_xs, prev_vs)
        return self.good_xs, self.good_vs, num_update
