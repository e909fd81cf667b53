"""Drop-in for the simulator part of `rlsolver.envs.env_MCPG` (env_MCPG.py:24-116 is the same
EnvMaxcut as env_L2A.py:24-116) plus the local-search driver loop of
`search_and_evaluate_local_search` (env_MCPG.py:408-491) as a reusable function."""
from __future__ import annotations

import torch as th

from ..methods.LocalSearch import LocalSearch
from ..methods.util_read_data import update_xs_by_vs
from .env_L2A import EnvMaxcut

TEN = th.Tensor


def local_search_outer_iteration(sim: EnvMaxcut, solver: LocalSearch, best_xs: TEN, best_vs: TEN,
                                 num_reset_flips: int = 16, num_searches: int = 16,
                                 num_iters: int = 64, num_spin: int = 4, noise_std: float = 0.3):
    """One outer iteration of env_MCPG.py:449-476: broadcast the best row to all envs, apply
    `num_reset_flips` uniform random flips per env, then run `random_search` `num_searches`
    times, merging into (best_xs, best_vs) each time.  RNG order as in the reference."""
    num_sims, num_nodes = best_xs.shape
    device = sim.device
    best_i = best_vs.argmax()
    best_xs[:] = best_xs[best_i]
    best_vs[:] = best_vs[best_i]
    xs = best_xs.clone()
    sim_ids = th.arange(num_sims, device=device)
    for _ in range(num_reset_flips):
        ids = th.randint(0, num_nodes, size=(num_sims,), device=device)
        xs[sim_ids, ids] = th.logical_not(xs[sim_ids, ids])
    solver.reset(xs)
    for _ in range(num_searches):
        good_xs, good_vs, _ = solver.random_search(num_iters=num_iters, num_spin=num_spin, noise_std=noise_std)
        update_xs_by_vs(best_xs, best_vs, good_xs, good_vs, if_maximize=sim.if_maximize)
    return best_xs, best_vs


__all__ = ["EnvMaxcut", "LocalSearch", "update_xs_by_vs", "local_search_outer_iteration"]
