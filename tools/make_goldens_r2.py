"""Round-2 fixtures from the UNMODIFIED reference (CPU, build container only): python tools/make_goldens_r2.py

* stale_ls_*.npz -- LocalSearch.random_search when the caller's good_vs is LARGER than the cut of its good_xs
  row for some rows (a stale value): the reference's final update_xs_by_vs keeps those rows' old spins unless
  the search reaches the stale value (rlsolver/methods/LocalSearch.py:53-86, merge at :85).
"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from make_goldens import OUT, Recorder, graph_cases  # noqa: E402
from rlsolver.envs import env_L2A  # noqa: E402
from rlsolver.methods import LocalSearch as ref_ls  # noqa: E402


def stale_case(name, mygraph, num_envs, seed):
    th.manual_seed(seed)
    sim = env_L2A.EnvMaxcut(mygraph=mygraph, if_bidirectional=False)
    n = sim.num_nodes
    num_spin = min(8, max(1, n // 8))
    solver = ref_ls.LocalSearch(simulator=sim, num_nodes=n)
    xs0 = sim.generate_xs_randomly(num_envs)
    vs0 = solver.reset(xs0.clone())
    out = {"edges": np.asarray(mygraph, dtype=np.int64), "xs0": xs0.numpy().copy(), "vs0": vs0.numpy().copy(),
           "num_spin": np.asarray(num_spin)}
    # every third row claims a value it does not have: +1 (often reached), +15 / +400 (never reached on these graphs)
    bump = th.zeros(num_envs, dtype=th.long)
    bump[0::3] = th.tensor([1, 15, 400] * num_envs)[: len(bump[0::3])]
    solver.good_vs = vs0 + bump
    out["stale_vs"] = solver.good_vs.numpy().copy()
    for tag, iters in (("a", 3), ("b", 2)):
        with Recorder("randn_like") as rec:
            rx, rv, nu = solver.random_search(num_iters=iters, num_spin=num_spin, noise_std=0.3)
        out[f"{tag}_noise"] = np.stack(rec.draws)
        out[f"{tag}_xs"], out[f"{tag}_vs"] = rx.numpy().copy(), rv.numpy().copy()
        out[f"{tag}_iters"] = np.asarray(iters)
    kept = int((out["a_xs"] == out["xs0"]).all(axis=1).sum())
    path = os.path.join(OUT, f"stale_ls_{name}_E{num_envs}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", kept, "rows kept their old spins after the first call")


def main():
    cases = graph_cases()
    stale_case("ba100", cases["ba100"], 48, seed=101)
    stale_case("multi67", cases["multi67"], 35, seed=102)


if __name__ == "__main__":
    main()
