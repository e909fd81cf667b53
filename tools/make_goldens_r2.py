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


def weighted(mygraph, values, seed):
    """The same edges with weights drawn from `values` (duplicates dropped: the reference's loader keeps one)."""
    rng = np.random.default_rng(seed)
    seen, out = set(), []
    for a, b, _ in mygraph:
        key = (min(a, b), max(a, b))
        if a != b and key not in seen:
            seen.add(key)
            out.append((a, b, int(rng.choice(values))))
    return out


def weighted_cut_case(name, mygraph, num_envs, seed):
    """Weighted objective known answers: obj_maxcut (rlsolver/methods/util_obj.py:31-39) on an nx.Graph with integer
    weights, and PISCO's tensor_core_energy / its x-gradient at T = 1 (rlsolver/envs/env_ISCO.py:436-444)."""
    import networkx as nx
    from rlsolver.envs import env_ISCO
    from rlsolver.methods import util_obj
    from rlsolver.methods.ISCO import util_maxcut
    import make_goldens_isco as mgi
    n = len({a for a, _, _ in mygraph} | {b for _, b, _ in mygraph})
    g = nx.Graph()
    g.add_nodes_from(range(n))
    for a, b, w in mygraph:
        g.add_edge(a, b, weight=w)
    rng = np.random.default_rng(seed)
    xs = rng.integers(0, 2, (num_envs, n)).astype(bool)
    cuts = np.asarray([util_obj.obj_maxcut(row.astype(int).tolist(), g) for row in xs], dtype=np.int64)
    tmp = f"/tmp/wcut_{name}.txt"
    mgi.write_graph(tmp, n, mygraph)
    for mod in (env_ISCO, util_maxcut):
        mod.BATCH_SIZE, mod.DEVICE = num_envs, th.device("cpu")
    params = util_maxcut.load_data(tmp)
    sampler = env_ISCO.PISCO_maxcut(params)
    pad = ((n + 7) // 8 * 8) - n
    x16 = th.nn.functional.pad(th.from_numpy(xs).to(th.float16), (0, pad))
    energy, grad = sampler.tensor_core_energy(x16, th.tensor(1.0))
    path = os.path.join(OUT, f"weighted_cut_{name}_E{num_envs}.npz")
    np.savez_compressed(path, edges=np.asarray(mygraph, dtype=np.int64), xs=xs, cuts=cuts,
                        pisco_energy=energy.detach().numpy(), pisco_grad=grad.detach().numpy()[:, :n])
    print("wrote", path, "cuts", cuts[:5], "energy", energy[:5].tolist())


def main():
    cases = graph_cases()
    if "wcut" in sys.argv[1:] or len(sys.argv) == 1:
        weighted_cut_case("ba100pm1", weighted(cases["ba100"], [-1, 1], 5), 40, 21)
        weighted_cut_case("hub50w3", weighted(cases["hub50"], [-3, -2, -1, 1, 2, 3], 6), 33, 22)
        weighted_cut_case("multi67w7", weighted(cases["multi67"], [-7, -4, 1, 2, 5, 7], 7), 65, 23)
    if "stale" in sys.argv[1:] or len(sys.argv) == 1:
        stale_case("ba100", cases["ba100"], 48, seed=101)
        stale_case("multi67", cases["multi67"], 35, seed=102)
    if "pisco" in sys.argv[1:] or len(sys.argv) == 1:
        # PISCO_maxcut with a weighted adjacency (rlsolver/envs/env_ISCO.py:365-448): Gset-style +-1 weights and
        # small integer weights; same recording as tools/make_goldens_isco.py (file names pisco_*w*.npz)
        import make_goldens_isco as mgi
        mgi.run_case("pisco", "ba100wpm1", weighted(cases["ba100"], [-1, 1], 5), 6, 8, 13)
        mgi.run_case("pisco", "hub50w3", weighted(cases["hub50"], [-3, -2, -1, 1, 2, 3], 6), 5, 6, 14)


if __name__ == "__main__":
    main()
