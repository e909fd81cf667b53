"""Generate tests/golden/maxcut_*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only:  python tools/make_goldens.py
Every RNG draw the reference makes inside the recorded calls is captured (by
wrapping torch.randn_like / torch.randperm for the duration of the call) so the
oracle and the CUDA path can replay exactly the same flip sequence.
"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.envs import env_L2A  # noqa: E402
from rlsolver.methods import LocalSearch as ref_ls  # noqa: E402
from rlsolver.methods import util as ref_util  # noqa: E402
from rlsolver.methods import util_read_data as ref_rd  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class Recorder:
    """Temporarily wrap a torch RNG entry point and keep what it returned."""

    def __init__(self, name):
        self.name, self.draws = name, []

    def __enter__(self):
        self.orig = getattr(th, self.name)

        def wrapped(*a, **k):
            out = self.orig(*a, **k)
            self.draws.append(out.clone().numpy())
            return out

        setattr(th, self.name, wrapped)
        return self

    def __exit__(self, *exc):
        setattr(th, self.name, self.orig)


def graph_cases():
    ref_data = os.path.join(ref_import.REF, "rlsolver", "data")
    ba = ref_rd.read_mygraph(os.path.join(ref_data, "syn_BA", "BA_100_ID0.txt"))
    toy = ref_rd.read_mygraph(os.path.join(ref_data, "gset", "gset_14.txt"))
    rng = np.random.default_rng(74)
    # random multigraph: duplicates + one self loop, every node touched, node ids dense
    n, m = 67, 180
    pairs = [(i, (i + 1) % n) for i in range(n)]
    while len(pairs) < m:
        a, b = rng.integers(0, n, 2)
        if a != b:
            pairs.append((int(min(a, b)), int(max(a, b))))
    pairs += pairs[5:9]            # 4 duplicate edges
    pairs.append((11, 11))         # self loop
    multi = [(a, b, 1) for a, b in pairs]
    # degree > 32 hub to exercise multi-chunk neighbour loops
    hub = [(0, j, 1) for j in range(1, 50)] + [(j, j + 1, 1) for j in range(1, 49)] + [(3, 40, 1), (7, 22, 1)]
    return {"ba100": ba, "toy14": toy, "multi67": multi, "hub50": hub}


def run_case(name, mygraph, bidir, num_envs, seed):
    th.manual_seed(seed)
    sim = env_L2A.EnvMaxcut(mygraph=mygraph, if_bidirectional=bidir)
    n = sim.num_nodes
    out = {
        "edges": np.asarray(mygraph, dtype=np.int64), "bidirectional": np.asarray(bidir),
        "num_nodes": np.asarray(n), "num_edges": np.asarray(sim.num_edges),
        "n0_ids": sim.n0_ids[0].numpy().copy(), "n1_ids": sim.n1_ids[0].numpy().copy(),
        "n0_num_n1": sim.n0_num_n1[0].numpy().copy(),
        "adjacency_bool": sim.adjacency_bool.numpy().copy(),
    }
    xs = sim.generate_xs_randomly(num_envs)
    out["xs"] = xs.numpy().copy()
    out["cut"] = sim.calculate_obj_values(xs).numpy().copy()
    if not bidir:
        out["cut_edges"] = sim.calculate_obj_values(xs, if_sum=False).numpy().copy()
    out["loop_nosum"] = sim.calculate_obj_values_for_loop(xs, if_sum=False).numpy().copy()
    out["loop_sum"] = sim.calculate_obj_values_for_loop(xs, if_sum=True).numpy().copy()

    num_spin = min(8, max(1, n // 8))
    # local_search_inplace with the () sentinel
    gx = xs.clone()
    with Recorder("randn_like") as rec:
        gx, gv = sim.local_search_inplace(gx, th.empty(()), num_iters=4, num_spin=num_spin, noise_std=0.3)
    out["ls_noise"] = np.stack(rec.draws)
    out["ls_num_iters"], out["ls_num_spin"] = np.asarray(4), np.asarray(num_spin)
    out["ls_xs"], out["ls_vs"] = gx.numpy().copy(), gv.numpy().copy()
    # a second call from the improved state with explicit good_vs
    gx2 = gx.clone()
    with Recorder("randn_like") as rec:
        gx2, gv2 = sim.local_search_inplace(gx2, gv.clone(), num_iters=3, num_spin=num_spin, noise_std=0.5)
    out["ls2_noise"] = np.stack(rec.draws)
    out["ls2_xs"], out["ls2_vs"] = gx2.numpy().copy(), gv2.numpy().copy()

    # LocalSearch.reset + random_search twice.  With a bidirectional simulator the reference
    # itself raises (float32 prev_vs vs int64 vs in index_put, LocalSearch.py:24) -> recorded as such.
    solver = ref_ls.LocalSearch(simulator=sim, num_nodes=n)
    xs_r = sim.generate_xs_randomly(num_envs)
    out["rs_xs0"] = xs_r.numpy().copy()
    vs_r = solver.reset(xs_r)
    out["rs_vs0"] = vs_r.numpy().copy()
    if bidir:
        try:
            solver.random_search(num_iters=1, num_spin=num_spin, noise_std=0.3)
            out["rs_raises"] = np.asarray(False)
        except RuntimeError:
            out["rs_raises"] = np.asarray(True)
    else:
        for tag, iters in (("rs1", 3), ("rs2", 2)):
            with Recorder("randn_like") as rec:
                rx, rv, nu = solver.random_search(num_iters=iters, num_spin=num_spin, noise_std=0.3)
            out[f"{tag}_noise"] = np.stack(rec.draws)
            out[f"{tag}_xs"], out[f"{tag}_vs"] = rx.numpy().copy(), rv.numpy().copy()
            out[f"{tag}_num_update"] = np.asarray(nu)
            out[f"{tag}_iters"] = np.asarray(iters)

    # select ops
    a_xs, b_xs = sim.generate_xs_randomly(num_envs), sim.generate_xs_randomly(num_envs)
    a_vs, b_vs = sim.calculate_obj_values(a_xs), sim.calculate_obj_values(b_xs)
    out["upd_xs0"], out["upd_vs0"] = a_xs.numpy().copy(), a_vs.numpy().copy()
    out["upd_xs1"], out["upd_vs1"] = b_xs.numpy().copy(), b_vs.numpy().copy()
    ret = ref_rd.update_xs_by_vs(a_xs, a_vs, b_xs, b_vs, if_maximize=True)
    out["upd_ret"] = np.asarray(ret)
    out["upd_xs_out"], out["upd_vs_out"] = a_xs.numpy().copy(), a_vs.numpy().copy()

    reps = 4
    sims = num_envs // reps
    p_xs = sim.generate_xs_randomly(reps * sims)
    p_vs = sim.calculate_obj_values(p_xs)
    pk_xs, pk_vs = ref_rd.pick_xs_by_vs(p_xs, p_vs, num_repeats=reps, if_maximize=True)
    out["pick_xs"], out["pick_vs"], out["pick_reps"] = p_xs.numpy().copy(), p_vs.numpy().copy(), np.asarray(reps)
    out["pick_xs_out"], out["pick_vs_out"] = pk_xs.numpy().copy(), pk_vs.numpy().copy()

    # evolutionary_replacement on distinct values (argsort ties are unspecified)
    e_xs = sim.generate_xs_randomly(num_envs)
    e_vs = th.randperm(num_envs) + 100
    out["evo_xs"], out["evo_vs"] = e_xs.numpy().copy(), e_vs.numpy().copy()
    low_k = max(1, num_envs // 8)
    with Recorder("randperm") as rec:
        ref_util.evolutionary_replacement(e_xs, e_vs, low_k=low_k, if_maximize=True)
    out["evo_perm"], out["evo_low_k"] = rec.draws[0], np.asarray(low_k)
    out["evo_xs_out"], out["evo_vs_out"] = e_xs.numpy().copy(), e_vs.numpy().copy()

    path = os.path.join(OUT, f"maxcut_{name}_{'bi' if bidir else 'uni'}_E{num_envs}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = graph_cases()
    plan = [("ba100", True, 37), ("ba100", False, 64), ("toy14", True, 40), ("toy14", False, 33),
            ("multi67", True, 64), ("multi67", False, 41), ("hub50", True, 36), ("hub50", False, 32)]
    for i, (name, bidir, e) in enumerate(plan):
        run_case(name, cases[name], bidir, e, seed=74 + i)


if __name__ == "__main__":
    main()
