"""Generate tests/golden/relaxed_*.npz by running the UNMODIFIED reference on CPU:
SimulatorMaxcut.get_objectives / get_objectives_using_for_loop / get_scores (rlsolver/envs/env_k_spin.py:164-197),
the gradient autograd derives from get_objectives, and PIGNN hamiltonian_maxcut (rlsolver/methods/PIGNN/util.py:4-8).
Build container only:  python tools/make_goldens_relaxed.py"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.envs import env_k_spin  # noqa: E402

pignn = ref_import.load_by_path("ref_pignn_util", "rlsolver/methods/PIGNN/util.py")

from make_goldens import graph_cases  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def case(name, mygraph, num_envs, seed):
    # SimulatorMaxcut asserts a simple directed edge set (its adjacency matrix keeps one entry per pair)
    seen, graph = set(), []
    for a, b, w in mygraph:
        if (a, b) in seen or a == b:
            continue
        seen.add((a, b))
        graph.append((a, b, w))
    n = len({a for a, _, _ in graph} | {b for _, b, _ in graph})
    sim = env_k_spin.SimulatorMaxcut(graph_name=name, gpu_id=-1, graph_tuple=(graph, n, len(graph)))
    th.manual_seed(seed)
    probs = sim.get_rand_probs(num_envs).requires_grad_(True)
    obj = sim.get_objectives(probs)
    gout = th.randn(num_envs)
    (obj * gout).sum().backward()
    loop = sim.get_objectives_using_for_loop(probs.detach())
    xs = sim.prob_to_bool(probs.detach())
    ham = pignn.hamiltonian_maxcut((sim.n0_ids[0].long(), sim.n1_ids[0].long()), probs.detach()[0])
    out = {"edges": np.asarray(graph, dtype=np.int64), "num_nodes": np.asarray(n), "probs": probs.detach().numpy(),
           "objectives": obj.detach().numpy(), "objectives_loop": loop.numpy(), "grad_out": gout.numpy(),
           "grad_probs": probs.grad.numpy(), "xs": xs.numpy(), "scores": sim.get_scores(xs).numpy(),
           "hamiltonian0": np.asarray(float(ham)), "n0_ids": sim.n0_ids[0].numpy(), "n1_ids": sim.n1_ids[0].numpy()}
    path = os.path.join(OUT, f"relaxed_{name}_E{num_envs}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "objectives[:3]", out["objectives"][:3])


if __name__ == "__main__":
    cases = graph_cases()
    for k, (name, mygraph) in enumerate(cases.items() if isinstance(cases, dict) else cases):
        case(name, mygraph, num_envs=37 + 20 * (k % 3), seed=100 + k)
