"""Generate tests/golden/ppo_*.npz and greedy_*.npz by running the UNMODIFIED reference on CPU:
env_PPO.EnvMaxcut reset/step trajectories (rlsolver/envs/env_PPO.py:63-126) and greedy_maxcut
(rlsolver/methods/greedy.py:33-78).  Build container only:  python tools/make_goldens_ppo.py"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.envs import env_PPO  # noqa: E402
from rlsolver.methods import util_read_data as ref_rd  # noqa: E402

from make_goldens import graph_cases  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class Args:
    def __init__(self, n, e, steps):
        self.num_nodes, self.num_envs, self.num_steps = n, e, steps


def ppo_case(name, mygraph, bidir, num_envs, seed, steps=13, num_steps=5):
    th.manual_seed(seed)
    n = len({a for a, _, _ in mygraph} | {b for _, b, _ in mygraph})
    env = env_PPO.EnvMaxcut(Args(n, num_envs, num_steps), mygraph=mygraph, if_bidirectional=bidir)
    xs0 = env.reset().clone()
    out = {"edges": np.asarray(mygraph, dtype=np.int64), "bidirectional": np.asarray(bidir), "num_nodes": np.asarray(n),
           "num_steps": np.asarray(num_steps), "xs0": xs0.numpy().copy(), "cut0": env.last_reward.numpy().copy()}
    acts, rews, dones, curs, obs = [], [], [], [], []
    for _ in range(steps):
        action = th.randint(0, n, (num_envs,))
        xs, reward, done, cur = env.step(action)
        acts.append(action.numpy().copy()), rews.append(reward.numpy().copy())
        dones.append(done.numpy().copy()), curs.append(cur.numpy().copy()), obs.append(xs.numpy().copy())
    out.update(actions=np.stack(acts), rewards=np.stack(rews), dones=np.stack(dones), cuts=np.stack(curs),
               obs=np.stack(obs))
    path = os.path.join(OUT, f"ppo_{name}_{'bi' if bidir else 'uni'}_E{num_envs}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def greedy_case(name, mygraph):
    import networkx as nx
    try:
        import matplotlib  # noqa: F401
    except ImportError:         # greedy.py -> util.py imports pyplot at module level; never used here
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
    try:
        from rlsolver.methods import greedy as ref_greedy
    except Exception as exc:                                   # heavy optional imports in greedy.py
        print("reference greedy import failed:", exc)
        raise
    n = len({a for a, _, _ in mygraph} | {b for _, b, _ in mygraph})
    g = nx.Graph()
    g.add_nodes_from(range(n))
    for a, b, w in mygraph:
        g.add_edge(a, b, weight=w)
    with contextlib.redirect_stdout(io.StringIO()):
        score, solution, scores = ref_greedy.greedy_maxcut(None, g)
    path = os.path.join(OUT, f"greedy_{name}.npz")
    np.savez_compressed(path, edges=np.asarray(mygraph, dtype=np.int64), num_nodes=np.asarray(n),
                        score=np.asarray(int(score)), solution=np.asarray(solution, dtype=np.int64),
                        scores=np.asarray([int(s) for s in scores], dtype=np.int64))
    print("wrote", path, "score", score, "flips", len(scores))


def main():
    cases = graph_cases()
    ppo_case("ba100", cases["ba100"], True, 37, 101)
    ppo_case("ba100", cases["ba100"], False, 64, 102)
    ppo_case("toy14", cases["toy14"], True, 33, 103)
    ppo_case("hub50", cases["hub50"], True, 40, 104)
    simple = {k: v for k, v in cases.items() if k in ("ba100", "toy14", "hub50")}      # greedy needs a simple graph
    for name, g in simple.items():
        greedy_case(name, g)


if __name__ == "__main__":
    main()
