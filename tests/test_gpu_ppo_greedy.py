"""CUDA path of the pattern-I env (env_PPO mirror, csrc/fields.cu step_flip) and the batched
greedy best-flip kernel against the reference-generated fixtures and the NumPy oracle."""
import os
import types

import numpy as np
import pytest
import torch as th

from conftest import golden_files
from oracle import maxcut as om
from synth import gset_like, random_graph

pytestmark = pytest.mark.gpu


def _edges(z):
    return [tuple(int(t) for t in row) for row in z["edges"]]


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("path", golden_files("ppo_"), ids=os.path.basename)
def test_env_ppo_golden_trajectory(path, cuda_device):
    from rlsolver_b200.envs.env_PPO import EnvMaxcut
    z = np.load(path)
    e = z["xs0"].shape[0]
    args = types.SimpleNamespace(num_nodes=int(z["num_nodes"]), num_envs=e, num_steps=int(z["num_steps"]))
    env = EnvMaxcut(args, mygraph=_edges(z), device=cuda_device, if_bidirectional=bool(z["bidirectional"]))
    env.reset()
    env.xs.copy_(th.from_numpy(z["xs0"]))                      # replay the reference's start state
    env.last_reward = env.calculate_obj_values().to(th.float)
    assert np.array_equal(_np(env.last_reward), z["cut0"])
    xs_obj = env.xs
    for t in range(z["actions"].shape[0]):
        xs, reward, done, cur = env.step(th.from_numpy(z["actions"][t]).to(cuda_device))
        assert xs is xs_obj and xs.dtype == th.float32
        assert np.array_equal(_np(xs), z["obs"][t]) and np.array_equal(_np(reward), z["rewards"][t])
        assert np.array_equal(_np(done), z["dones"][t]) and np.array_equal(_np(cur), z["cuts"][t])
        assert np.array_equal(_np(env.calculate_obj_values()).astype(np.float32), z["cuts"][t])
    assert env.num_bad_actions() == 0


def test_env_ppo_g22_vs_oracle(cuda_device):
    from rlsolver_b200.envs.env_PPO import EnvMaxcut
    edges = gset_like("G22")
    g = om.build_graph_store(edges, True)
    e = 4096
    args = types.SimpleNamespace(num_nodes=g.num_nodes, num_envs=e, num_steps=7)
    env = EnvMaxcut(args, mygraph=edges, device=cuda_device, if_bidirectional=True)
    th.manual_seed(3)
    xs = env.reset()
    ref = om.PPOEnv(g, 7)
    ref.reset(_np(xs) > 0)
    rng = np.random.default_rng(0)
    for t in range(9):
        a = rng.integers(0, g.num_nodes, e)
        _, reward, done, cur = env.step(th.from_numpy(a).to(cuda_device))
        _, r2, d2, c2 = ref.step(a)
        assert np.array_equal(_np(reward), r2) and np.array_equal(_np(done), d2) and np.array_equal(_np(cur), c2)
    assert np.array_equal(_np(env.xs), ref.xs)
    # out-of-range actions are flagged, not applied
    bad = th.full((e,), g.num_nodes, dtype=th.int64, device=cuda_device)
    _, reward, _, cur = env.step(bad)
    assert float(reward.abs().sum()) == 0 and np.array_equal(_np(cur), c2) and env.num_bad_actions() == e


@pytest.mark.parametrize("path", golden_files("greedy_"), ids=os.path.basename)
def test_greedy_golden(path, cuda_device):
    from rlsolver_b200.methods.greedy import greedy_maxcut
    z = np.load(path)
    score, solution, scores = greedy_maxcut(None, _edges(z), device=cuda_device)
    assert score == int(z["score"]) and solution == z["solution"].tolist() and scores == z["scores"].tolist()
    score2, _, scores2 = greedy_maxcut(3, _edges(z), device=cuda_device)       # num_steps cap
    assert scores2 == z["scores"].tolist()[:3] and score2 == scores2[-1]


@pytest.mark.parametrize("name,envs", [("G14", 256), ("G22", 96), ("R37", 45)])
def test_greedy_batched_vs_oracle(name, envs, cuda_device):
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    from rlsolver_b200.methods.greedy import greedy_maxcut_batched
    edges = random_graph(37, 90, seed=3) if name == "R37" else gset_like(name)
    g = om.build_graph_store(edges, True)
    sim = EnvMaxcut(mygraph=edges, device=cuda_device, if_bidirectional=True)
    rng = np.random.default_rng(11)
    xs_np = rng.integers(0, 2, size=(envs, g.num_nodes)).astype(bool)
    xs_np[0] = False
    want_xs, want_vs, want_flips = om.greedy_best_flip(g, xs_np)
    xs = th.from_numpy(xs_np.copy()).to(cuda_device)
    out, vs, flips = greedy_maxcut_batched(sim, xs)
    assert out is xs
    assert np.array_equal(_np(vs), want_vs) and np.array_equal(_np(flips), want_flips)
    assert np.array_equal(_np(xs), want_xs)
    assert np.array_equal(om.cut_values(g, _np(xs)), want_vs)
    # capped number of steps: exactly min(cap, flips) flips
    xs2 = th.from_numpy(xs_np.copy()).to(cuda_device)
    _, vs2, flips2 = greedy_maxcut_batched(sim, xs2, num_steps=5)
    assert np.array_equal(_np(flips2), np.minimum(want_flips, 5))
    assert np.array_equal(om.cut_values(g, _np(xs2)), _np(vs2))
