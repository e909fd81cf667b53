"""Eager launches of the round-2 kernels at the BASELINE shapes: the target of the ncu captures under profiles/.
python tools/profile_targets.py [step|peco|config3|isco|wmcpg ...]   (default: step peco)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch as th  # noqa: E402
from synth import gset_like, random_graph  # noqa: E402

dev = th.device("cuda:0")
want = sys.argv[1:] or ["step", "peco"]

if "step" in want:       # BASELINE config 2: one dREINFORCE local-search step, 3 times
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    sim = EnvMaxcut(mygraph=gset_like("G22"), device=dev, if_bidirectional=True)
    th.manual_seed(74)
    xs0 = sim.generate_xs_randomly(4096)
    for _ in range(3):
        gx, gv = sim.local_search_inplace(xs0.clone(), th.empty(()))
    th.cuda.synchronize()
    print("step: best cut", int(gv.max()))

if "peco" in want:       # BASELINE config 4: 10^6 ER-100 instances, compact layout, 4 steps
    from rlsolver_b200.envs import env_PECO as P
    e, n = 1 << 20, 100
    th.manual_seed(74)
    gg = P.RandomERGraphGenerator(n_spins=n, p_connection=0.15, edge_type=P.EdgeType.DISCRETE, num_envs=e, device=dev)
    env = P.SpinSystemFactory.get(gg, 2 * n, observables=P.ECO_PECO_OBSERVABLES, reward_signal=P.RewardSignal.BLS,
                                  extra_action=P.ExtraAction.NONE, optimisation_target=P.OptimisationTarget.CUT,
                                  spin_basis=P.SpinBasis.BINARY, norm_rewards=True, reversible_spins=True, device=dev,
                                  num_envs=e)
    for k in range(4):
        env.step(th.randint(0, n, (e,), device=dev), return_observation=False)
    th.cuda.synchronize()
    print("peco: best", float(env.get_best_cut().max()))

if "config3" in want:    # BASELINE config 3: one random_search(64, 4) on 16384 envs of the G70 shape
    from rlsolver_b200.envs.env_MCPG import EnvMaxcut, LocalSearch
    sim = EnvMaxcut(mygraph=gset_like("G70"), device=dev, if_bidirectional=False)
    ls = LocalSearch(sim, sim.num_nodes)
    th.manual_seed(74)
    ls.reset(sim.generate_xs_randomly(16384))
    for _ in range(2):
        ls.random_search(num_iters=64, num_spin=4)
    th.cuda.synchronize()
    print("config3: best", int(ls.good_vs.max()))

if "isco" in want:       # one PISCO step, 256 chains on the G22 shape
    from rlsolver_b200.envs import env_ISCO
    from rlsolver_b200.methods.ISCO import config_maxcut as cfg
    cfg.BATCH_SIZE, cfg.DEVICE = 256, dev
    edges = gset_like("G22")
    ef = th.tensor([a for a, _, _ in edges], device=dev)
    et = th.tensor([b for _, b, _ in edges], device=dev)
    s = env_ISCO.ISCO_maxcut({"num_nodes": 2000, "num_edges": len(edges), "edge_from": ef, "edge_to": et})
    x = s.random_gen_init_sample()
    for k in range(3):
        x, en, acc = s.step(x, th.full((256,), 8, device=dev), th.tensor(0.7, device=dev))
    th.cuda.synchronize()
    print("isco: energy", float(en.max()))
