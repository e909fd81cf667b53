"""BASELINE config 3 on one GPU's share: G70-shaped graph, 16384 envs, LocalSearch.random_search(num_iters=64,
num_spin=4) as in env_MCPG.py:449-476, fused RNG (64 mask arrays) against the explicit-noise path from the same seed."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as th  # noqa: E402
from synth import gset_like  # noqa: E402

from rlsolver_b200.envs.env_L2A import EnvMaxcut  # noqa: E402
from rlsolver_b200.methods.LocalSearch import LocalSearch  # noqa: E402

dev = th.device("cuda:0")
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
out = []
for fused in (True, False):
    sim = EnvMaxcut(mygraph=gset_like("G70"), device=dev, if_bidirectional=False)
    sim.fused_rng = fused
    th.manual_seed(7)
    ls = LocalSearch(sim, sim.num_nodes)
    ls.reset(sim.generate_xs_randomly(envs))
    th.cuda.synchronize()
    t = time.time()
    xs, vs, _ = ls.random_search(num_iters=64, num_spin=4)
    th.cuda.synchronize()
    dt = time.time() - t
    out.append((xs.clone(), vs.clone(), th.cuda.get_rng_state(dev)))
    print(f"fused_rng={fused}: random_search(64) on {envs} envs {dt * 1e3:.1f} ms, best cut {int(vs.max())}, "
          f"peak memory {th.cuda.max_memory_allocated(dev) / 2**30:.1f} GiB")
    del sim, ls
    th.cuda.empty_cache()
    th.cuda.reset_peak_memory_stats(dev)
print("identical:", th.equal(out[0][0], out[1][0]) and th.equal(out[0][1], out[1][1]) and th.equal(out[0][2], out[1][2]))
