"""One sampler_func call at the config-2 shape (target for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as th  # noqa: E402
from synth import gset_like  # noqa: E402

from rlsolver_b200.methods.MCPG import McpgData, metro_sampling, sampler_func  # noqa: E402

dev = th.device("cuda:0")
edges = gset_like("G22")
n = max(max(a, b) for a, b, _ in edges) + 1
data = McpgData(edges, n, dev)
total, rep = 512, 8
th.manual_seed(3)
probs = th.rand(n, device=dev) * 0.6 + 0.2
start = (th.rand(n, total * rep, device=dev) < 0.5).float()
xs = metro_sampling(probs, start, n // 10, dev)
for _ in range(3):
    out = sampler_func(data, xs, 8, total, rep)
th.cuda.synchronize()
print("levels", data.num_levels, "best", float(out[0].min()))
