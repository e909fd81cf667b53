"""A few eager local_search_inplace steps at the bench shape (target for ncu captures).
Usage: python tools/ls_one_step.py [G22] [4096] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch as th  # noqa: E402

from rlsolver_b200.envs.env_L2A import EnvMaxcut  # noqa: E402
from synth import gset_like  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "G22"
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = th.device("cuda:0")
sim = EnvMaxcut(mygraph=gset_like(name), device=dev, if_bidirectional=True)
th.manual_seed(74)
xs0 = sim.generate_xs_randomly(envs)
for _ in range(steps):
    xs = xs0.clone()
    gx, gv = sim.local_search_inplace(xs, th.empty(()))
th.cuda.synchronize()
print("best cut", int(gv.max()))
