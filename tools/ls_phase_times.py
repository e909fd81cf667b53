"""Phase timing of the pipelined local-search kernel (CTA 0): RLSB_LS_TIMES=1 python tools/ls_phase_times.py"""
import ctypes as C
import os
import sys

os.environ["RLSB_LS_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch as th  # noqa: E402
from synth import gset_like  # noqa: E402

import rlsolver_b200  # noqa: E402
from rlsolver_b200.envs.env_L2A import EnvMaxcut  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "G22"
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = th.device("cuda:0")
sim = EnvMaxcut(mygraph=gset_like(name), device=dev, if_bidirectional=True)
xs = sim.generate_xs_randomly(envs)
n_nodes = sim.num_nodes


def report(tag):
    th.cuda.synchronize()
    buf = (C.c_int64 * 64)()
    rlsolver_b200.lib().rlsb_ls_debug_times(buf)
    t = [int(v) for v in buf]
    waits = t[40:]
    chunks = t[22:38]
    t = t[:22]
    n = max(i for i, v in enumerate(t) if v) + 1
    t = t[:n]
    d = [t[i] - t[i - 1] for i in range(1, n)]
    print(tag, "total", t[-1] - t[0], "cycles; phases:", d, "| data waits:", [w for w in waits if w],
          "| pass-1 chunk deltas:", [chunks[i] - chunks[i - 1] for i in range(1, 16)], "first chunk after pass start stamp:",
          chunks[0] - t[2] if n > 3 else None)


for _ in range(3):
    sim.local_search_inplace(xs.clone(), th.empty(()))
report("fused thresh+8 iters+finish:")
st = sim.store
ws = st.ls_workspace(envs)
x2 = xs.clone()
vs = st.ls_begin(x2, None, 1, 0.3, ws)
nz = [th.randn((envs, n_nodes), device=dev) for _ in range(9)]
for _ in range(2):
    st.ls_thresh(envs, 1, nz[0], 8, ws)
    st.ls_search(vs, 1, nz[1:], True, x2, ws)
report("separate thresh, then 8 iters+finish:")
for _ in range(2):
    st.ls_run(vs, 1, nz[0], 8, nz[1:3], False, None, ws)
report("fused thresh + 2 iters:")
for _ in range(2):
    st.ls_run(vs, 1, nz[0], 8, [], False, None, ws)
report("thresh only:")
