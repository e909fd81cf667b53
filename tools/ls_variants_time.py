"""CUDA-event timing of ls_run variants (threshold pass / iterations / finish), no instrumentation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch as th  # noqa: E402
from synth import gset_like  # noqa: E402

from rlsolver_b200.envs.env_L2A import EnvMaxcut  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "G22"
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = th.device("cuda:0")
sim = EnvMaxcut(mygraph=gset_like(name), device=dev, if_bidirectional=True)
xs = sim.generate_xs_randomly(envs)
st = sim.store
ws = st.ls_workspace(envs)
x2 = xs.clone()
vs = st.ls_begin(x2, None, 1, 0.3, ws)
nz = [th.randn((envs, sim.num_nodes), device=dev) for _ in range(9)]
flush = th.empty(256 << 20, dtype=th.uint8, device=dev)


def timeit(tag, fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        th.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"{tag:44s} median {ts[len(ts) // 2]:8.1f} us  min {ts[0]:8.1f} us")


timeit("begin", lambda: st.ls_begin(x2, None, 1, 0.3, ws))
timeit("thresh only (pipe)", lambda: st.ls_run(vs, 1, nz[0], 8, [], False, None, ws))
timeit("thresh only (generic kernel)", lambda: st.ls_thresh(envs, 1, nz[0], 8, ws))
for k in (1, 2, 4, 8):
    timeit(f"thresh + {k} iters", lambda k=k: st.ls_run(vs, 1, nz[0], 8, nz[1:1 + k], False, None, ws))
for k in (1, 2, 4, 8):
    timeit(f"{k} iters", lambda k=k: st.ls_search(vs, 1, nz[1:1 + k], False, None, ws))
timeit("finish only", lambda: st.ls_search(vs, 1, [], True, x2, ws))
timeit("8 iters + finish", lambda: st.ls_search(vs, 1, nz[1:9], True, x2, ws))
timeit("thresh + 8 iters + finish", lambda: st.ls_run(vs, 1, nz[0], 8, nz[1:9], True, x2, ws))
