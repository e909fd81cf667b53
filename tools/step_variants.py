"""One captured local-search step (G22-shaped x 4096, the bench's headline) with the threshold draw in line on the one
stream, and on a second stream next to the begin kernel.  Both replay from the same generator state: the results must
be identical.
python tools/step_variants.py [G22:4096 ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as th  # noqa: E402
from synth import gset_like  # noqa: E402

from rlsolver_b200.envs.env_L2A import EnvMaxcut  # noqa: E402

dev = th.device("cuda:0")
flush = th.empty(256 << 20, dtype=th.uint8, device=dev)


def timed(fn, reps=40):
    out = []
    for _ in range(reps):
        flush.zero_()
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        th.cuda.synchronize()
        out.append(a.elapsed_time(b))
    out.sort()
    return out[len(out) // 2], out[0]


for spec in (sys.argv[1:] or ["G22:4096", "G70:16384"]):
    name, envs = spec.split(":")
    envs = int(envs)
    sim = EnvMaxcut(mygraph=gset_like(name), device=dev, if_bidirectional=name != "G70")
    spin = 8 if name != "G70" else 4
    th.manual_seed(74)
    xs0 = sim.generate_xs_randomly(envs)
    sentinel = th.tensor(0, device=dev)
    xs = xs0.clone()
    sim.local_search_inplace(xs, sentinel, 8, spin, 0.3)          # eager: creates the noise buffer / second stream
    th.cuda.synchronize()
    results = {}
    for tag, overlap in (("in line", False), ("second stream", True)):
        sim.overlap_threshold_draw = overlap
        th.manual_seed(75)
        xs.copy_(xs0)
        sim.store.rng_cursor_sync()
        g = th.cuda.CUDAGraph()
        with th.cuda.graph(g):
            out = sim.local_search_inplace(xs, sentinel, 8, spin, 0.3)
        th.cuda.synchronize()
        g.replay()
        th.cuda.synchronize()
        results[tag] = (out[0].clone(), out[1].clone())
        med, best = timed(g.replay)
        print(f"{name} x {envs}, threshold draw {tag:14s}: {med * 1e3:7.1f} us median, {best * 1e3:7.1f} us best", flush=True)
        del g
    ref = results["in line"]
    print("   results equal:", all(bool(th.equal(ref[0], r[0]) and th.equal(ref[1], r[1])) for r in results.values()), flush=True)
