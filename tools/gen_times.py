"""Mask generator alone: the one-draw-per-thread form (round 1) against the two-draws-per-thread group form
(round 2), same masks.  python tools/gen_times.py [G22:4096 ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as th  # noqa: E402
from synth import gset_like  # noqa: E402

from rlsolver_b200 import _lib, rng  # noqa: E402
from rlsolver_b200.envs.env_L2A import EnvMaxcut  # noqa: E402

dev = th.device("cuda:0")
flush = th.empty(256 << 20, dtype=th.uint8, device=dev)
GEN_PER_DRAW = 32


def timed(fn, reps=20):
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
    return out[len(out) // 2]


for spec in (sys.argv[1:] or ["G22:4096:8", "G70:16384:8", "G14:256:8", "G70:16384:64"]):
    name, envs, draws = spec.split(":")
    envs, draws = int(envs), int(draws)
    sim = EnvMaxcut(mygraph=gset_like(name), device=dev, if_bidirectional=name != "G70")
    st, n = sim.store, sim.num_nodes
    th.manual_seed(74)
    xs = sim.generate_xs_randomly(envs)
    ws = st.ls_workspace(envs)
    vs = st.ls_begin(xs, None, 1, 0.3, ws)
    st.ls_run(vs, 1, th.randn((envs, n), device=dev), 8 if name != "G70" else 4, [], False, None, ws)
    seed, offset, threads, iters = rng.peek(dev, envs * n)
    res = {}
    for tag, flag in (("per-draw", GEN_PER_DRAW), ("group", 0)):
        _lib.debug_flags(flag, GEN_PER_DRAW ^ flag if flag == 0 else 0)
        m = st.ls_noise_masks(envs, 1, draws, seed, offset, threads, iters, ws).clone()
        t = timed(lambda: st.ls_noise_masks(envs, 1, draws, seed, offset, threads, iters, ws))
        t_re = timed(lambda: st.ls_noise_masks(envs, 1, draws, seed, offset, threads, iters, ws, reuse_bound=True))
        res[tag] = (m, t, t_re)
        print(f"{name} x {envs}, {draws} draws, {tag:8s}: {t * 1e3:8.1f} us with the early-out pass, {t_re * 1e3:8.1f} us generator + memset only", flush=True)
    _lib.debug_flags(0, GEN_PER_DRAW)
    print("   masks equal:", bool(th.equal(res["per-draw"][0], res["group"][0])), flush=True)
