"""Per-kernel times of the fused-RNG local search (ls_begin, threshold pass, ls_noise_masks, ls_run_masks)
next to the explicit-noise path, for the BASELINE shapes.  Usage: python tools/ls_fused_times.py [G22:4096 ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch as th  # noqa: E402

from rlsolver_b200 import _lib, rng  # noqa: E402
from rlsolver_b200.envs.env_L2A import EnvMaxcut  # noqa: E402
from rlsolver_b200.graph_store import OpTimer  # noqa: E402
from synth import gset_like  # noqa: E402

dev = th.device("cuda:0")
flush = th.empty(256 << 20, dtype=th.uint8, device=dev)


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


for spec in (sys.argv[1:] or ["G22:4096", "G70:16384", "G14:256"]):
    name, envs = spec.split(":")
    envs = int(envs)
    sim = EnvMaxcut(mygraph=gset_like(name), device=dev, if_bidirectional=True)
    st, n = sim.store, sim.num_nodes
    th.manual_seed(74)
    xs0 = sim.generate_xs_randomly(envs)
    xs = xs0.clone()
    sentinel = th.empty(())
    print(f"== {name} x {envs} envs (N = {n})")
    for fused in (False, True):
        sim.fused_rng = fused

        def step():
            xs.copy_(xs0)
            sim.local_search_inplace(xs, sentinel)

        for _ in range(3):
            step()
        ms = timed(step)
        st.timer = OpTimer()
        for _ in range(10):
            flush.zero_()
            step()
        th.cuda.synchronize()
        spans = st.timer.summary()
        st.timer = None
        print(f"  fused_rng={fused}: local_search_inplace {ms * 1e3:.1f} us (eager, incl. torch randn);  "
              + ", ".join(f"{k} {v[1] / v[0] * 1e3:.1f}" for k, v in spans.items()))
    # pieces of the mask generator
    wsb = st.ls_workspace(envs)
    vs = st.ls_begin(xs0, None, 1, 0.3, wsb)
    noise0 = th.randn((envs, n), device=dev)
    st.ls_run(vs, 1, noise0, 8, [], False, None, wsb)
    seed, offset, threads, iters = rng.peek(dev, envs * n)
    t_randn = timed(lambda: th.randn((envs, n), device=dev))
    t_own = timed(lambda: st.torch_randn(envs * n, 1, seed, offset, threads, iters))
    _lib.debug_flags(_lib.DEBUG_PLAIN_MASKS, 0)
    t_plain = timed(lambda: st.ls_noise_masks(envs, 1, 8, seed, offset, threads, iters, wsb))
    _lib.debug_flags(0, _lib.DEBUG_PLAIN_MASKS)
    t_fast = timed(lambda: st.ls_noise_masks(envs, 1, 8, seed, offset, threads, iters, wsb))
    t_fast1 = timed(lambda: st.ls_noise_masks(envs, 1, 1, seed, offset, threads, iters, wsb))
    masks = st.ls_noise_masks(envs, 1, 8, seed, offset, threads, iters, wsb)
    off = int(st._lib.rlsb_ls_workspace_offset(st._h, envs, 7))
    words = st.ls_mask_words(envs)
    flips = sum(int(x) for x in (masks.view(th.uint8).cpu().numpy().reshape(8, -1)[:, :(envs * n + 7) // 8] != 0).sum(axis=1))
    print(f"  torch.randn {t_randn * 1e3:.1f} us, own randn kernel {t_own * 1e3:.1f} us, masks x8 plain {t_plain * 1e3:.1f} us, "
          f"early-out x8 {t_fast * 1e3:.1f} us (x1 {t_fast1 * 1e3:.1f} us), nonzero mask bytes {flips}, words/draw {words}")
    t_bits = timed(lambda: st.ls_run_masks(vs, masks, 8, False, None, wsb))
    t_bits_f = timed(lambda: st.ls_run_masks(vs, masks, 0, True, xs, wsb))
    print(f"  ls_run_masks 8 iterations {t_bits * 1e3:.1f} us, finish only {t_bits_f * 1e3:.1f} us")
