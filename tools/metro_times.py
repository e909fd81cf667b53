"""Kernel-level times of metro_sampling (split form) under the torch profiler-free CUDA-event timer.
Usage: python tools/metro_times.py [N] [C] [max_transfer]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402

import rlsolver_b200.methods.MCPG as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
c = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mt = int(sys.argv[3]) if len(sys.argv) > 3 else n // 10
dev = th.device("cuda:0")
th.manual_seed(1)
probs = th.rand(n, device=dev) * 0.6 + 0.2
start = (th.rand(n, c, device=dev) < 0.5).float()
for split in (True, False):
    M._METRO_SPLIT_MAX_BYTES = (8 << 30) if split else -1
    for _ in range(3):
        M.metro_sampling(probs, start, mt, dev)
    ts = []
    for _ in range(10):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        M.metro_sampling(probs, start, mt, dev)
        b.record()
        th.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"N={n} C={c} max_transfer={mt} split={split}: {ts[len(ts) // 2] * 1e3:.1f} us")
