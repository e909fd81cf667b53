"""Quick timing of the tensor-core QUBO Hamiltonian (config 5 shape: N=4096 dense float Q)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from rlsolver_b200.qubo import QuboModel

dev = th.device("cuda:0")
n = 4096
th.manual_seed(0)
u = th.randn(n, n, device=dev)
q = th.triu(u) + th.triu(u, 1).T
model = QuboModel(q)
for c in (1024, 8192):
    x = (th.randint(0, 2, (n, c), device=dev).float() * 2 - 1)
    for _ in range(3):
        e = model.energy(x)
    th.cuda.synchronize()
    evs = []
    for _ in range(10):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record(); e = model.energy(x); b.record(); evs.append((a, b))
    th.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    flop = 2.0 * n * n * c
    # torch fp32 reference on the same GPU
    for _ in range(2):
        r = (x * (q @ x)).sum(0)
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record(); r = (x * (q @ x)).sum(0); b.record(); th.cuda.synchronize()
    ref_ms = a.elapsed_time(b)
    err = ((e.double() - (x.double() * (q.double() @ x.double())).sum(0)).abs().max() / q.double().pow(2).sum().sqrt()).item()
    print(json.dumps({"N": n, "C": c, "ms": ms, "useful_TFLOPs": flop / ms / 1e9, "bf16_TFLOPs_3limb": 3 * flop / ms / 1e9,
                      "torch_fp32_ms": ref_ms, "max_err_over_scale": err}))

# sweeps (one Gauss-Seidel pass = N coordinate updates per chain): K split over several CTAs per chain group vs not
from rlsolver_b200 import _lib
for c in (1024, 2048, 8192):
    x = (th.randint(0, 2, (n, c), device=dev).float() * 2 - 1)
    row = {"N": n, "C": c}
    for tag, flag in (("split_k", 0), ("single_cta", _lib.DEBUG_QUBO_NO_SPLITK)):
        _lib.debug_flags(flag, _lib.DEBUG_QUBO_NO_SPLITK ^ flag)
        xs_ = x.clone()
        model.sweeps(xs_, 1)
        th.cuda.synchronize()
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            model.sweeps(xs_, 1)
        b.record()
        th.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        row[tag + "_ms"] = ms
        row[tag + "_bf16_TFLOPs_3limb"] = 3 * 2.0 * n * n * c / ms / 1e9
    _lib.debug_flags(0, _lib.DEBUG_QUBO_NO_SPLITK)
    print(json.dumps(row))
