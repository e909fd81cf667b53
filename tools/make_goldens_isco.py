"""Generate tests/golden/isco_*.npz by running the UNMODIFIED reference on CPU:
ISCO_maxcut / PISCO_maxcut .step trajectories (rlsolver/envs/env_ISCO.py:10-91, 365-448;
rlsolver/methods/ISCO/util.py) with every torch.rand draw recorded so the CUDA mirror can
replay them.  Build container only:  python tools/make_goldens_isco.py"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.envs import env_ISCO  # noqa: E402
from rlsolver.methods.ISCO import util as isco_util  # noqa: E402
from rlsolver.methods.ISCO import util_maxcut  # noqa: E402

from make_goldens import graph_cases  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def write_graph(path, n, edges):
    with open(path, "w") as f:
        f.write(f"{n} {len(edges)}\n")
        for a, b, w in edges:
            f.write(f"{a + 1} {b + 1} {w}\n")


def run_case(kind, name, mygraph, batch, steps, seed):
    n = len({a for a, _, _ in mygraph} | {b for _, b, _ in mygraph})
    tmp = f"/tmp/isco_{name}.txt"
    write_graph(tmp, n, mygraph)
    for mod in (env_ISCO, util_maxcut):
        mod.BATCH_SIZE = batch
        mod.DEVICE = th.device("cpu")
    params = util_maxcut.load_data(tmp)
    cls = env_ISCO.ISCO_maxcut if kind == "isco" else env_ISCO.PISCO_maxcut
    sampler = cls(params)
    th.manual_seed(seed)
    if kind == "isco":
        x = sampler.random_gen_init_sample(params)
    else:
        x = sampler.random_gen_init_sample()
        pad = ((n + 7) // 8 * 8) - n
        x = th.nn.functional.pad(x, (0, pad), mode="constant", value=0)
    draws = []
    orig_rand = th.rand

    def rec(*a, **k):
        t = orig_rand(*a, **k)
        draws.append(t.numpy().copy().ravel())
        return t

    xs, energies, accs, paths, temps = [x.float().numpy().copy()], [], [], [], []
    isco_util.torch.rand = rec
    th.rand = rec
    try:
        for step in range(steps):
            path_length = th.randint(1, 6, (batch,))
            temperature = th.tensor(1.0 - 0.9 * step / steps)
            x, energy, acc = sampler.step(x, path_length, temperature)
            xs.append(x.float().numpy().copy()), energies.append(energy.float().numpy().copy())
            accs.append(acc.float().numpy().copy()), paths.append(path_length.numpy().copy())
            temps.append(float(temperature))
    finally:
        th.rand = orig_rand
    ef, et = params["edge_from"].numpy(), params["edge_to"].numpy()
    adj = params["adj_matrix"].float().numpy()
    out = dict(edge_from=ef, edge_to=et, edge_w=adj[ef, et].astype(np.int64), num_nodes=np.asarray(n), xs=np.stack(xs), energies=np.stack(energies),
               accs=np.stack(accs), paths=np.stack(paths), temps=np.asarray(temps, dtype=np.float64),
               draws=np.concatenate(draws).astype(np.float32), draw_sizes=np.asarray([d.size for d in draws]))
    path = os.path.join(OUT, f"{kind}_{name}_B{batch}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "energies", energies[-1][:4], "acc", accs[-1][:4])


if __name__ == "__main__":
    cases = graph_cases()
    print(list(cases))
    for kind in ("isco", "pisco"):
        run_case(kind, "ba100", cases["ba100"], 6, 8, 3)
        run_case(kind, "toy14", cases["toy14"], 3, 6, 4)
