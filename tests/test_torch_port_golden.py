"""The torch restatement (oracle/torch_port.py: CPU-baseline arm and same-seed checker) against
the reference-generated fixtures, replaying the recorded RNG draws.  CPU only."""
import os

import numpy as np
import pytest
import torch as th

from oracle import torch_port as tp
from conftest import golden_files
from synth import replay

FILES = golden_files("maxcut_")


def _sim(z):
    edges = [tuple(int(t) for t in row) for row in z["edges"]]
    return tp.TorchSim(edges, bool(z["bidirectional"]))


@pytest.mark.parametrize("path", FILES, ids=os.path.basename)
def test_objective_and_local_search(path):
    z = np.load(path)
    sim = _sim(z)
    xs = th.from_numpy(z["xs"].copy())
    assert np.array_equal(sim.objective(xs).numpy(), z["cut"])
    a = sim.objective_for_loop(xs, if_sum=False).numpy()
    assert a.dtype == z["loop_nosum"].dtype and np.array_equal(a, z["loop_nosum"])
    with replay("randn_like", list(z["ls_noise"])):
        gx, gv = sim.local_search_inplace(xs, None, int(z["ls_num_iters"]), int(z["ls_num_spin"]), 0.3)
    assert np.array_equal(gx.numpy(), z["ls_xs"]) and np.array_equal(gv.numpy(), z["ls_vs"])


@pytest.mark.parametrize("path", [p for p in FILES if "_uni_" in p], ids=os.path.basename)
def test_random_search(path):
    z = np.load(path)
    sim = _sim(z)
    solver = tp.TorchLocalSearch(sim)
    solver.reset(th.from_numpy(z["rs_xs0"].copy()))
    for tag in ("rs1", "rs2"):
        with replay("randn_like", list(z[f"{tag}_noise"])):
            xs, vs, _ = solver.random_search(int(z[f"{tag}_iters"]), int(z["ls_num_spin"]), 0.3)
        assert np.array_equal(xs.numpy(), z[f"{tag}_xs"]) and np.array_equal(vs.numpy(), z[f"{tag}_vs"])
