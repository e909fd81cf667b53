"""oracle/qubo.py against the reference-generated fixtures (tools/make_goldens_qubo.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import golden_files
from oracle import qubo as oq

RTOL = 1e-5      # BASELINE.json: float-weighted QUBO within 1e-5 relative


@pytest.mark.parametrize("path", golden_files("qubo_"), ids=os.path.basename)
def test_energy_matches_reference(path):
    z = np.load(path)
    x = z["best"] if bool(z["binary"]) else 2 * z["best"] - 1          # sampling.py:345 returns (samples2 + 1) / 2
    e = oq.energy(z["Q"], x)
    assert np.allclose(e, z["max_res"], rtol=RTOL, atol=RTOL * oq.scale(z["Q"]))


@pytest.mark.parametrize("path", golden_files("qubo_"), ids=os.path.basename)
def test_oracle_sweeps_match_reference(path):
    """oracle.qubo.sweeps restates sampling.py:331-337 / 356-362: same post-sweep samples as the reference
    produced for every chain (float32 and float64 arithmetic agree on these fixtures)."""
    z = np.load(path)
    binary = bool(z["binary"])
    x0 = z["raw"] if binary else 2 * z["raw"] - 1
    want = z["all_samples"] if binary else 2 * z["all_samples"] - 1
    for dtype in (np.float32, np.float64):
        assert np.array_equal(oq.sweeps(z["Q"], x0, int(z["num_ls"]), binary, dtype=dtype), want)
    assert np.allclose(oq.energy(z["Q"], want), z["all_res"], rtol=1e-5, atol=1e-5 * oq.scale(z["Q"]))
