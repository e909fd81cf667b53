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
