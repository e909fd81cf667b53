"""Tensor-core QUBO Hamiltonian (csrc/qubo.cu, tcgen05) against the reference fixtures and the
float64 oracle.  Tolerance: 1e-5 relative (BASELINE.json), absolute floor 1e-5 * ||Q||_F."""
import os

import numpy as np
import pytest
import torch as th

from conftest import golden_files
from oracle import qubo as oq

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("path", golden_files("qubo_"), ids=os.path.basename)
def test_qubo_golden(path, cuda_device):
    from rlsolver_b200.qubo import QuboModel, qubo_values
    z = np.load(path)
    model = QuboModel(th.from_numpy(z["Q"]).to(cuda_device))
    x = z["best"] if bool(z["binary"]) else 2 * z["best"] - 1
    e = model.energy(th.from_numpy(x.astype(np.float32)).to(cuda_device))
    assert np.allclose(_np(e), z["max_res"], rtol=RTOL, atol=RTOL * oq.scale(z["Q"]))
    if "int" in path:                                   # integer Q: every partial sum is exact in fp32
        assert np.array_equal(_np(e), z["max_res"])


@pytest.mark.parametrize("n,c,kind", [(128, 256, "pm1"), (200, 300, "bin"), (1000, 777, "pm1"), (4096, 1024, "pm1")])
def test_qubo_vs_float64(n, c, kind, cuda_device):
    from rlsolver_b200.qubo import QuboModel
    rng = np.random.default_rng(n + c)
    u = rng.standard_normal((n, n)).astype(np.float32)
    q = (np.triu(u) + np.triu(u, 1).T).astype(np.float32)
    x = rng.integers(0, 2, (n, c)).astype(np.float32)
    if kind == "pm1":
        x = 2 * x - 1
        x[3, :] = 0                                       # a masked variable, as inside the reference's sweep
    model = QuboModel(th.from_numpy(q).to(cuda_device))
    e = _np(model.energy(th.from_numpy(x).to(cuda_device)))
    want = oq.energy(q, x)
    err = np.abs(e - want)
    assert (err <= RTOL * np.maximum(np.abs(want), oq.scale(q))).all(), float((err / oq.scale(q)).max())
    # linearity in Q (size-independent property): E(2Q) = 2 E(Q) exactly (power-of-two scaling)
    e2 = _np(QuboModel(th.from_numpy(2 * q).to(cuda_device)).energy(th.from_numpy(x).to(cuda_device)))
    assert np.array_equal(e2, 2 * e)
    # deterministic
    assert np.array_equal(_np(model.energy(th.from_numpy(x).to(cuda_device))), e)
