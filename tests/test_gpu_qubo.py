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


# ------------------------------------------------------------------ local-search sweeps (rlsb_qubo_sweeps)
@pytest.mark.parametrize("path", golden_files("qubo_"), ids=os.path.basename)
def test_qubo_sweeps_golden(path, cuda_device):
    """Blocked Gauss-Seidel on the tensor cores against the reference's N dependent GEMVs: same samples
    for every chain of the reference fixtures, and its energies / best-of-repeats / advantages."""
    from rlsolver_b200.qubo import QuboModel, qubo_values
    z = np.load(path)
    binary = bool(z["binary"])
    model = QuboModel(th.from_numpy(z["Q"]).to(cuda_device))
    x = th.from_numpy((z["raw"] if binary else 2 * z["raw"] - 1).astype(np.float32)).to(cuda_device).contiguous()
    model.sweeps(x, int(z["num_ls"]), binary=binary)
    want = z["all_samples"] if binary else 2 * z["all_samples"] - 1
    assert np.array_equal(_np(x), want)
    max_res, index, value = qubo_values(model, x, int(z["total_mcmc"]))
    tol = RTOL * oq.scale(z["Q"])
    assert np.allclose(_np(max_res), z["max_res"], rtol=RTOL, atol=tol)
    assert np.allclose(_np(value), z["value"], rtol=RTOL, atol=2 * tol)
    best = _np(x[:, index]) if binary else (_np(x[:, index]) + 1) / 2
    assert np.array_equal(best, z["best"])


@pytest.mark.parametrize("n,c,binary,integer", [(64, 128, False, True), (200, 300, True, True), (333, 130, False, False),
                                                (1000, 257, False, True), (1024, 512, True, False)])
def test_qubo_sweeps_vs_oracle(n, c, binary, integer, cuda_device):
    """Against the float64 restatement.  Integer Q: every dot product is exact, so all chains must match.
    Float Q: chains whose decisions all had a margin above the fp32 rounding error of a length-N dot
    product must match (the others may legitimately take the other branch of a near-tie)."""
    from rlsolver_b200.qubo import QuboModel
    rng = np.random.default_rng(7 * n + c)
    u = rng.integers(-40, 41, (n, n)).astype(np.float32) if integer else rng.standard_normal((n, n)).astype(np.float32)
    q = (np.triu(u) + np.triu(u, 1).T).astype(np.float32)
    x0 = rng.integers(0, 2, (n, c)).astype(np.float32)
    if not binary:
        x0 = 2 * x0 - 1
    sweeps = 2
    want = oq.sweeps(q, x0, sweeps, binary)
    x = th.from_numpy(x0.copy()).to(cuda_device)
    QuboModel(th.from_numpy(q).to(cuda_device)).sweeps(x, sweeps, binary=binary)
    same = (_np(x) == want).all(axis=0)
    if integer:
        assert same.all()
    else:
        margin = oq.sweep_margin(q, x0, sweeps, binary)
        safe = margin > 2e-5 * np.sqrt(n) * np.abs(q).max()        # ~100 x the rounding error of a dot product
        assert safe.mean() > 0.5 and same[safe].all() and same.mean() > 0.97, (float(safe.mean()), float(same.mean()))
    # a sweep never lowers the objective of a symmetric Q (coordinate ascent), any chain
    e0 = oq.energy(q, x0)
    e1 = oq.energy(q, _np(x))
    assert (e1 >= e0 - 1e-6 * oq.scale(q)).all()
