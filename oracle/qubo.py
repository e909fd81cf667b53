"""Float64 restatement of the dense QUBO objective of the reference.  TEST INFRASTRUCTURE ONLY.
rlsolver/methods/MCPG/sampling.py:339-340 / 364-365: res = Q @ X; value[c] = sum_i X[i,c] * res[i,c]
(x in {-1,+1} for mcpg_sampling_qubo, {0,1} for _qubo_bin; no 1/2 and no linear term in the code)."""
import numpy as np


def energy(q: np.ndarray, x: np.ndarray) -> np.ndarray:
    q64, x64 = q.astype(np.float64), x.astype(np.float64)
    return (x64 * (q64 @ x64)).sum(axis=0)


def scale(q: np.ndarray) -> float:
    """Standard deviation of x^T Q x over random sign vectors: the natural absolute scale of an
    energy (a relative tolerance on an energy that happens to be near zero is meaningless)."""
    return float(np.sqrt((q.astype(np.float64) ** 2).sum()))
