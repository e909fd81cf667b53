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


def sweeps(q: np.ndarray, x: np.ndarray, num_sweeps: int, binary: bool, dtype=np.float64) -> np.ndarray:
    """rlsolver/methods/MCPG/sampling.py:331-337 (x in {-1,+1}: x_i <- +1 if Q_i . x > 0 else -1, x_i
    zeroed first) and 356-362 (x in {0,1}: x_i <- [Q_i . x > -Q_ii / 2]), Gauss-Seidel over
    index 0..N-1, num_sweeps passes.  x: [N, C]; returns a new array.  dtype = the arithmetic of
    the dot products (float64 = the exact-sign reference, float32 = what torch computes)."""
    qq, xx = q.astype(dtype), x.astype(dtype).copy()
    n = qq.shape[0]
    for _ in range(num_sweeps):
        for i in range(n):
            xx[i] = 0
            res = qq[i] @ xx
            if binary:
                xx[i] = (res > -qq[i, i] / 2).astype(dtype)
            else:
                xx[i] = 2 * (res > 0).astype(dtype) - 1
    return xx.astype(np.float32)


def sweep_margin(q: np.ndarray, x: np.ndarray, num_sweeps: int, binary: bool) -> np.ndarray:
    """Per chain: the smallest |res - threshold| met along the float64 sweep -- chains whose margin is far
    above the fp32 rounding error of a length-N dot product must come out identical in any fp32 evaluation."""
    qq, xx = q.astype(np.float64), x.astype(np.float64).copy()
    n = qq.shape[0]
    margin = np.full(xx.shape[1], np.inf)
    for _ in range(num_sweeps):
        for i in range(n):
            xx[i] = 0
            res = qq[i] @ xx
            thr = -qq[i, i] / 2 if binary else 0.0
            margin = np.minimum(margin, np.abs(res - thr))
            xx[i] = (res > thr).astype(np.float64) if binary else 2 * (res > thr).astype(np.float64) - 1
    return margin
