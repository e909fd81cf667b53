"""Synthetic inputs shaped like the BASELINE.json configs (no Gset files ship offline)."""
import contextlib

import numpy as np
import torch as th


def random_graph(n: int, m: int, seed: int = 74):
    """Uniform random simple graph with exactly m edges and no isolated node."""
    rng = np.random.default_rng(seed)
    assert m >= (n + 1) // 2
    while True:
        seen = set()
        # a random perfect-ish matching first so every node has degree >= 1
        perm = rng.permutation(n)
        for k in range(0, n - 1, 2):
            a, b = int(perm[k]), int(perm[k + 1])
            seen.add((min(a, b), max(a, b)))
        if n % 2:
            a, b = int(perm[-1]), int(perm[0])
            seen.add((min(a, b), max(a, b)))
        while len(seen) < m:
            need = m - len(seen)
            a = rng.integers(0, n, size=2 * need + 16)
            b = rng.integers(0, n, size=2 * need + 16)
            for x, y in zip(a.tolist(), b.tolist()):
                if x != y:
                    seen.add((min(x, y), max(x, y)))
                    if len(seen) == m:
                        break
        edges = sorted(seen)
        return [(a, b, 1) for a, b in edges]


SHAPES = {"G14": (800, 4694), "G22": (2000, 19990), "G70": (10000, 9999)}


def gset_like(name: str, seed: int = 74):
    n, m = SHAPES[name]
    return random_graph(n, m, seed)


@contextlib.contextmanager
def replay(fn_name: str, draws, device=None):
    """Make torch.<fn_name> return the recorded draws in order (checking shapes)."""
    orig = getattr(th, fn_name)
    it = iter(draws)

    def fake(*args, **kw):
        arr = next(it)
        t = th.from_numpy(np.ascontiguousarray(arr)) if isinstance(arr, np.ndarray) else arr
        dev = kw.get("device", device)
        if dev is None and args and isinstance(args[0], th.Tensor):
            dev = args[0].device
        return t.to(dev) if dev is not None else t

    setattr(th, fn_name, fake)
    try:
        yield
    finally:
        setattr(th, fn_name, orig)
