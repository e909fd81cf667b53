"""EncoderBase64 mirror against the reference's own best-known-solution strings (format check: the
Gset graphs themselves are not shipped) and against a literal restatement of the reference code."""
import numpy as np
import torch as th

from rlsolver_b200.methods.util_evaluator import EncoderBase64

# rlsolver/methods/util_evaluator.py:258 (X_G14, 800 nodes, cut 3064)
X_G14 = (" 11Re2ycMx2zCiEhQl5ey$HyYnkUhDVE6KkPnuuhcWXwUO9Rn1fxrt_cn_g6iZFQex1YpwjD_j7KzbNN71qVekltv3QscNQJjrnrqHfsnOKWJzg9nJhZ$qh69"
         " $X_BvBQirx$i3F ")


def _ref_bool_to_str(x_bool, digits, string_len):            # util_evaluator.py:34-51, literally
    x_int = int(''.join([('1' if i else '0') for i in x_bool.tolist()]), 2)
    x_str = ""
    while True:
        remainder = x_int % 64
        x_str = digits[remainder] + x_str
        x_int //= 64
        if x_int == 0:
            break
    if len(x_str) > 120:
        x_str = '\n'.join([x_str[i:i + 120] for i in range(0, len(x_str), 120)])
    if len(x_str) > 64:
        x_str = f"\n{x_str}"
    return x_str.zfill(string_len)


def test_round_trip_reference_string():
    enc = EncoderBase64(800)
    x = enc.str_to_bool(X_G14)
    assert x.shape == (800,) and x.dtype == th.bool and 300 < int(x.sum()) < 500
    s = enc.bool_to_str(x)
    assert s.replace("\n", "") == X_G14.replace(" ", "")
    assert th.equal(enc.str_to_bool(s), x)


def test_matches_literal_reference_code():
    rng = np.random.default_rng(0)
    for n in (1, 5, 6, 7, 64, 100, 385, 800, 2000):
        enc = EncoderBase64(n)
        for _ in range(3):
            x = th.from_numpy(rng.integers(0, 2, n).astype(bool))
            assert enc.bool_to_str(x) == _ref_bool_to_str(x, enc.base_digits, enc.string_len)
            assert th.equal(enc.str_to_bool(enc.bool_to_str(x)), x)
        z = th.zeros(n, dtype=th.bool)
        assert enc.bool_to_str(z) == _ref_bool_to_str(z, enc.base_digits, enc.string_len)
        assert th.equal(enc.str_to_bool(enc.bool_to_str(z)), z)


def test_generate_mygraph_matches_reference():
    """methods/util_generate.generate_mygraph == the reference's (util_generate.py:75-93) for a seeded Python `random`
    (fixture: tools/make_goldens_generate.py)."""
    import os
    import random

    import numpy as np
    pytest = __import__("pytest")
    pytest.importorskip("networkx")
    from rlsolver_b200.methods.config import GraphType
    from rlsolver_b200.methods.util_generate import generate_mygraph
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "genmygraph.npz"))
    for gt in (GraphType.BA, GraphType.ER, GraphType.PL):
        random.seed(5)
        graph, n, m = generate_mygraph(gt, 30)
        assert [n, m] == z[gt.value + "_nm"].tolist()
        assert np.array_equal(np.asarray(graph, dtype=np.int64).reshape(-1, 3), z[gt.value])
