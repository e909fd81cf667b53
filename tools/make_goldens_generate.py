"""Generate tests/golden/genmygraph.npz by running the UNMODIFIED reference on CPU: generate_mygraph of
rlsolver/methods/util_generate.py:75-93 for the three graph families from a seeded Python `random`.
Build container only:  python tools/make_goldens_generate.py"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.methods import util_generate as ug  # noqa: E402
from rlsolver.methods.config import GraphType  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
out = {}
for gt in (GraphType.BA, GraphType.ER, GraphType.PL):
    random.seed(5)
    graph, n, m = ug.generate_mygraph(gt, 30)
    out[gt.value] = np.asarray(graph, dtype=np.int64).reshape(-1, 3)
    out[gt.value + "_nm"] = np.asarray([n, m])
np.savez_compressed(os.path.join(OUT, "genmygraph.npz"), **out)
print({k: v.shape for k, v in out.items()})
