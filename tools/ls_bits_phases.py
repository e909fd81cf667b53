"""Phase timing (clock64 of CTA 0) of the bit-mask tile kernel: RLSB_LS_TIMES=1 python tools/ls_bits_phases.py [G22] [4096]
Stamps: start | sweep structure landed | per iteration: candidate built, accepted | sweep done | final cut | unpack."""
import ctypes as C
import os
import sys

os.environ["RLSB_LS_TIMES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as th  # noqa: E402
from synth import gset_like  # noqa: E402

import rlsolver_b200  # noqa: E402
from rlsolver_b200 import _lib  # noqa: E402
from rlsolver_b200.envs.env_L2A import EnvMaxcut  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "G22"
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = th.device("cuda:0")
sim = EnvMaxcut(mygraph=gset_like(name), device=dev, if_bidirectional=True)
xs = sim.generate_xs_randomly(envs)
for full in ("0", "1"):
    _lib.debug_flags(*((_lib.DEBUG_FULL_CUT, 0) if full == "1" else (0, _lib.DEBUG_FULL_CUT)))
    for _ in range(3):
        sim.local_search_inplace(xs.clone(), th.empty(()))
    th.cuda.synchronize()
    buf = (C.c_int64 * 64)()
    rlsolver_b200.lib().rlsb_ls_debug_times(buf)
    t = [int(v) for v in buf]
    n = max(i for i, v in enumerate(t[:32]) if v) + 1
    d = [t[i] - t[i - 1] for i in range(1, n)]
    print(f"{name} x {envs}, full_cut={full}: total {t[n - 1] - t[0]} cycles; deltas {d}")
