import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch as th
from test_gpu_peco import _env_from_golden, _np
dev = th.device("cuda:0")
for name in ("peco_ba20_uniform_bls.npz", "peco_er20_discrete_bls.npz"):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    env, obs0 = _env_from_golden(z, name, dev)
    got, want = _np(env.state), z["state0"]
    bad = np.argwhere(got != want)
    print(name, "mismatches", len(bad), bad[:8].tolist())
    for e, r, j in bad[:8]:
        print("  ", e, r, j, repr(got[e, r, j]), repr(want[e, r, j]), "maxl", _np(env.max_local_reward_available_)[e], z["max_local"][e])
    print("  score eq", np.array_equal(_np(env.score), z["score0"]), "maxl eq", np.array_equal(_np(env.max_local_reward_available_), z["max_local"]))
