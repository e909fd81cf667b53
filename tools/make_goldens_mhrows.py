"""Generate tests/golden/mhrows_*.npz by running the UNMODIFIED reference on CPU:
metropolis_hastings_sampling_TNCO of rlsolver/envs/env_L2A.py:233-276 from a seeded generator; the fixture keeps the
inputs, the samples and four uniforms drawn right afterwards (the generator state the call leaves behind).
Build container only:  python tools/make_goldens_mhrows.py"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.envs import env_L2A as ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def case(name, sims, dim, repeats, num_iters, seed, sharp):
    g = th.Generator().manual_seed(seed)
    probs = th.rand((sims, dim), generator=g)
    if sharp:                                   # probabilities near 0 / 1: few accepts, several rounds
        probs = th.where(probs < 0.5, probs * 0.04 + 0.005, 1 - probs * 0.04)
    start = th.rand((sims, dim), generator=g) < probs
    th.manual_seed(seed + 1)
    out = ref.metropolis_hastings_sampling_TNCO(probs=probs, start_xs=start, num_repeats=repeats, num_iters=num_iters)
    after = th.rand(4)
    p = os.path.join(OUT, f"mhrows_{name}.npz")
    np.savez_compressed(p, probs=probs.numpy(), start=start.numpy(), repeats=np.asarray(repeats),
                        num_iters=np.asarray(num_iters), seed=np.asarray(seed + 1), out=out.numpy(), after=after.numpy())
    print("wrote", p, out.shape, "changed bits:", int((out != start.repeat(repeats, 1)).sum()))


def main():
    case("s6_d40_r3", 6, 40, 3, -1, 810, False)
    case("s5_d33_r2_sharp", 5, 33, 2, 6, 820, True)
    case("s3_d20_r4_one_iter", 3, 20, 4, 1, 830, False)


if __name__ == "__main__":
    main()
