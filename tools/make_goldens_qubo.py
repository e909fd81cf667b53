"""Generate tests/golden/qubo_*.npz by running the UNMODIFIED reference on CPU:
mcpg_sampling_qubo / mcpg_sampling_qubo_bin of rlsolver/methods/MCPG/sampling.py (323-370).
Build container only:  python tools/make_goldens_qubo.py"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def case(samp, name, fn, n, total_mcmc, repeat, seed, integer):
    th.manual_seed(seed)
    u = th.randn(n, n)
    if integer:
        u = th.randint(-50, 51, (n, n)).float()
    q = th.triu(u) + th.triu(u, 1).T
    data = {"Q": q, "nvar": n}
    c = total_mcmc * repeat
    probs = th.rand(n) * 0.6 + 0.2
    start = th.randint(0, 2, (n, c)).float()
    rng_state = th.get_rng_state()
    max_res, best, raw, value = fn(data, start, probs, 2, 3, total_mcmc, device=th.device("cpu"))
    # the same call with every chain as its own group hands back ALL chains after the sweeps
    # (samples[:, index] with index = arange(C)): the input / output pair of the local-search sweeps
    th.set_rng_state(rng_state)
    all_res, all_samples, raw2, _ = fn(data, start, probs, 2, 3, c, device=th.device("cpu"))
    assert th.equal(raw, raw2)
    np.savez_compressed(os.path.join(OUT, f"qubo_{name}.npz"), Q=q.numpy(), total_mcmc=np.asarray(total_mcmc),
                        max_res=max_res.numpy(), best=best.numpy(), value=value.numpy(),
                        binary=np.asarray(fn.__name__.endswith("_bin")), raw=raw.numpy(), num_ls=np.asarray(2),
                        all_res=all_res.numpy(), all_samples=all_samples.numpy())
    print("wrote qubo_" + name, "max_res[:3]", max_res[:3].tolist())


def main():
    samp = ref_import.load_by_path("ref_mcpg_sampling", "rlsolver/methods/MCPG/sampling.py",
                                   extra_sys_path=["rlsolver/methods/MCPG", "rlsolver/methods"])
    case(samp, "pm1_n96", samp.mcpg_sampling_qubo, 96, 12, 5, 401, False)
    case(samp, "bin_n70", samp.mcpg_sampling_qubo_bin, 70, 8, 4, 402, False)
    case(samp, "pm1_int_n130", samp.mcpg_sampling_qubo, 130, 6, 3, 403, True)


if __name__ == "__main__":
    main()
