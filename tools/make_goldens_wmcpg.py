"""Generate tests/golden/wmcpg_*.npz by running the UNMODIFIED reference on CPU: the WEIGHTED max-cut sampler
mcpg_sampling_maxcut of rlsolver/methods/MCPG/sampling.py:89-127 on data from maxcut_dataloader of
rlsolver/methods/MCPG/dataloader.py:53-103 (float `edge_attr`), with every torch.rand / randint draw recorded.
Weights: +-1 (Gset style), small integers, dyadic rationals (k/8: float weights whose sums are exact in float32,
so the decisions do not depend on the summation order of torch.mm) and arbitrary floats (checked to a tolerance,
on the expected cut of the returned states only).  Build container only:  python tools/make_goldens_wmcpg.py"""
import os
import sys
import tempfile

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from make_goldens import Recorder, graph_cases  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def with_weights(mygraph, values, seed):
    rng = np.random.default_rng(seed)
    return [(a, b, float(rng.choice(values))) for a, b, _ in mygraph if a != b]


def case(samp, loader, name, edges, total_mcmc, repeat, num_ls, change_times, seed):
    n = len({a for a, _, _ in edges} | {b for _, b, _ in edges})
    th.manual_seed(seed)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "g.txt")
        with open(path, "w") as fh:
            fh.write(f"{n} {len(edges)}\n")
            for a, b, w in edges:
                fh.write(f"{a + 1} {b + 1} {w!r}\n")
        data, num_nodes = loader.maxcut_dataloader(path, device=th.device("cpu"))
    c = total_mcmc * repeat
    probs = th.rand(n) * 0.6 + 0.2
    start = th.randint(0, 2, (n, c)).float()
    with Recorder("randint") as r_int, Recorder("rand") as r_u:
        vs_good, xs_good, start_out, value = samp.mcpg_sampling_maxcut(data, start, probs, num_ls, change_times,
                                                                       total_mcmc, device=th.device("cpu"))
    metro_iters = len(r_int.draws)
    out = dict(edges=np.asarray([(a, b) for a, b, _ in edges], dtype=np.int64),
               weights=np.asarray([w for _, _, w in edges], dtype=np.float32), num_nodes=np.asarray(n),
               total_mcmc=np.asarray(total_mcmc), repeat=np.asarray(repeat), num_ls=np.asarray(num_ls),
               change_times=np.asarray(change_times), order=data.sorted_degree_nodes.numpy().copy(),
               probs=probs.numpy().copy(), start=start.numpy().copy(),
               metro_idx=np.stack(r_int.draws), metro_u=np.stack(r_u.draws[:metro_iters]),
               ls_u=np.stack(r_u.draws[metro_iters:]), metro_out=start_out.numpy().copy(),
               vs_good=vs_good.numpy().copy(), xs_good=xs_good.numpy().copy(), value=value.numpy().copy(),
               edge_weight_sum=np.asarray(data.edge_weight_sum))
    p = os.path.join(OUT, f"wmcpg_{name}_T{total_mcmc}_R{repeat}.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, "metro iters", metro_iters, "ls draws", len(r_u.draws) - metro_iters, "vs_good[:4]", vs_good[:4].tolist())


def main():
    extra = ["rlsolver/methods/MCPG", "rlsolver/methods"]
    samp = ref_import.load_by_path("ref_mcpg_sampling", "rlsolver/methods/MCPG/sampling.py", extra_sys_path=extra)
    loader = ref_import.load_by_path("ref_mcpg_dataloader", "rlsolver/methods/MCPG/dataloader.py", extra_sys_path=extra)
    cases = graph_cases()
    case(samp, loader, "ba100pm1", with_weights(cases["ba100"], [-1.0, 1.0], 1), 16, 5, 3, 10, 501)
    case(samp, loader, "hub50int", with_weights(cases["hub50"], [-3.0, -1.0, 1.0, 2.0, 5.0], 2), 12, 3, 2, 5, 502)
    case(samp, loader, "multi67dyadic", with_weights(cases["multi67"], [k / 8 for k in range(-12, 13) if k], 3), 9, 4, 2, 4, 503)
    case(samp, loader, "toy14unit", with_weights(cases["toy14"], [1.0], 4), 7, 3, 1, 2, 504)
    case(samp, loader, "ba100float", [(a, b, float(np.float32(w))) for (a, b, _), w in
                                      zip(cases["ba100"], np.random.default_rng(5).normal(size=len(cases["ba100"])))],
         8, 4, 2, 6, 505)


if __name__ == "__main__":
    main()
