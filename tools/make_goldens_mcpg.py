"""Generate tests/golden/mcpg_*.npz and subset_*.npz by running the UNMODIFIED reference on CPU:
metro_sampling + sampler_func of rlsolver/methods/MCPG.py and sub_set_sampling of
rlsolver/methods/L2A/transformer.py, with every torch.rand / randint / rand_like draw recorded.
Build container only:  python tools/make_goldens_mcpg.py"""
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


def write_graph(path, edges, n):
    with open(path, "w") as fh:
        fh.write(f"{n} {len(edges)}\n")
        for a, b, w in edges:
            fh.write(f"{a + 1} {b + 1} {w}\n")


def mcpg_case(mcpg, name, edges, total_mcmc, repeat, num_ls, max_transfer, seed):
    n = len({a for a, _, _ in edges} | {b for _, b, _ in edges})
    th.manual_seed(seed)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "g.txt")
        write_graph(path, edges, n)
        data, num_nodes = mcpg.maxcut_dataloader(path, device=th.device("cpu"))
    c = total_mcmc * repeat
    probs = th.rand(n) * 0.6 + 0.2                              # Simpler clamps to [0.2, 0.8] (MCPG.py:183)
    start = th.randint(0, 2, (n, c)).float()
    with Recorder("randint") as r_int, Recorder("rand") as r_u:
        xs_sample = mcpg.metro_sampling(probs, start.clone(), max_transfer, device=th.device("cpu"))
    with Recorder("rand") as r_ls:
        vs_good, xs_good, value = mcpg.sampler_func(data, xs_sample, num_ls, total_mcmc, repeat, device=th.device("cpu"))
    out = dict(edges=np.asarray(edges, dtype=np.int64), num_nodes=np.asarray(n), total_mcmc=np.asarray(total_mcmc),
               repeat=np.asarray(repeat), num_ls=np.asarray(num_ls), max_transfer=np.asarray(max_transfer),
               order=data.sorted_degree_nodes.numpy().copy(), probs=probs.numpy().copy(), start=start.numpy().copy(),
               metro_idx=np.stack(r_int.draws), metro_u=np.stack(r_u.draws), xs_sample=xs_sample.numpy().copy(),
               ls_u=np.stack(r_ls.draws), vs_good=vs_good.numpy().copy(), xs_good=xs_good.numpy().copy(),
               value=value.numpy().copy())
    p = os.path.join(OUT, f"mcpg_{name}_T{total_mcmc}_R{repeat}.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, "metro iters", len(r_int.draws), "ls draws", len(r_ls.draws))


def subset_case(tr, name, s, n, repeats, top_k, seed):
    th.manual_seed(seed)
    start = th.randint(0, 2, (s, n), dtype=th.bool)
    probs = th.rand((s, n), dtype=th.float32)
    with Recorder("rand_like") as rec:
        xs, probs_out = tr.sub_set_sampling(probs=probs, start_xs=start, num_repeats=repeats, top_k=top_k)
    det = th.abs(probs - 0.5)
    top_values, top_ids = th.topk(det, k=min(top_k, n), largest=False, dim=1)
    out = dict(start=start.numpy().copy(), probs=probs.numpy().copy(), repeats=np.asarray(repeats), top_k=np.asarray(top_k),
               u=np.stack(rec.draws) if rec.draws else np.zeros((0, s * repeats), np.float32), xs=xs.numpy().copy(),
               probs_out=probs_out.numpy().copy(), top_ids=top_ids.numpy().copy(), top_values=top_values.numpy().copy())
    p = os.path.join(OUT, f"subset_{name}.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, "draws", len(rec.draws))


def main():
    mcpg = ref_import.load_by_path("ref_mcpg_single", "rlsolver/methods/MCPG.py", extra_sys_path=["rlsolver/methods"])
    cases = graph_cases()
    mcpg_case(mcpg, "ba100", cases["ba100"], total_mcmc=16, repeat=5, num_ls=3, max_transfer=10, seed=201)
    mcpg_case(mcpg, "toy14", cases["toy14"], total_mcmc=9, repeat=4, num_ls=2, max_transfer=2, seed=202)
    mcpg_case(mcpg, "hub50", cases["hub50"], total_mcmc=40, repeat=2, num_ls=1, max_transfer=5, seed=203)
    # transformer.py does a bare `from config import ConfigGraph`: L2A/config.py must win
    sys.modules.pop("config", None)
    sys.path.insert(0, os.path.join(ref_import.REF, "rlsolver", "methods", "L2A"))
    from rlsolver.methods.L2A import transformer as tr
    subset_case(tr, "S8_N37", 8, 37, 4, 9, 301)
    subset_case(tr, "S5_N64", 5, 64, 7, 64, 302)
    subset_case(tr, "S3_N10_k0", 3, 10, 2, 0, 303)


if __name__ == "__main__":
    main()
