"""Per-kernel measurement of the SURVEY.md section 8(d) table on the five BASELINE.json configurations.

For every kernel row (K1 .. K7) and every configuration it applies to: CUDA-event time through the
public Python API (L2 flushed between repetitions), env-steps/s, the row's algorithmic bytes (or
flops) per launch and the fraction of the measured peak (MEASURED_PEAKS.json).  One JSON line per
measurement on stdout; `--md FILE` also writes a table.  Synthetic inputs, seeds as in bench.py.

    python tools/bench_configs.py --md profiles/r01e_configs.md > profiles/r01e_configs.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch as th  # noqa: E402
from synth import gset_like  # noqa: E402

ROWS = []


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1400.0, "fallback"


HBM, TENSOR, PEAK_SRC = peaks()


def timeit(fn, flush, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        th.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(config, kernel, what, ms, env_steps, alg_bytes=None, flops=None, note=""):
    row = {"config": config, "kernel": kernel, "what": what, "ms": round(ms, 4),
           "env_steps_per_s": env_steps / (ms * 1e-3) if env_steps else None}
    if alg_bytes is not None:
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        row.update(bound="hbm", algorithmic_bytes=int(alg_bytes), achieved_gbs=round(gbs, 1), peak_gbs=HBM,
                   frac=round(gbs / HBM, 4))
    if flops is not None:
        tf = flops / (ms * 1e-3) / 1e12
        row.update(bound="tensor", flops=flops, achieved_tflops=round(tf, 1), peak_tflops=TENSOR,
                   frac=round(tf / TENSOR, 4))
    row["peak_source"] = PEAK_SRC
    if note:
        row["note"] = note
    ROWS.append(row)
    print(json.dumps(row), flush=True)


def maxcut_config(tag, name, envs, dev, flush, with_samplers=False):
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    from rlsolver_b200.envs.env_PPO import EnvMaxcut as EnvPPO
    from rlsolver_b200.methods.LocalSearch import LocalSearch
    edges = gset_like(name)
    sim = EnvMaxcut(mygraph=edges, device=dev, if_bidirectional=True)
    st = sim.store
    n, m, np_ = sim.num_nodes, sim.num_edges, st.padded_nodes
    th.manual_seed(74)
    xs = sim.generate_xs_randomly(envs)
    packed = st.pack(xs)
    out = th.empty((envs,), dtype=th.int64, device=dev)
    cfg = f"{tag}: {name}-shaped N={n} M={m}, {envs} envs"
    # K1' / K1: objective
    ms = timeit(lambda: sim.calculate_obj_values(xs), flush)
    report(cfg, "K1' cut_eval (bool rows in)", "calculate_obj_values", ms, envs, envs * n + 8 * envs + 4 * m)
    ms = timeit(lambda: st.cut_eval_packed(packed, envs, out), flush)
    report(cfg, "K1 cut_eval (packed in)", "cut_eval_packed", ms, envs, envs * np_ // 8 + 8 * envs + 4 * m,
           note="working set is L2/SM resident: bounded by launch latency + integer ALU, not HBM")
    # relaxed objective (env_k_spin.SimulatorMaxcut.get_objectives) and its gradient
    from rlsolver_b200.relaxed import relaxed_cut
    probs = th.rand((envs, n), device=dev)
    ms = timeit(lambda: relaxed_cut(st, probs), flush)
    report(cfg, "relaxed objective -(p0+p1-2p0p1) (fp32 rows in)", "SimulatorMaxcut.get_objectives", ms, envs,
           4 * envs * n + 4 * m + 4 * envs)
    pg = probs.clone().requires_grad_(True)
    gout = th.ones(envs, device=dev)

    def fwd_bwd():
        pg.grad = None
        (relaxed_cut(st, pg) * gout).sum().backward()
    ms = timeit(fwd_bwd, flush)
    report(cfg, "relaxed objective forward + backward", "get_objectives(...).backward()", ms, envs,
           3 * 4 * envs * n + 8 * m + 8 * envs, note="includes torch's (obj * g).sum() and its backward")
    del probs, pg
    # K2: exhaustive single-flip pass (N env-steps per env)
    vs = st.cut_eval_packed(packed, envs)
    pk = packed.clone()

    def sweep():
        pk.copy_(packed)
        st.flip_sweep(pk, vs)
    ms_copy = timeit(lambda: pk.copy_(packed), flush)
    ms = timeit(sweep, flush) - ms_copy
    report(cfg, "K2 flip sweep", "flip_sweep (N candidate flips per env)", ms, envs * n,
           2 * envs * np_ // 8 + 4 * (n + 1) + 2 * st.num_full * 2 + 16 * envs,
           note="latency bound: one barrier per dependency level of the Gauss-Seidel order")
    # K3: noisy iterations (ls_run: threshold + 8 iterations + sweep)
    cb = 1 if max(st.max_listed_degree, st.max_full_degree) <= 255 else 2
    ws = st.ls_workspace(envs)
    x2 = xs.clone()
    vs2 = st.ls_begin(x2, None, 1, 0.3, ws)
    nz = [th.randn((envs, n), device=dev) for _ in range(9)]
    ms_begin = timeit(lambda: st.ls_begin(x2, None, 1, 0.3, ws), flush)
    report(cfg, "ls_begin (pack + cut + cross counts + spread)", "ls_begin", ms_begin, envs,
           envs * n + envs * np_ // 8 + cb * envs * np_ + 8 * envs)
    it_bytes = 4 * envs * n + cb * envs * np_
    ms8 = timeit(lambda: st.ls_run(vs2, 1, nz[0], 8, nz[1:9], False, None, ws), flush)
    report(cfg, "K3 noisy iterations", "ls_run: threshold pass + 8 iterations", ms8, 9 * envs,
           9 * it_bytes + 2 * envs * np_ // 8 + 20 * envs)
    ms_full = timeit(lambda: st.ls_run(vs2, 1, nz[0], 8, nz[1:9], True, x2, ws), flush)
    report(cfg, "K3+K2 local_search_inplace kernel", "ls_run: threshold + 8 iterations + sweep + unpack", ms_full,
           envs * (9 + n), 9 * it_bytes + 2 * envs * np_ // 8 + envs * n + 20 * envs)
    del nz
    # K3 without noise tensors: the generator recomputed in place (noise_masks.cu) + the bit-mask tile kernel
    from rlsolver_b200 import rng as _rng
    if st.ls_mask_words(envs) >= 0:
        seed, offset, threads, iters = _rng.peek(dev, envs * n)
        ms_gen = timeit(lambda: st.ls_noise_masks(envs, 1, 8, seed, offset, threads, iters, ws), flush)
        gen_bytes = envs * np_ + envs * n // 2 + 8 * (envs * n // 2 + 2 * envs * n // 8)
        report(cfg, "K3 mask generator (8 draws, early-out)", "ls_noise_masks", ms_gen, 8 * envs, gen_bytes,
               note=f"issue bound (Philox); stands in for {8 * 4 * envs * n / 1e6:.0f} MB of float32 noise = "
                    f"{8 * 4 * envs * n / (ms_gen * 1e-3) / 1e9:.0f} GB/s noise-equivalent")
        masks = st.ls_noise_masks(envs, 1, 8, seed, offset, threads, iters, ws)
        ms_bits = timeit(lambda: st.ls_run_masks(vs2, masks, 8, True, x2, ws), flush)
        report(cfg, "K3+K2 bit-mask tile kernel", "ls_run_masks: 8 iterations + sweep + unpack", ms_bits, envs * (8 + n),
               2 * envs * np_ // 8 + 8 * envs * n // 8 + envs * n + 16 * envs,
               note="working set is SM resident: latency / issue bound, not HBM")
    # whole reference call: RNG consumed in place (default) and as explicit torch.randn tensors
    ms = timeit(lambda: sim.local_search_inplace(xs.clone(), th.empty(())), flush)
    report(cfg, "local_search_inplace (public API, generator consumed in place)", "EnvMaxcut.local_search_inplace", ms,
           envs * (9 + n))
    sim.fused_rng = False
    ms = timeit(lambda: sim.local_search_inplace(xs.clone(), th.empty(())), flush)
    report(cfg, "local_search_inplace (public API, explicit torch.randn tensors)", "EnvMaxcut.local_search_inplace "
           "(fused_rng=False)", ms, envs * (9 + n))
    sim.fused_rng = True
    if not with_samplers:
        uni = EnvMaxcut(mygraph=edges, device=dev, if_bidirectional=False)   # as env_MCPG.py:416 builds it
        ls = LocalSearch(uni, n)
        ls.reset(xs.clone())
        ms = timeit(lambda: ls.random_search(num_iters=8, num_spin=4), flush)
        report(cfg, "LocalSearch.random_search(8)", "random_search", ms, envs * (8 + n))
    # K4: pattern-I step on the shared graph
    class A:
        num_nodes, num_envs, num_steps = n, envs, 1 << 30
    env = EnvPPO(A, mygraph=edges, device=dev, if_bidirectional=True)
    env.reset()
    acts = [th.randint(0, n, (envs,), device=dev) for _ in range(8)]
    k = [0]

    def step():
        env.step(acts[k[0] % 8])
        k[0] += 1
    ms = timeit(step, flush)
    dbar = 2.0 * m / n
    report(cfg, "K4 pattern-I step_flip", "env_PPO.EnvMaxcut.step", ms, envs, envs * (4 * dbar + 24),
           note="scattered 32-byte sectors of the float32 observation rows; launch-latency floor ~3 us")
    # greedy best flip (a19)
    g0 = th.zeros((min(envs, 4096), n), dtype=th.bool, device=dev)
    x = g0.clone()

    def greedy():
        x.copy_(g0)
        st.greedy_best_flip(x, n, True)
    try:
        ms = timeit(greedy, flush, reps=3, warm=1)
        report(cfg, "greedy best-flip to local optimum", "greedy_best_flip from all-zeros", ms, None)
    except NotImplementedError as exc:      # resident gains of 10000 nodes x 32 envs exceed shared memory
        print(json.dumps({"config": cfg, "kernel": "greedy best-flip", "unsupported": str(exc)[:120]}), flush=True)
    if with_samplers:
        from rlsolver_b200.methods.L2A.transformer import sub_set_sampling
        from rlsolver_b200.methods.MCPG import McpgData, metro_sampling, sampler_func
        data = McpgData(edges, n, dev)
        total, rep = 512, 8
        chains = total * rep
        probs = th.rand(n, device=dev) * 0.6 + 0.2
        start = (th.rand(n, chains, device=dev) < 0.5).float()
        import rlsolver_b200.methods.MCPG as _M
        ms = timeit(lambda: metro_sampling(probs, start, n // 10, dev), flush)
        iters = 5 * (n // 10)
        cap, _M._METRO_SPLIT_MAX_BYTES = _M._METRO_SPLIT_MAX_BYTES, -1
        ms_two = timeit(lambda: metro_sampling(probs, start, n // 10, dev), flush)
        _M._METRO_SPLIT_MAX_BYTES = cap
        report(cfg, "K6 metro_sampling, two-pass chain kernel (round 1e form)", f"metro_sampling(max_transfer_time={n // 10})",
               ms_two, chains * iters, note="count pass + apply pass, two Philox blocks inside every dependent iteration")
        report(cfg, "K6 metro_sampling", f"metro_sampling(max_transfer_time={n // 10}), <= {iters} iterations", ms,
               chains * iters, 2 * 4 * n * chains, note="split form: draws of all iterations in parallel, one pass of the chain, surplus moves undone")
        xs_s = metro_sampling(probs, start, n // 10, dev)
        ms = timeit(lambda: sampler_func(data, xs_s, 8, total, rep), flush)
        report(cfg, "K5 MCPG sweeps", "sampler_func(num_ls=8): 8 Gauss-Seidel sweeps + expected cut", ms,
               chains * 8 * n, 2 * 4 * n * chains + 4 * chains)
        sims, reps_ = 64, 64
        p2 = th.rand(sims, n, device=dev)
        s2 = th.rand(sims, n, device=dev) < 0.5
        ms = timeit(lambda: sub_set_sampling(p2, s2, reps_, n // 4), flush)
        report(cfg, "sub_set_sampling (dREINFORCE)", f"top_k={n // 4}, {sims} sims x {reps_} repeats", ms,
               sims * reps_ * (n // 4), sims * reps_ * n * 2)


def peco_config(dev, flush, envs):
    from rlsolver_b200.envs import env_PECO as P
    n = 100
    for kind in ("BA", "ER"):
        gg = (P.RandomBAGraphGenerator(n_spins=n, m_insertion_edges=4, edge_type=P.EdgeType.DISCRETE, num_envs=envs,
                                       device=dev) if kind == "BA" else
              P.RandomERGraphGenerator(n_spins=n, p_connection=0.15, edge_type=P.EdgeType.DISCRETE, num_envs=envs,
                                       device=dev))
        env = P.SpinSystemFactory.get(gg, 2 * n, observables=P.ECO_PECO_OBSERVABLES, reward_signal=P.RewardSignal.BLS,
                                      extra_action=P.ExtraAction.NONE, optimisation_target=P.OptimisationTarget.CUT,
                                      spin_basis=P.SpinBasis.BINARY, norm_rewards=True, memory_length=None,
                                      horizon_length=None, stag_punishment=None, basin_reward=None,
                                      reversible_spins=True, device=dev, num_envs=envs)
        acts = [th.randint(0, n, (envs,), device=dev) for _ in range(8)]
        k = [0]

        def step():
            env.step(acts[k[0] % 8], return_observation=False)
            k[0] += 1
        ms = timeit(step, flush, reps=7, warm=2)
        nobs = len(P.ECO_PECO_OBSERVABLES)
        # one matrix row (4N) + fields r/w (8N) + observable rows touched (~4 rows r/w) + scalars
        alg = envs * (4 * n + 8 * n + 2 * 4 * n * 4 + 40)
        dens = float((env.matrix[:1024] != 0).float().sum() / 1024 / n)
        report(f"config 4: {kind}-100 per-env graphs, {envs} envs (mean degree {dens:.1f})",
               "K4' PECO step (dense fp32 matrix row, resident fields, 7 observables)",
               "SpinSystemUnbiased.step(random actions), no observation concat", ms, envs, alg,
               note=f"state [E,{nobs},N] f32 + matrix [E,N,N] f32 = {envs * (nobs + n) * n * 4 / 2**30:.1f} GiB resident")
        del env, gg, acts
        th.cuda.empty_cache()


def qubo_config(dev, flush):
    from rlsolver_b200.qubo import QuboModel
    n = 4096
    th.manual_seed(0)
    u = th.randn(n, n, device=dev)
    q = th.triu(u) + th.triu(u, 1).T
    model = QuboModel(q)
    for c in (1024, 8192):
        x = th.randint(0, 2, (n, c), device=dev).float() * 2 - 1
        ms = timeit(lambda: model.energy(x), flush)
        report(f"config 5: dense QUBO N={n}, {c} chains", "K7 qubo_energy (3-limb bf16 tcgen05)", "QuboModel.energy",
               ms, c, flops=3 * 2.0 * n * n * c, note="flops = bf16 issued (3 limbs); useful fp32-accurate = 1/3")
        xs_ = x.clone()
        ms = timeit(lambda: model.sweeps(xs_, 1), flush, reps=3, warm=1)
        report(f"config 5: dense QUBO N={n}, {c} chains", "K7' qubo_sweeps (blocked Gauss-Seidel, 3-limb bf16 tcgen05)",
               "QuboModel.sweeps(1 sweep = N coordinate updates per chain)", ms, c * n, flops=3 * 2.0 * n * n * c,
               note="one launch per 64-row block (64 launches); reference = N dependent GEMVs")
        ref = timeit(lambda: (x * (q @ x)).sum(0), flush, reps=3, warm=1)
        report(f"config 5: dense QUBO N={n}, {c} chains", "torch fp32 (x*(Q@x)).sum(0) on the same GPU", "torch", ref, c)


def write_md(path):
    with open(path, "w") as f:
        f.write("| config | kernel | call | ms | env-steps/s | algorithmic MB | achieved | frac of peak | note |\n")
        f.write("|---|---|---|---|---|---|---|---|---|\n")
        for r in ROWS:
            ach = (f"{r['achieved_gbs']} GB/s" if "achieved_gbs" in r else
                   f"{r['achieved_tflops']} TFLOP/s" if "achieved_tflops" in r else "")
            mb = f"{r['algorithmic_bytes'] / 1e6:.2f}" if "algorithmic_bytes" in r else ""
            es = f"{r['env_steps_per_s']:.3g}" if r.get("env_steps_per_s") else ""
            f.write(f"| {r['config']} | {r['kernel']} | {r['what']} | {r['ms']} | {es} | {mb} | {ach} | "
                    f"{r.get('frac', '')} | {r.get('note', '')} |\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", default=None)
    ap.add_argument("--peco-envs", type=int, default=1 << 18)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = th.device("cuda:0")
    flush = th.empty(256 << 20, dtype=th.uint8, device=dev)
    want = set(a.only.split(",")) if a.only else {"0", "1", "2", "3", "4"}
    if "0" in want:
        maxcut_config("config 1", "G14", 256, dev, flush)
    if "1" in want:
        maxcut_config("config 2", "G22", 4096, dev, flush, with_samplers=True)
    if "2" in want:
        maxcut_config("config 3 (one GPU's share x8)", "G70", 16384, dev, flush)
    if "3" in want:
        peco_config(dev, flush, a.peco_envs)
    if "4" in want:
        qubo_config(dev, flush)
    if a.md:
        write_md(a.md)


if __name__ == "__main__":
    main()
