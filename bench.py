#!/usr/bin/env python
"""Benchmark of the max-cut environment hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--configs 1,3,4,5|none]

Headline workload (BASELINE.json configs[1]): G22-shaped max-cut (2000 nodes, 19990 edges, synthetic,
seed 74), 4096 environments per GPU, one dREINFORCE sampling step of local search =
`EnvMaxcut.local_search_inplace(xs, ())` with the reference defaults (8 noisy multi-flip
iterations + the 2000-node single-flip pass).  1 env-step = one (environment, candidate move)
whose cut value is produced: E * (1 + num_iters + N) per step.

One JSON line on rank 0.  `value` = env-steps/s with inputs resident in HBM, CUDA-event timed,
L2 flushed between steps; `e2e` = the same through the public API from pinned HOST buffers
(H2D of the spins, D2H of spins + values inside the timed region).  The same line carries
  * `configs`: the other four BASELINE configurations (1: G14 x 256, 3: G70 x 16384 strong-scaled over the
    ranks, 4: 10^6 per-env BA/ER-100 graphs, 5: dense QUBO N = 4096 with 8192 chains split over the ranks),
    each with its own time, env-steps/s, roofline entry and clocks;
  * `gpu_reference`: the torch restatement of the reference's algorithm (oracle/torch_port.py: index-gather
    objective, one full re-evaluation per candidate flip) on the SAME GPU, the FULL headline workload;
  * `cpu_baseline`: the same restatement on the host cores on a bounded sample.
`--impl reference` times the CPU restatement alone.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "maxcut_env_steps_per_sec"
UNIT = "env-steps/s"
GRAPH, NUM_ENVS, NUM_ITERS, NUM_SPIN, NOISE_STD = "G22", 4096, 8, 8, 0.3
SETTLE_STEPS = 1000     # untimed, identical on every rank


def workload_name(envs):
    return (f"G22-shaped maxcut (2000 nodes, 19990 edges, synthetic seed 74), {envs} envs/GPU, "
            f"local_search_inplace: {NUM_ITERS} noisy multi-flip iters + full single-flip pass (dREINFORCE sampling step)")


def env_steps_per_call(envs, n, num_iters, sweep_nodes):
    return envs * (1 + num_iters + sweep_nodes)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tensor_burst": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor": 1400.0, "tensor_burst": 1650.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        """Clock summary of the samples taken in [t0, t1] (the sampler keeps running)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.rows):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self, t0, t1):
        if self.proc is None:
            return self.window(t0, t1)
        time.sleep(0.15)
        out = self.window(t0, t1)
        self.proc.terminate()
        return out


# ----------------------------------------------------------------------------- reference arm (host cores)
def cpu_reference_run(steps, warmup, budget_s=3.0):
    """The reference's algorithm on the host cores: torch restatement (index-gather objective,
    one full re-evaluation per candidate flip).  Each step is a bounded sample of the workload:
    256 of the envs, all noisy iterations, and the first `sweep_nodes` nodes of the single-flip
    pass, sized so a step takes about `budget_s` seconds."""
    import torch as th
    from oracle import torch_port as tp
    from synth import gset_like

    th.set_num_threads(os.cpu_count() or 1)
    edges = gset_like(GRAPH)
    sim = tp.TorchSim(edges, True, device="cpu")
    envs = 256
    th.manual_seed(74)
    xs0 = sim.random_xs(envs)
    sim.objective(xs0)                                   # builds the [E, Md] index tensors
    t = time.perf_counter()
    for _ in range(3):
        sim.objective(xs0)
    t_eval = (time.perf_counter() - t) / 3
    sweep_nodes = int(max(8, min(sim.num_nodes, budget_s / max(t_eval * 1.3, 1e-4) - (2 + NUM_ITERS))))
    per_call = env_steps_per_call(envs, sim.num_nodes, NUM_ITERS, sweep_nodes)
    times = []
    for it in range(warmup + steps):
        xs = xs0.clone()
        t = time.perf_counter()
        sim.local_search_inplace(xs, None, NUM_ITERS, NUM_SPIN, NOISE_STD, sweep_nodes=sweep_nodes)
        dt = time.perf_counter() - t
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    value = per_call * len(times) / total
    full_per_call = env_steps_per_call(NUM_ENVS, sim.num_nodes, NUM_ITERS, sim.num_nodes)
    sample = (f"{envs} envs of the same graph, {NUM_ITERS} noisy iters + first {sweep_nodes} of 2000 sweep nodes "
              f"per step ({per_call} env-steps), torch CPU ops as the reference uses them")
    basis = (f"every env-step of this path is one full objective evaluation of one env in the reference, so the rate is "
             f"per evaluation: the full workload ({NUM_ENVS} envs, all {sim.num_nodes} sweep nodes = {full_per_call} "
             f"env-steps per step) would take {full_per_call / value:.1f} s per step at this rate; sampled because it "
             f"is {full_per_call / per_call:.0f}x the sample")
    return {"value": value, "unit": UNIT, "cores": th.get_num_threads(), "kind": "port", "sample": sample,
            "ms_per_step": 1e3 * total / len(times), "steps": len(times), "same_config_basis": basis,
            "full_workload_seconds_per_step_extrapolated": full_per_call / value}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1) if args.warmup else 0
    r = cpu_reference_run(args.steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": workload_name(NUM_ENVS), "cpu_sample": r["sample"],
                       "same_config_basis": r["same_config_basis"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "same_config_basis",
                                               "full_workload_seconds_per_step_extrapolated")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------- shared helpers of the B200 arm
class Ctx:
    """Per-process state shared by the headline and the per-config blocks."""

    def __init__(self, args):
        import torch as th
        import torch.distributed as dist
        self.th, self.dist, self.args = th, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not th.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (rlsolver_b200 has no CPU fallback)")
        th.cuda.set_device(self.local)
        self.dev = th.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = th.empty(256 << 20, dtype=th.uint8, device=self.dev)       # > 126 MB L2
        self.peaks = peaks()
        self.clocks = ClockSampler(self.local) if self.rank == 0 else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.th.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.th.tensor([x], dtype=self.th.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps, warm=3, flush=True, prepare=None):
        """CUDA-event time of `fn` per repetition (ms, list).  `prepare` runs untimed before each repetition;
        the L2 is evicted between repetitions (256 MiB write) unless flush=False."""
        th = self.th
        for _ in range(warm):
            if prepare:
                prepare()
            fn()
        self.barrier()
        evs = []
        for _ in range(reps):
            if prepare:
                prepare()
            if flush:
                self.flush.zero_()
            a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        self.barrier()
        return [a.elapsed_time(b) for a, b in evs]

    def job_ms(self, ms_list):
        """Per-repetition time of the whole job: sum over the repetitions, max over the ranks, / repetitions."""
        return self.max_over_ranks(sum(ms_list)) / len(ms_list)

    def window_clocks(self, t0, t1):
        return self.clocks.window(t0, t1) if self.clocks else None


_T0 = time.time()


def log(msg):
    """Progress marker on stderr (stdout carries the one JSON line): tells where a run that timed out was."""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def roof(bound, achieved, peak, unit, kernel, traffic=None, **extra):
    d = {"kernel": kernel, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
         "frac": achieved / peak if peak else None, "traffic": traffic}
    d.update(extra)
    return d


def dram_traffic():
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def ls_alg_bytes(st, envs, n, m, num_iters):
    """Algorithmic HBM bytes per launch group of the local-search path (DESIGN.md "Kernels")."""
    np_ = st.padded_nodes
    cb = 1 if max(st.max_listed_degree, st.max_full_degree) <= 255 else 2   # cross-count bytes
    graph_b = 4 * m + 2 * st.num_full + 12 * np_
    return {
        "ls_search": 4 * envs * n + cb * envs * np_ + 2 * envs * np_ // 8 + 20 * envs + 8 * np_,
        "ls_thresh": 4 * envs * n + cb * envs * np_ + 4 * envs + 8 * np_,
        # early-out pass (cross counts in, one byte per Box-Muller pair out) + per draw those bytes in, one bit per
        # element out (zeroed, then OR-ed).  The float32 noise itself (4 B per element and draw) never exists.
        "ls_noise_masks": envs * np_ + envs * n // 2 + num_iters * (envs * n // 2 + 2 * envs * n // 8) + 4 * envs
                          + 8 * np_,
        # packed tile in/out, one mask bit per element and iteration, bool rows + values out, graph once
        "ls_run_masks": 2 * envs * np_ // 8 + num_iters * envs * n // 8 + envs * n + 16 * envs + graph_b,
        "ls_fused_search": 2 * envs * np_ // 8 + envs * n + envs * np_ + 16 * envs + graph_b,
        "torch_randn": 4 * envs * n,
        "rng_cursor_advance": 16,
        "ls_begin": envs * n + envs * np_ // 8 + cb * envs * np_ + 8 * envs + graph_b,
        "pack_spins": envs * n + envs * np_ // 8,
        "unpack_spins": envs * n + envs * np_ // 8,
        "cut_eval_packed": envs * np_ // 8 + 8 * envs + 4 * m,
        "cut_eval": envs * n + 8 * envs + 4 * m,
        "greedy_best_flip": 2 * envs * n + 8 * envs + graph_b,
        "select_rows": 16 * envs,
    }


def per_kernel_pass(ctx, store, fn, reps, prepare=None):
    """CUDA events around every call of this library made by `fn` (eager launches): {op: ms per launch},
    {op: share of the summed kernel time}."""
    from rlsolver_b200.graph_store import OpTimer
    store.timer = OpTimer()
    for _ in range(reps):
        if prepare:
            prepare()
        ctx.flush.zero_()
        fn()
    spans = store.timer.summary()
    store.timer = None
    total = sum(v[1] for v in spans.values()) or 1e-9
    return ({k: v[1] / v[0] for k, v in spans.items()}, {k: round(v[1] / total, 4) for k, v in spans.items()},
            {k: v[0] / reps for k, v in spans.items()})


# ----------------------------------------------------------------------------- config 1: G14 x 256
def config1(ctx):
    """check_local_search_maxcut (rlsolver/envs/env_L2A.py:178-230) shape: calculate_obj_values x100,
    local_search_inplace x4, then greedy best-flip to a local optimum, 256 envs on the G14 shape."""
    th = ctx.th
    from synth import gset_like
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    edges = gset_like("G14")
    sim = EnvMaxcut(mygraph=edges, device=ctx.dev, if_bidirectional=True)
    st, n, m, envs = sim.store, sim.num_nodes, sim.num_edges, 256
    th.manual_seed(74)
    xs0 = sim.generate_xs_randomly(envs)
    xs = xs0.clone()
    sentinel = th.empty(())

    def block():
        for _ in range(100):
            sim.calculate_obj_values(xs)
        gx, gv = xs, sentinel
        for _ in range(4):
            gx, gv = sim.local_search_inplace(gx, gv, NUM_ITERS, NUM_SPIN, NOISE_STD)
        st.greedy_best_flip(gx, n, True)

    t0 = time.time()
    ms = ctx.timed(block, reps=10, warm=3, prepare=lambda: xs.copy_(xs0))
    t1 = time.time()
    ms_job = sum(ms) / len(ms)
    steps = envs * (100 + 4 * (1 + NUM_ITERS + n))            # greedy's flips are data dependent: not counted
    kms, share, per_rep = per_kernel_pass(ctx, st, block, 3, prepare=lambda: xs.copy_(xs0))
    alg = ls_alg_bytes(st, envs, n, m, NUM_ITERS)
    dom = max(share, key=share.get)
    gbs = alg.get(dom, 0) / (kms[dom] * 1e-3) / 1e9
    return {"workload": f"G14-shaped maxcut (N={n}, M={m}, synthetic seed 74), {envs} envs: calculate_obj_values x100 + "
                        f"local_search_inplace x4 + greedy best-flip to a local optimum (check_local_search_maxcut)",
            "n_gpus": 1, "ms": ms_job, "env_steps": steps, "env_steps_per_s": steps / (ms_job * 1e-3),
            "roofline": roof("hbm", gbs, ctx.peaks["hbm"], "GB/s", dom, algorithmic_bytes_per_launch=alg.get(dom),
                             ms_per_launch=kms[dom], share_of_block=share,
                             note="8 tiles on 148 SMs: every kernel of this configuration sits at the launch / latency "
                                  "floor (working set 205 KB, L2 resident); the HBM fraction is reported for the record"),
            "kernel_us": {k: round(v * 1e3, 2) for k, v in kms.items()}, "launches_per_block": per_rep,
            "clocks": ctx.window_clocks(t0, t1)}


# ----------------------------------------------------------------------------- config 3: G70 x 16384, strong scaling
def config3(ctx, inner=4, outer=2):
    """env_MCPG.py:449-476 (search_and_evaluate_local_search): broadcast the best row, 16 random flips per env,
    then `inner` x {LocalSearch.random_search(64, 4) + update_xs_by_vs}; one best-cut exchange per outer
    iteration.  16384 envs in total, split over the ranks (strong scaling)."""
    th = ctx.th
    from synth import gset_like
    from rlsolver_b200.dist import best_allreduce
    from rlsolver_b200.envs.env_MCPG import EnvMaxcut, LocalSearch, update_xs_by_vs
    total_envs = 16384
    envs = total_envs // ctx.world
    edges = gset_like("G70")
    sim = EnvMaxcut(mygraph=edges, device=ctx.dev, if_bidirectional=False)     # as env_MCPG.py:435 builds it
    st, n, m = sim.store, sim.num_nodes, sim.num_edges
    solver = LocalSearch(sim, n)
    th.manual_seed(74 + ctx.rank)
    best_xs = sim.generate_xs_randomly(envs)
    best_vs = sim.calculate_obj_values(best_xs)
    sim_ids = th.arange(envs, device=ctx.dev)
    iters, spin = 64, 4

    def outer_iteration():
        # the exchange: global best row (all-gather of 8 + N byte records), broadcast to every env of every rank
        cut, gid, row = best_allreduce(best_vs, best_xs, ctx.rank, ctx.world, envs)
        best_xs[:] = row
        best_vs[:] = cut
        xs = best_xs.clone()
        for _ in range(16):
            ids = th.randint(0, n, size=(envs,), device=ctx.dev)
            xs[sim_ids, ids] = th.logical_not(xs[sim_ids, ids])
        solver.reset(xs)
        for _ in range(inner):
            solver.random_search(num_iters=iters, num_spin=spin)
            update_xs_by_vs(best_xs, best_vs, solver.good_xs, solver.good_vs, True)

    t0 = time.time()
    ms = ctx.timed(outer_iteration, reps=outer, warm=1, flush=False)
    t1 = time.time()
    ms_job = ctx.job_ms(ms)
    steps = total_envs * (1 + inner * (iters + n))
    kms, share, per_rep = per_kernel_pass(ctx, st, lambda: solver.random_search(num_iters=iters, num_spin=spin), 1)
    alg = ls_alg_bytes(st, envs, n, m, iters)
    dom = max(share, key=share.get)
    gbs = alg.get(dom, 0) / (kms[dom] * 1e-3) / 1e9
    gen = next((k for k in ("ls_noise_masks", "ls_fused_search") if k in kms), None)
    noise_eq = iters * 4 * envs * n / (kms[gen] * 1e-3) / 1e9 if gen else None
    return {"workload": f"G70-shaped maxcut (N={n}, M={m}, synthetic seed 74), {total_envs} envs over {ctx.world} GPU(s) "
                        f"({envs}/GPU): outer iteration of env_MCPG.py:449-476 with {inner} x "
                        f"LocalSearch.random_search(num_iters=64, num_spin=4) (the reference runs 16) + one best-cut "
                        f"exchange",
            "n_gpus": ctx.world, "scaling": "strong", "ms": ms_job, "env_steps": steps,
            "env_steps_per_s": steps / (ms_job * 1e-3),
            "exchange": "best_record kernel + ncclAllGather of world x (8 + N) bytes + best_pick kernel, once per outer iteration",
            "roofline": roof("hbm", gbs, ctx.peaks["hbm"], "GB/s", dom, algorithmic_bytes_per_launch=alg.get(dom),
                             ms_per_launch=kms[dom], share_of_random_search=share, noise_equivalent_gbs=noise_eq,
                             traffic=dram_traffic().get("config3_" + dom),
                             note="one random_search(64, 4) call: the mask generator recomputes 64 draws of randn "
                                  "[E, N] in registers (issue bound); noise_equivalent_gbs = the float32 noise a "
                                  "streaming implementation reads over the same time"),
            "kernel_ms": {k: round(v, 4) for k, v in kms.items()}, "clocks": ctx.window_clocks(t0, t1)}


# ----------------------------------------------------------------------------- config 4: 10^6 BA / ER instances
def config4(ctx, total_envs):
    """Distribution-wise pattern-I step: one graph per env (BA m=4 and ER p=0.15, N=100, +-1 weights), uniformly random
    actions.  Two step/reward patterns on the same graphs:
      * "s2v" -- what the config names (train_S2V.py:37-47, ECO_S2V/src/envs/spinsystem.py): irreversible spins from all
        +1, the single SPIN_STATE observable, DENSE reward = delta cut / N; every env flips its spins in a random order;
      * "eco" -- the batched PECO environment's training configuration (train_PECO.py:34-44, spinsystem_PECO.py:306-486):
        reversible spins, seven observables, BLS reward."""
    th = ctx.th
    from rlsolver_b200.envs import env_PECO as P
    n = 100
    envs = total_envs // ctx.world
    out = {}
    for kind in ("BA", "ER"):
        for style in ("s2v", "eco"):
            t_gen = time.time()
            if kind == "BA":
                gg = P.RandomBAGraphGenerator(n_spins=n, m_insertion_edges=4, edge_type=P.EdgeType.DISCRETE, num_envs=envs,
                                              device=ctx.dev)
            else:
                gg = P.RandomERGraphGenerator(n_spins=n, p_connection=0.15, edge_type=P.EdgeType.DISCRETE, num_envs=envs,
                                              device=ctx.dev)
            th.manual_seed(74 + ctx.rank)
            s2v = style == "s2v"
            env = P.SpinSystemFactory.get(gg, n if s2v else 2 * n,
                                          observables=P.S2V_OBSERVABLES if s2v else P.ECO_PECO_OBSERVABLES,
                                          reward_signal=P.RewardSignal.DENSE if s2v else P.RewardSignal.BLS,
                                          extra_action=P.ExtraAction.NONE, optimisation_target=P.OptimisationTarget.CUT,
                                          spin_basis=P.SpinBasis.BINARY, norm_rewards=True, memory_length=None,
                                          horizon_length=None, stag_punishment=None, basin_reward=None,
                                          reversible_spins=not s2v, device=ctx.dev, num_envs=envs)
            th.cuda.synchronize()
            gen_s = time.time() - t_gen
            if s2v:         # a random order of every env's spins: no spin is flipped twice (what the agent's mask enforces)
                order = th.rand((envs, n), device=ctx.dev).argsort(dim=1)[:, :32].t().contiguous()
                acts = [order[i] for i in range(32)]
                del order
            else:
                acts = [th.randint(0, n, (envs,), device=ctx.dev) for _ in range(8)]
            k = [0]

            def step():
                env.step(acts[k[0] % len(acts)], return_observation=False)
                k[0] += 1

            t0 = time.time()
            ms = ctx.timed(step, reps=20, warm=3, flush=False)         # state >> L2: every step streams from HBM
            t1 = time.time()
            assert k[0] <= len(acts) or not s2v
            ms_job = ctx.job_ms(ms)
            alg = env.step_algorithmic_bytes() if hasattr(env, "step_algorithmic_bytes") else \
                envs * (4 * n + 8 * n + 2 * 4 * n * 4 + 40)
            gbs = alg / (ms_job * 1e-3) / 1e9
            deg = float(env.mean_degree()) if hasattr(env, "mean_degree") else \
                float((env.matrix[:1024] != 0).float().sum() / min(envs, 1024) / n)
            what = ("S2V-DQN pattern: irreversible spins, SPIN_STATE observable, DENSE reward / N, random flip order" if s2v
                    else "PECO pattern: reversible spins, 7 observables, BLS reward, random actions")
            out[kind if not s2v else kind + "_s2v"] = {
                "workload": f"{kind}-100 per-env graphs (+-1 weights), {total_envs} envs over {ctx.world} GPU(s) "
                            f"({envs}/GPU), mean degree {deg:.1f}: SpinSystemUnbiased.step, {what}",
                "n_gpus": ctx.world, "scaling": "strong", "ms": ms_job, "env_steps": total_envs,
                "env_steps_per_s": total_envs / (ms_job * 1e-3), "graph_generation_s": round(gen_s, 2),
                "state_layout": getattr(env, "state_layout", "dense float32 matrix [E,N,N] + state [E,7,N]"),
                "roofline": roof("hbm", gbs, ctx.peaks["hbm"], "GB/s", "peco_step",
                                 algorithmic_bytes_per_launch=alg, bytes_per_env_step=alg / envs,
                                 traffic=dram_traffic().get("config4_peco_step_" + kind)),
                "clocks": ctx.window_clocks(t0, t1)}
            if ctx.world == 1 and ctx.rank == 0 and not ctx.args.no_cpu_baseline and kind == "ER" and not s2v:
                out["cpu_baseline"] = peco_cpu_baseline(env, n)
            del env, gg, acts
            th.cuda.empty_cache()
    return out


def peco_cpu_baseline(env, n, sample_envs=4096, steps=10):
    """Pattern-I CPU baseline: the NumPy restatement of the reference's SpinSystemUnbiased.step
    (oracle/peco.py: every local field recomputed with a batched matmul, the state cloned, like the reference) on a
    sample of the same graphs, host cores."""
    import numpy as np
    from oracle import peco as op
    eco = [op.SPIN, op.IMM, op.TSF, op.DSCORE, op.DSTATE, op.GREEDY, op.TERM]
    from rlsolver_b200.envs.env_PECO import CompactGraphs
    c = env._compact
    sub = CompactGraphs(c.adj[:sample_envs].contiguous(),
                        None if c.sgn is None else (c.sgn if c.sgn.dim() == 2 else c.sgn[:sample_envs].contiguous()), n)
    matrix = sub.dense().cpu().numpy()
    spins = env.state[:sample_envs, 0, :].cpu().numpy() if env.num_envs <= (1 << 17) else \
        (2.0 * np.random.default_rng(0).integers(0, 2, (sample_envs, n)) - 1.0).astype(np.float32)
    ref = op.SpinSystem(matrix, spins, eco, 2 * n, op.BLS, True, scalar_div_as_cuda=True)
    rng = np.random.default_rng(1)
    acts = [rng.integers(0, n, sample_envs) for _ in range(steps + 1)]
    ref.step(acts[0])
    t = time.perf_counter()
    for k in range(steps):
        ref.step(acts[1 + k])
    dt = time.perf_counter() - t
    return {"value": sample_envs * steps / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{sample_envs} of the envs x {steps} steps, oracle/peco.py (NumPy float32, batched matmul per step "
                      f"like spinsystem_PECO.py:306-486)"}


# ----------------------------------------------------------------------------- config 5: dense QUBO
def config5(ctx, total_chains=8192):
    """Dense float QUBO, N = 4096: x^T Q x for every chain on the tensor cores (MCPG/sampling.py:339-340) and one
    coordinate-ascent sweep (sampling.py:331-337); chains split over the ranks, Q replicated."""
    th = ctx.th
    from rlsolver_b200.qubo import QuboModel
    n = 4096
    chains = total_chains // ctx.world
    th.manual_seed(0)
    u = th.randn(n, n, device=ctx.dev)
    q = th.triu(u) + th.triu(u, 1).T
    model = QuboModel(q)
    th.manual_seed(1 + ctx.rank)
    x = th.randint(0, 2, (n, chains), device=ctx.dev).float() * 2 - 1
    out = {}
    t0 = time.time()
    ms = ctx.timed(lambda: model.energy(x), reps=10, warm=3)
    ms_e = ctx.job_ms(ms)
    xs_ = x.clone()
    ms = ctx.timed(lambda: model.sweeps(xs_, 1), reps=3, warm=1)
    ms_s = ctx.job_ms(ms)
    ms = ctx.timed(lambda: (x * (q @ x)).sum(0), reps=3, warm=1)
    ms_t = ctx.job_ms(ms)
    t1 = time.time()
    flops = 3 * 2.0 * n * n * chains                      # bf16 issued per GPU: 3 limbs of Q
    for tag, ms_k, steps, what in (("energy", ms_e, total_chains, "QuboModel.energy (x^T Q x per chain)"),
                                   ("sweep", ms_s, total_chains * n, "QuboModel.sweeps(1): N coordinate updates per chain")):
        tf = flops / (ms_k * 1e-3) / 1e12
        out[tag] = {"workload": f"dense QUBO N={n} fp32 (3 exact bf16 limbs), {total_chains} chains over {ctx.world} GPU(s) "
                                f"({chains}/GPU): {what}",
                    "n_gpus": ctx.world, "scaling": "strong", "ms": ms_k, "env_steps": steps,
                    "env_steps_per_s": steps / (ms_k * 1e-3),
                    "roofline": roof("tensor", tf, ctx.peaks["tensor"], "TFLOP/s", "qubo_" + tag + "_kernel",
                                     flops_issued_per_launch=flops, useful_tflops=tf / 3,
                                     peak_note="sustained cuBLAS bf16 (MEASURED_PEAKS.json bf16_tflops_sustained); "
                                               "useful fp32-accurate flops are 1/3 of the issued bf16 flops")}
    out["torch_fp32_same_gpu_ms"] = ms_t
    out["clocks"] = ctx.window_clocks(t0, t1)
    return out


# ----------------------------------------------------------------------------- same-GPU reference arm
def gpu_reference(ctx, envs):
    """The reference's algorithm, op for op in torch (oracle/torch_port.py), on the SAME B200: the FULL headline
    workload (4096 envs, all 8 iterations, all 2000 sweep nodes).  Isolates 'GPU vs CPU' from the algorithm."""
    th = ctx.th
    from oracle import torch_port as tp
    from synth import gset_like
    edges = gset_like(GRAPH)
    sim = tp.TorchSim(edges, True, device=ctx.dev)
    th.manual_seed(74)
    xs0 = sim.random_xs(envs)
    times = []
    for it in range(2):
        xs = xs0.clone()
        th.cuda.synchronize()
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        sim.local_search_inplace(xs, None, NUM_ITERS, NUM_SPIN, NOISE_STD)
        b.record()
        th.cuda.synchronize()
        if it >= 1:
            times.append(a.elapsed_time(b))
    ms = sum(times) / len(times)
    per_call = env_steps_per_call(envs, sim.num_nodes, NUM_ITERS, sim.num_nodes)
    peak_mem = th.cuda.max_memory_allocated(ctx.dev)
    del sim
    th.cuda.empty_cache()
    return {"impl": "oracle/torch_port.py (torch ops as the reference uses them: 3 int64 [E, Md] index tensors, one full "
                    "objective evaluation per candidate flip) on the same GPU, full workload, 1 warm-up + 1 timed step",
            "ms_per_step": ms, "value": per_call / (ms * 1e-3), "unit": UNIT, "envs": envs,
            "peak_memory_gb": round(peak_mem / 2 ** 30, 2)}


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    ctx = Ctx(args)
    th, dist = ctx.th, ctx.dist
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    from synth import gset_like

    import rlsolver_b200
    from rlsolver_b200.dist import BestExchange, PeerBestExchange
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    from rlsolver_b200.host_pipeline import HostPipeline

    rlsolver_b200.build()
    log(f"world {world}: library ready")

    envs = args.envs
    edges = gset_like(GRAPH)
    sim = EnvMaxcut(mygraph=edges, device=dev, if_bidirectional=True)
    n = sim.num_nodes
    th.manual_seed(74 + rank)
    xs0 = sim.generate_xs_randomly(envs)
    xs = xs0.clone()
    sentinel = th.empty(())
    flush = ctx.flush
    small_flush_steps = flush[:160 << 20]                          # > 126 MB L2 (timed when the exchange is pipelined)
    barrier = ctx.barrier
    # the path's only exchange: one kernel over NVLink peer memory (PeerBestExchange); NCCL all-gather form on request
    # (--exchange nccl, --pipelined-exchange) or when CUDA IPC between the ranks is not available on the box
    exchange, exchange_kind = None, None
    if world > 1:
        if args.exchange == "peer" and not args.pipelined_exchange:
            try:
                exchange, exchange_kind = PeerBestExchange(n, rank, world, envs, dev), "peer"
            except RuntimeError as exc:      # raised on every rank alike (the constructor agrees on the outcome)
                log(f"peer exchange not available: {str(exc)[:200]}")
        if exchange is None:
            exchange, exchange_kind = BestExchange(n, rank, world, envs, dev), "nccl"
    in_graph = exchange_kind == "peer"       # a plain kernel launch: it is captured with the local search
    ex_stream = th.cuda.Stream(device=dev) if world > 1 else None
    posted = th.cuda.Event() if world > 1 else None

    def local_part():
        return sim.local_search_inplace(xs, sentinel, NUM_ITERS, NUM_SPIN, NOISE_STD)

    def step_body():
        """The local part of a step: fixed launch sequence, captured into a CUDA graph below."""
        gx, gv = local_part()
        return gx, gv, (exchange(gv, gx) if in_graph else None)

    state = {"graph": None, "out": None}

    def one_step():
        """One step: the local search and -- with several ranks -- the path's only exchange (best cut, its argmax, the
        winner's spins) right behind it on the same stream.  The peer-memory exchange is one kernel and part of the
        captured graph.  The NCCL form stays OUTSIDE the graph: capturing the all-gather hung both ranks on this
        torch / NCCL build (round 2, N = 2), so it is three eager launches on preallocated buffers (record kernel,
        ncclAllGather, pick kernel), no host sync."""
        if state["graph"] is not None:
            state["graph"].replay()
            gx, gv, best = state["out"]
        else:
            gx, gv, best = step_body()
        if exchange is not None and not in_graph:
            if not args.pipelined_exchange:
                best = exchange(gv, gx)
            else:
                # pipelined: the record kernel (the only part that reads this step's xs / vs) is ordered before the
                # next step; all-gather + pick run on their own stream under the next step's local search.  The
                # timed region ends only after the last exchange has finished (timed_steps joins the stream).
                cur = th.cuda.current_stream(dev)
                ex_stream.wait_stream(cur)
                with th.cuda.stream(ex_stream):
                    exchange.post(gv, gx)
                    posted.record(ex_stream)
                    best = exchange.finish()
                cur.wait_event(posted)
        return gx, gv, best

    def timed_steps(k):
        """Per-step CUDA-event times.  With the pipelined exchange the steps are timed as ONE region (first start to
        the end of the last exchange) and the total is spread evenly: the L2-evicting write between steps is then
        inside the timed region."""
        if exchange is not None and args.pipelined_exchange:
            cur = th.cuda.current_stream(dev)
            a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            xs.copy_(xs0)
            a.record()
            for i in range(k):
                if i:
                    xs.copy_(xs0)
                    small_flush_steps.zero_()
                one_step()
            cur.wait_stream(ex_stream)
            b.record()
            th.cuda.synchronize()
            return [a.elapsed_time(b) / k] * k
        evs = []
        for _ in range(k):
            xs.copy_(xs0)
            flush.zero_()                                           # evict L2 between steps (untimed)
            a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            a.record()
            one_step()
            b.record()
            evs.append((a, b))
        th.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    def try_capture():
        """The local part of the step is a fixed sequence of launches of this library's kernels: capture it once into
        a CUDA graph so its host cost is one replay.  The device-resident generator cursor keeps the random stream
        identical to eager execution."""
        if args.no_graph:
            return "disabled (--no-graph)"
        try:
            side = th.cuda.Stream(device=dev)
            side.wait_stream(th.cuda.current_stream(dev))
            with th.cuda.stream(side):
                for _ in range(3):
                    step_body()
            th.cuda.current_stream(dev).wait_stream(side)
            th.cuda.synchronize()
            sim.store.rng_cursor_sync()        # the captured kernels read the generator state from device memory
            g = th.cuda.CUDAGraph()
            counted = sim.store.launch_count
            with th.cuda.graph(g):
                out = step_body()
            th.cuda.synchronize()
            # launches of this library inside the captured step (the threshold draw and the cursor update are this
            # library's kernels here; in eager mode the draw is torch's own call)
            state["graph_launches"] = sim.store.launch_count - counted
            state["graph"], state["out"] = g, out
            return "captured (local search" + (")" if world == 1 else " + peer-memory best-cut exchange)" if in_graph else
                                               "); best-cut exchange eager behind the replay")
        except Exception as exc:                                   # noqa: BLE001 - eager remains correct
            state["graph"] = None
            th.cuda.synchronize()
            return f"eager ({type(exc).__name__}: {str(exc)[:80]})"

    # warm-up: W steps plus a fixed number of settle steps (~0.5 s of load so the clocks settle).  The
    # count must not depend on wall time: every rank has to issue the same sequence of collectives.
    t_w = time.time()
    launches_before = sim.store.launch_count
    timed_steps(1)
    launches_per_step = sim.store.launch_count - launches_before + (0 if world == 1 else 1 if in_graph else 2)
    log("first eager step done")
    graph_status = try_capture()
    log(f"graph: {graph_status}")
    if world > 1:       # all ranks must agree on the mode (the collective sequence is part of it)
        flag = th.tensor([1 if state["graph"] is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and state["graph"] is not None:
            state["graph"], graph_status = None, "eager (another rank could not capture)"
    for _ in range(max(3, args.warmup) + SETTLE_STEPS):
        timed_steps(1)
    barrier()
    t0 = time.time()
    ms = timed_steps(args.steps)
    barrier()
    t1 = time.time()
    log("timed steps done")
    if state["graph"] is not None and state.get("graph_launches"):
        launches_per_step = state["graph_launches"] + (0 if world == 1 else 1 if in_graph else 2)
    launches = launches_per_step * args.steps          # kernels of this library per step (counted while capturing it)
    total_ms = ctx.max_over_ranks(sum(ms))
    per_call = env_steps_per_call(envs, n, NUM_ITERS, n)
    value = per_call * world * args.steps / (total_ms * 1e-3)

    # exchange cost alone (multi-GPU), eager, CUDA events
    exch_us = None
    if world > 1:
        gx, gv = xs, sim.calculate_obj_values(xs)
        ex_ms = ctx.timed(lambda: exchange(gv, gx), reps=50, warm=5, flush=False)
        exch_us = 1e3 * ctx.job_ms(ex_ms)
        if exchange_kind == "peer":
            calls, timeouts = exchange.status()
            if timeouts:
                raise RuntimeError(f"peer exchange: {timeouts} polls timed out in {calls} calls -- the ranks did not make "
                                   f"the same sequence of calls")

    # ---- end to end from pinned host buffers through the public API.  Every step copies its spins from pinned host
    # memory to the device, runs the step, and copies spins + values back.  Two host layouts: `bool` = the
    # reference's bool [E, N] rows (one byte per spin, what EnvMaxcut.local_search_inplace takes), and `packed` =
    # the library's packed tiles (uint32 [E/32, Np], one bit per spin: rlsb_ls_begin_packed in, the workspace's
    # packed section out), through EnvMaxcut.local_search_packed.  For each: `serial` (copy in, compute, copy out,
    # one event pair per step) and `pipelined` (rlsolver_b200.host_pipeline.HostPipeline: three streams, three device
    # buffers, one captured graph of the step per buffer, H2D(i+1) | step(i) | D2H(i-1); one event pair around all K
    # steps, fill and drain included).  The headline e2e is pipelined / bool (the
    # reference-facing call); the packed figures are reported next to it.
    cur = th.cuda.current_stream(dev)
    small_flush = flush[:160 << 20]                                # > 126 MB L2, inside the timed region
    pk0 = sim.store.pack(xs0)

    def make_layout(kind):
        if kind == "bool":
            h_in = xs0.cpu().pin_memory()
            d_in = th.empty_like(xs0)

            def run(src):
                xs.copy_(src)
                gx, gv, _ = one_step()
                return gx, gv
        else:
            h_in = pk0.cpu().pin_memory()
            d_in = th.empty_like(pk0)

            def run(src):
                pk, gv = sim.local_search_packed(src, NUM_ITERS, NUM_SPIN, NOISE_STD)
                if exchange is not None:
                    exchange.packed(gv, pk, sim.store)
                return pk, gv
        h_out = th.empty_like(h_in).pin_memory()
        h_vs = th.empty((envs,), dtype=th.int64).pin_memory()
        return h_in, h_out, h_vs, d_in, run

    def e2e_measure(kind):
        h_in, h_out, h_vs, d_in, run = make_layout(kind)

        def serial_step():
            d_in.copy_(h_in, non_blocking=True)
            gx, gv = run(d_in)
            h_out.copy_(gx, non_blocking=True)
            h_vs.copy_(gv, non_blocking=True)

        for _ in range(3):
            serial_step()
        barrier()
        ser = []
        for _ in range(args.steps):
            flush.zero_()
            a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            a.record()
            serial_step()
            b.record()
            th.cuda.synchronize()
            ser.append(a.elapsed_time(b))
        barrier()
        serial_ms = ctx.max_over_ranks(sum(ser)) / args.steps

        # the pipelined form is the package's HostPipeline: three device buffers, one captured graph of the step per
        # buffer (L2-evicting write + local search in place + the peer-memory exchange), H2D / step / D2H on three streams
        if kind == "bool":
            in_post = (lambda res, vs: exchange(vs, res)) if in_graph else None
            eager_post = (lambda res, vs: exchange(vs, res)) if exchange is not None and not in_graph else None
        else:
            in_post = (lambda res, vs: exchange.packed(vs, res, sim.store)) if in_graph else None
            eager_post = (lambda res, vs: exchange.packed(vs, res, sim.store)) if exchange is not None and not in_graph else None
        pipe = HostPipeline(sim, envs, layout=kind, num_iters=NUM_ITERS, num_spin=NUM_SPIN, noise_std=NOISE_STD,
                            pre_step=small_flush.zero_, in_graph_post=in_post, eager_post=eager_post)

        def pipelined(k):
            ta, tb = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            ta.record(cur)
            for _ in range(k):
                pipe.submit(h_in, h_out, h_vs)
            pipe.join()
            tb.record(cur)
            th.cuda.synchronize()
            return ta.elapsed_time(tb)

        pipelined(4)
        barrier()
        pipe_ms = ctx.max_over_ranks(pipelined(args.steps)) / args.steps
        barrier()
        nbytes = h_in.numel() * h_in.element_size()
        return {"ms_per_step": pipe_ms, "value": per_call * world / (pipe_ms * 1e-3), "serial_ms_per_step": serial_ms,
                "serial_value": per_call * world / (serial_ms * 1e-3), "h2d_bytes_per_step": nbytes,
                "d2h_bytes_per_step": nbytes + 8 * envs}

    e2e_bool = e2e_measure("bool")
    log("e2e (bool rows) done")
    e2e_packed = e2e_measure("packed")
    log("e2e (packed) done")
    t_headline_end = time.time()

    # ---- per-kernel pass (CUDA events around every launch of this library) for the roofline
    saved_graph, state["graph"] = state["graph"], None          # per-kernel events need eager launches
    kernel_ms, share, _ = per_kernel_pass(ctx, sim.store, lambda: step_body(), args.steps,
                                          prepare=lambda: xs.copy_(xs0))
    state["graph"] = saved_graph
    dom = max(share, key=share.get)
    alg_bytes = ls_alg_bytes(sim.store, envs, n, sim.num_edges, NUM_ITERS)
    peak, peak_src = ctx.peaks["hbm"], ctx.peaks["source"] + " hbm_gbs"
    achieved = alg_bytes[dom] / (kernel_ms[dom] * 1e-3) / 1e9
    gen = next((k for k in ("ls_noise_masks", "ls_fused_search") if k in kernel_ms), None)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": dram_traffic().get(dom), "peak_source": peak_src,
                "ms_per_launch": kernel_ms[dom], "algorithmic_bytes_per_launch": alg_bytes[dom],
                "share_of_step": share,
                "all_kernels_gbs": {k: alg_bytes[k] / (kernel_ms[k] * 1e-3) / 1e9 for k in kernel_ms if k in alg_bytes},
                "pass": "separate K-step pass with CUDA events around each call of this library, single stream",
                "note": ("the step does not stream noise: torch's Philox/Box-Muller stream is recomputed in registers "
                         "(issue-bound, see profiles/), the algorithmic HBM bytes of that kernel are the early-out bytes "
                         "and the mask bits; noise_equivalent_gbs = the 4 B per element and draw a streaming "
                         "implementation reads, over the same time"),
                "noise_equivalent_gbs": (NUM_ITERS * 4 * envs * n / (kernel_ms[gen] * 1e-3) / 1e9 if gen else None),
                "whole_step_noise_equivalent_frac": ((1 + NUM_ITERS) * 4 * envs * n / (total_ms / args.steps * 1e-3)
                                                     / 1e9 / peak)}

    log("per-kernel pass done")
    cpu = gpu_ref = None
    if world == 1 and not args.no_cpu_baseline:
        if rank == 0:
            r = cpu_reference_run(steps=3, warmup=1, budget_s=4.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "same_config_basis",
                                     "full_workload_seconds_per_step_extrapolated")}
            try:
                gpu_ref = gpu_reference(ctx, envs)
                gpu_ref["device_timed_ratio"] = value / gpu_ref["value"]
            except Exception as exc:                               # noqa: BLE001
                gpu_ref = {"unavailable": f"{type(exc).__name__}: {str(exc)[:120]}"}
                th.cuda.empty_cache()

    # ---- the other BASELINE configurations
    want = set() if args.configs == "none" else set(args.configs.split(","))
    configs = {}

    def attempt(name, fn):
        try:
            barrier()
            log(f"{name} ...")
            configs[name] = fn()
        except Exception as exc:                                   # noqa: BLE001 - one config must not sink the line
            configs[name] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
            th.cuda.empty_cache()

    del flush
    if "1" in want and world == 1:
        attempt("config1_G14_256", lambda: config1(ctx))
    if "3" in want:
        attempt("config3_G70_16384", lambda: config3(ctx))
    if "5" in want:
        attempt("config5_QUBO_4096_8192", lambda: config5(ctx))
    if "4" in want:
        attempt("config4_BA_ER_100_1M", lambda: config4(ctx, args.peco_envs))
    clock_info = ctx.clocks.stop(t_w, t_headline_end) if ctx.clocks else None

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32 bit-packed spins / int64 values / f32 noise",
                "data": "synthetic",
                "config": {"workload": workload_name(envs), "envs_per_gpu": envs, "nodes": n, "edges": sim.num_edges,
                           "env_steps_per_step_per_gpu": per_call, "l2": "flushed between steps (256 MiB write)", "cuda_graph": graph_status,
                           "rng": "torch's CUDA Philox stream, 1+8 draws of randn [E,N] f32 per step inside the timed region, "
                                  "recomputed in place (only thresholds / flip bits leave the kernels)",
                           "multi_gpu": ("env batch sharded, graph replicated, one best-cut exchange per step: ONE kernel over "
                                         "NVLink peer memory (every rank stores its 8+N B record into each peer's mailbox, "
                                         "raises an arrival word, polls its own, picks the winner), captured in the step's "
                                         "CUDA graph, no NCCL call, no host sync" if exchange_kind == "peer" else
                                         "env batch sharded, graph replicated, one best-cut exchange per step behind the "
                                         "graph replay: best_record kernel + ncclAllGather of world x (8+N) B + "
                                         "best_pick kernel on preallocated buffers, no host sync; "
                                         + ("serial behind every step" if not args.pipelined_exchange else
                                            "pipelined: only the record kernel is ordered before the next step, all-gather "
                                            "+ pick run on a second stream under it; K steps timed as one region that "
                                            "ends after the last exchange, the 160 MiB L2-evicting write per step included")),
                           "exchange": exchange_kind, "exchange_us": exch_us},
                "e2e": {"value": e2e_bool["value"], "unit": UNIT, "h2d_bytes_per_step": e2e_bool["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": e2e_bool["d2h_bytes_per_step"], "ms_per_step": e2e_bool["ms_per_step"],
                        "layout": "bool [E, N] rows (the reference's layout), EnvMaxcut.local_search_inplace",
                        "mode": ("pipelined (rlsolver_b200.host_pipeline.HostPipeline): H2D(i+1) | step(i) | D2H(i-1) on three "
                                 "streams, three device buffers, the step a captured CUDA graph per buffer that runs in "
                                 "place on it; one event pair around all K steps (fill + drain and a 160 MiB L2-evicting "
                                 "write per step included)"),
                        "serial_ms_per_step": e2e_bool["serial_ms_per_step"], "serial_value": e2e_bool["serial_value"],
                        "packed": dict(e2e_packed, layout="packed tiles uint32 [E/32, Np] (one bit per spin), "
                                                          "EnvMaxcut.local_search_packed / rlsb_ls_begin_packed")},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "gpu_reference": gpu_ref,
                "configs": configs, "clocks": clock_info}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


class _QuietStdout:
    """Everything libraries write to fd 1 while the bench runs (NCCL's version banner, ...) goes to stderr;
    stdout carries exactly the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def emit(line):
    _RESULT.append(json.dumps(line))


_RESULT = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=NUM_ENVS)
    ap.add_argument("--configs", default="1,3,4,5", help="other BASELINE configs to measure ('none' to skip)")
    ap.add_argument("--peco-envs", type=int, default=1 << 20, help="total envs of config 4 (10^6 instances)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipelined-exchange", action="store_true",
                    help="N > 1: order only the record kernel before the next step and run all-gather + pick under it "
                         "(K steps timed as one region, L2-evicting writes inside); default: serial behind every step")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="N > 1: the best-cut exchange as one kernel over NVLink peer memory (default) or as record kernel + "
                         "ncclAllGather + pick kernel")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    with _QuietStdout():
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    for text in _RESULT:
        print(text, flush=True)


if __name__ == "__main__":
    main()
