"""Round-2 parity tests of the CUDA path (through the C ABI / the reference-mirroring classes):

* the generator streaming NEXT TO the tile kernel (rlsb_ls_fused_search) against the sequential form and the oracle;
* the BASELINE config-3 shape at full size (G70-like, 16384 envs) against the NumPy oracle on sampled rows;
* packed-tile entry points (rlsb_ls_begin_packed, packed best-cut record) against the bool-row ones;
* the RNG first-use self-check;
* a 2-rank NCCL run against the single-GPU run of each shard (needs 2 GPUs: `gpurun --gpus 2`).
Integer / bool results must be bit-exact.  Needs a B200: run with `-m gpu`.
"""
import os
import socket

import numpy as np
import pytest
import torch as th

from oracle import maxcut as om
from synth import gset_like, random_graph

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _rand_xs(e, n, seed):
    rng = np.random.default_rng(seed)
    xs = rng.integers(0, 2, size=(e, n)).astype(bool)
    xs[:, 0] = False
    return xs


# ------------------------------------------------------------------ generator next to the tile kernel
@pytest.mark.parametrize("name,envs,iters,bidir", [("G22", 4096, 8, True), ("G14", 256, 8, True), ("G70", 9600, 5, False),
                                                   ("N37", 45, 3, False), ("N100", 33, 7, True), ("G22", 1, 8, True),
                                                   ("G22", 4097, 2, False), ("G14", 6000, 9, True)])
def test_overlapped_generator_equals_sequential(name, envs, iters, bidir, cuda_device):
    """rlsb_ls_fused_search (generator on the side stream, tile CTAs waiting per group of draws) against
    rlsb_ls_noise_masks followed by rlsb_ls_run_masks: same states, same values, same generator state."""
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    from rlsolver_b200.methods.LocalSearch import LocalSearch
    edges = (random_graph(37, 90, seed=3) if name == "N37" else random_graph(100, 320, seed=8) if name == "N100"
             else gset_like(name))
    out = []
    for overlap in (False, True):
        sim = EnvMaxcut(mygraph=edges, device=cuda_device, if_bidirectional=bidir)
        sim.store.overlap_generator = overlap
        th.manual_seed(5)
        xs = sim.generate_xs_randomly(envs)
        xs, vs = sim.local_search_inplace(xs, th.empty(()), num_iters=iters, num_spin=6, noise_std=0.3)
        xs, vs = sim.local_search_inplace(xs, vs, num_iters=iters + 1, num_spin=3, noise_std=0.4)
        res = [xs.clone(), vs.clone()]
        if not bidir:
            solver = LocalSearch(sim, sim.num_nodes)
            solver.reset(xs.clone())
            x2, v2, _ = solver.random_search(num_iters=iters, num_spin=4)
            res += [x2.clone(), v2.clone()]
        res.append(th.cuda.get_rng_state(cuda_device))
        out.append(res)
        assert th.equal(sim.calculate_obj_values(res[0]), res[1])
    assert all(th.equal(a, b) for a, b in zip(*out))


def test_overlapped_generator_many_calls_and_graph_replay(cuda_device):
    """Back-to-back fused calls reuse the side stream, the events and the counters; replays of a captured call
    continue the random stream like eager calls."""
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    edges = gset_like("G22")
    sims = [EnvMaxcut(mygraph=edges, device=cuda_device, if_bidirectional=True) for _ in range(2)]
    sims[0].store.overlap_generator = False
    sims[1].store.overlap_generator = True
    envs = 1024
    th.manual_seed(11)
    x0 = sims[0].generate_xs_randomly(envs)
    sentinel = th.empty(())
    # eager reference: 5 consecutive calls of the sequential form
    th.manual_seed(3)
    xa = x0.clone()
    for _ in range(5):
        xa, va = sims[0].local_search_inplace(xa, sentinel)
    state_a = th.cuda.get_rng_state(cuda_device)
    # overlapped form: 2 eager calls, then one captured call replayed 3 times
    th.manual_seed(3)
    xb = x0.clone()
    for _ in range(2):
        sims[1].local_search_inplace(xb, sentinel)
    sims[1].store.rng_cursor_sync()
    side = th.cuda.Stream(device=cuda_device)
    side.wait_stream(th.cuda.current_stream(cuda_device))
    g = th.cuda.CUDAGraph()
    with th.cuda.stream(side):
        with th.cuda.graph(g, stream=side):
            gx, gv = sims[1].local_search_inplace(xb, sentinel)
    th.cuda.current_stream(cuda_device).wait_stream(side)
    th.cuda.synchronize()
    for _ in range(3):
        g.replay()
    th.cuda.synchronize()
    sims[1].store.rng_cursor_commit()
    assert th.equal(xa, gx) and th.equal(va, gv)
    assert th.equal(th.cuda.get_rng_state(cuda_device), state_a)


# ------------------------------------------------------------------ config 3 at full size vs the oracle
def _g70_setup(cuda_device, envs):
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    edges = gset_like("G70")
    g = om.build_graph_store(edges, False)
    sim = EnvMaxcut(mygraph=edges, device=cuda_device, if_bidirectional=False)
    xs_np = _rand_xs(envs, g.num_nodes, 70)
    rows = np.arange(7, envs, 128)                     # 128 sampled envs, every tile position hit
    return g, sim, xs_np, rows


def _sampled_draws(cuda_device, envs, n, count, rows, seed):
    """The rows `rows` of the `count` draws torch.randn((envs, n)) returns from `seed` (one draw resident at a time)."""
    th.manual_seed(seed)
    idx = th.from_numpy(rows).to(cuda_device)
    out = []
    for _ in range(count):
        out.append(th.randn((envs, n), device=cuda_device)[idx].cpu().numpy())
    return out, th.cuda.get_rng_state(cuda_device)


def test_local_search_inplace_g70_16384_vs_oracle(cuda_device):
    """BASELINE config 3 shape at full size: G70-like (10000 nodes), 16384 envs, local_search_inplace with the
    defaults.  The oracle replays 128 sampled envs with the whole batch's spread (oracle.maxcut.batch_rd_std);
    every other row is checked through the size-independent properties (value == cut of the row, not worse)."""
    envs = 16384
    g, sim, xs_np, rows = _g70_setup(cuda_device, envs)
    n = g.num_nodes
    draws, after = _sampled_draws(cuda_device, envs, n, 9, rows, seed=74)
    th.manual_seed(74)
    xs = th.from_numpy(xs_np.copy()).to(cuda_device)
    gx, gv = sim.local_search_inplace(xs, th.empty(()))
    assert th.equal(th.cuda.get_rng_state(cuda_device), after)
    rd = om.batch_rd_std(g, xs_np, 1, 0.3)
    want_xs, want_vs = om.local_search_inplace(g, xs_np[rows].copy(), None, draws, 8, 8, 0.3, literal=False, rd_std=rd)
    got_xs, got_vs = _np(gx), _np(gv)
    assert np.array_equal(got_xs[rows], want_xs) and np.array_equal(got_vs[rows], want_vs)
    assert np.array_equal(om.cut_values(g, got_xs), got_vs)
    assert (got_vs >= om.cut_values(g, xs_np)).all()


def test_random_search_64_4_g70_16384_vs_oracle(cuda_device):
    """The inner call of config 3's loop (env_MCPG.py:463): LocalSearch.random_search(num_iters=64, num_spin=4) on
    16384 envs of the G70 shape, 128 sampled envs replayed by the oracle with the 64 draws torch returns."""
    from rlsolver_b200.methods.LocalSearch import LocalSearch
    envs = 16384
    g, sim, xs_np, rows = _g70_setup(cuda_device, envs)
    n = g.num_nodes
    draws, after = _sampled_draws(cuda_device, envs, n, 64, rows, seed=75)
    th.manual_seed(75)
    solver = LocalSearch(sim, n)
    vs0 = solver.reset(th.from_numpy(xs_np.copy()).to(cuda_device))
    assert np.array_equal(_np(vs0), om.cut_values(g, xs_np))
    gx, gv, _ = solver.random_search(num_iters=64, num_spin=4)
    assert th.equal(th.cuda.get_rng_state(cuda_device), after)
    rd = om.batch_rd_std(g, xs_np, 2, 0.3)
    ref = om.LocalSearch(g)
    ref.reset(xs_np[rows].copy())
    want_xs, want_vs, _ = ref.random_search(draws, 64, 4, 0.3, literal=False, rd_std=rd)
    got_xs, got_vs = _np(gx), _np(gv)
    assert np.array_equal(got_xs[rows], want_xs) and np.array_equal(got_vs[rows], want_vs)
    assert np.array_equal(om.cut_values(g, got_xs), got_vs)
    assert (got_vs >= om.cut_values(g, xs_np)).all()


# ------------------------------------------------------------------ packed-tile entry points
@pytest.mark.parametrize("name,envs", [("G22", 4096), ("G14", 250), ("N37", 45)])
def test_local_search_packed_equals_bool_rows(name, envs, cuda_device):
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    edges = random_graph(37, 90, seed=3) if name == "N37" else gset_like(name)
    sim = EnvMaxcut(mygraph=edges, device=cuda_device, if_bidirectional=True)
    th.manual_seed(2)
    x0 = sim.generate_xs_randomly(envs)
    th.manual_seed(9)
    xa, va = sim.local_search_inplace(x0.clone(), th.empty(()))
    state = th.cuda.get_rng_state(cuda_device)
    th.manual_seed(9)
    pk, vb = sim.local_search_packed(sim.store.pack(x0), num_sims=envs)
    assert th.equal(th.cuda.get_rng_state(cuda_device), state)
    assert th.equal(va, vb) and th.equal(sim.store.unpack(pk.contiguous(), envs), xa)
    # second call starting from the packed result with the values handed in (`pk` is a view of the simulator's
    # workspace: copy it before the next local-search call overwrites it)
    pk_copy = pk.clone()
    th.manual_seed(10)
    xa2, va2 = sim.local_search_inplace(xa.clone(), va.clone(), num_iters=3, num_spin=5)
    th.manual_seed(10)
    pk2, vb2 = sim.local_search_packed(pk_copy, 3, 5, num_sims=envs, good_vs=vb)
    assert th.equal(va2, vb2) and th.equal(sim.store.unpack(pk2.contiguous(), envs), xa2)


def test_best_exchange_object_matches_best_allreduce(cuda_device):
    from rlsolver_b200.dist import BestExchange, best_allreduce
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    th.manual_seed(0)
    for envs, edges in ((37, random_graph(100, 320, seed=8)), (4096, gset_like("G22")), (1, random_graph(37, 90, seed=3))):
        sim = EnvMaxcut(mygraph=edges, device=cuda_device, if_bidirectional=False)
        n = sim.num_nodes
        xs = sim.generate_xs_randomly(envs)
        vs = sim.calculate_obj_values(xs) - 500            # negative values too
        if envs > 3:
            vs[3] = vs[envs - 1] = 9000                    # tie -> lowest global env id
        want = best_allreduce(vs, xs, rank=3, world=1, envs_per_rank=envs)
        ex = BestExchange(n, rank=3, world=1, envs_per_rank=envs, device=cuda_device)
        got = ex(vs, xs)
        assert all(th.equal(a, b) for a, b in zip(want, got))
        got = ex.packed(vs, sim.store.pack(xs), sim.store)
        assert all(th.equal(a, b) for a, b in zip(want, got))


# ------------------------------------------------------------------ RNG self-check
def test_rng_self_check_detects_a_foreign_geometry(cuda_device, monkeypatch):
    """rng.self_check passes on this torch build and raises when the modelled call geometry is not torch's."""
    from rlsolver_b200 import rng
    rng._CHECKED.discard(cuda_device.index)
    state = th.cuda.get_rng_state(cuda_device)
    rng.self_check(cuda_device)
    assert cuda_device.index in rng._CHECKED
    assert th.equal(th.cuda.get_rng_state(cuda_device), state)          # the caller's generator is untouched
    rng._CHECKED.discard(cuda_device.index)
    real = rng._max_grid(cuda_device)
    monkeypatch.setitem(rng._MAX_GRID, cuda_device.index, real // 2)   # a torch that sized its grids differently
    with pytest.raises(RuntimeError, match="rlsolver_b200.rng"):
        rng.self_check(cuda_device)
    monkeypatch.setitem(rng._MAX_GRID, cuda_device.index, real)
    rng.self_check(cuda_device)


# ------------------------------------------------------------------ 2 ranks over NCCL
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, envs, out):
    import torch.distributed as dist
    from rlsolver_b200.dist import BestExchange, best_allreduce
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    th.cuda.set_device(rank)
    dev = th.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    edges = gset_like("G22")
    sim = EnvMaxcut(mygraph=edges, device=dev, if_bidirectional=True)
    th.manual_seed(74 + rank)
    xs = sim.generate_xs_randomly(envs)
    xs, vs = sim.local_search_inplace(xs, th.empty(()))
    cut, gid, row = best_allreduce(vs, xs, rank, world, envs)
    ex = BestExchange(sim.num_nodes, rank, world, envs, dev)
    cut2, gid2, row2 = ex(vs, xs)
    cut2, gid2, row2 = int(cut2), int(gid2), row2.clone()      # views of the exchange's buffers: keep the values
    # what bench.py does per step: the local search replayed from a CUDA graph, the exchange eager behind it
    sim.store.rng_cursor_sync()
    side = th.cuda.Stream(device=dev)
    side.wait_stream(th.cuda.current_stream(dev))
    g = th.cuda.CUDAGraph()
    xg = xs.clone()
    with th.cuda.stream(side):
        with th.cuda.graph(g, stream=side):
            gx, gv = sim.local_search_inplace(xg, th.empty(()))
    th.cuda.current_stream(dev).wait_stream(side)
    g.replay()
    best = ex(gv, gx)
    th.cuda.synchronize()
    out[rank] = {"xs": xs.cpu(), "vs": vs.cpu(), "cut": int(cut), "gid": int(gid), "row": row.cpu(),
                 "cut2": cut2, "gid2": gid2, "row2": row2.cpu(),
                 "g_vs": gv.cpu(), "g_xs": gx.cpu(), "g_cut": int(best[0]), "g_gid": int(best[1]), "g_row": best[2].cpu()}
    dist.destroy_process_group()


@pytest.mark.skipif(th.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_shards_equal_single_gpu_runs():
    """Rank r's (xs, vs) equal the single-GPU run of shard r with seed 74 + r (bit-exactness is per shard,
    SURVEY.md 8e) and best_allreduce / BestExchange return the global argmax on both ranks, also behind a CUDA-graph
    replay of the local search."""
    import torch.multiprocessing as mp
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    world, envs = 2, 1024
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_nccl_worker, args=(world, _free_port(), envs, out), nprocs=world, join=True)
        res = [out[r] for r in range(world)]
    dev = th.device("cuda:0")
    sim = EnvMaxcut(mygraph=gset_like("G22"), device=dev, if_bidirectional=True)
    all_vs = []
    for r in range(world):
        th.manual_seed(74 + r)
        xs = sim.generate_xs_randomly(envs)
        xs, vs = sim.local_search_inplace(xs, th.empty(()))
        assert th.equal(xs.cpu(), res[r]["xs"]) and th.equal(vs.cpu(), res[r]["vs"])
        all_vs.append(vs.cpu())
    flat = th.cat(all_vs)
    best = int(flat.max())
    gid = int((flat == best).nonzero()[0])
    want_row = res[gid // envs]["xs"][gid % envs]
    for r in range(world):
        for tag in ("", "2"):
            assert res[r]["cut" + tag] == best and res[r]["gid" + tag] == gid and th.equal(res[r]["row" + tag], want_row)
        # graph replay: a second local search from the first one's result, then the exchange, same on both ranks
        assert res[r]["g_cut"] == res[0]["g_cut"] and res[r]["g_gid"] == res[0]["g_gid"]
        assert th.equal(res[r]["g_row"], res[0]["g_row"])
    g_flat = th.cat([res[r]["g_vs"] for r in range(world)])
    assert res[0]["g_cut"] == int(g_flat.max()) and res[0]["g_gid"] == int((g_flat == g_flat.max()).nonzero()[0])
    owner = res[0]["g_gid"] // envs
    assert th.equal(res[0]["g_row"], res[owner]["g_xs"][res[0]["g_gid"] % envs])


# ------------------------------------------------------------------ the exchange as one kernel over peer memory
def test_peer_exchange_single_rank_equals_best_allreduce(cuda_device):
    """world = 1: the mailbox kernel alone (record, arrival word, pick) against the torch formulation, bool rows and
    packed tiles, negative values and ties (lowest env id wins), many calls (both banks), inside a CUDA graph."""
    from rlsolver_b200.dist import PeerBestExchange, best_allreduce
    from rlsolver_b200.graph_store import GraphStore
    st = GraphStore(gset_like("G14"), True, device=cuda_device)
    n, envs = st.num_nodes, 300
    ex = PeerBestExchange(n, 0, 1, envs, cuda_device)
    gen = th.Generator(device="cpu").manual_seed(5)
    vs_buf = th.zeros((envs,), dtype=th.int64, device=cuda_device)
    xs_buf = th.zeros((envs, n), dtype=th.bool, device=cuda_device)
    graph = None
    for it in range(7):
        vs = th.randint(-50, 50, (envs,), generator=gen).to(cuda_device)
        if it == 3:
            vs[:] = 7                                   # all tied: env 0 wins
        xs = (th.rand((envs, n), generator=gen) < 0.5).to(cuda_device)
        want = best_allreduce(vs, xs, 0, 1, envs)
        got = ex(vs, xs)
        assert int(got[0]) == int(want[0]) and int(got[1]) == int(want[1]) and th.equal(got[2], want[2])
        got = ex.packed(vs, st.pack(xs), st)
        assert int(got[0]) == int(want[0]) and int(got[1]) == int(want[1]) and th.equal(got[2], want[2])
        vs_buf.copy_(vs), xs_buf.copy_(xs)
        if graph is None:
            th.cuda.synchronize()
            graph = th.cuda.CUDAGraph()
            with th.cuda.graph(graph):
                g_out = ex(vs_buf, xs_buf)
        graph.replay()
        assert int(g_out[0]) == int(want[0]) and int(g_out[1]) == int(want[1]) and th.equal(g_out[2], want[2])
    calls, timeouts = ex.status()
    assert calls == 7 * 3 and timeouts == 0
    ex.close()


def _peer_worker(rank, world, port, envs, out):
    import time
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RLSB_PEER_TIMEOUT_MS="8000")
    from rlsolver_b200.dist import BestExchange, PeerBestExchange
    from rlsolver_b200.graph_store import GraphStore
    th.cuda.set_device(rank)
    dev = th.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    st = GraphStore(gset_like("G22"), True, device=dev)
    n = st.num_nodes
    ref = BestExchange(n, rank, world, envs, dev)
    ex = PeerBestExchange(n, rank, world, envs, dev)
    vs_buf = th.zeros((envs,), dtype=th.int64, device=dev)
    xs_buf = th.zeros((envs, n), dtype=th.bool, device=dev)
    graph, rows, ok = None, [], True
    for it in range(12):
        gen = th.Generator(device="cpu").manual_seed(1000 * it + rank)
        vs = th.randint(-1000, 1000, (envs,), generator=gen).to(dev)
        if it == 4:
            vs[:] = 3                                   # every env of every rank tied: global env 0 wins
        xs = (th.rand((envs, n), generator=gen) < 0.5).to(dev)
        if it == 6 and rank == 1:
            th.cuda.synchronize()
            time.sleep(0.5)                             # a late rank: the others poll
        want = tuple(t.clone() for t in ref(vs, xs))
        got = tuple(t.clone() for t in ex(vs, xs))
        got_p = tuple(t.clone() for t in ex.packed(vs, st.pack(xs), st))
        vs_buf.copy_(vs), xs_buf.copy_(xs)
        if graph is None:
            th.cuda.synchronize()
            graph = th.cuda.CUDAGraph()
            with th.cuda.graph(graph):
                g_out = ex(vs_buf, xs_buf)
        graph.replay()
        got_g = tuple(t.clone() for t in g_out)
        for g in (got, got_p, got_g):
            ok = ok and int(g[0]) == int(want[0]) and int(g[1]) == int(want[1]) and bool(th.equal(g[2], want[2]))
        rows.append((int(want[0]), int(want[1])))
    calls, timeouts = ex.status()
    out[rank] = {"ok": ok, "rows": rows, "calls": calls, "timeouts": timeouts}
    ex.close()
    dist.destroy_process_group()


@pytest.mark.skipif(th.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_peer_exchange_equals_all_gather_exchange():
    """PeerBestExchange (one kernel, mailboxes in peer memory) against BestExchange (NCCL all-gather) on the same
    data: bool rows, packed tiles, a captured replay, a tie across ranks, a rank that arrives half a second late."""
    import torch.multiprocessing as mp
    world, envs = 2, 512
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_peer_worker, args=(world, _free_port(), envs, out), nprocs=world, join=True)
        res = [out[r] for r in range(world)]
    for r in range(world):
        assert res[r]["ok"] and res[r]["timeouts"] == 0 and res[r]["calls"] == 12 * 3
        assert res[r]["rows"] == res[0]["rows"]
    assert res[0]["rows"][4] == (3, 0)
    assert len({gid // envs for _, gid in res[0]["rows"]}) == 2        # both ranks won at least once


# ------------------------------------------------------------------ integer-weighted objective (row W)
from conftest import golden_files  # noqa: E402


@pytest.mark.parametrize("path", golden_files("weighted_cut_"), ids=os.path.basename)
def test_weighted_cut_and_fields_match_reference(path, cuda_device):
    """rlsb_cut_eval_weighted / rlsb_node_fields_weighted against obj_maxcut and PISCO's energy / gradient computed by
    the reference (tools/make_goldens_r2.py): +-1 weights (two buckets), small integers (|w| bit buckets)."""
    from rlsolver_b200.graph_store import GraphStore
    z = np.load(path)
    edges = [tuple(int(t) for t in row) for row in z["edges"]]
    n = z["xs"].shape[1]
    st = GraphStore(edges, True, device=cuda_device, num_nodes=n)
    assert st.weighted
    xs = th.from_numpy(z["xs"]).to(cuda_device)
    envs = xs.shape[0]
    assert np.array_equal(_np(st.cut_eval_weighted(xs=xs)), z["cuts"])
    packed = st.pack(xs)
    assert np.array_equal(_np(st.cut_eval_weighted(packed=packed, num_envs=envs)), z["cuts"])
    fields = _np(st.node_fields_weighted(packed, envs))[:, :n].astype(np.int64)
    assert np.array_equal(fields, om.node_fields_weighted(edges, n, z["xs"]))
    # the unweighted entry points keep ignoring weights, like the reference's EnvMaxcut (env_L2A.py:54-66)
    unit = om.build_graph_store(edges, True)
    assert np.array_equal(_np(st.cut_eval(xs)), om.cut_values(unit, z["xs"]))


@pytest.mark.parametrize("values,envs", [((-1, 1), 4096), ((-5, -1, 1, 2, 6), 333), ((1, 1, 1, 3), 64)])
def test_weighted_cut_g22_shape_vs_oracle(values, envs, cuda_device):
    from rlsolver_b200.graph_store import GraphStore
    rng = np.random.default_rng(len(values))
    edges = [(a, b, int(rng.choice(values))) for a, b, _ in gset_like("G22")]
    n = 2000
    st = GraphStore(edges, False, device=cuda_device, num_nodes=n)
    xs_np = _rand_xs(envs, n, 3)
    xs = th.from_numpy(xs_np).to(cuda_device)
    assert np.array_equal(_np(st.cut_eval_weighted(xs=xs)), om.cut_values_weighted(edges, xs_np))
    got = _np(st.node_fields_weighted(st.pack(xs), envs))[:, :n].astype(np.int64)
    assert np.array_equal(got, om.node_fields_weighted(edges, n, xs_np))
    with pytest.raises(ValueError):
        GraphStore(gset_like("G14"), False, device=cuda_device).cut_eval_weighted(xs=th.zeros((2, 800), dtype=th.bool,
                                                                                              device=cuda_device))


# ------------------------------------------------------------------ weighted MCPG sampler (a12, float edge_attr)
@pytest.mark.parametrize("path", golden_files("wmcpg_"), ids=os.path.basename)
def test_weighted_mcpg_sampler_matches_reference(path, cuda_device):
    """mcpg_sampling_maxcut (MCPG/sampling.py:89-127) with the reference's recorded draws replayed: +-1, integer and
    dyadic weights bit-exact; arbitrary float weights within 1e-5 on every chain whose decisions agree."""
    from rlsolver_b200.methods.MCPG import WeightedMcpgData, mcpg_sampling_maxcut
    z = np.load(path)
    n, t = int(z["num_nodes"]), int(z["total_mcmc"])
    data = WeightedMcpgData(z["edges"], z["weights"], n, cuda_device)
    assert np.array_equal(data.sorted_degree_nodes.numpy(), z["order"])
    assert abs(data.edge_weight_sum - float(z["edge_weight_sum"])) < 1e-6
    dev = cuda_device
    explicit = (th.from_numpy(z["metro_idx"]).to(dev), th.from_numpy(z["metro_u"]).to(dev))
    vs_good, xs_good, start, value = mcpg_sampling_maxcut(
        data, th.from_numpy(z["start"]).to(dev), th.from_numpy(z["probs"]).to(dev), int(z["num_ls"]),
        int(z["change_times"]), t, dev, _explicit=explicit, _explicit_u=th.from_numpy(z["ls_u"]).to(dev))
    assert np.array_equal(_np(start), z["metro_out"])
    if "float" in os.path.basename(path):
        same = (_np(xs_good) == z["xs_good"]).all(axis=0)
        assert same.mean() >= 0.8
        np.testing.assert_allclose(_np(vs_good)[same], z["vs_good"][same], rtol=1e-5, atol=1e-5)
        return
    assert np.array_equal(_np(xs_good), z["xs_good"]) and np.array_equal(_np(vs_good), z["vs_good"])
    np.testing.assert_allclose(_np(value), z["value"], rtol=0, atol=1e-4)


def test_weighted_mcpg_sampler_same_seed_vs_oracle(cuda_device):
    """Generator consumed in place: the kernel's decisions from torch's Philox stream equal the oracle fed with the
    draws torch.rand returns for the same seed, and the generator ends where num_ls * N rand(C) calls leave it."""
    from oracle import mcpg as oq
    from rlsolver_b200.methods.MCPG import WeightedMcpgData, mcpg_sampling_maxcut
    rng = np.random.default_rng(4)
    edges = np.asarray([(a, b) for a, b, _ in random_graph(300, 1500, seed=9)], dtype=np.int64)
    weights = rng.choice([-1.0, 1.0, 0.5, -2.0, 3.0], size=len(edges)).astype(np.float32)
    n, t, rep, num_ls, change = 300, 64, 4, 3, 6
    c = t * rep
    data = WeightedMcpgData(edges, weights, n, cuda_device)
    th.manual_seed(77)
    probs = th.rand(n, device=cuda_device) * 0.6 + 0.2
    start = th.randint(0, 2, (n, c), device=cuda_device).float()
    th.manual_seed(78)
    vs_good, xs_good, metro_out, value = mcpg_sampling_maxcut(data, start, probs, num_ls, change, t, cuda_device)
    after = th.cuda.get_rng_state(cuda_device)
    # replay: the metro part again from the same seed (it consumes a data-dependent number of draws), then the
    # num_ls * N uniform draws of the sweeps
    from rlsolver_b200.methods.MCPG import metro_sampling
    th.manual_seed(78)
    again = metro_sampling(probs, start.clone(), change, cuda_device)
    assert th.equal(again, metro_out)
    draws = np.stack([_np(th.rand(c, device=cuda_device)) for _ in range(num_ls * n)])
    assert th.equal(th.cuda.get_rng_state(cuda_device), after)
    want = oq.weighted_sampler(n, edges, weights, data.sorted_degree_nodes.numpy(), _np(metro_out), num_ls, t, draws)
    assert np.array_equal(_np(xs_good), want[1]) and np.array_equal(_np(vs_good), want[0])


# ------------------------------------------------------------------ host-resident batches (HostPipeline)
@pytest.mark.parametrize("layout", ["bool", "packed"])
def test_host_pipeline_equals_eager_calls(layout, cuda_device):
    """Seven host batches through HostPipeline (3 buffers, one graph per buffer, three streams) == seven eager
    local_search calls from the same seed, batch for batch; torch's generator ends where the eager calls leave it."""
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    from rlsolver_b200.host_pipeline import HostPipeline
    sim = EnvMaxcut(mygraph=gset_like("G14"), device=cuda_device, if_bidirectional=True)
    st, envs, k = sim.store, 512, 7
    gen = th.Generator().manual_seed(11)
    batches = [(th.rand((envs, sim.num_nodes), generator=gen) < 0.5) for _ in range(k)]
    th.manual_seed(901)
    want = []
    for xb in batches:
        xs, vs = sim.local_search_inplace(xb.to(cuda_device), th.empty(()))
        want.append((xs.cpu(), vs.cpu()))
    end_state = th.cuda.get_rng_state(cuda_device)

    pipe = HostPipeline(sim, envs, layout=layout)
    th.manual_seed(901)
    pipe.sync_rng()
    if layout == "bool":
        h_in = [xb.pin_memory() for xb in batches]
    else:
        h_in = [st.pack(xb.to(cuda_device)).cpu().pin_memory() for xb in batches]
    h_out = [th.empty_like(t).pin_memory() for t in h_in]
    h_vs = [th.empty((envs,), dtype=th.int64).pin_memory() for _ in range(k)]
    for i in range(k):
        pipe.submit(h_in[i], h_out[i], h_vs[i])
    pipe.drain()
    pipe.commit_rng()
    for i in range(k):
        got_xs = h_out[i] if layout == "bool" else st.unpack(h_out[i].to(cuda_device), envs).cpu()
        assert th.equal(got_xs, want[i][0]), f"batch {i}"
        assert th.equal(h_vs[i], want[i][1]), f"batch {i}"
    assert th.equal(th.cuda.get_rng_state(cuda_device), end_state)


# ------------------------------------------------------------------ evolutionary_replacement (row a8)
def test_evolutionary_replacement_equals_reference_indexing(cuda_device):
    """rlsb_copy_rows behind the mirror against the reference's two fancy-index assignments (util.py:87-94) from the
    same generator state, G22 x 4096 with many tied values.  (Minimising, the reference indexes out of range -- on CUDA a
    device-side assert; the mirror raises IndexError on the host instead.)"""
    from rlsolver_b200.methods.util import evolutionary_replacement
    e, n, low_k = 4096, 2000, 512
    g = th.Generator(device=cuda_device).manual_seed(3)
    xs = th.rand((e, n), device=cuda_device, generator=g) < 0.5
    vs = th.randint(13000, 13040, (e,), device=cuda_device, generator=g)
    want_xs, want_vs = xs.clone(), vs.clone()
    th.manual_seed(77)
    ids = want_vs.argsort()
    top_ids, low_ids = ids[:-low_k], ids[-low_k:]
    replace_ids = top_ids[th.randperm(e - low_k, device=cuda_device)[:low_k]]
    want_xs[replace_ids] = want_xs[low_ids]
    want_vs[replace_ids] = want_vs[low_ids]
    end = th.cuda.get_rng_state(cuda_device)
    th.manual_seed(77)
    evolutionary_replacement(xs, vs, low_k, True)
    assert th.equal(xs, want_xs) and th.equal(vs, want_vs)
    assert th.equal(th.cuda.get_rng_state(cuda_device), end)
    with pytest.raises(IndexError):
        evolutionary_replacement(xs, vs, low_k, False)


# ------------------------------------------------------------------ row-major Metropolis (a11, TNCO variant)
@pytest.mark.parametrize("sims,dim,repeats,num_iters,sharp", [(6, 40, 3, -1, False), (5, 33, 2, 6, True), (3, 20, 4, 1, False),
                                                              (64, 2000, 8, -1, False), (16, 1500, 4, 40, True),
                                                              (7, 3000, 5, 0, False)])
def test_row_major_metropolis_same_seed_same_samples(sims, dim, repeats, num_iters, sharp, cuda_device):
    """metropolis_hastings_sampling_TNCO on the kernels == the torch restatement of env_L2A.py:233-276 on the same
    device and seed: samples and the generator state afterwards (1 to 4 rounds, early stop inside a round, a stop
    target of zero = exactly one column visited)."""
    from oracle import torch_port as tp
    from rlsolver_b200.envs.env_L2A import metropolis_hastings_sampling_TNCO
    g = th.Generator(device=cuda_device).manual_seed(5)
    probs = th.rand((sims, dim), device=cuda_device, generator=g)
    if sharp:
        probs = th.where(probs < 0.5, probs * 0.04 + 0.005, 1 - probs * 0.04)
    start = th.rand((sims, dim), device=cuda_device, generator=g) < probs
    start0 = start.clone()
    th.manual_seed(99)
    want = tp.metropolis_hastings_sampling_tnco(probs, start, repeats, num_iters)
    end = th.cuda.get_rng_state(cuda_device)
    th.manual_seed(99)
    got = metropolis_hastings_sampling_TNCO(probs=probs, start_xs=start, num_repeats=repeats, num_iters=num_iters)
    assert th.equal(got, want)
    assert th.equal(th.cuda.get_rng_state(cuda_device), end)
    assert th.equal(start, start0)            # the start rows are not modified


# ------------------------------------------------------------------ threshold-only pass: warp per env == tile pipeline
@pytest.mark.parametrize("name,envs,spin,bidir", [("G22", 4096, 8, True), ("G70", 2048 + 17, 4, False), ("G14", 300, 8, True)])
def test_threshold_rows_kernel_equals_pipelined_kernel(name, envs, spin, bidir, cuda_device):
    """rlsb_ls_run(thresh only): the warp-per-env kernel over the row-major counts and the pipelined tile kernel
    (RLSB_DEBUG_THRESH_PIPE) produce the same float32 thresholds, bit for bit, and they equal torch.kthvalue of the
    reference's expression evaluated with torch ops on the same device."""
    from rlsolver_b200 import _lib
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    sim = EnvMaxcut(mygraph=gset_like(name), device=cuda_device, if_bidirectional=bidir)
    st, n = sim.store, sim.num_nodes
    th.manual_seed(5)
    xs = sim.generate_xs_randomly(envs)
    ws = st.ls_workspace(envs)
    vs = st.ls_begin(xs, None, 1, 0.3, ws)
    noise = th.randn((envs, n), device=cuda_device)
    # every second row ascending: all of its later values beat the bound taken from its first 256 (the kernel's
    # parking columns overflow and the row is redone with every value inserted)
    noise[::2] = th.arange(n, device=cuda_device, dtype=th.float32)[None, :] * 0.01 + noise[::2] * 1e-3
    got = []
    for flag in (0, _lib.DEBUG_THRESH_PIPE):
        _lib.debug_flags(flag, _lib.DEBUG_THRESH_PIPE ^ flag)
        st.ls_section(ws, envs, "thresh").fill_(float("nan"))
        st.ls_run(vs, 1, noise, spin, [], False, None, ws)
        got.append(st.ls_section(ws, envs, "thresh").clone())
    _lib.debug_flags(0, _lib.DEBUG_THRESH_PIPE)
    assert th.equal(got[0], got[1]) and not bool(th.isnan(got[0]).any())
    # the reference's expression (env_L2A.py:90-96) in torch on the same tensors
    vs_raw = sim.calculate_obj_values_for_loop(xs, if_sum=False)
    wsr = sim.n0_num_n1 - (2 if bidir else 1) * vs_raw
    ws_std = wsr.max(dim=0, keepdim=True)[0] - wsr.min(dim=0, keepdim=True)[0]
    spin_rand = wsr + noise * (ws_std.float() * 0.3)
    want = th.kthvalue(spin_rand, k=n - spin, dim=1)[0]
    assert th.equal(got[0], want)
