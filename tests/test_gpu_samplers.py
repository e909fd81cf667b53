"""CUDA samplers (csrc/samplers.cu through the MCPG / L2A mirrors) against
  (1) the reference-generated fixtures, replaying the recorded draws,
  (2) torch's own generator: the in-kernel Philox stream must equal torch.rand / torch.randint,
  (3) the reference algorithm (oracle) fed with draws taken from torch's generator at the same seed:
      "same seeds, same flip sequence" for the fused kernels."""
import ctypes as C
import os

import numpy as np
import pytest
import torch as th

from conftest import golden_files
from oracle import mcpg as oq
from synth import gset_like

pytestmark = pytest.mark.gpu


def _edges(z):
    return [tuple(int(t) for t in row) for row in z["edges"]]


def _np(t):
    return t.detach().cpu().numpy()


def _data(edges, n, dev):
    from rlsolver_b200.methods.MCPG import McpgData
    return McpgData(edges, n, dev)


# ------------------------------------------------------------------ (2) the Philox stream is torch's
@pytest.mark.parametrize("numel", [1, 255, 4096, 303104, 303105, 1300000, 5000001])
def test_philox_stream_equals_torch(numel, cuda_device):
    from rlsolver_b200 import _lib, rng
    lib = _lib.lib()
    calls = 3
    th.manual_seed(1234)
    th.rand(17, device=cuda_device)                              # move the offset off zero
    seed, offset, threads, iters = rng.peek(cuda_device, numel)
    want_u = th.stack([th.rand(numel, device=cuda_device) for _ in range(calls)])
    after = rng.generator(cuda_device).get_offset()
    got_u = th.empty((calls, numel), dtype=th.float32, device=cuda_device)
    _lib.check(lib.rlsb_torch_rand(seed, offset, threads, iters, calls, numel, C.c_void_p(got_u.data_ptr()), None))
    assert th.equal(got_u, want_u)
    assert after == offset + 4 * iters * calls                  # rng.advance bookkeeping
    # randint shares the stream layout
    seed, offset, threads, iters = rng.peek(cuda_device, numel)
    want_i = th.stack([th.randint(0, 2000, (numel,), device=cuda_device) for _ in range(calls)])
    got_i = th.empty((calls, numel), dtype=th.int64, device=cuda_device)
    _lib.check(lib.rlsb_torch_randint(seed, offset, threads, iters, calls, numel, 2000,
                                      C.c_void_p(got_i.data_ptr()), None))
    assert th.equal(got_i, want_i)
    rng.advance(cuda_device, numel, 0)
    before = rng.generator(cuda_device).get_offset()
    rng.advance(cuda_device, numel, 5)
    th.manual_seed(1234)
    assert rng.generator(cuda_device).get_offset() == 0 and before > 0


# ------------------------------------------------------------------ (1) reference goldens, replayed draws
@pytest.mark.parametrize("path", golden_files("mcpg_"), ids=os.path.basename)
def test_mcpg_golden_replay(path, cuda_device):
    from rlsolver_b200.methods.MCPG import metro_sampling, sampler_func
    z = np.load(path)
    n, t, r = int(z["num_nodes"]), int(z["total_mcmc"]), int(z["repeat"])
    data = _data(_edges(z), n, cuda_device)
    assert np.array_equal(data.sorted_degree_nodes.numpy(), z["order"])
    dev = cuda_device
    # pad the recorded draws to the full 5*max_transfer iterations: the kernel must stop where the reference did
    tmax = 5 * int(z["max_transfer"])
    idx = np.zeros((tmax, t * r), np.int64)
    u = np.zeros((tmax, t * r), np.float32)                      # u = 0 would accept everything if ever read
    idx[:z["metro_idx"].shape[0]], u[:z["metro_u"].shape[0]] = z["metro_idx"], z["metro_u"]
    xs_sample = metro_sampling(th.from_numpy(z["probs"]).to(dev), th.from_numpy(z["start"]).to(dev),
                               int(z["max_transfer"]), device=dev,
                               _explicit=(th.from_numpy(idx), th.from_numpy(u)))
    assert np.array_equal(_np(xs_sample), z["xs_sample"])
    vs_good, xs_good, value = sampler_func(data, th.from_numpy(z["xs_sample"]).to(dev), int(z["num_ls"]), t, r,
                                           _explicit_u=th.from_numpy(z["ls_u"]))
    assert np.array_equal(_np(vs_good), z["vs_good"]) and np.array_equal(_np(xs_good), z["xs_good"])
    assert np.allclose(_np(value), z["value"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("path", golden_files("subset_"), ids=os.path.basename)
def test_subset_golden_replay(path, cuda_device):
    from rlsolver_b200.methods.L2A.transformer import sub_set_sampling
    z = np.load(path)
    dev = cuda_device
    xs, probs = sub_set_sampling(th.from_numpy(z["probs"]).to(dev), th.from_numpy(z["start"]).to(dev), int(z["repeats"]),
                                 int(z["top_k"]), _explicit_u=th.from_numpy(z["u"]) if z["u"].shape[0] else None)
    assert np.array_equal(_np(xs), z["xs"]) and np.array_equal(_np(probs), z["probs_out"])


# ------------------------------------------------------------------ (3) same seed as the reference's call sequence
def _draws(fn, count, dev):
    return np.stack([_np(fn()) for _ in range(count)]) if count else None


# the last case has more than 4 * 148 tiles of 32 chains: sampler_func then computes its tie-breaks inside the sweep
# kernel instead of taking them from the pre-pass
@pytest.mark.parametrize("name,total_mcmc,repeat,num_ls,max_transfer", [("G14", 64, 8, 2, 20), ("G22", 512, 8, 1, 40),
                                                                          ("G14", 2400, 8, 2, 5)])
def test_mcpg_same_seed_as_reference_calls(name, total_mcmc, repeat, num_ls, max_transfer, cuda_device):
    """Run the fused kernels from a seed; then rewind the generator, draw what the reference's torch
    calls would have drawn (randint/rand per Metropolis iteration, rand(C) per node visit) and feed
    them to the reference algorithm (oracle): states and generator offsets must coincide."""
    from rlsolver_b200 import rng
    from rlsolver_b200.methods.MCPG import metro_sampling, sampler_func
    dev = cuda_device
    edges = gset_like(name)
    n = max(max(a, b) for a, b, _ in edges) + 1
    data = _data(edges, n, dev)
    c = total_mcmc * repeat
    th.manual_seed(5)
    probs = (th.rand(n, device=dev) * 0.6 + 0.2)
    start = th.randint(0, 2, (n, c), device=dev).float()
    gen = rng.generator(dev)
    off0 = gen.get_offset()
    xs_sample = metro_sampling(probs, start, max_transfer, device=dev)
    off1 = gen.get_offset()
    vs_good, xs_good, value = sampler_func(data, xs_sample, num_ls, total_mcmc, repeat)
    off2 = gen.get_offset()
    # --- the reference's call sequence from the same generator state
    gen.set_offset(off0)
    idx, u = [], []
    for _ in range(5 * max_transfer):
        idx.append(_np(th.randint(low=0, high=n, size=[c], device=dev)))
        u.append(_np(th.rand(c, device=dev)))
    want_sample, iters = oq.metro_sampling(_np(probs), _np(start), max_transfer, np.stack(idx), np.stack(u))
    assert np.array_equal(_np(xs_sample), want_sample)
    gen.set_offset(off0)
    for _ in range(iters):
        th.randint(low=0, high=n, size=[c], device=dev), th.rand(c, device=dev)
    assert gen.get_offset() == off1
    ls_u = np.stack([_np(th.rand(c, device=dev)) for _ in range(num_ls * n)])
    assert gen.get_offset() == off2
    w_vs, w_xs, w_value, _, _ = oq.sampler_func(n, edges, data.sorted_degree_nodes.numpy(), want_sample, num_ls,
                                                total_mcmc, repeat, ls_u)
    assert np.array_equal(_np(vs_good), w_vs) and np.array_equal(_np(xs_good), w_xs)
    assert np.allclose(_np(value), w_value, rtol=0, atol=1e-3)
    # cut values are consistent with the states
    from oracle import maxcut as om
    g = om.build_graph_store(edges, False)
    assert np.array_equal(om.cut_values(g, _np(xs_good).T > 0), _np(vs_good).astype(np.int64))


def test_subset_same_seed_as_reference_calls(cuda_device):
    from rlsolver_b200 import rng
    from rlsolver_b200.methods.L2A.transformer import sub_set_sampling
    dev = cuda_device
    s, n, repeats, top_k = 64, 2000, 64, 500                      # the dREINFORCE shapes of demo_instance.py
    th.manual_seed(11)
    start = th.randint(0, 2, (s, n), dtype=th.bool, device=dev)
    probs = th.rand((s, n), dtype=th.float32, device=dev)
    gen = rng.generator(dev)
    off0 = gen.get_offset()
    xs, _ = sub_set_sampling(probs, start, repeats, top_k)
    off1 = gen.get_offset()
    gen.set_offset(off0)
    det = th.abs(probs - 0.5)
    top_values, top_ids = th.topk(det, k=top_k, largest=False, dim=1)
    u = np.stack([_np(th.rand_like(top_values[:, 0].repeat(repeats))) for _ in range(top_k)])
    assert gen.get_offset() == off1
    want = oq.sub_set_sampling(_np(top_ids), _np(top_values), _np(start), repeats, u)
    assert np.array_equal(_np(xs), want)


@pytest.mark.parametrize("n,c,max_transfer,lo,hi", [(2000, 4096, 200, 0.2, 0.8), (37, 33, 6, 0.05, 0.5),
                                                    (800, 1000, 40, 0.4, 0.6), (100, 70, 3, 0.45, 0.55)])
def test_metro_split_equals_two_pass(n, c, max_transfer, lo, hi, cuda_device, monkeypatch):
    """The split form of metro_sampling (draws in parallel, one pass of the chain, surplus moves undone) against
    the two-pass kernel from the same generator state: same samples, same number of executed iterations (the
    generator ends at the same offset) -- with early stops (probabilities near 0.5 accept almost every move) and
    runs that use all 5 * max_transfer iterations."""
    import rlsolver_b200.methods.MCPG as M
    from rlsolver_b200 import rng
    dev = cuda_device
    th.manual_seed(11)
    probs = th.rand(n, device=dev) * (hi - lo) + lo
    start = th.randint(0, 2, (n, c), device=dev).float()
    gen = rng.generator(dev)
    off0 = gen.get_offset()
    got = M.metro_sampling(probs, start, max_transfer, device=dev)
    off_split = gen.get_offset()
    gen.set_offset(off0)
    monkeypatch.setattr(M, "_METRO_SPLIT_MAX_BYTES", -1)
    want = M.metro_sampling(probs, start, max_transfer, device=dev)
    assert gen.get_offset() == off_split
    assert th.equal(got, want)
    assert not th.equal(got, start)
