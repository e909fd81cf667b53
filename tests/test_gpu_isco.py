"""ISCO_maxcut / PISCO_maxcut mirrors (rlsolver_b200/envs/env_ISCO.py) against trajectories of the
UNMODIFIED reference (tools/make_goldens_isco.py, CPU) with its uniform draws replayed."""
import glob
import os

import numpy as np
import pytest
import torch as th

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Replay:
    """Stands in for torch.rand: hands out the recorded draws in order."""

    def __init__(self, draws, sizes, device):
        self.draws, self.sizes, self.device, self.at, self.k = draws, sizes, device, 0, 0

    def __call__(self, *shape, **kw):
        shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        n = int(np.prod(shape)) if shape else 1
        assert n == int(self.sizes[self.k]), f"draw {self.k}: asked {n}, recorded {self.sizes[self.k]}"
        out = th.from_numpy(self.draws[self.at:self.at + n].copy()).reshape(shape).to(self.device)
        self.at += n
        self.k += 1
        return out


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*isco_*.npz"))), ids=os.path.basename)
def test_isco_step_matches_reference(path, cuda_device, monkeypatch):
    from rlsolver_b200.envs import env_ISCO
    from rlsolver_b200.methods.ISCO import config_maxcut as cfg
    z = np.load(path)
    pisco = os.path.basename(path).startswith("pisco")
    n = int(z["num_nodes"])
    batch = z["xs"].shape[1]
    monkeypatch.setattr(cfg, "BATCH_SIZE", batch)
    monkeypatch.setattr(cfg, "DEVICE", cuda_device)
    ef, et = th.from_numpy(z["edge_from"]).to(cuda_device), th.from_numpy(z["edge_to"]).to(cuda_device)
    params = {"num_nodes": n, "num_edges": len(z["edge_from"]), "edge_from": ef, "edge_to": et}
    if pisco:
        npad = (n + 7) // 8 * 8
        A = th.zeros((npad, npad), dtype=th.float16, device=cuda_device)
        w = th.from_numpy(z["edge_w"]).to(cuda_device).to(th.float16) if "edge_w" in z.files else 1   # weighted goldens
        A[ef, et] = w
        A[et, ef] = w
        params["adj_matrix"] = A
        sampler = env_ISCO.PISCO_maxcut(params)
    else:
        sampler = env_ISCO.ISCO_maxcut(params)
    replay = Replay(z["draws"], z["draw_sizes"], cuda_device)
    monkeypatch.setattr(env_ISCO.th, "rand", replay)
    x = th.from_numpy(z["xs"][0]).to(cuda_device)
    x = x.to(th.float16) if pisco else x
    steps = z["energies"].shape[0]
    for k in range(steps):
        path_length = th.from_numpy(z["paths"][k]).to(cuda_device)
        temperature = th.tensor(float(z["temps"][k]), dtype=th.float32, device=cuda_device)
        # the integer core is exact: the energy of the incoming state is the cut count (ISCO)
        if not pisco:
            cut = sampler.model(x, th.tensor(1.0, device=cuda_device)).cpu().numpy()
            bits = z["xs"][k][:, :n] != 0
            want = (bits[:, z["edge_from"]] ^ bits[:, z["edge_to"]]).sum(axis=1)
            assert np.array_equal(cut, want.astype(np.float32))
        x, energy, acc = sampler.step(x, path_length, temperature)
        # same sites chosen, same accept decisions -> identical next state; floats to 1e-5
        assert np.array_equal(x.float().cpu().numpy(), z["xs"][k + 1]), f"step {k}: state differs"
        np.testing.assert_allclose(energy.float().cpu().numpy(), z["energies"][k], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(acc.float().cpu().numpy(), z["accs"][k], rtol=2e-4, atol=1e-30)
    assert replay.k == len(z["draw_sizes"])


@pytest.mark.parametrize("kind,nodes,edges_n,batch", [("isco", 2000, 19990, 64), ("pisco", 800, 4694, 33), ("pisco_w", 300, 1500, 17),
                                                      ("isco", 37, 90, 5)])
def test_step_kernels_vs_torch_restatement(kind, nodes, edges_n, batch, cuda_device, monkeypatch):
    """csrc/isco.cu (rlsb_isco_propose / rlsb_isco_accept) against the torch restatement of the reference's helper
    functions (oracle/isco.py: two sorts, gather / scatter / cumsum) at Gset-like sizes, fed the same uniform draws:
    same chosen sites, same accepts -> same next states; log-probabilities to 1e-5."""
    from oracle import isco as oi
    from synth import random_graph
    from rlsolver_b200.envs import env_ISCO
    from rlsolver_b200.methods.ISCO import config_maxcut as cfg
    dev = cuda_device
    monkeypatch.setattr(cfg, "BATCH_SIZE", batch)
    monkeypatch.setattr(cfg, "DEVICE", dev)
    edges = random_graph(nodes, edges_n, seed=12)
    rng = np.random.default_rng(1)
    ef = th.tensor([a for a, _, _ in edges], device=dev)
    et = th.tensor([b for _, b, _ in edges], device=dev)
    params = {"num_nodes": nodes, "num_edges": len(edges), "edge_from": ef, "edge_to": et}
    if kind.startswith("pisco"):
        npad = (nodes + 7) // 8 * 8
        A = th.zeros((npad, npad), dtype=th.float16, device=dev)
        w = th.from_numpy(rng.choice([-2, -1, 1, 3], size=len(edges))).to(dev).to(th.float16) if kind == "pisco_w" else 1
        A[ef, et] = w
        A[et, ef] = w
        params["adj_matrix"] = A
        sampler = env_ISCO.PISCO_maxcut(params)
        x = sampler.random_gen_init_sample()
        x = th.nn.functional.pad(x, (0, npad - nodes))
    else:
        sampler = env_ISCO.ISCO_maxcut(params)
        x = sampler.random_gen_init_sample()
    th.manual_seed(3)
    for step in range(6):
        temperature = th.tensor(1.0 - 0.15 * step, device=dev)
        path_length = th.randint(1, 9, (batch,), device=dev)
        u1 = th.rand(x.shape, device=dev)
        u2 = th.rand((batch,), device=dev)
        want = oi.step(lambda s: sampler.get_local_dist(s, temperature), x, path_length, u1, u2)
        draws = iter([u1, u2])
        monkeypatch.setattr(env_ISCO.th, "rand", lambda *a, **k: next(draws))
        got_x, got_e, got_acc = sampler.step(x, path_length, temperature)
        monkeypatch.undo()
        monkeypatch.setattr(cfg, "BATCH_SIZE", batch)
        monkeypatch.setattr(cfg, "DEVICE", dev)
        assert th.equal(got_x.float(), want[0].float()), step
        np.testing.assert_allclose(got_e.cpu().numpy(), (want[1] * temperature).cpu().numpy(), rtol=1e-5, atol=1e-5)
        # log_acc is a difference of energies of several hundred: float32 leaves ~1e-4 absolute in the log domain
        want_log = want[2].cpu().numpy()
        seen = want_log > -80.0                        # below that exp() underflows
        np.testing.assert_allclose(np.log(got_acc.cpu().numpy()[seen]), want_log[seen], rtol=0, atol=2e-3)
        x = got_x
