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
    from rlsolver_b200.methods.ISCO import util as isco_util
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
    monkeypatch.setattr(isco_util.th, "rand", replay)
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
