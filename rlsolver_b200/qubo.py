"""Dense QUBO Hamiltonian x^T Q x for batches of chains on the B200 tensor cores (csrc/qubo.cu).

`QuboModel(Q)` holds the device-side bf16 limb split of a float32 `Q [N, N]`; `energy(X)` evaluates
`sum(X * (Q @ X), dim=0)` for `X [N, C]` -- the "compute value" lines of mcpg_sampling_qubo /
mcpg_sampling_qubo_bin (rlsolver/methods/MCPG/sampling.py:339-340, 364-365).  `qubo_values` is the
tail of those two functions (per-chain best over the repeats and the advantage)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch as th

from . import _lib
from .graph_store import _ptr, _stream_ptr, on_device, require_cuda

TEN = th.Tensor


class QuboModel:
    def __init__(self, Q: TEN):
        if Q.dim() != 2 or Q.shape[0] != Q.shape[1]:
            raise ValueError(f"Q must be square, got {tuple(Q.shape)}")
        self.device = require_cuda(Q.device)
        self.Q = Q.to(th.float32).contiguous()
        self.nvar = int(Q.shape[0])
        self._lib = _lib.lib()
        handle = C.c_void_p()
        with on_device(self.device):
            _lib.check(self._lib.rlsb_qubo_create(_ptr(self.Q), self.nvar, self.device.index, C.byref(handle),
                                                  _stream_ptr(self.device)), "qubo_create")
        self._h = handle
        self._ws = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.rlsb_qubo_destroy(h)
            except Exception:
                pass

    def energy(self, X: TEN) -> TEN:
        """X: float32 [N, C] with entries in {-1, 0, +1} (exact in bf16).  Returns float32 [C]."""
        if X.dim() != 2 or X.shape[0] != self.nvar:
            raise IndexError(f"X must be [{self.nvar}, C], got {tuple(X.shape)}")
        if X.device != self.device:
            raise RuntimeError(f"X is on {X.device}, the model is on {self.device}")
        X = X.to(th.float32).contiguous()
        c = int(X.shape[1])
        out = th.empty((c,), dtype=th.float32, device=self.device)
        need = int(self._lib.rlsb_qubo_workspace_bytes(self._h, c))
        if self._ws is None or self._ws.numel() < need:
            self._ws = th.empty((need,), dtype=th.uint8, device=self.device)
        with on_device(self.device):
            _lib.check(self._lib.rlsb_qubo_energy(self._h, _ptr(X), c, _ptr(out), _ptr(self._ws),
                                                  _stream_ptr(self.device)), "qubo_energy")
        return out

    def sweeps(self, X: TEN, num_sweeps: int, binary: bool = False) -> TEN:
        """The local-search sweeps of mcpg_sampling_qubo(_bin) (sampling.py:331-337 / 356-362), in place on
        X float32 [N, C] (contiguous): num_sweeps Gauss-Seidel passes x_i <- rule(Q_i . x with x_i zeroed)."""
        if X.dim() != 2 or X.shape[0] != self.nvar:
            raise IndexError(f"X must be [{self.nvar}, C], got {tuple(X.shape)}")
        if X.device != self.device or X.dtype != th.float32 or not X.is_contiguous():
            raise TypeError(f"X must be a contiguous float32 tensor on {self.device} (it is updated in place)")
        c = int(X.shape[1])
        need = int(self._lib.rlsb_qubo_workspace_bytes(self._h, c))
        if self._ws is None or self._ws.numel() < need:
            self._ws = th.empty((need,), dtype=th.uint8, device=self.device)
        with on_device(self.device):
            _lib.check(self._lib.rlsb_qubo_sweeps(self._h, _ptr(self.Q), _ptr(X), c, int(num_sweeps), int(bool(binary)),
                                                  _ptr(self._ws), _stream_ptr(self.device)), "qubo_sweeps")
        return X


def _sampling(model: QuboModel, start_result: TEN, probs: TEN, num_ls: int, change_times: int,
              total_mcmc_num: int, binary: bool):
    from .methods.MCPG import metro_sampling
    raw_samples = metro_sampling(probs, start_result.clone(), change_times, model.device)
    samples = raw_samples.clone() if binary else raw_samples * 2 - 1
    model.sweeps(samples, num_ls, binary=binary)
    max_res, index, value = qubo_values(model, samples, total_mcmc_num)
    best = samples[:, index] if binary else (samples[:, index] + 1) / 2
    return max_res, best, raw_samples, value


def mcpg_sampling_qubo(data, start_result: TEN, probs: TEN, num_ls: int, change_times: int, total_mcmc_num: int,
                       device=None):
    """rlsolver/methods/MCPG/sampling.py:323-346 on the tensor-core kernels.  data = {'Q', 'nvar'} or a
    QuboModel (reuse it across calls: the bf16 limb split of Q is done once)."""
    model = data if isinstance(data, QuboModel) else QuboModel(data['Q'].to(device or start_result.device))
    return _sampling(model, start_result, probs, num_ls, change_times, total_mcmc_num, False)


def mcpg_sampling_qubo_bin(data, start_result: TEN, probs: TEN, num_ls: int, change_times: int, total_mcmc_num: int,
                           device=None):
    """rlsolver/methods/MCPG/sampling.py:349-370 (x in {0, 1}, threshold -Q_ii / 2)."""
    model = data if isinstance(data, QuboModel) else QuboModel(data['Q'].to(device or start_result.device))
    return _sampling(model, start_result, probs, num_ls, change_times, total_mcmc_num, True)


def qubo_values(model: QuboModel, samples: TEN, total_mcmc_num: int) -> Tuple[TEN, TEN, TEN]:
    """sampling.py:339-346: (max_res [T], index [T] of the best repeat per chain, advantage [C])."""
    res_sample = model.energy(samples)
    index = th.argmax(res_sample.reshape((-1, total_mcmc_num)), dim=0)
    index = th.arange(total_mcmc_num, device=samples.device) + index * total_mcmc_num
    return res_sample[index], index, -(res_sample - th.mean(res_sample.float()))
