"""Drop-ins for `rlsolver.envs.env_ISCO.ISCO_maxcut` (env_ISCO.py:10-91) and `PISCO_maxcut`
(env_ISCO.py:365-448): one Metropolis-Hastings step of the path-auxiliary discrete sampler on
B chains.

What the reference computes per step, twice (for x and for the proposal y):
  * ISCO : energy = #cut edges / T by three [B, M] gathers, its gradient by autograd + vmap;
  * PISCO: energy = -1/4 s^T A s / T by a dense fp16 matmul against the padded adjacency, gradient
           by autograd.
Both gradients are integers in disguise: with d = 2x-1, (1-2x_i) * grad_i / 2 = d_i * sum_j A_ij d_j
/ (2T) = (same-side neighbours - other-side neighbours) / (2T).  Here the integer part --
cut[b] and cross[b][i] for all chains and nodes -- comes from the bit-packed tile kernels behind
the C ABI (rlsb_pack_spins, rlsb_node_cross_counts, rlsb_cut_eval_packed); the float tail (one
division by T, log-softmax, the Gumbel top-k path proposal, the without-replacement renormalisation
and the MH accept: 2 sorts, an argsort and ~15 elementwise launches per step in the reference,
rlsolver/methods/ISCO/util.py:3-75) is two kernels with one CTA per chain (csrc/isco.cu:
rlsb_isco_propose, rlsb_isco_accept).  The uniform draws are the reference's two torch.rand calls,
made here in the same order and shape, so a seeded or replayed stream reproduces its choices.

Numerics: PISCO's energy and flip gains reproduce the reference's fp16/fp32 roundings exactly
(small integers, one division); ISCO's autograd accumulates +-0.5/T per edge with atomics, so its
gains carry a few ulp of order-dependent noise in the reference itself -- parity there is to
1e-6 relative (tests/test_gpu_isco.py).  ISCO_maxcut takes the edge list and ignores weights like the reference;
PISCO_maxcut's adjacency may carry integer weights (rlsb_cut_eval_weighted / rlsb_node_fields_weighted)."""
from __future__ import annotations

import torch as th

from .. import _lib
from ..graph_store import GraphStore, _ptr, _stream_ptr, on_device, require_cuda
from ..methods.ISCO import config_maxcut as cfg

TEN = th.Tensor


class _PathAuxMaxcut:
    """Shared step logic; subclasses define how the integer fields become (energy, flip gain)."""

    def __init__(self, params_dict):
        self.batch_size = cfg.BATCH_SIZE
        self.device = require_cuda(params_dict.get('device', cfg.DEVICE))
        self.chain_length = cfg.CHAIN_LENGTH
        self.init_temperature = th.tensor(cfg.INIT_TEMPERATURE, device=self.device)
        self.final_temperature = th.tensor(cfg.FINAL_TEMPERATURE, device=self.device)
        self.max_num_nodes = params_dict['num_nodes']
        self.num_edges = params_dict['num_edges']
        self.edge_from = params_dict['edge_from']
        self.edge_to = params_dict['edge_to']
        ef, et = self.edge_from.tolist(), self.edge_to.tolist()
        # bidirectional store: the listed neighbours of a node are all its neighbours
        self.store = GraphStore([(a, b, 1) for a, b in zip(ef, et)], True, device=self.device,
                                num_nodes=self.max_num_nodes)
        self._deg = th.from_numpy(self.store.listed_degree_numpy()).to(self.device)[None, :].long()

    # ---- integer core on the tile kernels
    def _int_fields(self, sample: TEN):
        """sample [B, >=N] with entries {0,1} -> (cut int64 [B], d_i * sum_j A_ij d_j int64 [B, N])."""
        n = self.max_num_nodes
        bits = (sample[:, :n] != 0).contiguous()
        b = bits.shape[0]
        packed = self.store.pack(bits)
        cross, _, _ = self.store.cross_counts(packed, b, want_minmax=False)
        cut = self.store.cut_eval_packed(packed, b)
        cross = cross[:, :n].to(th.long) & 0xFFFF
        return cut, self._deg - 2 * cross

    # ---- one MH step on the kernels (env_ISCO.py:27-77; util.py:3-75)
    _pisco = False

    def _raw_fields(self, sample: TEN):
        """(cut int64 [B], per-site cross counts as the kernels take them, weighted flag, row stride)."""
        n = self.max_num_nodes
        bits = (sample[:, :n] != 0).contiguous()
        b = bits.shape[0]
        packed = self.store.pack(bits)
        cross, _, _ = self.store.cross_counts(packed, b, want_minmax=False)
        return self.store.cut_eval_packed(packed, b), cross, 0, self.store.padded_nodes

    def _deg32(self) -> TEN:
        d = getattr(self, "_deg_i32", None)
        if d is None:
            d = self._deg_i32 = self._deg.reshape(-1).to(th.int32).contiguous()
        return d

    def _temperature(self, temperature) -> TEN:
        t = temperature if isinstance(temperature, th.Tensor) else th.tensor(float(temperature))
        return t.to(device=self.device, dtype=th.float32).reshape(1).contiguous()

    def proposal(self, x, path_length, temperature):
        """get_local_dist(x) + multinomial + the flips.  Returns (ll_x [B], y, trajectory) with trajectory =
        {'ll_x2y' [B], 'selected_idx': {'sites' int32 [B, kmax]: the chosen sites in order, -1 padded}}."""
        x = x.contiguous()
        b, ld = x.shape
        lib = _lib.lib()
        cut, cross, weighted, cross_ld = self._raw_fields(x)
        t = self._temperature(temperature)
        pl = path_length.to(device=self.device, dtype=th.int64).contiguous()
        u = th.rand((b, ld), device=self.device)                   # the draw of gumbel() (util.py:4)
        kmax = int(min(ld, max(1, int(pl.max().item())))) if b else 1
        sel = th.empty((b, kmax), dtype=th.int32, device=self.device)
        y = th.empty_like(x)
        ll_x = th.empty((b,), dtype=th.float32, device=self.device)
        ll_x2y = th.empty_like(ll_x)
        with on_device(self.device):
            _lib.check(lib.rlsb_isco_propose(_ptr(x), _ptr(y), int(x.dtype == th.float16), _ptr(cross), weighted, cross_ld,
                                             _ptr(self._deg32()), _ptr(cut), int(self._pisco), _ptr(t), _ptr(pl), _ptr(u),
                                             _ptr(sel), kmax, _ptr(ll_x), _ptr(ll_x2y), self.max_num_nodes, ld, b,
                                             _stream_ptr(self.device)), "isco_propose")
        return ll_x, y, {'ll_x2y': ll_x2y, 'selected_idx': {'sites': sel}, '_t': t}

    def step(self, x, path_length, temperature):
        x = x.contiguous()
        ll_x, y, trajectory = self.proposal(x, path_length, temperature)
        b, ld = x.shape
        cut_y, cross_y, weighted, cross_ld = self._raw_fields(y)
        u = th.rand((b,), device=self.device)                      # the draw of bernoulli_logp() (util.py:64)
        sel = trajectory['selected_idx']['sites']
        energy = th.empty((b,), dtype=th.float32, device=self.device)
        acc = th.empty_like(energy)
        with on_device(self.device):
            _lib.check(_lib.lib().rlsb_isco_accept(_ptr(x), _ptr(y), int(x.dtype == th.float16), _ptr(cross_y), weighted,
                                                   cross_ld, _ptr(self._deg32()), _ptr(cut_y), int(self._pisco),
                                                   _ptr(trajectory['_t']), _ptr(u), _ptr(sel), sel.shape[1], _ptr(ll_x),
                                                   _ptr(trajectory['ll_x2y']), _ptr(energy), _ptr(acc),
                                                   self.max_num_nodes, ld, b, _stream_ptr(self.device)), "isco_accept")
        return y, energy, acc


class ISCO_maxcut(_PathAuxMaxcut):
    def random_gen_init_sample(self, params_dict=None):
        return th.bernoulli(th.full((cfg.BATCH_SIZE, self.max_num_nodes), 0.5, device=self.device))

    def model(self, sample, temperature):
        """env_ISCO.py:79-86: number of cut edges / temperature (sample [N] or [B, N])."""
        cut, _ = self._int_fields(sample if sample.dim() == 2 else sample[None, :])
        energy = cut.float() / temperature
        return energy if sample.dim() == 2 else energy[0]

    def get_local_dist(self, sample, temperature):
        cut, gain2 = self._int_fields(sample)
        energy_x = cut.float() / temperature
        score_change_x = (gain2.float() / temperature) / 2
        return energy_x, th.log_softmax(score_change_x, dim=-1)


class PISCO_maxcut(_PathAuxMaxcut):
    _pisco = True

    def __init__(self, params_dict):
        super().__init__(params_dict)
        self.adj_matrix = params_dict['adj_matrix']
        self.sum_A = th.sum(self.adj_matrix)
        a = self.adj_matrix[:self.max_num_nodes, :self.max_num_nodes].float()
        if not bool((a == a.round()).all()):
            raise NotImplementedError("PISCO_maxcut: integer edge weights only (the packed-spin kernels count "
                                      "weighted edges exactly)")
        self._wdeg = None
        if not bool(((a == 0) | (a == 1)).all()):
            # weighted adjacency (Gset's +-1 instances): the energy is -1/4 s^T A s with A the WEIGHT matrix
            # (env_ISCO.py:436-444), so the integer core is the weighted cut and the weighted local fields
            iu = th.triu(a, diagonal=1).nonzero()
            w = a[iu[:, 0], iu[:, 1]].long()
            graph = th.stack([iu[:, 0], iu[:, 1], w], dim=1).cpu().numpy()
            self.store = GraphStore(graph, True, device=self.device, num_nodes=self.max_num_nodes)
            self._wdeg = a.sum(dim=1).long()[None, :]

    def _raw_fields(self, sample: TEN):
        if self._wdeg is None:
            return super()._raw_fields(sample)
        n = self.max_num_nodes
        bits = (sample[:, :n] != 0).contiguous()
        b = bits.shape[0]
        packed = self.store.pack(bits)
        return (self.store.cut_eval_weighted(packed=packed, num_envs=b), self.store.node_fields_weighted(packed, b), 1,
                self.store.padded_nodes)

    def _deg32(self) -> TEN:
        if self._wdeg is None:
            return super()._deg32()
        d = getattr(self, "_deg_i32", None)
        if d is None:
            d = self._deg_i32 = self._wdeg.reshape(-1).to(th.int32).contiguous()
        return d

    def _int_fields(self, sample: TEN):
        if self._wdeg is None:
            return super()._int_fields(sample)
        n = self.max_num_nodes
        bits = (sample[:, :n] != 0).contiguous()
        b = bits.shape[0]
        packed = self.store.pack(bits)
        cut = self.store.cut_eval_weighted(packed=packed, num_envs=b)
        cross = self.store.node_fields_weighted(packed, b)[:, :n].long()
        return cut, self._wdeg - 2 * cross

    def random_gen_init_sample(self):
        return th.bernoulli(th.full((cfg.BATCH_SIZE, self.max_num_nodes), 0.5, device=self.device)).to(th.float16)

    def tensor_core_energy(self, sample, temperature):
        """env_ISCO.py:436-444: (-1/4 s^T A s) / T and its x-gradient / T, with the reference's
        roundings: the quadratic form is rounded to fp16 once, the gradient -A d is an exact fp16
        integer, both are divided by T in fp32."""
        cut, gain2 = self._int_fields(sample)
        n, npad = self.max_num_nodes, sample.shape[1]
        quad = gain2.sum(dim=1).to(th.float16)                       # s^T A s = 2M - 4 cut
        energy_x = (-0.25 * quad).to(th.float) / temperature
        d = (sample[:, :n].to(th.float) * 2 - 1)
        grad = th.zeros((sample.shape[0], npad), dtype=th.float, device=sample.device)
        grad[:, :n] = -(gain2.to(th.float) * d)                      # -(A d)_i = -(d_i * gain2_i)
        return energy_x, grad / temperature

    def get_local_dist(self, sample, temperature):
        energy_x, grad_x = self.tensor_core_energy(sample, temperature)
        delta_x = 1 - sample * 2
        score_change_x = (delta_x * grad_x) / 2
        return energy_x, th.log_softmax(score_change_x, dim=-1)

    def check_tensor(self, tensor):
        if th.isinf(tensor).any():
            raise ValueError(" contains inf values!")


__all__ = ["ISCO_maxcut", "PISCO_maxcut"]
