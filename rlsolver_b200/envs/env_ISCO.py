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
division by T, log_softmax, the Gumbel top-k path proposal and the MH accept) is the reference's
own sequence of torch ops, so a replayed uniform stream reproduces its choices.

Numerics: PISCO's energy and flip gains reproduce the reference's fp16/fp32 roundings exactly
(small integers, one division); ISCO's autograd accumulates +-0.5/T per edge with atomics, so its
gains carry a few ulp of order-dependent noise in the reference itself -- parity there is to
1e-6 relative (tests/test_gpu_isco.py).  ISCO_maxcut takes the edge list and ignores weights like the reference;
PISCO_maxcut's adjacency may carry integer weights (rlsb_cut_eval_weighted / rlsb_node_fields_weighted)."""
from __future__ import annotations

import torch as th

from ..graph_store import GraphStore, require_cuda
from ..methods.ISCO import config_maxcut as cfg
from ..methods.ISCO.util import mh_step, multinomial, noreplacement_sampling_renormalize

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

    def step(self, x, path_length, temperature):
        ll_x, y, trajectory = self.proposal(x, path_length, temperature)
        ll_x2y = trajectory['ll_x2y']
        ll_y, ll_y2x = self.ll_y2x(trajectory, y, temperature)
        log_acc = th.clamp(ll_y + ll_y2x - ll_x - ll_x2y, max=0.0)
        y = self.select_sample(log_acc, x, y)
        return y, ll_y * temperature, log_acc.exp()

    def proposal(self, x, path_length, temperature):
        ll_x, log_prob = self.get_local_dist(x, temperature)
        selected_idx, ll_selected = multinomial(log_prob, path_length)
        mask = selected_idx['selected_mask']
        y = x * (1 - mask) + mask * (1 - x)
        return ll_x, y, {'ll_x2y': th.sum(ll_selected, dim=-1), 'selected_idx': selected_idx}

    def ll_y2x(self, forward_trajectory, y, temperature):
        ll_y, log_prob = self.get_local_dist(y, temperature)
        selected_mask = forward_trajectory['selected_idx']['selected_mask']
        backwd_idx = th.argsort(forward_trajectory['selected_idx']['perturbed_ll'], dim=-1)
        log_prob = th.where(selected_mask.bool(), log_prob, th.tensor(-1e18, device=log_prob.device))
        backwd_ll = th.gather(log_prob, dim=-1, index=backwd_idx)
        backwd_mask = th.gather(selected_mask, dim=-1, index=backwd_idx)
        ll_backwd = noreplacement_sampling_renormalize(backwd_ll)
        zero = th.tensor(0.0, device=log_prob.device)
        return ll_y, th.sum(th.where(backwd_mask.bool(), ll_backwd, zero), dim=-1)

    def select_sample(self, log_acc, x, y):
        y, _ = mh_step(log_acc, x, y)
        return y


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
