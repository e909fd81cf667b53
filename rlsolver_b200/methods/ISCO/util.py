"""Path-auxiliary sampling helpers of ISCO (rlsolver/methods/ISCO/util.py:3-75), restated.

These are small dense float32 ops on [B, N] tensors (sort / cumsum / exp / log); they stay torch
calls in the reference's order so a replayed uniform stream gives the reference's choices.  The
part of an ISCO step that touches the graph -- energies and per-node flip gains -- comes from the
sm_100a kernels (see envs/env_ISCO.py)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch as th

TEN = th.Tensor


def gumbel(loc: TEN) -> TEN:
    """util.py:3-5: loc + Gumbel(0, 1) noise from one uniform draw of loc's shape."""
    u = th.rand(loc.shape, device=loc.device)
    return loc - th.log(-th.log(u))


def log1mexp(x: TEN) -> TEN:
    """util.py:7-10: log(1 - exp(-|x|)), switching formula at -0.693 as the reference does."""
    neg = -th.abs(x)
    return th.where(neg > -0.693, th.log(-th.expm1(neg)), th.log1p(-th.exp(neg)))


def noreplacement_sampling_renormalize(ll_idx: TEN, dim: int = -1) -> TEN:
    """util.py:12-17: log-probability of drawing the items in this order without replacement."""
    top = th.max(ll_idx, dim=dim, keepdim=True).values
    weight = th.exp(ll_idx - top)
    taken_before = th.log(th.cumsum(weight, dim=dim) - weight) + top
    return th.clamp(ll_idx - log1mexp(taken_before), max=0.0)


def multinomial(log_prob: TEN, path_length: TEN) -> Tuple[Dict[str, TEN], TEN]:
    """util.py:19-60: Gumbel top-k choice of path_length[b] sites per chain.

    Returns ({'selected_mask' int32 [B,N], 'perturbed_ll' [B,N]}, ll_selected [B,N])."""
    num_classes = log_prob.shape[-1]
    perturbed = gumbel(log_prob)
    ascending, _ = th.sort(perturbed)
    threshold = th.gather(ascending, 1, (num_classes - path_length).unsqueeze(1))
    mask = (perturbed >= threshold.expand_as(perturbed)).int()
    order = th.argsort(-perturbed, dim=-1)
    ll_in_order = noreplacement_sampling_renormalize(th.gather(log_prob, dim=-1, index=order))
    ll_selected = th.zeros_like(ll_in_order)
    ll_selected.scatter_(1, order.view(-1, num_classes), ll_in_order.view(-1, num_classes))
    ll_selected = ll_selected.view(log_prob.shape) * mask
    return {'selected_mask': mask, 'perturbed_ll': perturbed}, ll_selected


def bernoulli_logp(log_prob: TEN) -> TEN:
    """util.py:62-65."""
    noise = th.rand(log_prob.shape, device=log_prob.device)
    return th.log(noise + 1e-24) < log_prob


def mh_step(log_prob: TEN, current_sample: TEN, new_sample: TEN) -> Tuple[TEN, TEN]:
    """util.py:67-75: Metropolis-Hastings accept per chain."""
    accept = bernoulli_logp(log_prob)
    return th.where(accept.unsqueeze(-1).expand_as(new_sample), new_sample, current_sample), accept
