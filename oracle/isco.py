"""Torch restatement of the path-auxiliary sampling helpers of ISCO (rlsolver/methods/ISCO/util.py:3-75) and of the
step built on them (rlsolver/envs/env_ISCO.py:27-77).  TEST INFRASTRUCTURE ONLY: the product path runs
csrc/isco.cu (rlsb_isco_propose / rlsb_isco_accept); the tests feed both the same uniform draws and compare."""
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



def step(log_prob_of, x: TEN, path_length: TEN, u_gumbel: TEN, u_accept: TEN):
    """One MH step (env_ISCO.py:27-77) given `log_prob_of(state) -> (energy [B], log_prob [B, N])` and the two uniform
    draws the reference makes (gumbel: [B, N], bernoulli_logp: [B]).  Returns (next state, ll_y, log_acc, ll_x2y,
    ll_y2x, selected mask)."""
    ll_x, log_prob = log_prob_of(x)
    num_classes = log_prob.shape[-1]
    perturbed = log_prob - th.log(-th.log(u_gumbel))
    ascending, _ = th.sort(perturbed)
    threshold = th.gather(ascending, 1, (num_classes - path_length).unsqueeze(1))
    mask = (perturbed >= threshold.expand_as(perturbed)).int()
    order = th.argsort(-perturbed, dim=-1)
    ll_in_order = noreplacement_sampling_renormalize(th.gather(log_prob, dim=-1, index=order))
    ll_selected = th.zeros_like(ll_in_order)
    ll_selected.scatter_(1, order, ll_in_order)
    ll_x2y = th.sum(ll_selected * mask, dim=-1)
    y = x * (1 - mask) + mask * (1 - x)
    ll_y, log_prob_y = log_prob_of(y)
    backwd_idx = th.argsort(perturbed, dim=-1)
    log_prob_y = th.where(mask.bool(), log_prob_y, th.tensor(-1e18, device=x.device))
    backwd_ll = th.gather(log_prob_y, dim=-1, index=backwd_idx)
    backwd_mask = th.gather(mask, dim=-1, index=backwd_idx)
    ll_backwd = noreplacement_sampling_renormalize(backwd_ll)
    ll_y2x = th.sum(th.where(backwd_mask.bool(), ll_backwd, th.tensor(0.0, device=x.device)), dim=-1)
    log_acc = th.clamp(ll_y + ll_y2x - ll_x - ll_x2y, max=0.0)
    accept = th.log(u_accept + 1e-24) < log_acc
    nxt = th.where(accept.unsqueeze(-1).expand_as(y), y, x)
    return nxt, ll_y, log_acc, ll_x2y, ll_y2x, mask
