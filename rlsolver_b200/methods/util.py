"""Graph helpers of rlsolver/methods/util.py used on the hot path:
calc_num_nodes_in_mygraph (35-40), evolutionary_replacement (87-94), build_adjacency_bool (343-370)."""
from __future__ import annotations

import numpy as np
import torch as th

from .config import MyGraph

TEN = th.Tensor


def calc_num_nodes_in_mygraph(mygraph: MyGraph) -> int:
    """Number of distinct endpoints -- isolated nodes do not count (reference behaviour)."""
    if len(mygraph) == 0:
        return 0
    arr = np.asarray([(a, b) for a, b, _ in mygraph], dtype=np.int64)
    return int(np.unique(arr).size)


def build_adjacency_bool(mygraph: MyGraph, num_nodes: int = 0, if_bidirectional: bool = False) -> TEN:
    if num_nodes == 0:
        num_nodes = calc_num_nodes_in_mygraph(mygraph)
    adj = np.zeros((num_nodes, num_nodes), dtype=bool)
    arr = np.asarray([(a, b) for a, b, _ in mygraph], dtype=np.int64)
    adj[arr[:, 0], arr[:, 1]] = True
    if if_bidirectional:
        adj |= adj.T
    return th.from_numpy(adj)


def evolutionary_replacement(xs: TEN, vs: TEN, low_k: int, if_maximize: bool):
    """Overwrite `low_k` random non-elite rows with the `low_k` best rows (in place).
    The argsort (its order among equal values) and the one randperm(E - low_k) on xs.device are torch's, as in the
    reference; the row copy is one kernel (rlsb_copy_rows) for bool CUDA rows."""
    num_sims = xs.shape[0]
    ids = vs.argsort()
    top_ids, low_ids = (ids[:-low_k], ids[-low_k:]) if if_maximize else (ids[:low_k], ids[low_k:])
    if not if_maximize and top_ids.shape[0] < num_sims - low_k:
        # minimising, the reference indexes its low_k-row `top_ids` with a permutation of E - low_k > low_k numbers:
        # IndexError on the CPU, a device-side assert (which poisons the CUDA context) on the GPU.  Same failure, raised
        # on the host before anything is launched.
        raise IndexError(f"index out of range: a permutation of {num_sims - low_k} indexes {top_ids.shape[0]} rows "
                         f"(rlsolver/methods/util.py:91-92 with if_maximize=False)")
    replace_ids = top_ids[th.randperm(num_sims - low_k, device=xs.device)[:low_k]]
    if (xs.is_cuda and xs.dtype == th.bool and xs.dim() == 2 and xs.is_contiguous() and vs.dtype == th.int64
            and vs.is_contiguous() and replace_ids.numel() == low_ids.numel()):
        from .. import _lib
        from ..graph_store import on_device
        bad = th.zeros((1,), dtype=th.int32, device=xs.device)
        replace_ids, low_ids = replace_ids.contiguous(), low_ids.contiguous()
        with on_device(xs.device):
            _lib.check(_lib.lib().rlsb_copy_rows(xs.data_ptr(), vs.data_ptr(), replace_ids.data_ptr(), low_ids.data_ptr(),
                                                 replace_ids.numel(), num_sims, xs.shape[1], bad.data_ptr(),
                                                 th.cuda.current_stream(xs.device).cuda_stream), "copy_rows")
        return
    xs[replace_ids] = xs[low_ids]          # other layouts / devices: the reference's own indexing (and its errors)
    vs[replace_ids] = vs[low_ids]
