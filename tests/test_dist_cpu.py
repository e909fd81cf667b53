"""World-size-2 gloo test of the best-cut exchange (host logic of the N>1 path)."""
import os
import socket

import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from rlsolver_b200.dist import best_allreduce, decode_key, local_best_key


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = th.Generator().manual_seed(5)
    all_vs = th.randint(0, 50, (world * 8,), generator=g)
    all_vs[3] = all_vs[11] = 99                      # tie across ranks -> lowest global id wins
    all_xs = th.randint(0, 2, (world * 8, 13), generator=g).bool()
    vs, xs = all_vs[rank * 8:(rank + 1) * 8], all_xs[rank * 8:(rank + 1) * 8]
    cut, gid, row = best_allreduce(vs, xs, rank, world, 8)
    ok = cut == 99 and gid == 3 and th.equal(row, all_xs[3])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_best_allreduce_world2():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert all(out[r] for r in range(world)) and len(out) == world


def test_key_roundtrip():
    vs = th.tensor([5, 9, 9, 1])
    k = int(local_best_key(vs, rank=2, envs_per_rank=4).item())
    assert decode_key(k) == (9, 2 * 4 + 1)
