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


def test_negative_values_order_below_positive_ones():
    """Weighted cuts / QUBO energies can be negative: the key is biased so they rank below positive values
    (kernel and torch path use the same rule, csrc/select.cu best_key)."""
    vs = th.tensor([-7, -2, -2, -100])
    assert decode_key(int(local_best_key(vs, rank=1, envs_per_rank=4).item())) == (-2, 4 + 1)
    vs = th.tensor([-7, 3, -2, 0])
    assert decode_key(int(local_best_key(vs, rank=0, envs_per_rank=4).item())) == (3, 1)
    xs = th.arange(4 * 5).reshape(4, 5).remainder(2).bool()
    cut, gid, row = best_allreduce(th.tensor([-7, -3, -3, -9]), xs, rank=0, world=1, envs_per_rank=4)
    assert int(cut) == -3 and int(gid) == 1 and th.equal(row, xs[1])
    # saturation to the int32 range
    big = th.tensor([2 ** 40, 5])
    assert decode_key(int(local_best_key(big, 0, 2).item())) == (2 ** 31 - 1, 0)


def test_peer_exchange_needs_cuda_devices():
    """The peer-memory exchange is a CUDA kernel over NVLink mailboxes: on a CPU device it refuses loudly (the gloo /
    torch path for CPU tensors is best_allreduce)."""
    import pytest
    from rlsolver_b200.dist import PeerBestExchange
    with pytest.raises(RuntimeError, match="CUDA"):
        PeerBestExchange(10, 0, 1, 4, th.device("cpu"))


def test_host_pipeline_rejects_bad_arguments():
    import pytest
    from rlsolver_b200.host_pipeline import HostPipeline
    with pytest.raises(ValueError):
        HostPipeline(None, 4, layout="rows")
    with pytest.raises(ValueError):
        HostPipeline(None, 4, depth=1)
