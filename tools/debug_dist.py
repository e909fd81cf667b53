import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as th, torch.distributed as dist
from rlsolver_b200.dist import best_allreduce, _local_record
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
th.cuda.set_device(local); dev = th.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
vs = th.randint(12000, 13000, (4096,), device=dev); xs = th.randint(0, 2, (4096, 2000), device=dev).bool()
def timeit(fn, reps=200):
    for _ in range(20): fn()
    dist.barrier(); th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    t = time.perf_counter(); a.record()
    for _ in range(reps): fn()
    b.record(); th.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, (time.perf_counter() - t) / reps * 1e6
rec = _local_record(vs, xs, rank, 4096)
gathered = th.empty((world, rec.numel()), dtype=th.uint8, device=dev)
print(rank, "record kernel us (gpu, wall)", timeit(lambda: _local_record(vs, xs, rank, 4096)), flush=True)
print(rank, "all_gather us", timeit(lambda: dist.all_gather_into_tensor(gathered, rec.unsqueeze(0))), flush=True)
print(rank, "best_allreduce us", timeit(lambda: best_allreduce(vs, xs, rank, world, 4096)), flush=True)
k = th.zeros(1, dtype=th.int64, device=dev)
print(rank, "all_reduce(8B) us", timeit(lambda: dist.all_reduce(k, op=dist.ReduceOp.MAX)), flush=True)
dist.destroy_process_group()
