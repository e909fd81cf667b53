import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as th, torch.distributed as dist
from rlsolver_b200.dist import local_best_key, decode_key, best_allreduce
from rlsolver_b200.envs.env_L2A import EnvMaxcut
from synth import gset_like
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
th.cuda.set_device(local); dev = th.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sim = EnvMaxcut(mygraph=gset_like("G22"), device=dev, if_bidirectional=True)
th.manual_seed(74 + rank)
xs = sim.generate_xs_randomly(4096)
gx, gv = sim.local_search_inplace(xs, th.empty(()), 8, 8, 0.3)
th.cuda.synchronize()
print(rank, "gv", gv.dtype, gv.shape, int(gv.min()), int(gv.max()), gv.device, flush=True)
key = local_best_key(gv, rank, 4096)
print(rank, "key", hex(int(key.item())), flush=True)
print(rank, "best", best_allreduce(gv, gx, rank, world, 4096)[:2], flush=True)
dist.destroy_process_group()
