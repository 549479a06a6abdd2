"""torchrun --nproc-per-node N tools/batch_dist.py [count] : cfg5 (BASELINE.json configs[4]) sharded one problem set per
GPU -- no data-path collective; aggregate LP/s = count / max over ranks of the wall time (barrier on both sides)."""
import json, os, sys, time
sys.path.insert(0, '/root/repo')
import torch, torch.distributed as dist
from abip_b200 import problems
from abip_b200.dist import lp_solve_batch_sharded, shard_indices
rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
os.environ['ABIP_GPU_DEVICE'] = str(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
count = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mine = set(shard_indices(count, world, rank))
probs = [problems.random_lp(500, 2000, 5, seed=5000 + i) if i in mine else None for i in range(count)]  # only the shard is generated
lp_solve_batch_sharded([p for p in probs[:8 * world]], dict(tol=1e-4, verbose=0), gather=False)            # warm-up
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
res = lp_solve_batch_sharded(probs, dict(tol=1e-4, verbose=0), gather=False)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], device='cuda')
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
ok = torch.tensor([sum(r[3]['status'] == 'Solved' for r in res.values())], device='cuda')
dist.all_reduce(ok)
if rank == 0:
    print(json.dumps({'config': 'cfg5 batch, one problem set per GPU', 'count': count, 'n_gpus': world, 'wall_s_max': round(float(dt), 3),
                      'lp_per_s': round(count / float(dt), 1), 'solved': int(ok)}))
dist.destroy_process_group()
