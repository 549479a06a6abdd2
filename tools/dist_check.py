"""torchrun --nproc-per-node N tools/dist_check.py [case] : multi-GPU solve vs the single-GPU engine."""
import os, sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import torch, torch.distributed as dist
from abip_b200 import problems, lp_solve
from abip_b200.dist import LpSolverDist
rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
os.environ['ABIP_GPU_DEVICE'] = str(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
case = sys.argv[1] if len(sys.argv) > 1 else 'small'
p = {'small': lambda: problems.random_lp(200, 700, 4, seed=3), 'mcf': lambda: problems.mcf_lp(4, 40, 200, 6, 300, seed=6),
     'cfg1': problems.cfg1, 'cfg2s': lambda: problems.cfg2(scale=0.05), 'cfg2': problems.cfg2,
     'cfg4s': lambda: problems.cfg4(scale=float(os.environ.get('CFG4_SCALE', '0.25'))), 'cfg4': problems.cfg4}[case]()
print('rank', rank, 'init...', flush=True)
sol = LpSolverDist(p.csc(), dict(tol=1e-4, verbose=int(os.environ.get('VERB', 0))), min_nnz_per_gpu=0)  # force sharding of the small test problems
print('rank', rank, 'connected', flush=True)
dist.barrier(); torch.cuda.synchronize()
t0 = time.time()
x, y, s, info = sol.solve(p.b, p.c)
dt = time.time() - t0
print('rank', rank, 'solved', info['status'], dt, flush=True)
x2, y2, s2, info2 = sol.solve(p.b, p.c)
if rank == 0:
    A = p.csc()
    pres = np.linalg.norm(A @ x - p.b) / (1 + np.linalg.norm(p.b)); dres = np.linalg.norm(A.T @ y + s - p.c) / (1 + np.linalg.norm(p.c))
    print(json.dumps({'case': case, 'world': dist.get_world_size(), 'm': p.m, 'n': p.n, 'nnz': p.nnz, 'status': info['status'],
                      'ipm': info['ipm_iter'], 'admm': info['admm_iter'], 'pobj': info['pobj'], 'pres_cpu': pres, 'dres_cpu': dres,
                      'solve_ms': info['solve_time_ms'], 'solve2_ms': info2['solve_time_ms'], 'repeat_identical': bool(np.array_equal(x, x2)),
                      'cg': info['stats']['n_cg_iters']}))
sol.close()
if case in ('small', 'mcf', 'cfg1', 'cfg2s') and rank == 0:
    x1, y1, s1, i1 = lp_solve(p.csc(), p.b, p.c, dict(tol=1e-4, verbose=0))
    print('single-GPU:', i1['status'], i1['ipm_iter'], i1['admm_iter'], i1['pobj'], 'x diff', np.max(np.abs(x - x1)) / np.max(np.abs(x1)))
dist.barrier()
dist.destroy_process_group()
