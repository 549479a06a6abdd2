"""cfg5 family: batch of independent small LPs (m=500, n=2000) on one GPU; throughput vs concurrency / CTAs."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
from abip_b200 import problems, lp_solve_batch, lp_solve
count = int(sys.argv[1]) if len(sys.argv) > 1 else 64
probs = [problems.random_lp(500, 2000, 5, seed=5000 + i) for i in range(count)]
t = time.time(); x, y, s, info = lp_solve(probs[0].csc(), probs[0].b, probs[0].c, dict(tol=1e-4, verbose=0)); print('single full-grid solve: %.3fs' % (time.time() - t), info['status'], info['admm_iter'], info['pobj'], flush=True)
cfgs = [tuple(int(v) for v in a.split('x')) for a in sys.argv[2:]] or [(1, 148), (1, 8), (8, 16), (16, 8), (18, 8), (32, 4), (36, 4), (64, 2)]
for conc, ctas in cfgs:
    t = time.time()
    res = lp_solve_batch(probs, dict(tol=1e-4, verbose=0), concurrency=conc, ctas_per_problem=ctas)
    dt = time.time() - t
    ok = sum(r[3]['status'] == 'Solved' for r in res)
    print(json.dumps({'count': count, 'concurrency': conc, 'ctas': ctas, 'wall_s': round(dt, 3), 'lp_per_s': round(count / dt, 1), 'solved': ok,
                      'admm_iter_0': res[0][3]['admm_iter'], 'pobj_0': res[0][3]['pobj']}), flush=True)
