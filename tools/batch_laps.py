import sys, time, os
sys.path.insert(0, '/root/repo')
from abip_b200 import problems, lp_solve_batch
probs = [problems.random_lp(500, 2000, 5, seed=5000 + i) for i in range(24)]
lp_solve_batch(probs[:16], dict(tol=1e-4, verbose=0), concurrency=16)
t=time.time(); r = lp_solve_batch(probs[16:20], dict(tol=1e-4, verbose=1), concurrency=1); print('4 problems sequential, verbose (host outer loop): %.1f ms each' % ((time.time()-t)*250))
t=time.time(); r = lp_solve_batch(probs[20:24], dict(tol=1e-4, verbose=0), concurrency=1); print('4 problems sequential, device outer loop: %.1f ms each' % ((time.time()-t)*250), [x[3]['admm_iter'] for x in r], [x[3].get('setup_time_ms') for x in r], [x[3].get('solve_time_ms') for x in r])
