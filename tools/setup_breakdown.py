"""Host set-up breakdown of abip_gpu_init on cfg2 (engine description printed by the verbose init) + e2e of lp_solve."""
import sys, time
sys.path.insert(0, '/root/repo')
from abip_b200 import problems, lp_solve
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
p = problems.cfg2(scale=scale)
A = p.csc()
lp_solve(A, p.b, p.c, dict(tol=1e-4, verbose=0))
for rep in range(2):
    t = time.time()
    x, y, ss, info = lp_solve(A, p.b, p.c, dict(tol=1e-4, verbose=1 if rep == 1 else 0))
    print('lp_solve e2e %.3f s  setup %.1f ms  solve %.1f ms' % (time.time() - t, info.get('setup_time_ms', -1), info.get('solve_time_ms', -1)))
