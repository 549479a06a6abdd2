"""Small batch through the lock-step executor with a watchdog (debugging aid)."""
import sys, time, faulthandler
sys.path.insert(0, '/root/repo')
faulthandler.dump_traceback_later(int(sys.argv[3]) if len(sys.argv) > 3 else 50, exit=True)
from abip_b200 import problems, lp_solve_batch
count = int(sys.argv[1]) if len(sys.argv) > 1 else 8
conc = int(sys.argv[2]) if len(sys.argv) > 2 else count
probs = [problems.random_lp(500, 2000, 5, seed=5000 + i) for i in range(count)]
lp_solve_batch(probs[:16], dict(tol=1e-4, verbose=0), concurrency=16, ctas_per_problem=1)  # context + module load
t = time.time()
res = lp_solve_batch(probs, dict(tol=1e-4, verbose=0), concurrency=conc, ctas_per_problem=1)
dt = time.time() - t
print('count', count, 'conc', conc, 'wall %.3f s' % dt, 'engine wall %.3f s = %.1f LP/s' % (res[0][3]['batch_wall_s'], count / res[0][3]['batch_wall_s']), '%.1f LP/s' % (count / dt), 'solved', sum(r[3]['status'] == 'Solved' for r in res),
      'admm', [r[3]['admm_iter'] for r in res[:8]], 'pobj0 %.10f' % res[0][3]['pobj'], flush=True)
