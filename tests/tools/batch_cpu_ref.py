"""Reference CPU rate on the cfg5 family (BASELINE.json configs[4]): the reference's own C solver (oracle/_ref,
compiled unmodified, indirect path, one thread) looped over the problems, as one process per core would do."""
import sys, time, json
sys.path.insert(0, '/root/repo')
from abip_b200 import problems
from oracle import ref_lp
count = int(sys.argv[1]) if len(sys.argv) > 1 else 32
probs = [problems.random_lp(500, 2000, 5, seed=5000 + i) for i in range(count)]
t = time.perf_counter()
its, ok = 0, 0
for p in probs:
    r = ref_lp.solve(p, which='indirect', eps=1e-4)
    its += r['admm_iter']; ok += r['status'] == 'Solved'
dt = time.perf_counter() - t
print(json.dumps({'count': count, 'wall_s': round(dt, 3), 'lp_per_s_per_core': round(count / dt, 2), 'solved': ok,
                  'mean_admm_iter': its / count}))
