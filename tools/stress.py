import os, sys, hashlib
sys.path.insert(0, '/root/repo')
import numpy as np
from abip_b200 import problems, lp_solve
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg1'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
p = {'cfg1': problems.cfg1, 'small': lambda: problems.random_lp(200, 700, 4, seed=3),
     'mcf': lambda: problems.mcf_lp(4, 40, 200, 6, 300, seed=6), 'cfg2s': lambda: problems.cfg2(scale=0.05)}[name]()
seen = {}
for r in range(reps):
    x, y, s, info = lp_solve(p.csc(), p.b, p.c, dict(tol=1e-4, verbose=0))
    h = hashlib.md5(x.tobytes() + y.tobytes() + s.tobytes()).hexdigest()
    seen.setdefault((info['admm_iter'], h), 0)
    seen[(info['admm_iter'], h)] += 1
print(name, p.m, p.n, p.nnz, seen)
