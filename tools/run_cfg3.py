"""cfg3 (BASELINE.json configs[2]): SOCP/QCP with 10k second-order cones, n = 500k, nnz(A) = 10M, sparse PSD Q."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from abip_b200 import problems
from abip_b200.qcp import qcp_solve_raw
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
t = time.time(); p = problems.cfg3(scale=scale); print('gen %.1fs' % (time.time() - t), p.m, p.n, p.A.nnz, p.Q.nnz, flush=True)
for rep in range(2):
    x, y, s, info = qcp_solve_raw(p.A, p.Q, p.b, p.c, p.K, eps_p=1e-4, eps_d=1e-4, eps_g=1e-4, verbose=0)
    print({k: v for k, v in info.items()}, flush=True)
Qx = p.Q @ x
print('pres', np.max(np.abs(p.A @ x - p.b)) / (1 + max(np.max(np.abs(p.A @ x)), np.max(np.abs(p.b)))),
      'dres', np.max(np.abs(Qx - p.A.T @ y + p.c - s)) / (1 + max(np.max(np.abs(Qx)), np.max(np.abs(p.c)))),
      'pobj', 0.5 * x @ Qx + p.c @ x, 'dobj', -0.5 * x @ Qx + p.b @ y)
xs = x.reshape(-1, 50)
print('min SOC margin', np.min(xs[:, 0] - np.linalg.norm(xs[:, 1:], axis=1)))
