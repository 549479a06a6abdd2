import sys
sys.path.insert(0, '/root/repo')
from abip_b200 import problems, LpSolver
p = problems.cfg2(scale=0.02)
s = LpSolver(p.csc(), dict(tol=1e-4, verbose=0))
x, y, z, info = s.solve(p.b, p.c)
print(info)
