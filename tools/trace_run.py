import os, sys
sys.path.insert(0, '/root/repo')
from abip_b200 import problems, lp_solve
p = problems.cfg1()
os.environ['ABIP_GPU_TRACE'] = '/tmp/gpu_trace.txt'
x,y,s,info = lp_solve(p.csc(), p.b, p.c, dict(tol=1e-4, verbose=0))
print(open('/tmp/gpu_trace.txt').read()[:6000])
