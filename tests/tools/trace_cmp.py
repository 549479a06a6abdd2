import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from abip_b200 import problems, lp_solve
from oracle import lp_oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg1'
p = problems.cfg1() if name == 'cfg1' else problems.random_lp(200,700,4,seed=3)
os.environ['ABIP_GPU_TRACE'] = '/tmp/gpu_trace.txt'
x,y,s,info = lp_solve(p.csc(), p.b, p.c, dict(tol=1e-4, verbose=0))
o = O.solve(p.csc(), p.b, p.c, O.Settings(eps=1e-4), trace=True)
print(info['admm_iter'], o.admm_iter)
g = [l.split() for l in open('/tmp/gpu_trace.txt')]
gi = [l for l in g if l[0]=='it']
for a, b in zip(gi, o.trace):
    print('gpu i=%s j=%s k=%s mu=%.3e beta=%.6f cg=%s q=%.6e avg=%s | orc i=%d j=%d k=%d mu=%.3e beta=%.6f cg=%d q=%.6e' % (a[1],a[2],a[3],float(a[4]),float(a[5]),a[6],float(a[7]),a[8], *b))
