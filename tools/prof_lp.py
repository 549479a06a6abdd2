"""Profiling driver: a few ADMM iterations + one BB round of cfg2 through the step ABI (for ncu)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from abip_b200 import problems, _capi, api
import ctypes as C
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
p = problems.cfg2(scale=scale)
L = _capi.lib()
H = api.CscHolder((p.m, p.n, p.Ap.copy(), p.Ai.copy(), p.Ax.copy()))
st = _capi.default_settings(verbose=0)
sc = _capi.ABIPScaling()
L.abip_normalize_A(C.byref(H.c), C.byref(st), C.byref(sc))
D = np.ctypeslib.as_array(sc.D, shape=(p.m,)).copy(); E = np.ctypeslib.as_array(sc.E, shape=(p.n,)).copy()
b = p.b / D; c = p.c / E
b *= sc.mean_norm_col_A / max(np.linalg.norm(b), 1e-3); c *= sc.mean_norm_row_A / max(np.linalg.norm(c), 1e-3)
e = api.LpEngine((p.m, p.n, H.Ap, H.Ai, H.Ax))
print(e.describe())
e.set_problem(b, c, D, E)
e.cold_start(1.0, 1.0); e.outer_prologue(0)
t = time.time()
for j in range(iters):
    s = e.admm_iter(j, 40 + j, 1e-2, 1.0)
    print(j, 'cg', s[0], 'tol %.3e res %.3e' % (s[2], s[3]))
print('avg ms/iter', (time.time() - t) / iters * 1e3)
e.bb_begin()
s = e.bb_round(0, 50, 1e-2, 1.0)
print('bb cg', s[0], s[1])
if hasattr(L, 'abipgpu_lp_phase_times'):
    out = np.zeros(32)
    L.abipgpu_lp_phase_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    L.abipgpu_lp_phase_times(e.e, out.ctypes.data_as(C.POINTER(C.c_double)), 0)
    names = ['rhs', 'rhsB', 'S1(A+AT)', 'S2(A)', 'L1(AT)', 'L2(A)', 'L3(upd)', 'L4(p)', 'S4(AT)', 'prox', 'qnorm']
    if out[16:].sum() > 0:
        for i, nm in enumerate(names):
            if out[16 + i] > 0:
                print('phase %-9s count %6d  avg %8.2f us  total %8.2f ms' % (nm, out[16 + i], out[i] / out[16 + i] / 1e3, out[i] / 1e6))
