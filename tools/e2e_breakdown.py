import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, ctypes as C
from abip_b200 import problems, _capi, api
p = problems.cfg2()
A = p.csc()
L = _capi.lib()
for rep in range(3):
    t0 = time.perf_counter()
    H = api.CscHolder(A)
    t1 = time.perf_counter()
    pp, st = api._lp_settings(dict(tol=1e-4, verbose=0))
    b = np.ascontiguousarray(p.b); c = np.ascontiguousarray(p.c)
    d = _capi.ABIPData(H.m, H.n, C.pointer(H.c), api._fp(b), api._fp(c), float(H.nnz) / (float(H.m) * float(H.n)), C.pointer(st))
    sol = _capi.ABIPSolution(); info = _capi.ABIPInfo()
    w = L.abip_gpu_init(C.byref(d), C.byref(info))
    t2 = time.perf_counter()
    L.abip_gpu_solve(w, C.byref(d), C.byref(sol), C.byref(info))
    t3 = time.perf_counter()
    L.abip_gpu_finish(w)
    t4 = time.perf_counter()
    print('holder %.3f init %.3f (reported setup %.3f) solve %.3f (reported %.3f) finish %.3f total %.3f' % (t1-t0, t2-t1, info.setup_time/1e3, t3-t2, info.solve_time/1e3, t4-t3, t4-t0), flush=True)
from abip_b200 import lp_solve
for rep in range(3):
    t0 = time.perf_counter()
    x, y, s, info = lp_solve(A, p.b, p.c, dict(tol=1e-4, verbose=0), want_stats=True)
    print('lp_solve total %.3f (setup %.3f solve %.3f)' % (time.perf_counter() - t0, info['setup_time_ms']/1e3, info['solve_time_ms']/1e3), flush=True)
import torch
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for rep in range(2):
    t0 = time.perf_counter()
    x, y, s, info = lp_solve(A, p.b, p.c, dict(tol=1e-4, verbose=0), want_stats=True)
    print('with torch ctx: lp_solve total %.3f (setup %.3f solve %.3f)' % (time.perf_counter() - t0, info['setup_time_ms']/1e3, info['solve_time_ms']/1e3), flush=True)
