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
    names = ['rhs', 'rhsB', 'S1(A+AT)', 'S2(A)', 'L1(AT)', 'L2(A)', 'L3+L4(upd)', '-', 'S4(AT)', 'prox', 'qnorm']
    if out[16:].sum() > 0:
        for i, nm in enumerate(names):
            if out[16 + i] > 0:
                print('phase %-9s count %6d  avg %8.2f us  total %8.2f ms' % (nm, out[16 + i], out[i] / out[16 + i] / 1e3, out[i] / 1e6))
if hasattr(L, 'abipgpu_lp_warp_times') and out[16:].sum() > 0:
    W = C.c_int(0)
    buf = np.zeros(2 * 148 * 32 * 4 + 1024)
    L.abipgpu_lp_warp_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.abipgpu_lp_warp_times(e.e, buf.ctypes.data_as(C.POINTER(C.c_double)), C.byref(W))
    W = W.value
    import os
    os.makedirs('gpurun_out', exist_ok=True)
    np.save('gpurun_out/warp_times.npy', buf[:2 * W + W // 16])
    calls = out[16 + 4]
    for which, nm in ((0, "A' pass"), (1, 'A pass')):
        t = buf[which * W:(which + 1) * W] / max(calls, 1) / 1e3   # us per call per warp
        import re
        wpc = int(re.search(r'x (\d+) threads', e.describe()).group(1)) // 32
        cta = t.reshape(-1, wpc)
        print('%s per-warp busy us: min %.1f p10 %.1f mean %.1f p90 %.1f max %.1f | per-CTA max: min %.1f mean %.1f max %.1f | per-CTA mean: min %.1f max %.1f'
              % (nm, t.min(), np.percentile(t, 10), t.mean(), np.percentile(t, 90), t.max(), cta.max(1).min(), cta.max(1).mean(),
                 cta.max(1).max(), cta.mean(1).min(), cta.mean(1).max()))
        srt = np.argsort(cta.max(1))
        print('   slowest CTAs', srt[-6:], cta.max(1)[srt[-6:]].round(1), 'fastest', srt[:4], cta.max(1)[srt[:4]].round(1))
if hasattr(L, 'abipgpu_lp_spmv_prof'):
    pr = (C.c_ulonglong * 16)()
    L.abipgpu_lp_spmv_prof.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
    L.abipgpu_lp_spmv_prof(e.e, pr, 0)
    for b, nm in ((0, "A'"), (8, 'A ')):
        nch = max(pr[b + 4], 1)
        print('%s chunk loop, cycles per chunk per warp: wait %.0f  gather+mul %.0f  rowsum+epi %.0f  issue %.0f  (chunks %d)'
              % (nm, pr[b] / nch, pr[b + 1] / nch, pr[b + 2] / nch, pr[b + 3] / nch, pr[b + 4]))
