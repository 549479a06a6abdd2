"""Host-side mirror of the reference's user entry `[x, y, s, info] = abip(data, K, params)`
(scripts/matlab/abip.m:1-30, abip_lpsolve.m, abip_get_params.m) on top of the C ABI.

MATLAB is not available in this environment, so this Python layer plays the role of the MATLAB scripts +
mex gateway (src/abip-lp/mexfile/abip_mex.c:83-424): it marshals CSC + settings into ABIPData and calls
abip_gpu_main.  All numerical work happens in libabip_gpu.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import time
import numpy as np
import scipy.sparse as sp

from . import _capi

# vector / scalar ids of include/abip_gpu.h
VEC = dict(U=0, V=1, UT=2, UPREV=3, USUM=4, VSUM=5, UAVGC=6, VAVGC=7, UAVG=8, VAVG=9, H=10, G=11, M=12,
           BB_UPREV=13, BB_VPREV=14, BB_U=15, BB_V=16, BB_UNEXT=17, BB_VNEXT=18, BB_UT=19, BB_UTNEXT=20)
SC = dict(CG_ITS=0, CG_ITS2=1, CG_TOL=2, CG_RES=3, S_PR=4, W_AX=5, W_PR=6, BTY=7, UU_Y=8, S_DR=9, W_ATYS=10,
          W_DR=11, CTX=12, UU_X=13, VV=14, TAU=15, KAP=16, AVG_BASE=20, HAS_AVG=36, BB_UTUT=40, BB_UTV=41,
          BB_UU=42, BB_VV=43, BB_UV=44, MIN_XS=48, SUM_XS=49, VEC_NORM2=50)


def get_params() -> dict:
    """scripts/matlab/abip_get_params.m."""
    return dict(verbose=1, normalize=1, pcg=1, max_admm_iter=1000000, max_ipm_iter=500, timelimit=3600,
                tol=1e-3, solver=-1,
                lpalg=dict(restart_thresh=100000, restart_freq=1000, feasopt=0, scaling_method=1, half_update=0),
                qcpalg=dict(rho_primal=1.0, rho_dual=1e-6, admm_tol_factor=1.0))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_long))


class CscHolder:
    """Keeps the numpy buffers behind an ABIPMatrix alive."""

    def __init__(self, A):
        if sp.issparse(A):
            A = A.tocsc()
            A.sort_indices()
            m, n = A.shape
            Ap, Ai, Ax = A.indptr, A.indices, A.data
        else:
            m, n, Ap, Ai, Ax = A
        self.m, self.n = int(m), int(n)
        self.Ap = np.ascontiguousarray(Ap, dtype=np.int64)
        self.Ai = np.ascontiguousarray(Ai, dtype=np.int64)
        self.Ax = np.ascontiguousarray(Ax, dtype=np.float64)
        self.nnz = int(self.Ap[-1])
        self.c = _capi.ABIPMatrix(_fp(self.Ax), _ip(self.Ai), _ip(self.Ap), self.m, self.n)


def _lp_settings(params: dict | None, **raw):
    """abip_lpsolve.m:37-60 parameter translation (using the names the mex actually reads)."""
    p = get_params()
    if params:
        for k, v in params.items():
            if isinstance(v, dict) and isinstance(p.get(k), dict):
                p[k].update(v)
            else:
                p[k] = v
    lp = p["lpalg"]
    st = _capi.default_settings(
        verbose=int(p["verbose"]), normalize=int(p["normalize"]), max_admm_iters=int(p["max_admm_iter"]),
        max_ipm_iters=int(p["max_ipm_iter"]), max_time=float(p["timelimit"]), eps=float(p["tol"]),
        origin_rescale=int(lp["scaling_method"] == 3), pc_ruiz_rescale=int(lp["scaling_method"] == 1),
        qp_rescale=int(lp["scaling_method"] == 2), restart_thresh=int(lp["restart_thresh"]),
        restart_fre=int(lp["restart_freq"]), pfeasopt=int(lp["feasopt"]), half_update=int(lp["half_update"]))
    for k, v in raw.items():
        if not hasattr(st, k):
            raise KeyError(f"unknown ABIP setting {k!r}")
        setattr(st, k, v)
    return p, st


def lp_solve(A, b, c, params: dict | None = None, want_stats: bool = False, **raw_settings):
    """ABIP-LP on the GPU engine.  A: scipy sparse or (m, n, Ap, Ai, Ax) CSC.  Returns (x, y, s, info)."""
    L = _capi.lib()
    p, st = _lp_settings(params, **raw_settings)
    if not p.get("pcg", 1):
        raise ValueError("abip_b200 implements the indirect (pcg=1) path only; the direct LDL' path of the "
                         "reference is out of scope")
    H = CscHolder(A)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    if b.shape != (H.m,) or c.shape != (H.n,):
        raise ValueError("dimension mismatch between A, b and c")
    d = _capi.ABIPData(H.m, H.n, C.pointer(H.c), _fp(b), _fp(c), float(H.nnz) / (float(H.m) * float(H.n)),
                       C.pointer(st))
    sol = _capi.ABIPSolution()
    info = _capi.ABIPInfo()
    t0 = time.perf_counter()
    w = L.abip_gpu_init(C.byref(d), C.byref(info))
    stats = _capi.ABIPGpuStats()
    if w:
        L.abip_gpu_solve(w, C.byref(d), C.byref(sol), C.byref(info))
        L.abip_gpu_get_stats(w, C.byref(stats))
        L.abip_gpu_finish(w)
    else:
        info.status_val = -4
        info.status = b"Failure"
    wall = time.perf_counter() - t0
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    out = {}
    for name, ln in (("x", H.n), ("y", H.m), ("s", H.n)):
        ptr = getattr(sol, name)
        out[name] = np.ctypeslib.as_array(ptr, shape=(ln,)).copy() if ptr else np.full(ln, np.nan)
        if ptr:
            libc.free(C.cast(ptr, C.c_void_p))
    res = dict(status=info.status.decode(), status_val=int(info.status_val), ipm_iter=int(info.ipm_iter),
               admm_iter=int(info.admm_iter), pres=info.res_pri, dres=info.res_dual, gap=info.rel_gap,
               pobj=info.pobj, dobj=info.dobj, res_infeas=info.res_infeas, res_unbdd=info.res_unbdd,
               setup_time_ms=info.setup_time, solve_time_ms=info.solve_time, time=wall, solver="abip-lp-b200")
    if want_stats:
        res["stats"] = {f: getattr(stats, f) for f, _ in stats._fields_}
    return out["x"], out["y"], out["s"], res


class LpSolver:
    """ABIP(init) / ABIP(solve) / ABIP(finish) life cycle (src/abip-lp/include/abip.h:116-124): A is scaled and
    uploaded once, solve() may be called repeatedly with the data resident in HBM."""

    def __init__(self, A, params: dict | None = None, **raw_settings):
        self.L = _capi.lib()
        self.p, self.st = _lp_settings(params, **raw_settings)
        self.H = CscHolder(A)
        self.info = _capi.ABIPInfo()
        self._b = np.zeros(self.H.m)
        self._c = np.zeros(self.H.n)
        self.d = _capi.ABIPData(self.H.m, self.H.n, C.pointer(self.H.c), _fp(self._b), _fp(self._c),
                                float(self.H.nnz) / (float(self.H.m) * float(self.H.n)), C.pointer(self.st))
        self.w = self.L.abip_gpu_init(C.byref(self.d), C.byref(self.info))
        if not self.w:
            raise RuntimeError("abip_gpu_init failed")
        self.setup_time_ms = self.info.setup_time
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [C.c_void_p]

    def solve(self, b, c):
        self._b[:] = b
        self._c[:] = c
        sol = _capi.ABIPSolution()
        info = _capi.ABIPInfo()
        self.L.abip_gpu_solve(self.w, C.byref(self.d), C.byref(sol), C.byref(info))
        stats = _capi.ABIPGpuStats()
        self.L.abip_gpu_get_stats(self.w, C.byref(stats))
        out = {}
        for name, ln in (("x", self.H.n), ("y", self.H.m), ("s", self.H.n)):
            ptr = getattr(sol, name)
            out[name] = np.ctypeslib.as_array(ptr, shape=(ln,)).copy() if ptr else np.full(ln, np.nan)
            if ptr:
                self._libc.free(C.cast(ptr, C.c_void_p))
        res = dict(status=info.status.decode(), status_val=int(info.status_val), ipm_iter=int(info.ipm_iter),
                   admm_iter=int(info.admm_iter), pres=info.res_pri, dres=info.res_dual, gap=info.rel_gap,
                   pobj=info.pobj, dobj=info.dobj, solve_time_ms=info.solve_time,
                   stats={f: getattr(stats, f) for f, _ in stats._fields_})
        return out["x"], out["y"], out["s"], res

    def close(self):
        if self.w:
            self.L.abip_gpu_finish(self.w)
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lp_solve_batch(problems, params: dict | None = None, concurrency: int = 192, ctas_per_problem: int = 1,
                   **raw_settings):
    """Batch of independent LPs on the current GPU (abip_gpu_batch_main).  problems: iterable of objects with
    csc()/m/n/b/c (abip_b200.problems.LPProblem) or (A, b, c) tuples.  Returns a list of (x, y, s, info).
    ctas_per_problem <= 1 (default): lock-step mode, `concurrency` problems in flight, one CTA each, one batched launch
    per step; >= 2: one persistent grid of that many CTAs per problem, `concurrency` host threads."""
    L = _capi.lib()
    p, st0 = _lp_settings(params, **raw_settings)
    keep, datas = [], []
    for pr in problems:
        A, b, c = (pr.csc(), pr.b, pr.c) if hasattr(pr, "csc") else pr
        H = CscHolder(A)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        st = _capi.ABIPSettings.from_buffer_copy(st0)
        d = _capi.ABIPData(H.m, H.n, C.pointer(H.c), _fp(b), _fp(c), float(H.nnz) / (float(H.m) * float(H.n)),
                           C.pointer(st))
        keep.append((H, b, c, st))
        datas.append(d)
    n = len(datas)
    ptrs = (C.POINTER(_capi.ABIPData) * n)(*[C.pointer(d) for d in datas])
    sols = (_capi.ABIPSolution * n)()
    infos = (_capi.ABIPInfo * n)()
    t0 = time.perf_counter()
    nfail = L.abip_gpu_batch_main(ptrs, sols, infos, n, int(concurrency), int(ctas_per_problem))
    wall = time.perf_counter() - t0
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    out = []
    for i in range(n):
        H = keep[i][0]
        vec = {}
        for name, ln in (("x", H.n), ("y", H.m), ("s", H.n)):
            ptr = getattr(sols[i], name)
            vec[name] = np.ctypeslib.as_array(ptr, shape=(ln,)).copy() if ptr else np.full(ln, np.nan)
            if ptr:
                libc.free(C.cast(ptr, C.c_void_p))
        inf = infos[i]
        out.append((vec["x"], vec["y"], vec["s"],
                    dict(status=inf.status.decode(), status_val=int(inf.status_val), ipm_iter=int(inf.ipm_iter),
                         admm_iter=int(inf.admm_iter), pres=inf.res_pri, dres=inf.res_dual, gap=inf.rel_gap,
                         pobj=inf.pobj, dobj=inf.dobj, solve_time_ms=inf.solve_time, batch_wall_s=wall,
                         batch_failed=int(nfail))))
    return out


def abip(data: dict, K: dict, params: dict | None = None):
    """[x, y, s, info] = abip(data, K, params) -- scripts/matlab/abip.m:1-30.

    data: {'A': sparse m x n, 'b': [m], 'c': [n]} (+ 'Q' for QCP); K: cone dict ('l' for LP; 'f','z','q','rq'
    select the QCP solver)."""
    params = params or get_params()
    if any(k in K for k in ("f", "q", "rq", "z")) or params.get("solver", -1) == 1 or data.get("Q") is not None:
        from . import qcp  # noqa: WPS433  (QCP engine)
        return qcp.qcp_solve(data, K, params)
    if "l" not in K:
        raise ValueError("Invalid conic format for LP")  # abip_lpsolve.m:8-11
    x, y, s, info = lp_solve(data["A"], data["b"], data["c"], params)
    info["pobj"] = float(np.dot(data["c"], x))  # abip_lpsolve.m:27-28
    info["dobj"] = float(np.dot(data["b"], y))
    return x, y, s, info


class LinSysPlugin:
    """The reference's linsys 'priv' plugin through the C ABI (init / solve / accum / free), host pointers.
    Mirrors how src/abip.c drives linsys/indirect.c; used by the parity tests."""

    def __init__(self, A, **raw_settings):
        self.L = _capi.lib()
        self.H = CscHolder(A)
        self.st = _capi.default_settings(**raw_settings)
        self.p = self.L.abip_init_lin_sys_work(C.byref(self.H.c), C.byref(self.st))
        if not self.p:
            raise RuntimeError("abip_init_lin_sys_work failed (no usable CUDA device?)")

    def solve(self, b, s=None, it=0):
        b = np.ascontiguousarray(b, dtype=np.float64).copy()
        sp_ = None if s is None else np.ascontiguousarray(s, dtype=np.float64)
        rc = self.L.abip_solve_lin_sys(C.byref(self.H.c), C.byref(self.st), self.p, _fp(b),
                                       _fp(sp_) if sp_ is not None else None, int(it))
        if rc < 0:
            raise RuntimeError("abip_solve_lin_sys failed")
        return b

    def accum_by_A(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64).copy()
        self.L.abip_accum_by_A(C.byref(self.H.c), self.p, _fp(x), _fp(y))
        return y

    def accum_by_Atrans(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64).copy()
        self.L.abip_accum_by_Atrans(C.byref(self.H.c), self.p, _fp(x), _fp(y))
        return y

    def close(self):
        if self.p:
            self.L.abip_free_lin_sys_work(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LpEngine:
    """Device-resident step functions (group (3) of include/abip_gpu.h) for step-level parity tests."""

    def __init__(self, A_scaled, device: int = 0, **raw_settings):
        self.L = _capi.lib()
        self.H = CscHolder(A_scaled)
        self.m, self.n = self.H.m, self.H.n
        self.l = self.m + self.n + 1
        self.st = _capi.default_settings(**raw_settings)
        self.e = self.L.abipgpu_lp_create(self.m, self.n, _ip(self.H.Ap), _ip(self.H.Ai), _fp(self.H.Ax),
                                          C.byref(self.st), device)
        if not self.e:
            raise RuntimeError("abipgpu_lp_create failed (no usable CUDA device?)")

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed")

    def set_problem(self, b, c, D=None, E=None):
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        Dp = _fp(np.ascontiguousarray(D, dtype=np.float64)) if D is not None else None
        Ep = _fp(np.ascontiguousarray(E, dtype=np.float64)) if E is not None else None
        self._keep = (b, c, D, E)
        self._ck(self.L.abipgpu_lp_set_problem(self.e, _fp(b), _fp(c), Dp, Ep), "set_problem")

    def g_th(self):
        return float(self.L.abipgpu_lp_g_th(self.e))

    def cold_start(self, mu=1.0, beta=1.0):
        self._ck(self.L.abipgpu_lp_cold_start(self.e, mu, beta), "cold_start")

    def outer_prologue(self, avg_criterion=0):
        self._ck(self.L.abipgpu_lp_outer_prologue(self.e, int(avg_criterion)), "outer_prologue")

    def admm_iter(self, j, k, mu, beta):
        sc = np.zeros(_capi.SC_COUNT)
        self._ck(self.L.abipgpu_lp_admm_iter(self.e, int(j), int(k), float(mu), float(beta), _fp(sc)), "admm_iter")
        return sc

    def mu_stats(self, avg_criterion=0):
        sc = np.zeros(_capi.SC_COUNT)
        self._ck(self.L.abipgpu_lp_mu_stats(self.e, int(avg_criterion), _fp(sc)), "mu_stats")
        return sc

    def reinit(self, indx, sigma, avg_criterion=0):
        self._ck(self.L.abipgpu_lp_reinit(self.e, int(indx), float(sigma), int(avg_criterion)), "reinit")

    def bb_begin(self):
        self._ck(self.L.abipgpu_lp_bb_begin(self.e), "bb_begin")

    def bb_round(self, carry, k, mu, beta_prev):
        sc = np.zeros(_capi.SC_COUNT)
        self._ck(self.L.abipgpu_lp_bb_round(self.e, int(carry), int(k), float(mu), float(beta_prev), _fp(sc)),
                 "bb_round")
        return sc

    def solve_vec(self, rhs_id, warm_id, it):
        sc = np.zeros(_capi.SC_COUNT)
        self._ck(self.L.abipgpu_lp_solve_vec(self.e, int(rhs_id), int(warm_id), int(it), _fp(sc)), "solve_vec")
        return sc

    def get(self, name, length=None):
        vid = VEC[name]
        n = length if length is not None else (self.m + self.n if name in ("H", "G") else
                                               self.m if name == "M" else self.l)
        out = np.zeros(n)
        self._ck(self.L.abipgpu_lp_get_vec(self.e, vid, _fp(out), n), "get_vec")
        return out

    def set(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        self._ck(self.L.abipgpu_lp_set_vec(self.e, VEC[name], _fp(arr), arr.size), "set_vec")

    def spmv(self, x, trans=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n if trans else self.m)
        self._ck(self.L.abipgpu_lp_spmv(self.e, int(trans), _fp(x), _fp(y)), "spmv")
        return y

    def describe(self):
        buf = C.create_string_buffer(1280)
        self.L.abipgpu_lp_describe(self.e, buf, 1280)
        return buf.value.decode()

    def close(self):
        if self.e:
            self.L.abipgpu_lp_destroy(self.e)
            self.e = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
