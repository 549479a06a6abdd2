"""ABIP-QCP on the GPU engine: host-side mirror of abip_qcpsolve.m / the abip_qcp mex gateway
(scripts/matlab/abip_qcpsolve.m, src/abip-qcp/mex/abip_qcp_mex.c:109-525) over the C ABI (abip_qcp_gpu)."""
from __future__ import annotations

import ctypes as C
import time
import numpy as np
import scipy.sparse as sp

from . import _capi


class QcpMatrix(C.Structure):
    _fields_ = [("x", C.POINTER(C.c_double)), ("i", C.POINTER(C.c_int)), ("p", C.POINTER(C.c_int)), ("m", C.c_int),
                ("n", C.c_int)]


class QcpCone(C.Structure):
    _fields_ = [("q", C.POINTER(C.c_int)), ("qsize", C.c_int), ("rq", C.POINTER(C.c_int)), ("rqsize", C.c_int),
                ("f", C.c_int), ("z", C.c_int), ("l", C.c_int)]


class QcpSettings(C.Structure):
    _fields_ = [("normalize", C.c_int), ("scale_E", C.c_int), ("scale_bc", C.c_int), ("scale", C.c_double),
                ("rho_x", C.c_double), ("rho_y", C.c_double), ("rho_tau", C.c_double), ("max_ipm_iters", C.c_int),
                ("max_admm_iters", C.c_int), ("eps", C.c_double), ("eps_p", C.c_double), ("eps_d", C.c_double),
                ("eps_g", C.c_double), ("eps_inf", C.c_double), ("eps_unb", C.c_double), ("err_dif", C.c_double),
                ("alpha", C.c_double), ("cg_rate", C.c_double), ("use_indirect", C.c_int),
                ("inner_check_period", C.c_int), ("outer_check_period", C.c_int), ("verbose", C.c_int),
                ("linsys_solver", C.c_int), ("prob_type", C.c_int), ("time_limit", C.c_double), ("psi", C.c_double),
                ("origin_scaling", C.c_int), ("ruiz_scaling", C.c_int), ("pc_scaling", C.c_int)]


class QcpData(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("A", C.POINTER(QcpMatrix)), ("Q", C.POINTER(QcpMatrix)),
                ("b", C.POINTER(C.c_double)), ("c", C.POINTER(C.c_double)), ("lambda_", C.c_double),
                ("stgs", C.POINTER(QcpSettings))]


class QcpInfo(C.Structure):
    _fields_ = [("status", C.c_char * 32), ("status_val", C.c_int), ("ipm_iter", C.c_int), ("admm_iter", C.c_int),
                ("pobj", C.c_double), ("dobj", C.c_double), ("res_pri", C.c_double), ("res_dual", C.c_double),
                ("rel_gap", C.c_double), ("res_infeas", C.c_double), ("res_unbdd", C.c_double),
                ("setup_time", C.c_double), ("solve_time", C.c_double), ("avg_linsys_time", C.c_double),
                ("avg_cg_iters", C.c_double)]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _mat(M):
    M = sp.csc_matrix(M)
    M.sort_indices()
    keep = (np.ascontiguousarray(M.data, dtype=np.float64), np.ascontiguousarray(M.indices, dtype=np.int32),
            np.ascontiguousarray(M.indptr, dtype=np.int32))
    return QcpMatrix(_dp(keep[0]), _ip(keep[1]), _ip(keep[2]), M.shape[0], M.shape[1]), keep


def _bind():
    L = _capi.lib()
    if not getattr(L, "_qcp_bound", False):
        L.abip_qcp_gpu.restype = C.c_int
        L.abip_qcp_gpu.argtypes = [C.POINTER(QcpData), C.POINTER(_capi.ABIPSolution), C.POINTER(QcpInfo),
                                   C.POINTER(QcpCone)]
        L.abip_qcp_gpu_set_default_settings.argtypes = [C.POINTER(QcpData)]
        L.abip_qcp_gpu_set_default_settings.restype = None
        L.abip_qcp_gpu_last_counters.argtypes = [C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long),
                                                 C.POINTER(C.c_double)]
        L.abip_qcp_gpu_last_counters.restype = None
        L.abip_qcp_scale_data.restype = None
        L.abip_qcp_scale_data.argtypes = [C.POINTER(QcpMatrix), C.POINTER(QcpMatrix), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(QcpCone), C.POINTER(QcpSettings),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double)]
        L._qcp_bound = True
    return L


def make_cone(K: dict):
    q = np.ascontiguousarray(K.get("q", []) or [], dtype=np.int32)
    rq = np.ascontiguousarray(K.get("rq", []) or [], dtype=np.int32)
    cone = QcpCone(_ip(q) if q.size else None, int(q.size), _ip(rq) if rq.size else None, int(rq.size),
                   int(K.get("f", 0) or 0), int(K.get("z", 0) or 0), int(K.get("l", 0) or 0))
    return cone, (q, rq)


def default_settings(**over) -> QcpSettings:
    L = _bind()
    st = QcpSettings()
    d = QcpData()
    d.stgs = C.pointer(st)
    L.abip_qcp_gpu_set_default_settings(C.byref(d))
    for k, v in over.items():
        if not hasattr(st, k):
            raise KeyError(f"unknown ABIP-QCP setting {k!r}")
        setattr(st, k, v)
    return st


def qcp_solve_raw(A, Q, b, c, K: dict, **settings):
    """abip(d, sol, info, K) of the reference (source/abip.c:1335) on the GPU engine."""
    L = _bind()
    st = default_settings(**settings)
    Am, keepA = _mat(A)
    Qm, keepQ = (None, None) if Q is None else _mat(Q)
    m, n = Am.m, Am.n
    b = np.ascontiguousarray(b, dtype=np.float64).copy()
    c = np.ascontiguousarray(c, dtype=np.float64).copy()
    cone, keepK = make_cone(K)
    d = QcpData(m, n, C.pointer(Am), C.pointer(Qm) if Qm is not None else None, _dp(b), _dp(c), 0.0, C.pointer(st))
    sol = _capi.ABIPSolution()
    info = QcpInfo()
    t0 = time.perf_counter()
    L.abip_qcp_gpu(C.byref(d), C.byref(sol), C.byref(info), C.byref(cone))
    wall = time.perf_counter() - t0
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    out = {}
    for name, ln in (("x", n), ("y", m), ("s", n)):
        ptr = getattr(sol, name)
        out[name] = np.ctypeslib.as_array(ptr, shape=(ln,)).copy() if ptr else np.full(ln, np.nan)
        if ptr:
            libc.free(C.cast(ptr, C.c_void_p))
    cnt = [C.c_long(), C.c_long(), C.c_long(), C.c_double()]
    L.abip_qcp_gpu_last_counters(*[C.byref(x) for x in cnt])
    res = dict(status=info.status.decode(), status_val=int(info.status_val), ipm_iter=int(info.ipm_iter),
               admm_iter=int(info.admm_iter), pres=info.res_pri, dres=info.res_dual, gap=info.rel_gap, pobj=info.pobj,
               dobj=info.dobj, setup_time_ms=info.setup_time, solve_time_ms=info.solve_time, time=wall,
               avg_cg_iters=info.avg_cg_iters, n_iter=cnt[0].value, n_cg=cnt[1].value, n_inner=cnt[2].value,
               kernel_ms=cnt[3].value, solver="abip-qcp-b200")
    return out["x"], out["y"], out["s"], res


def qcp_solve(data: dict, K: dict, params: dict | None = None):
    """abip_qcpsolve.m: parameter translation (:27-55) + call."""
    from .api import get_params
    p = get_params()
    if params:
        for k, v in params.items():
            if isinstance(v, dict) and isinstance(p.get(k), dict):
                p[k].update(v)
            else:
                p[k] = v
    qa = p["qcpalg"]
    return qcp_solve_raw(data["A"], data.get("Q"), data["b"], data["c"], K, verbose=int(p["verbose"]),
                         normalize=int(p["normalize"]), max_admm_iters=int(p["max_admm_iter"]),
                         max_ipm_iters=int(p["max_ipm_iter"]), time_limit=float(p["timelimit"]),
                         eps_p=float(p["tol"]), eps_d=float(p["tol"]), eps_g=float(p["tol"]),
                         rho_x=float(qa["rho_primal"]), rho_y=float(qa["rho_dual"]), psi=float(qa["admm_tol_factor"]))


QSC = dict(CG_ITS=0, INNER_ITS=1, TAU_T=2, CG_RES=3, S_DIFF=4, S_QU=5, S_VO=6, UMU=7, YB=8, XC=9, XQX=10, AXD2=11,
           QXE2=12, ATYS_E2=13, AXB_INF=14, AXB_D_INF=15, AX_D_INF=16, RESD_INF=17, RESD_E_INF=18, QX_E_INF=19, TAU=20,
           VO_TAU=21, A_COEF=22)


class QcpEngine:
    """Device-resident QCP step functions (abipgpu_qcp_*) on already scaled data; for step-level parity tests."""

    def __init__(self, A, Q, b, c, D, E, K, rho_x=1.0, rho_y=1e-6, rho_tau=1.0, alpha=1.8, rtol=1e-8, device=0):
        L = _bind()
        self.L = L
        L.abipgpu_qcp_create.restype = C.c_void_p
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.abipgpu_qcp_create.argtypes = [C.c_int, C.c_int, ip, ip, dp, ip, ip, dp, dp, dp, dp, dp, ip, C.c_int, ip,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.c_double, C.c_double, C.c_int]
        L.abipgpu_qcp_destroy.argtypes = [C.c_void_p]
        L.abipgpu_qcp_destroy.restype = None
        L.abipgpu_qcp_iter.argtypes = [C.c_void_p, C.c_long, C.c_double, C.c_double, dp]
        L.abipgpu_qcp_solve_vec.argtypes = [C.c_void_p, dp, dp, C.c_double, dp]
        L.abipgpu_qcp_get_vec.argtypes = [C.c_void_p, C.c_int, dp, C.c_long]
        L.abipgpu_qcp_set_vec.argtypes = [C.c_void_p, C.c_int, dp, C.c_long]
        L.abipgpu_qcp_a_coef.restype = C.c_double
        L.abipgpu_qcp_a_coef.argtypes = [C.c_void_p]
        Am, self._kA = _mat(A)
        self.m, self.n = Am.m, Am.n
        self.l = self.m + self.n + 1
        if Q is not None:
            Qm, self._kQ = _mat(Q)
            qa = (Qm.p, Qm.i, Qm.x)
        else:
            qa = (None, None, None)
        cone, self._kK = make_cone(K)
        self._v = [np.ascontiguousarray(x, dtype=np.float64) for x in (b, c, D, E)]
        self.e = L.abipgpu_qcp_create(self.m, self.n, Am.p, Am.i, Am.x, qa[0], qa[1], qa[2], _dp(self._v[0]),
                                      _dp(self._v[1]), _dp(self._v[2]), _dp(self._v[3]), cone.q, cone.qsize, cone.rq,
                                      cone.rqsize, cone.f, cone.z, cone.l, rho_x, rho_y, rho_tau, alpha, rtol, device)
        if not self.e:
            raise RuntimeError("abipgpu_qcp_create failed (no usable CUDA device?)")

    def a_coef(self):
        return float(self.L.abipgpu_qcp_a_coef(self.e))

    def iter(self, k, mu, beta):
        sc = np.zeros(32)
        if self.L.abipgpu_qcp_iter(self.e, int(k), float(mu), float(beta), _dp(sc)) != 0:
            raise RuntimeError("abipgpu_qcp_iter failed")
        return sc

    def solve_vec(self, vec, warm=None, rtol=1e-10):
        vec = np.ascontiguousarray(vec, dtype=np.float64).copy()
        w = None if warm is None else np.ascontiguousarray(warm, dtype=np.float64)
        sc = np.zeros(32)
        if self.L.abipgpu_qcp_solve_vec(self.e, _dp(vec), _dp(w) if w is not None else None, float(rtol), _dp(sc)) != 0:
            raise RuntimeError("abipgpu_qcp_solve_vec failed")
        return vec, sc

    def solve_nspace(self, vec, warm_x=None, rtol=1e-10, max_iter=None):
        """solve_qcp_linsys through the reference's n-space qcp_pcg (abipgpu_qcp_solve_nspace)"""
        vec = np.ascontiguousarray(vec, dtype=np.float64).copy()
        w = None if warm_x is None else np.ascontiguousarray(warm_x, dtype=np.float64)
        sc = np.zeros(32)
        fn = self.L.abipgpu_qcp_solve_nspace
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_long, C.POINTER(C.c_double)]
        fn.restype = C.c_int
        if fn(self.e, _dp(vec), _dp(w) if w is not None else None, float(rtol), int(max_iter or 4 * self.n + 50), _dp(sc)) != 0:
            raise RuntimeError("abipgpu_qcp_solve_nspace failed")
        return vec, sc

    def get(self, name):
        out = np.zeros(self.l if name != "r" else self.m + self.n)
        if self.L.abipgpu_qcp_get_vec(self.e, {"u": 0, "v": 1, "ut": 2, "r": 3}[name], _dp(out), out.size) != 0:
            raise RuntimeError("get_vec failed")
        return out

    def set(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        if self.L.abipgpu_qcp_set_vec(self.e, {"u": 0, "v": 1}[name], _dp(arr), arr.size) != 0:
            raise RuntimeError("set_vec failed")

    def close(self):
        if self.e:
            self.L.abipgpu_qcp_destroy(self.e)
            self.e = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
