"""TEST INFRASTRUCTURE -- ctypes driver for the *unmodified* reference ABIP-LP solver compiled by
oracle/Makefile into oracle/_ref/ (never imported by the product package).

Struct layouts follow /root/reference/src/abip-lp/include/abip.h:23-105 and linsys/amatrix.h:10-17 with
-DDLONG (abip_int = long).  Driver obligations are those of the mex gateway
(mexfile/abip_mex.c:181,320-341,362): defaults, max_time, pfeasopt, warm_start, sp.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

c_int = C.c_long     # abip_int under -DDLONG
c_float = C.c_double  # abip_float


class ABIPMatrix(C.Structure):
    _fields_ = [("x", C.POINTER(c_float)), ("i", C.POINTER(c_int)), ("p", C.POINTER(c_int)),
                ("m", c_int), ("n", c_int)]


class ABIPSettings(C.Structure):
    _fields_ = [("normalize", c_int), ("pfeasopt", c_int), ("scale", c_float), ("rho_y", c_float),
                ("sparsity_ratio", c_float), ("max_ipm_iters", c_int), ("max_admm_iters", c_int),
                ("max_time", c_float), ("eps", c_float), ("alpha", c_float), ("cg_rate", c_float),
                ("adaptive", c_int), ("eps_cor", c_float), ("eps_pen", c_float),
                ("dynamic_sigma", c_float), ("dynamic_x", c_float), ("dynamic_eta", c_float),
                ("restart_fre", c_int), ("restart_thresh", c_int), ("verbose", c_int),
                ("warm_start", c_int), ("adaptive_lookback", c_int), ("origin_rescale", c_int),
                ("pc_ruiz_rescale", c_int), ("qp_rescale", c_int), ("ruiz_iter", c_int),
                ("hybrid_mu", c_int), ("hybrid_thresh", c_float), ("dynamic_sigma_second", c_float),
                ("half_update", c_int), ("avg_criterion", c_int)]


class ABIPData(C.Structure):
    _fields_ = [("m", c_int), ("n", c_int), ("A", C.POINTER(ABIPMatrix)), ("b", C.POINTER(c_float)),
                ("c", C.POINTER(c_float)), ("sp", c_float), ("stgs", C.POINTER(ABIPSettings))]


class ABIPSolution(C.Structure):
    _fields_ = [("x", C.POINTER(c_float)), ("y", C.POINTER(c_float)), ("s", C.POINTER(c_float))]


class ABIPInfo(C.Structure):
    _fields_ = [("status", C.c_char * 32), ("status_val", c_int), ("ipm_iter", c_int),
                ("admm_iter", c_int), ("pobj", c_float), ("dobj", c_float), ("res_pri", c_float),
                ("res_dual", c_float), ("rel_gap", c_float), ("res_infeas", c_float),
                ("res_unbdd", c_float), ("setup_time", c_float), ("solve_time", c_float)]


class ABIPScaling(C.Structure):
    _fields_ = [("D", C.POINTER(c_float)), ("E", C.POINTER(c_float)), ("mean_norm_row_A", c_float),
                ("mean_norm_col_A", c_float)]


def available(which: str = "indirect") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libabip_{which}_ref.so"))


_libs: dict = {}


def load(which: str = "indirect"):
    if which not in _libs:
        lib = C.CDLL(os.path.join(REF_DIR, f"libabip_{which}_ref.so"), mode=os.RTLD_LOCAL)
        lib.abip_main.restype = c_int
        lib.abip_main.argtypes = [C.POINTER(ABIPData), C.POINTER(ABIPSolution), C.POINTER(ABIPInfo)]
        lib.abip_set_default_settings.argtypes = [C.POINTER(ABIPData)]
        lib.abip_set_default_settings.restype = None
        _libs[which] = lib
    return _libs[which]


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefMatrix:
    """Keeps numpy buffers alive behind an ABIPMatrix."""

    def __init__(self, m, n, Ap, Ai, Ax):
        self.Ap = np.ascontiguousarray(Ap, dtype=np.int64)
        self.Ai = np.ascontiguousarray(Ai, dtype=np.int64)
        self.Ax = np.ascontiguousarray(Ax, dtype=np.float64)
        self.c = ABIPMatrix(_ptr(self.Ax, c_float), _ptr(self.Ai, c_int), _ptr(self.Ap, c_int), m, n)


def default_settings(lib) -> ABIPSettings:
    st = ABIPSettings()
    d = ABIPData()
    d.stgs = C.pointer(st)
    lib.abip_set_default_settings(C.byref(d))
    st.max_time = 3600.0   # abip_mex.c:320-326
    st.pfeasopt = 0        # abip_mex.c:335-341
    st.warm_start = 0
    return st


def solve(prob, which: str = "indirect", **overrides):
    """Run reference abip_main on an LPProblem; returns dict(x,y,s,info...)."""
    lib = load(which)
    st = default_settings(lib)
    st.verbose = 0
    for k, v in overrides.items():
        if not hasattr(st, k):
            raise KeyError(k)
        setattr(st, k, v)
    A = RefMatrix(prob.m, prob.n, prob.Ap, prob.Ai, prob.Ax)
    b = np.ascontiguousarray(prob.b, dtype=np.float64).copy()
    c = np.ascontiguousarray(prob.c, dtype=np.float64).copy()
    d = ABIPData(prob.m, prob.n, C.pointer(A.c), _ptr(b, c_float), _ptr(c, c_float),
                 float(prob.nnz) / (float(prob.m) * float(prob.n)), C.pointer(st))
    sol = ABIPSolution()
    info = ABIPInfo()
    status = lib.abip_main(C.byref(d), C.byref(sol), C.byref(info))
    out = {"status_val": int(status), "status": info.status.decode(), "ipm_iter": int(info.ipm_iter),
           "admm_iter": int(info.admm_iter), "pobj": info.pobj, "dobj": info.dobj,
           "res_pri": info.res_pri, "res_dual": info.res_dual, "rel_gap": info.rel_gap,
           "res_infeas": info.res_infeas, "res_unbdd": info.res_unbdd,
           "setup_time_ms": info.setup_time, "solve_time_ms": info.solve_time}
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for name, ln in (("x", prob.n), ("y", prob.m), ("s", prob.n)):
        p = getattr(sol, name)
        out[name] = np.ctypeslib.as_array(p, shape=(ln,)).copy() if p else None
        if p:
            libc.free(C.cast(p, C.c_void_p))
    return out
