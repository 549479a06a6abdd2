"""TEST INFRASTRUCTURE -- ctypes driver for the *unmodified* reference ABIP-QCP solver compiled by oracle/Makefile
(with stub MKL headers) into oracle/_ref/libabip_qcp_ref.so.  Struct layouts follow
/root/reference/src/abip-qcp/include/abip.h:67-165 (abip_int = int: make_abip_qcp.m does not set DLONG).
Only linsys_solver = 1 (vendored QDLDL) is usable: the shipped pcg dispatch (linsys.c:1158-1165) is broken
(SURVEY.md section 8c)."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libabip_qcp_ref.so")
c_int = C.c_int
c_float = C.c_double


class ABIPMatrix(C.Structure):
    _fields_ = [("x", C.POINTER(c_float)), ("i", C.POINTER(c_int)), ("p", C.POINTER(c_int)), ("m", c_int), ("n", c_int)]


class ABIPSettings(C.Structure):
    _fields_ = [("normalize", c_int), ("scale_E", c_int), ("scale_bc", c_int), ("scale", c_float), ("rho_x", c_float),
                ("rho_y", c_float), ("rho_tau", c_float), ("max_ipm_iters", c_int), ("max_admm_iters", c_int),
                ("eps", c_float), ("eps_p", c_float), ("eps_d", c_float), ("eps_g", c_float), ("eps_inf", c_float),
                ("eps_unb", c_float), ("err_dif", c_float), ("alpha", c_float), ("cg_rate", c_float),
                ("use_indirect", c_int), ("inner_check_period", c_int), ("outer_check_period", c_int),
                ("verbose", c_int), ("linsys_solver", c_int), ("prob_type", c_int), ("time_limit", c_float),
                ("psi", c_float), ("origin_scaling", c_int), ("ruiz_scaling", c_int), ("pc_scaling", c_int)]


class ABIPData(C.Structure):
    _fields_ = [("m", c_int), ("n", c_int), ("A", C.POINTER(ABIPMatrix)), ("Q", C.POINTER(ABIPMatrix)),
                ("b", C.POINTER(c_float)), ("c", C.POINTER(c_float)), ("lambda_", c_float),
                ("stgs", C.POINTER(ABIPSettings))]


class ABIPCone(C.Structure):
    _fields_ = [("q", C.POINTER(c_int)), ("qsize", c_int), ("rq", C.POINTER(c_int)), ("rqsize", c_int),
                ("f", c_int), ("z", c_int), ("l", c_int)]


class ABIPSolution(C.Structure):
    _fields_ = [("x", C.POINTER(c_float)), ("y", C.POINTER(c_float)), ("s", C.POINTER(c_float))]


class ABIPInfo(C.Structure):
    _fields_ = [("status", C.c_char * 32), ("status_val", c_int), ("ipm_iter", c_int), ("admm_iter", c_int),
                ("pobj", c_float), ("dobj", c_float), ("res_pri", c_float), ("res_dual", c_float),
                ("rel_gap", c_float), ("res_infeas", c_float), ("res_unbdd", c_float), ("setup_time", c_float),
                ("solve_time", c_float), ("avg_linsys_time", c_float), ("avg_cg_iters", c_float)]


def available() -> bool:
    return os.path.exists(LIB)


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB, mode=os.RTLD_LOCAL)
        _lib.abip.restype = c_int
        _lib.abip.argtypes = [C.POINTER(ABIPData), C.POINTER(ABIPSolution), C.POINTER(ABIPInfo), C.POINTER(ABIPCone)]
        _lib.abip_set_default_settings.argtypes = [C.POINTER(ABIPData)]
        _lib.abip_set_default_settings.restype = None
    return _lib


def _mat(M):
    M = M.tocsc()
    M.sort_indices()
    keep = (np.ascontiguousarray(M.data, dtype=np.float64), np.ascontiguousarray(M.indices, dtype=np.int32),
            np.ascontiguousarray(M.indptr, dtype=np.int32))
    s = ABIPMatrix(keep[0].ctypes.data_as(C.POINTER(c_float)), keep[1].ctypes.data_as(C.POINTER(c_int)),
                   keep[2].ctypes.data_as(C.POINTER(c_int)), M.shape[0], M.shape[1])
    return s, keep


def solve(prob, **overrides):
    """prob: abip_b200.problems.QCPProblem.  Runs the reference abip() with linsys_solver = 1."""
    import io
    lib = load()
    A, keepA = _mat(prob.A)
    Q, keepQ = (None, None)
    if prob.Q is not None:
        Q, keepQ = _mat(prob.Q)
    b = np.ascontiguousarray(prob.b, dtype=np.float64).copy()
    c = np.ascontiguousarray(prob.c, dtype=np.float64).copy()
    st = ABIPSettings()
    d = ABIPData(prob.m, prob.n, C.pointer(A), C.pointer(Q) if Q is not None else None,
                 b.ctypes.data_as(C.POINTER(c_float)), c.ctypes.data_as(C.POINTER(c_float)), 0.0, C.pointer(st))
    lib.abip_set_default_settings(C.byref(d))
    st.linsys_solver = 1
    st.prob_type = 2      # mex/abip_qcp_mex.c:436 (enum QCP)
    st.verbose = 0
    st.time_limit = 600.0
    for k, v in overrides.items():
        if not hasattr(st, k):
            raise KeyError(k)
        setattr(st, k, v)
    q = np.ascontiguousarray(prob.K.get("q", []), dtype=np.int32)
    rq = np.ascontiguousarray(prob.K.get("rq", []), dtype=np.int32)
    K = ABIPCone(q.ctypes.data_as(C.POINTER(c_int)) if q.size else None, int(q.size),
                 rq.ctypes.data_as(C.POINTER(c_int)) if rq.size else None, int(rq.size),
                 int(prob.K.get("f", 0)), int(prob.K.get("z", 0)), int(prob.K.get("l", 0)))
    sol = ABIPSolution()
    info = ABIPInfo()
    status = lib.abip(C.byref(d), C.byref(sol), C.byref(info), C.byref(K))
    out = {"status_val": int(status), "status": info.status.decode(), "ipm_iter": int(info.ipm_iter),
           "admm_iter": int(info.admm_iter), "pobj": info.pobj, "dobj": info.dobj, "res_pri": info.res_pri,
           "res_dual": info.res_dual, "rel_gap": info.rel_gap, "setup_time_ms": info.setup_time,
           "solve_time_ms": info.solve_time}
    for name, ln in (("x", prob.n), ("y", prob.m), ("s", prob.n)):
        p = getattr(sol, name)
        out[name] = np.ctypeslib.as_array(p, shape=(ln,)).copy() if p else None
    return out
