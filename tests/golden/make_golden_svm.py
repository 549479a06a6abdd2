"""Generates tests/golden/svm_golden.json from the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libabip_qcp_ref.so): abip() with prob_type = SVM (enum value 1, include/abip.h:13), cone K = {rq: [n + 2],
l: 2 + 2 m + 2 n} as set by mex/abip_ml_mex.c:332-336, lambda = C, linsys_solver = 1 (QDLDL).  Run in the build container:
    python tests/golden/make_golden_svm.py"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from abip_b200.problems import SVM_CASES as CASES  # noqa: E402
from oracle import ref_qcp as R  # noqa: E402


def ref_svm(X, y, Cpar, eps, qp=False):
    """qp: prob_type = SVMQP (enum value 3; K = {f: n + 1, l: 2 m}, lambda = 1 / (m C)) instead of SVM."""
    lib = R.load()
    m, n = X.shape
    A, keep = R._mat(X.copy())  # the reference scales A by the labels in place
    b = y.copy()
    c = np.zeros(4 + 3 * n + 2 * m)
    st = R.ABIPSettings()
    d = R.ABIPData(m, n, C.pointer(A), None, b.ctypes.data_as(C.POINTER(C.c_double)),
                   c.ctypes.data_as(C.POINTER(C.c_double)), 1.0 / (m * float(Cpar)) if qp else float(Cpar), C.pointer(st))
    lib.abip_set_default_settings(C.byref(d))
    st.linsys_solver = 1
    st.prob_type = 3 if qp else 1
    st.verbose = 0
    st.time_limit = 600.0
    st.eps_p = st.eps_d = st.eps_g = eps
    rq = np.array([2 + n], dtype=np.int32)
    K = R.ABIPCone(None, 0, rq.ctypes.data_as(C.POINTER(C.c_int)), 1, 0, 0, 2 + 2 * m + 2 * n)
    if qp:
        K = R.ABIPCone(None, 0, None, 0, n + 1, 0, 2 * m)
    sol, info = R.ABIPSolution(), R.ABIPInfo()
    lib.abip(C.byref(d), C.byref(sol), C.byref(info), C.byref(K))
    w = np.ctypeslib.as_array(sol.x, shape=(n,)).copy()
    b0 = float(np.ctypeslib.as_array(sol.y, shape=(1,))[0])
    return w, b0, info


if __name__ == "__main__":
    out = {}
    for name, make in CASES.items():
        X, y, Cpar = make()
        w, b0, info = ref_svm(X, y, Cpar, 1e-5)
        xi = np.maximum(0.0, 1.0 - y * (X @ w + b0))
        out[name] = {"status": info.status.decode(), "ipm_iter": int(info.ipm_iter), "admm_iter": int(info.admm_iter),
                     "pobj": info.pobj, "objective": 0.5 * float(w @ w) + Cpar * float(xi.sum()), "w": w.tolist(), "b": b0}
        print(name, out[name]["status"], out[name]["admm_iter"], out[name]["pobj"], out[name]["objective"])
        w, b0, info = ref_svm(X, y, Cpar, 1e-5, qp=True)
        xi = np.maximum(0.0, 1.0 - y * (X @ w + b0))
        out[name + "_qp"] = {"status": info.status.decode(), "ipm_iter": int(info.ipm_iter), "admm_iter": int(info.admm_iter),
                             "pobj": info.pobj, "objective": 0.5 * float(w @ w) + Cpar * float(xi.sum()), "w": w.tolist(), "b": b0}
        print(name + "_qp", out[name + "_qp"]["status"], out[name + "_qp"]["admm_iter"], out[name + "_qp"]["pobj"], out[name + "_qp"]["objective"])
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "svm_golden.json"), "w"))
