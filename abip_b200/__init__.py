"""abip_b200 -- B200-native engine for ABIP's inner ADMM iteration (indirect / pcg=1 path).

Product layout: csrc/ (hand-written sm_100a CUDA kernels + the C ABI of include/abip_gpu.h) and the
host-side mirror of the reference's `abip(data, K, params)` entry (api.py).  No CPU fallback.
"""
from .api import abip, lp_solve, lp_solve_batch, get_params, LinSysPlugin, LpEngine, LpSolver  # noqa: F401
from . import problems  # noqa: F401
from .lasso import lasso_solve, lasso_cone_program  # noqa: F401
from .svm import svm_solve, svm_cone_program, svm_qp_solve, svm_qp_program  # noqa: F401

__all__ = ["abip", "lp_solve", "get_params", "LinSysPlugin", "LpEngine", "LpSolver", "lp_solve_batch", "problems",
           "lasso_solve", "lasso_cone_program", "svm_solve", "svm_cone_program",
           "svm_qp_solve", "svm_qp_program"]
