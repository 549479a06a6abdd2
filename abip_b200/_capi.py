"""ctypes binding of libabip_gpu.so (the C ABI declared in include/abip_gpu.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C abip_b200/csrc`.  There is no CPU
fallback: importing a compute entry point without the library raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABIP_GPU_LIB") or os.path.join(_HERE, "csrc", "libabip_gpu.so")  # env override: tuning builds

c_int = C.c_long      # abip_int  (-DDLONG layout)
c_float = C.c_double  # abip_float
SC_COUNT = 64


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


class ABIPGpuStats(C.Structure):
    _fields_ = [("n_admm_launch", c_int), ("n_bb_launch", c_int), ("n_solves", c_int),
                ("n_cg_iters", c_int), ("n_spmv_A", c_int), ("n_spmv_AT", c_int),
                ("n_kernel_launches", c_int), ("admm_kernel_ms", c_float), ("bb_kernel_ms", c_float),
                ("alg_bytes", c_float), ("h2d_bytes", c_float), ("d2h_bytes", c_float),
                ("alg_bytes_admm", c_float), ("alg_bytes_bb", c_float), ("solve_event_ms", c_float)]


# every symbol include/abip_gpu.h declares (checked by tests/test_capi_symbols.py)
DECLARED_SYMBOLS = [
    "abip_init_lin_sys_work", "abip_solve_lin_sys", "abip_free_lin_sys_work", "abip_accum_by_Atrans",
    "abip_accum_by_A", "abip_validate_lin_sys", "abip_get_lin_sys_method", "abip_get_lin_sys_summary",
    "abip_normalize_A", "abip_un_normalize_A", "abip_free_A_matrix", "abip_copy_A_matrix",
    "abip_gpu_set_default_settings", "abip_gpu_init", "abip_gpu_solve", "abip_gpu_finish", "abip_gpu_main",
    "abip_gpu_get_stats", "abip_gpu_init_dist", "abip_gpu_comm_export", "abip_gpu_comm_connect", "abip_gpu_partition", "abip_gpu_column_partition", "abip_gpu_batch_main",
    "abipgpu_lp_create", "abipgpu_lp_create_scaling", "abipgpu_lp_destroy", "abipgpu_lp_set_problem", "abipgpu_lp_cold_start",
    "abipgpu_lp_outer_prologue", "abipgpu_lp_admm_iter", "abipgpu_lp_mu_stats", "abipgpu_lp_reinit",
    "abipgpu_lp_clamp_v", "abipgpu_lp_bb_begin", "abipgpu_lp_bb_round", "abipgpu_lp_solve_vec", "abipgpu_lp_get_vec",
    "abipgpu_lp_set_vec", "abipgpu_lp_g_th", "abipgpu_lp_spmv", "abipgpu_lp_describe", "abipgpu_plan_debug", "abipgpu_plan_debug_cost",
    # ABIP-QCP (bound in abip_b200/qcp.py)
    "abip_qcp_gpu_set_default_settings", "abip_qcp_gpu", "abip_qcp_gpu_last_counters", "abip_qcp_scale_data",
    "abipgpu_qcp_create", "abipgpu_qcp_destroy", "abipgpu_qcp_iter", "abipgpu_qcp_solve_vec", "abipgpu_qcp_solve_nspace", "abipgpu_qcp_get_vec",
    "abipgpu_qcp_set_vec", "abipgpu_qcp_a_coef", "abipgpu_qcp_counters",
]

_lib = None


def lib():
    """Load libabip_gpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (nvcc, sm_100a). abip_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
    P = C.POINTER
    vp = C.c_void_p
    fp = P(c_float)
    ip = P(c_int)
    sig = {
        "abip_init_lin_sys_work": (vp, [P(ABIPMatrix), P(ABIPSettings)]),
        "abip_solve_lin_sys": (c_int, [P(ABIPMatrix), P(ABIPSettings), vp, fp, fp, c_int]),
        "abip_free_lin_sys_work": (None, [vp]),
        "abip_accum_by_Atrans": (None, [P(ABIPMatrix), vp, fp, fp]),
        "abip_accum_by_A": (None, [P(ABIPMatrix), vp, fp, fp]),
        "abip_validate_lin_sys": (c_int, [P(ABIPMatrix)]),
        "abip_get_lin_sys_method": (vp, [P(ABIPMatrix), P(ABIPSettings)]),
        "abip_get_lin_sys_summary": (vp, [vp, P(ABIPInfo)]),
        "abip_normalize_A": (None, [P(ABIPMatrix), P(ABIPSettings), P(ABIPScaling)]),
        "abip_un_normalize_A": (None, [P(ABIPMatrix), P(ABIPSettings), P(ABIPScaling)]),
        "abip_free_A_matrix": (None, [P(ABIPMatrix)]),
        "abip_copy_A_matrix": (c_int, [P(P(ABIPMatrix)), P(ABIPMatrix)]),
        "abip_gpu_set_default_settings": (None, [P(ABIPData)]),
        "abip_gpu_init": (vp, [P(ABIPData), P(ABIPInfo)]),
        "abip_gpu_solve": (c_int, [vp, P(ABIPData), P(ABIPSolution), P(ABIPInfo)]),
        "abip_gpu_finish": (None, [vp]),
        "abip_gpu_main": (c_int, [P(ABIPData), P(ABIPSolution), P(ABIPInfo)]),
        "abip_gpu_get_stats": (None, [vp, P(ABIPGpuStats)]),
        "abip_gpu_init_dist": (vp, [P(ABIPData), P(ABIPInfo), c_int, c_int]),
        "abip_gpu_comm_export": (c_int, [vp, vp]),
        "abip_gpu_comm_connect": (c_int, [vp, vp]),
        "abip_gpu_partition": (None, [vp, P(c_int), P(c_int)]),
        "abip_gpu_column_partition": (None, [c_int, ip, c_int, c_int, P(c_int), P(c_int)]),
        "abip_gpu_batch_main": (c_int, [P(P(ABIPData)), P(ABIPSolution), P(ABIPInfo), c_int, c_int, c_int]),
        "abipgpu_lp_create": (vp, [c_int, c_int, ip, ip, fp, P(ABIPSettings), C.c_int]),
        "abipgpu_lp_create_scaling": (vp, [c_int, c_int, ip, ip, fp, P(ABIPSettings), C.c_int, fp, fp, fp, fp]),
        "abipgpu_lp_destroy": (None, [vp]),
        "abipgpu_lp_set_problem": (C.c_int, [vp, fp, fp, fp, fp]),
        "abipgpu_lp_cold_start": (C.c_int, [vp, c_float, c_float]),
        "abipgpu_lp_outer_prologue": (C.c_int, [vp, C.c_int]),
        "abipgpu_lp_admm_iter": (C.c_int, [vp, c_int, c_int, c_float, c_float, fp]),
        "abipgpu_lp_mu_stats": (C.c_int, [vp, C.c_int, fp]),
        "abipgpu_lp_reinit": (C.c_int, [vp, C.c_int, c_float, C.c_int]),
        "abipgpu_lp_clamp_v": (C.c_int, [vp]),
        "abipgpu_lp_bb_begin": (C.c_int, [vp]),
        "abipgpu_lp_bb_round": (C.c_int, [vp, C.c_int, c_int, c_float, c_float, fp]),
        "abipgpu_lp_solve_vec": (C.c_int, [vp, C.c_int, C.c_int, c_int, fp]),
        "abipgpu_lp_get_vec": (C.c_int, [vp, C.c_int, fp, c_int]),
        "abipgpu_lp_set_vec": (C.c_int, [vp, C.c_int, fp, c_int]),
        "abipgpu_lp_g_th": (c_float, [vp]),
        "abipgpu_lp_spmv": (C.c_int, [vp, C.c_int, fp, fp]),
        "abipgpu_lp_describe": (None, [vp, C.c_char_p, c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def default_settings(**overrides) -> ABIPSettings:
    st = ABIPSettings()
    d = ABIPData()
    d.stgs = C.pointer(st)
    lib().abip_gpu_set_default_settings(C.byref(d))
    for k, v in overrides.items():
        if not hasattr(st, k):
            raise KeyError(f"unknown ABIP setting {k!r}")
        setattr(st, k, v)
    return st
