"""GPU parity tests (through the C ABI) of the ABIP-LP engine against the oracle (oracle/lp_oracle.py, a numpy
restatement pinned to the compiled reference) and the golden fixtures produced by the reference itself.

Tolerances: FP64 everywhere; kernels differ from the oracle only by summation order, so single steps agree to
~1e-10 relative; whole solves follow north_star: same status, residuals <= eps, objective within 1e-6 relative,
ADMM iteration count within 5%.
"""
import json
import math
import os

import numpy as np
import pytest

from abip_b200 import problems
from abip_b200.api import LinSysPlugin, LpEngine, SC, lp_solve, abip
from oracle import lp_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "lp_golden.json")


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / (np.max(np.abs(b)) + 1e-300))


def make_work(p, **kw):
    st = O.Settings(**kw)
    return O.Work(p.csc(), p.b, p.c, st), st


def make_engine(w, **kw):
    A = w.A.tocsc()
    A.sort_indices()
    e = LpEngine(A, **kw)
    e.set_problem(w.b, w.c, w.D, w.E)
    return e


PROBLEMS = {
    "rand_200x700": lambda: problems.random_lp(200, 700, 4, seed=3),
    "rand_1x9": lambda: problems.random_lp(1, 9, 1, seed=4),
    "rand_37x1000_dense_rows": lambda: problems.random_lp(37, 1000, 20, seed=5),     # rows ~540 nnz -> long-row path
    "mcf_small": lambda: problems.mcf_lp(4, 40, 200, 6, 300, seed=6),               # mixed short/long rows
}


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_spmv_and_adjoint(name):
    p = PROBLEMS[name]()
    A = p.csc()
    e = LpEngine(A)
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(p.n), rng.standard_normal(p.m)
    ax, aty = e.spmv(x), e.spmv(y, trans=True)
    assert rel(ax, A @ x) < 1e-13
    assert rel(aty, A.T @ y) < 1e-13
    # adjoint identity between the two stored CSR copies, linearity
    assert abs(np.dot(ax, y) - np.dot(x, aty)) <= 1e-11 * (np.linalg.norm(ax) * np.linalg.norm(y) + 1)
    assert rel(e.spmv(2.5 * x), 2.5 * ax) < 1e-14
    # determinism: bit-identical on repeat
    assert np.array_equal(ax, e.spmv(x))
    e.close()


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_plugin_accum_and_solve(name):
    """linsys plugin symbols with host pointers vs oracle LinSys (indirect.c:222-242, 393-434)."""
    p = PROBLEMS[name]()
    w, st = make_work(p)
    A = w.A.tocsc(); A.sort_indices()
    plug = LinSysPlugin(A)
    rng = np.random.default_rng(1)
    x, y0 = rng.standard_normal(p.n), rng.standard_normal(p.m)
    assert rel(plug.accum_by_A(x, y0), y0 + A @ x) < 1e-13
    x0 = rng.standard_normal(p.n)
    assert rel(plug.accum_by_Atrans(y0, x0), x0 + A.T @ y0) < 1e-13
    for it, warm in ((-1, None), (0, rng.standard_normal(p.m)), (7, rng.standard_normal(p.m)), (400, np.zeros(p.m))):
        b = rng.standard_normal(p.m + p.n)
        ref = b.copy()
        its = w.p.solve(ref, warm, it)
        got = plug.solve(b, warm, it)
        # both stop on |r| < tol: compare through the KKT residual instead of elementwise when its differ
        K_res = np.concatenate([st.rho_y * got[:p.m] + A @ got[p.m:], A.T @ got[:p.m] - got[p.m:]]) - b
        tol = max(np.linalg.norm(b[:p.m]) * (1e-9 if it < 0 else 0.1 / (it + 1.0) ** 2), 1e-7)
        assert np.linalg.norm(K_res[:p.m]) < 1.5 * tol + 1e-12
        assert np.linalg.norm(K_res[p.m:]) < 1e-9 * (1 + np.linalg.norm(got))
        assert rel(got, ref) < 50 * tol / (np.linalg.norm(ref) + 1e-300) + 1e-9
    plug.close()


@pytest.mark.parametrize("name", ["rand_200x700", "mcf_small", "rand_37x1000_dense_rows"])
def test_set_problem_g(name):
    p = PROBLEMS[name]()
    w, _ = make_work(p)
    e = make_engine(w)
    assert rel(e.get("H"), w.h) < 1e-15
    assert rel(e.get("G"), w.g) < 1e-7
    assert abs(e.g_th() - w.g_th) < 1e-7 * abs(w.g_th)
    assert rel(e.get("M"), w.p.M) < 1e-14
    e.close()


@pytest.mark.parametrize("half", [0, 1])
@pytest.mark.parametrize("name", ["rand_200x700", "mcf_small"])
def test_admm_iterations_stepwise(name, half):
    """30 inner iterations, same (mu, beta): every vector and every reduced scalar against the oracle."""
    p = PROBLEMS[name]()
    w, st = make_work(p, half_update=half)
    e = make_engine(w, half_update=half)
    e.cold_start(1.0, 1.0)
    e.outer_prologue(0)
    mu, beta = 0.37, 1.3
    w.mu, w.beta = mu, beta
    for j in range(30):
        w.u_prev[:] = w.u
        its = O.project_lin_sys(w, j + 5)
        if half:
            O.half_update_dual_vars(w); O.project_barrier_dual(w)
        else:
            O.project_barrier(w); O.update_dual_vars(w)
        O.restart_vars(w, j, j + 5)
        O.compute_avg(w, j)
        sc = e.admm_iter(j, j + 5, mu, beta)
        assert abs(sc[SC["CG_ITS"]] - its) <= 1
        for nm, ref in (("UT", w.u_t), ("U", w.u), ("V", w.v), ("UAVGC", w.u_avgcon), ("VAVGC", w.v_avgcon)):
            assert rel(e.get(nm), ref) < 2e-6, (j, nm)   # CG stops at a tolerance, not at machine precision
        q, nrm = O.q_norm_parts(w, w.u, w.v)
        g = sc[SC["S_PR"]:SC["S_PR"] + 13]
        q_gpu = g[0] + g[5] + (g[3] - g[8] - g[12]) ** 2
        n_gpu = 1 + math.sqrt(g[4] + g[9] + g[11] ** 2 + g[10] + g[12] ** 2)
        assert abs(math.sqrt(q_gpu) / n_gpu - math.sqrt(q) / nrm) < 1e-5 * (math.sqrt(q) / nrm) + 1e-9
        assert bool(sc[SC["HAS_AVG"]]) == ((j + 1) % 10 == 0)
        if (j + 1) % 10 == 0:
            qa, na = O.q_norm_parts(w, w.u_avgcon, w.v_avgcon)
            g = sc[SC["AVG_BASE"]:SC["AVG_BASE"] + 13]
            qa_gpu = g[0] + g[5] + (g[3] - g[8] - g[12]) ** 2
            assert abs(qa_gpu - qa) < 1e-5 * qa + 1e-12
        # re-synchronise so that stopping-rule jitter of CG cannot accumulate
        e.set("U", w.u); e.set("V", w.v); e.set("USUM", w.u_sumcon); e.set("VSUM", w.v_sumcon)
    # weighted residual sums against calc_residuals
    r = O.Residuals()
    O.calc_residuals(w, r, 0, 30)
    g = sc[SC["S_PR"]:SC["S_PR"] + 13]
    nrm = st.scale * w.sc_c * w.sc_b
    res_pri = math.sqrt(g[2]) / (w.sc_b * st.scale) / (1 + w.nm_b) / abs(g[11])
    res_dual = math.sqrt(g[7]) / (w.sc_c * st.scale) / (1 + w.nm_c) / abs(g[11])
    assert abs(res_pri - r.res_pri) < 1e-5 * r.res_pri
    assert abs(res_dual - r.res_dual) < 1e-5 * r.res_dual
    assert abs(g[3] / nrm - r.bt_y_by_tau) < 1e-6 * abs(r.bt_y_by_tau) + 1e-12
    e.close()


def test_restart_path():
    """restart_vars firing (abip.c:601-628) with a tiny threshold/frequency."""
    p = PROBLEMS["rand_200x700"]()
    kw = dict(restart_thresh=3, restart_fre=4)
    w, st = make_work(p, **kw)
    e = make_engine(w, **kw)
    e.cold_start(1.0, 1.0)
    e.outer_prologue(0)
    w.mu, w.beta = 0.5, 1.0
    for j in range(13):
        w.u_prev[:] = w.u
        O.project_lin_sys(w, j)
        O.project_barrier(w); O.update_dual_vars(w)
        O.restart_vars(w, j, j)
        O.compute_avg(w, j)
        e.admm_iter(j, j, w.mu, w.beta)
        for nm, ref in (("U", w.u), ("V", w.v), ("UAVGC", w.u_avgcon)):
            assert rel(e.get(nm), ref) < 1e-5, (j, nm)
    e.close()


@pytest.mark.parametrize("name", ["rand_200x700", "mcf_small"])
def test_reinit_mu_stats_bb(name):
    p = PROBLEMS[name]()
    w, st = make_work(p)
    e = make_engine(w)
    e.cold_start(1.0, 1.0)
    e.outer_prologue(0)
    w.mu, w.beta = 0.8, 1.0
    for j in range(6):
        w.u_prev[:] = w.u
        O.project_lin_sys(w, j); O.project_barrier(w); O.update_dual_vars(w); O.compute_avg(w, j)
        e.admm_iter(j, j, w.mu, w.beta)
    e.set("U", w.u); e.set("V", w.v)
    sc = e.mu_stats(0)
    xs = w.u[w.m:] * w.v[w.m:]
    assert abs(sc[SC["MIN_XS"]] - xs.min()) < 1e-14 and abs(sc[SC["SUM_XS"]] - xs.sum()) < 1e-9 * xs.sum()
    w.sigma = 0.64
    for idx in (0, 1):
        O.reinitialize_vars(w, idx); e.reinit(idx, w.sigma, 0)
    assert rel(e.get("U"), w.u) < 1e-15 and rel(e.get("V"), w.v) < 1e-15
    # Barzilai-Borwein rounds: the 5 inner products per round and the resulting beta sequence
    tr = []
    O.update_adapt_params(w, 6, trace=tr)
    e.bb_begin()
    beta_prev, carry, betas = 1.0, 0, []
    for _ in range(len(tr)):
        sc = e.bb_round(carry, 6, w.mu, beta_prev)
        beta = O.bb_beta_from_scalars(st, sc[SC["BB_UTUT"]], sc[SC["BB_UTV"]], sc[SC["BB_UU"]], sc[SC["BB_VV"]],
                                      sc[SC["BB_UV"]], beta_prev)
        betas.append(beta)
        d = abs(beta - beta_prev)
        if 0 < d <= st.eps_pen:
            break
        elif d > st.eps_pen:
            beta_prev, carry = beta, 1
        else:
            carry = 2
    assert len(betas) == len(tr)
    assert np.allclose(betas, tr, rtol=2e-3), (betas, tr)
    e.close()


def _check_solution(p, x, y, s, info, eps):
    """size-independent properties: residuals of the returned point recomputed on the CPU from the ORIGINAL data"""
    A = p.csc()
    pres = np.linalg.norm(A @ x - p.b) / (1 + np.linalg.norm(p.b))
    dres = np.linalg.norm(A.T @ y + s - p.c) / (1 + np.linalg.norm(p.c))
    pobj, dobj = float(p.c @ x), float(p.b @ y)
    gap = abs(pobj - dobj) / (1 + abs(pobj) + abs(dobj))
    assert pres < 1.5 * eps and dres < 1.5 * eps and gap < 1.5 * eps, (pres, dres, gap)
    assert x.min() > -1e-9 and s.min() > -1e-9
    assert abs(pobj - info["pobj"]) < 1e-8 * (1 + abs(pobj))


@pytest.mark.parametrize("name,eps", [("rand_200x700", 1e-4), ("mcf_small", 1e-4), ("rand_1x9", 1e-3),
                                      ("rand_37x1000_dense_rows", 1e-4)])
def test_full_solve_vs_oracle(name, eps):
    p = PROBLEMS[name]()
    o = O.solve(p.csc(), p.b, p.c, O.Settings(eps=eps))
    x, y, s, info = lp_solve(p.csc(), p.b, p.c, dict(tol=eps, verbose=0))
    assert info["status_val"] == o.status_val == 1
    assert abs(info["admm_iter"] - o.admm_iter) <= max(2, 0.05 * o.admm_iter), (info["admm_iter"], o.admm_iter)
    assert abs(info["pobj"] - o.pobj) <= 1e-6 * (1 + abs(o.pobj)) + 2 * eps * eps
    assert max(info["pres"], info["dres"], info["gap"]) < eps
    _check_solution(p, x, y, s, info, eps)


def test_cfg1_against_reference_golden():
    """BASELINE.json configs[0] against the numbers the compiled reference produced (tests/golden/make_golden.py)."""
    gold = json.load(open(GOLD))["cfg1"]
    p = problems.cfg1()
    x, y, s, info = abip({"A": p.csc(), "b": p.b, "c": p.c}, {"l": p.n}, dict(tol=1e-4, verbose=0, pcg=1))
    assert info["status"] == gold["status"]
    assert abs(info["admm_iter"] - gold["admm_iter"]) <= 0.05 * gold["admm_iter"]
    assert abs(info["pobj"] - gold["pobj"]) <= 1e-6 * abs(gold["pobj"]) + 1e-8
    assert max(info["pres"], info["dres"], info["gap"]) < 1e-4
    assert rel(x[:16], np.array(gold["x_head"])) < 1e-3
    _check_solution(p, x, y, s, info, 1e-4)


def test_infeasible_and_validation():
    """error behaviour of the entry: m > n is rejected like the reference's validate (abip.c:1661-1665)."""
    p = problems.random_lp(30, 20, 3, seed=9) if False else None
    import scipy.sparse as sp
    A = sp.random(8, 5, density=0.6, random_state=1, format="csc")
    x, y, s, info = lp_solve(A, np.ones(8), np.ones(5), dict(verbose=0))
    assert info["status_val"] == -4 and np.isnan(x).all()


def test_dropin_reference_core_with_gpu_plugin():
    """The UNMODIFIED reference solver core (abip.c, adaptive.c, ...) linked against our linsys plugin instead of
    linsys/indirect.c + common.c (oracle/Makefile target libabip_gpuplug_ref.so): host-pointer drop-in."""
    from oracle import ref_lp
    if not ref_lp.available("gpuplug"):
        pytest.skip("oracle/_ref/libabip_gpuplug_ref.so not built")
    gold = json.load(open(GOLD))["rand_200x700"]
    p = PROBLEMS["rand_200x700"]()
    r = ref_lp.solve(p, which="gpuplug", eps=1e-4)
    assert r["status"] == gold["status"]
    assert abs(r["admm_iter"] - gold["admm_iter"]) <= max(2, 0.05 * gold["admm_iter"])
    assert abs(r["pobj"] - gold["pobj"]) <= 1e-6 * abs(gold["pobj"]) + 1e-7
    assert max(r["res_pri"], r["res_dual"], r["rel_gap"]) < 1e-4


GOLD_CASES = {
    "rand_200x700_eps1e-3": lambda: problems.random_lp(200, 700, 4, seed=3),
    "rand_200x700_half": lambda: problems.random_lp(200, 700, 4, seed=3),
    "rand_200x700_noadapt": lambda: problems.random_lp(200, 700, 4, seed=3),
    "rand_200x700_nonorm": lambda: problems.random_lp(200, 700, 4, seed=3),
    "rand_1x9": lambda: problems.random_lp(1, 9, 1, seed=4),
    "cfg5_lp_0": lambda: problems.random_lp(500, 2000, 5, seed=5000),
    "cfg2_scale0.01": lambda: problems.cfg2(scale=0.01),
}


@pytest.mark.parametrize("name", list(GOLD_CASES))
def test_settings_variants_against_reference_golden(name):
    """half_update / adaptive=0 / normalize=0 / eps variants and the cfg2 / cfg5 families against the outputs of the
    compiled reference (tests/golden/lp_golden.json)."""
    g = json.load(open(GOLD))[name]
    p = GOLD_CASES[name]()
    eps = g["settings"]["eps"]
    x, y, s, info = lp_solve(p.csc(), p.b, p.c, dict(verbose=0), **g["settings"])
    assert info["status"] == g["status"]
    assert abs(info["admm_iter"] - g["admm_iter"]) <= max(2, 0.05 * g["admm_iter"]), (info["admm_iter"], g["admm_iter"])
    assert abs(info["pobj"] - g["pobj"]) <= 1e-6 * abs(g["pobj"]) + 2 * eps * eps
    assert max(info["pres"], info["dres"], info["gap"]) < eps
    _check_solution(p, x, y, s, info, eps)


@pytest.mark.parametrize("kw", [dict(), dict(origin_rescale=1), dict(qp_rescale=1, pc_ruiz_rescale=0), dict(scale=2.0)])
def test_device_equilibration_is_bit_identical_to_host(kw):
    """abipgpu_lp_create_scaling (device-side common.c:150-565) against abip_normalize_A (host C++) and the oracle."""
    import ctypes as C
    from abip_b200 import _capi, api
    L = _capi.lib()
    p = problems.mcf_lp(4, 40, 200, 6, 300, seed=6)
    st = _capi.default_settings(verbose=0, **kw)
    H = api.CscHolder((p.m, p.n, p.Ap.copy(), p.Ai.copy(), p.Ax.copy()))
    sc = _capi.ABIPScaling()
    L.abip_normalize_A(C.byref(H.c), C.byref(st), C.byref(sc))      # host reference (scales H in place)
    Dh = np.ctypeslib.as_array(sc.D, shape=(p.m,)).copy()
    Eh = np.ctypeslib.as_array(sc.E, shape=(p.n,)).copy()
    Hu = api.CscHolder((p.m, p.n, p.Ap.copy(), p.Ai.copy(), p.Ax.copy()))
    D, E = np.zeros(p.m), np.zeros(p.n)
    mr, mc = C.c_double(), C.c_double()
    e = L.abipgpu_lp_create_scaling(p.m, p.n, api._ip(Hu.Ap), api._ip(Hu.Ai), api._fp(Hu.Ax), C.byref(st), 0,
                                    api._fp(D), api._fp(E), C.byref(mr), C.byref(mc))
    assert e
    assert np.array_equal(D, Dh) and np.array_equal(E, Eh)
    assert mr.value == sc.mean_norm_row_A and mc.value == sc.mean_norm_col_A
    # the scaled matrix on the device: A' y through the engine == scaled host matrix times y
    import scipy.sparse as sp
    As = sp.csc_matrix((H.Ax, H.Ai, H.Ap), shape=(p.m, p.n))
    y = np.random.default_rng(0).standard_normal(p.m)
    out = np.zeros(p.n)
    assert L.abipgpu_lp_spmv(e, 1, api._fp(y), api._fp(out)) == 0
    assert rel(out, As.T @ y) < 1e-13
    x = np.random.default_rng(1).standard_normal(p.n)
    out2 = np.zeros(p.m)
    assert L.abipgpu_lp_spmv(e, 0, api._fp(x), api._fp(out2)) == 0
    assert rel(out2, As @ x) < 1e-13
    L.abipgpu_lp_destroy(e)


@pytest.mark.parametrize("mode", ["lockstep", "grids"])
def test_batch_of_small_lps_matches_oracle(mode):
    """abip_gpu_batch_main (BASELINE.json configs[4] family): lock-step mode (one CTA per problem, all problems of the
    batch advanced by one k_batch launch per step) and the older one-grid-per-problem mode, each problem against the
    oracle (status, ADMM iterations within 5 %, objective 1e-6) and against residuals recomputed on the CPU."""
    from abip_b200 import lp_solve_batch
    probs = [problems.random_lp(60 + 7 * i, 200 + 31 * i, 4, seed=900 + i) for i in range(10)]
    probs.append(problems.mcf_lp(3, 20, 80, 4, 100, seed=11))      # long rows inside a one-CTA engine
    conc, ctas = (len(probs), 1) if mode == "lockstep" else (4, 8)
    res = lp_solve_batch(probs, dict(tol=1e-4, verbose=0), concurrency=conc, ctas_per_problem=ctas)
    assert len(res) == len(probs)
    for p, (x, y, s, info) in zip(probs, res):
        o = O.solve(p.csc(), p.b, p.c, O.Settings(eps=1e-4))
        assert info["status"] == o.status == "Solved", (p.name, info["status"], o.status)
        assert abs(info["admm_iter"] - o.admm_iter) <= max(2, 0.05 * o.admm_iter)
        assert abs(info["pobj"] - o.pobj) <= 1e-6 * (1 + abs(o.pobj))
        _check_solution(p, x, y, s, info, 1e-4)


# ---------------------------------------------------------------------------------------------------------
# Benchmark-scale parity: fixtures produced by the compiled reference, OpenMP build
# (tests/golden/make_golden_large.py): cfg2 at scale 0.1 / 0.25 / 1.0 (the bench workload, 5.0 M nonzeros, 196 ADMM
# iterations), cfg4 at scale 0.05, and the first 64 problems of cfg5 through the batch executor.
# ---------------------------------------------------------------------------------------------------------
GOLD_LARGE = os.path.join(os.path.dirname(__file__), "golden", "lp_golden_large.json")
GOLD_CFG5 = os.path.join(os.path.dirname(__file__), "golden", "cfg5_golden.json")
LARGE_CASES = {
    "cfg2_scale0.1": lambda: problems.cfg2(scale=0.1),
    "cfg2_scale0.25": lambda: problems.cfg2(scale=0.25),
    "cfg4_scale0.05": lambda: problems.cfg4(scale=0.05),
    "cfg2_full": lambda: problems.cfg2(scale=1.0),
}


@pytest.mark.parametrize("name", list(LARGE_CASES))
def test_benchmark_scale_against_reference_golden(name):
    g = json.load(open(GOLD_LARGE))[name]
    p = LARGE_CASES[name]()
    assert (p.m, p.n, p.nnz) == (g["m"], g["n"], g["nnz"])
    x, y, s, info = lp_solve(p.csc(), p.b, p.c, dict(tol=1e-4, verbose=0))
    assert info["status"] == g["status"]
    assert info["ipm_iter"] == g["ipm_iter"]
    assert abs(info["admm_iter"] - g["admm_iter"]) <= 0.05 * g["admm_iter"], (info["admm_iter"], g["admm_iter"])
    assert abs(info["pobj"] - g["pobj"]) <= 1e-6 * abs(g["pobj"]) + 2e-8, (info["pobj"], g["pobj"])
    assert max(info["pres"], info["dres"], info["gap"]) < 1e-4
    assert rel(x[:16], np.array(g["x_head"])) < 1e-3 and rel(y[:16], np.array(g["y_head"])) < 1e-3
    assert abs(np.linalg.norm(x) - g["x_norm"]) < 1e-4 * g["x_norm"]
    _check_solution(p, x, y, s, info, 1e-4)


def test_cfg5_batch_slice_against_reference_golden():
    """64 problems of BASELINE.json configs[4] through abip_gpu_batch_main (lock-step executor), each against the
    reference's own result: status, ADMM iterations within 5 %, objective 1e-6."""
    from abip_b200 import lp_solve_batch
    gold = json.load(open(GOLD_CFG5))
    probs = problems.cfg5_batch(len(gold))
    res = lp_solve_batch(probs, dict(tol=1e-4, verbose=0), concurrency=64, ctas_per_problem=1)
    assert len(res) == len(gold)
    for p, g, (x, y, s, info) in zip(probs, gold, res):
        assert info["status"] == g["status"], (p.name, info["status"], g["status"])
        assert abs(info["admm_iter"] - g["admm_iter"]) <= max(2, 0.05 * g["admm_iter"]), (p.name, info["admm_iter"], g["admm_iter"])
        pobj = float(p.c @ x)
        assert abs(pobj - g["pobj"]) <= 1e-6 * abs(g["pobj"]) + 2e-8, (p.name, pobj, g["pobj"])
        assert abs(np.linalg.norm(x) - g["x_norm"]) < 1e-3 * g["x_norm"]


@pytest.mark.parametrize("raw", [dict(), dict(hybrid_mu=0, dynamic_sigma=0.0), dict(hybrid_mu=0, dynamic_sigma=0.3),
                                 dict(hybrid_mu=0, dynamic_sigma=-0.5), dict(adaptive=0), dict(restart_thresh=40)])
def test_batch_device_outer_loop_equals_host_outer_loop(raw, monkeypatch):
    """The device-resident OUTER loop of the batch engine (k_batch kind BATCH_SOLVE: inner loops, convergence checks, mu rules
    of src/abip.c:753-992 and their selection :2251-2277, re-initialisation, BB search) takes the branches of the host loop:
    same status, ADMM / IPM iteration counts and objective as with ABIP_GPU_BATCH_HOST_OUTER=1, for the default hybrid selection
    (dynamic rule, then the LOQO rule with min / sum of u_i v_i reduced inside the launch), the table rule, the LOQO rule and
    the dynamic rule alone, without the adaptive search, and with the hand-over to the host at the restart threshold."""
    from abip_b200 import lp_solve_batch
    probs = [problems.random_lp(60, 200, 4, seed=900 + i) for i in range(10)]
    par = dict(tol=1e-4, verbose=0)
    monkeypatch.setenv("ABIP_GPU_BATCH_HOST_OUTER", "1")
    ref = lp_solve_batch(probs, par, concurrency=10, ctas_per_problem=1, **raw)
    monkeypatch.delenv("ABIP_GPU_BATCH_HOST_OUTER")
    dev = lp_solve_batch(probs, par, concurrency=10, ctas_per_problem=1, **raw)
    for (xr, yr, sr, ir), (xd, yd, sd, idv) in zip(ref, dev):
        assert (idv["status"], idv["ipm_iter"], idv["admm_iter"]) == (ir["status"], ir["ipm_iter"], ir["admm_iter"])
        assert abs(idv["pobj"] - ir["pobj"]) <= 1e-9 * max(1.0, abs(ir["pobj"]))
        assert np.max(np.abs(xd - xr)) <= 1e-9 * max(1.0, np.max(np.abs(xr)))
    assert any(r[3]["status"] == "Solved" for r in dev)


def test_sigint_handler_is_restored_after_a_batch():
    """the SIGINT listener is process-global but a batch runs one solve per host thread; it is
    reference-counted, so the host's own handler (Python's KeyboardInterrupt) is back in place afterwards."""
    import signal
    from abip_b200 import lp_solve_batch
    before = signal.getsignal(signal.SIGINT)
    probs = [problems.random_lp(40, 120, 3, seed=70 + i) for i in range(12)]
    res = lp_solve_batch(probs, dict(tol=1e-3, verbose=0), concurrency=12, ctas_per_problem=1)
    assert all(r[3]["status_val"] == 1 for r in res)
    assert signal.getsignal(signal.SIGINT) is before
    x, y, s, info = lp_solve(probs[0].csc(), probs[0].b, probs[0].c, dict(tol=1e-3, verbose=0))
    assert signal.getsignal(signal.SIGINT) is before


@pytest.mark.parametrize("name", ["rand_200x700", "cfg1"])
def test_trajectory_without_resync_matches_oracle(name, tmp_path):
    """The whole solve, iteration by iteration, with NO re-synchronisation of the iterates from the oracle: the trace of
    the host loop (ABIP_GPU_TRACE: outer / inner / global iteration, mu, beta, CG iterations, Q-norm criterion) against the
    oracle's own trace.  Drift is allowed to show: the iteration structure must be identical, mu exact, beta and the
    criterion to the accuracy the inexact solves allow."""
    p = problems.cfg1() if name == "cfg1" else PROBLEMS[name]()
    tf = str(tmp_path / "trace.txt")
    os.environ["ABIP_GPU_TRACE"] = tf
    try:
        x, y, s, info = lp_solve(p.csc(), p.b, p.c, dict(tol=1e-4, verbose=0))
    finally:
        del os.environ["ABIP_GPU_TRACE"]
    o = O.solve(p.csc(), p.b, p.c, O.Settings(eps=1e-4), trace=True)
    gpu = [ln.split() for ln in open(tf) if ln.startswith("it ")]
    assert info["admm_iter"] == o.admm_iter
    assert len(gpu) == len(o.trace)
    cg_diff = 0
    for a, b in zip(gpu, o.trace):
        assert (int(a[1]), int(a[2]), int(a[3])) == (b[0], b[1], b[2])
        assert abs(float(a[4]) - b[3]) <= 1e-9 * abs(b[3])                       # mu
        assert abs(float(a[5]) - b[4]) <= 1e-4 * abs(b[4])                       # beta (BB search on inexact solves)
        assert abs(float(a[7]) - b[6]) <= 1e-3 * abs(b[6]) + 1e-12               # Q-norm criterion
        cg_diff += int(a[6]) != b[5]
    assert cg_diff <= 0.05 * len(gpu) + 1
