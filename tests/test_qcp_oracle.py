"""CPU tests: the numpy QCP oracle against the golden outputs of the unmodified reference (direct QDLDL path), the
host-side QCP scaling of the C-ABI library against the oracle, and the engine's m-space Schur formulation (restated
in the oracle) against the exact solve."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from abip_b200 import problems, qcp
from oracle import qcp_oracle as O
from oracle import ref_qcp

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "qcp_golden.json")))
CASES = {
    "toy_qcp": lambda: problems.toy_qcp(),
    "mixed_cones_q": lambda: problems.random_qcp(30, 6, 5, n_rsoc=3, rsoc_dim=4, n_free=4, n_lin=10, seed=1),
    "socp_noq": lambda: problems.random_qcp(40, 10, 6, n_lin=20, seed=2, with_q=False),
    "qp_lin_only": lambda: problems.random_qcp(50, 0, 0, n_lin=150, seed=3),
    "soc_dim1_and_big": lambda: problems.random_qcp(20, 1, 60, n_lin=5, seed=4),
    "rsoc_only": lambda: problems.random_qcp(25, 0, 0, n_rsoc=12, rsoc_dim=5, seed=6),
    "cfg3_scale0.003": lambda: problems.cfg3(scale=0.003),
}


def _st(**kw):
    return O.Settings(eps_p=1e-4, eps_d=1e-4, eps_g=1e-4, **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_golden(name):
    g = GOLD[name]
    p = CASES[name]()
    o = O.solve(p.A, p.Q, p.b, p.c, p.K, _st())
    assert (o.status, o.ipm_iter, o.admm_iter) == (g["status"], g["ipm_iter"], g["admm_iter"])
    assert abs(o.pobj - g["pobj"]) <= 1e-8 * (1 + abs(g["pobj"]))
    assert abs(o.res_pri - g["res_pri"]) <= 1e-5 * g["res_pri"] + 1e-12
    assert np.allclose(o.x[:16], g["x_head"][:len(o.x[:16])], rtol=1e-6, atol=1e-8)


def test_toy_problem_known_answer():
    """The reference's only test problem (test/test_abip_install.m:32-43); values recorded in SURVEY.md section 4."""
    g = GOLD["toy_qcp"]
    assert (g["status"], g["ipm_iter"], g["admm_iter"]) == ("Solved", 7, 60)
    assert abs(g["pobj"] - (-9.84029712e-01)) < 1e-8 and abs(g["dobj"] - (-9.84045633e-01)) < 1e-8
    assert np.allclose(g["x_head"][:8], [0.04652, 0.04510, 0.01136, 0.34251, 0.06148, 0.20521, -2.16128, 2.00618], atol=1e-5)


@pytest.mark.skipif(not ref_qcp.available(), reason="oracle/_ref/libabip_qcp_ref.so not built")
def test_oracle_matches_compiled_reference_live():
    p = problems.random_qcp(24, 4, 4, n_rsoc=2, rsoc_dim=5, n_free=3, n_lin=12, seed=21)
    r = ref_qcp.solve(p, eps_p=1e-4, eps_d=1e-4, eps_g=1e-4)
    o = O.solve(p.A, p.Q, p.b, p.c, p.K, _st())
    assert (o.status_val, o.ipm_iter, o.admm_iter) == (r["status_val"], r["ipm_iter"], r["admm_iter"])
    assert np.max(np.abs(o.x - r["x"])) <= 1e-8 * (1 + np.max(np.abs(r["x"])))


@pytest.mark.parametrize("name", ["mixed_cones_q", "socp_noq", "cfg3_scale0.003"])
def test_schur_formulation_tracks_direct_path(name):
    """Engine design check on the CPU: eliminating x (m-space Schur PCG, rtol 1e-8) gives the ADMM iteration counts
    of the exact solve; the reference's n-space normal equations have condition ~ 1/rho_y."""
    p = CASES[name]()
    d = O.solve(p.A, p.Q, p.b, p.c, p.K, _st())
    s = O.solve(p.A, p.Q, p.b, p.c, p.K, _st(), linsys="schur", pcg_rtol=1e-8)
    assert (s.status, s.ipm_iter, s.admm_iter) == (d.status, d.ipm_iter, d.admm_iter)
    assert abs(s.pobj - d.pobj) <= 1e-6 * abs(d.pobj)
    w = O.Work(p.A, p.Q, p.b, p.c, p.K, _st(), "pcg", 1e-8)   # n-space qcp_pcg restatement (linsys.c:755-851)
    wd = O.Work(p.A, p.Q, p.b, p.c, p.K, _st(), "direct")
    rng = np.random.default_rng(0)
    b = rng.standard_normal(p.m + p.n)
    b1, b2 = b.copy(), b.copy()
    w.solve_linsys(b1, None, 0)
    wd.solve_linsys(b2, None, 0)
    assert np.max(np.abs(b1 - b2)) / np.max(np.abs(b2)) > 1e-6   # relative residual 1e-8, error orders larger


@pytest.mark.parametrize("kw", [dict(), dict(pc_scaling=1), dict(ruiz_scaling=0), dict(origin_scaling=0)])
def test_host_scaling_matches_oracle(kw):
    L = qcp._bind()
    p = CASES["mixed_cones_q"]()
    As, Qs, bs, cs, Do, Eo, sbo, sco = O.scaling_data(p.A, p.Q, p.b, p.c, p.K, O.Settings(**kw))
    st = qcp.default_settings(**kw)
    A, kA = qcp._mat(p.A.copy())
    Qm, kQ = qcp._mat(p.Q.copy())
    b, c = p.b.copy(), p.c.copy()
    cone, _keep = qcp.make_cone(p.K)
    D, E = np.zeros(p.m), np.zeros(p.n)
    sb, sc = C.c_double(), C.c_double()
    L.abip_qcp_scale_data(C.byref(A), C.byref(Qm), qcp._dp(b), qcp._dp(c), C.byref(cone), C.byref(st), qcp._dp(D),
                          qcp._dp(E), C.byref(sb), C.byref(sc))
    for got, ref in ((kA[0], As.data), (kQ[0], Qs.data), (b, bs), (c, cs), (D, Do), (E, Eo)):
        assert np.allclose(got, ref, rtol=1e-13, atol=1e-15)
    assert abs(sb.value - sbo) < 1e-15


def test_host_scaling_threads_are_bit_identical(monkeypatch):
    """abip_qcp_scale_data on several host threads (problems of >= 400k nonzeros) returns bit-for-bit what one thread
    returns: columns are independent, the row maxima of the Ruiz sweeps do not depend on the order, row sums stay serial."""
    from abip_b200 import problems
    L = qcp._bind()
    p = problems.cfg3(scale=0.05)   # n = 25,000, nnz(A) ~ 500k
    assert p.A.nnz + p.Q.nnz >= 400000
    out = {}
    for thr in ("1", "8"):
        monkeypatch.setenv("ABIP_GPU_HOST_THREADS", thr)
        st = qcp.default_settings(pc_scaling=1)
        A, kA = qcp._mat(p.A.copy())
        Qm, kQ = qcp._mat(p.Q.copy())
        b, c = p.b.copy(), p.c.copy()
        cone, _keep = qcp.make_cone(p.K)
        D, E = np.zeros(p.m), np.zeros(p.n)
        sb, sc = C.c_double(), C.c_double()
        L.abip_qcp_scale_data(C.byref(A), C.byref(Qm), qcp._dp(b), qcp._dp(c), C.byref(cone), C.byref(st), qcp._dp(D),
                              qcp._dp(E), C.byref(sb), C.byref(sc))
        out[thr] = (kA[0].copy(), kQ[0].copy(), b, c, D, E, sb.value, sc.value)
    for u, v in zip(out["1"][:6], out["8"][:6]):
        assert np.array_equal(u, v)
    assert out["1"][6:] == out["8"][6:]


def test_cone_prox_properties():
    """Barrier prox outputs are strictly inside their cones and reduce to the orthant formula in 1-D."""
    rng = np.random.default_rng(1)
    for _ in range(20):
        t = rng.standard_normal(7) * 3
        lam = float(rng.uniform(1e-4, 2))
        x = O.soc_prox(t, lam)
        assert x[0] > np.linalg.norm(x[1:])
        y = O.rsoc_prox(t, lam, 1.0)
        assert 2 * y[0] * y[1] > y[2:] @ y[2:] and y[0] > 0 and y[1] > 0
    t = np.array([-3.0, 0.0, 2.5])
    x = O.positive_orthant_prox(t, 0.3)
    assert (x > 0).all() and np.allclose(x * (x - t), 0.3)   # x (x - t) = lambda
