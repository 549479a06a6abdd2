"""MPS reader + standard-form preprocessing (abip_b200/mps.py, restating MATLAB mpsread + scripts/bench-lp/preprocess.m)
and the command-line entry (abip_b200/cli.py, the `bin/abip-indirect <mps> ...` call of
scripts/bench-lp/run_all_abip-binary-nobar-indirect.sh:50).  CPU part: the transformation is checked against an
independent LP solver (scipy HiGHS) on the general and on the standard form; GPU part: the CLI end to end."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from abip_b200 import mps, problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_general_mps(path, rng, m=14, n=22, fixed_format=False):
    """Random feasible bounded LP with every row type, ranges and every bound type; returns nothing (file only)."""
    A = sp.random(m, n, density=0.35, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csr")
    A.data = np.round(rng.standard_normal(A.nnz) * 4) / 2 + 0.25
    x0 = rng.uniform(0.5, 2.0, n)                       # interior point -> feasible by construction
    ax = A @ x0
    types = rng.choice(list("LGE"), size=m)
    rhs = np.where(types == "L", ax + 1.0, np.where(types == "G", ax - 1.0, ax))
    ranged = {i: float(rng.choice([-1, 1]) * rng.uniform(1.5, 3.0)) for i in range(m) if rng.random() < 0.3}
    for i, R in ranged.items():                         # keep x0 inside the ranged interval
        if types[i] == "E":
            rhs[i] = ax[i] - 0.5 * abs(R) if R >= 0 else ax[i] + 0.5 * abs(R)
    cost = np.round(rng.standard_normal(n) * 3) / 2
    kinds = rng.choice(["", "UP", "LO", "LOUP", "FX", "FR", "MI"], size=n, p=[.3, .2, .1, .2, .05, .1, .05])
    sep = "  " if fixed_format else " "
    with open(path, "w") as fh:
        fh.write("NAME          GENLP\n* a comment line\nROWS\n N  COST\n")
        for i in range(m):
            fh.write(f" {types[i]}{sep}R{i}\n")
        fh.write("COLUMNS\n")
        Ac = A.tocsc()
        for j in range(n):
            items = [("COST", cost[j])] if cost[j] != 0 else []
            items += [(f"R{Ac.indices[k]}", Ac.data[k]) for k in range(Ac.indptr[j], Ac.indptr[j + 1])]
            if not items:
                items = [("COST", 0.0)]
            for k in range(0, len(items), 2):            # two entries per line, like fixed-format writers do
                fh.write(f"    X{j}" + "".join(f"{sep}{r}{sep}{float(v)!r}" for r, v in items[k:k + 2]) + "\n")
        fh.write("RHS\n    RHS  COST  -2.5\n")
        for i in range(m):
            fh.write(f"    RHS  R{i}  {float(rhs[i])!r}\n")
        fh.write("RANGES\n")
        for i, R in ranged.items():
            fh.write(f"    RNG  R{i}  {R!r}\n")
        fh.write("BOUNDS\n")
        for j, k in enumerate(kinds):
            if "LO" in k:
                fh.write(f" LO BND X{j} {float(x0[j] - 1.5)!r}\n")
            if "UP" in k:
                fh.write(f" UP BND X{j} {float(x0[j] + 1.0)!r}\n")
            if k == "FX":
                fh.write(f" FX BND X{j} {float(x0[j])!r}\n")
            if k == "FR":
                fh.write(f" FR BND X{j}\n")
            if k == "MI":
                fh.write(f" MI BND X{j}\n UP BND X{j} {float(x0[j] + 2.0)!r}\n")
        fh.write("ENDATA\n")


def _bounded(g):
    """box the free directions so that HiGHS has a finite optimum to compare (same box on both forms)"""
    lb = np.where(np.isfinite(g.lb), g.lb, -50.0)
    ub = np.where(np.isfinite(g.ub), g.ub, 50.0)
    return lb, ub


@pytest.mark.parametrize("seed", range(6))
def test_standard_form_matches_general_form_under_highs(tmp_path, seed):
    rng = np.random.default_rng(100 + seed)
    f = str(tmp_path / "g.mps")
    write_general_mps(f, rng, fixed_format=bool(seed % 2))
    g = mps.read_mps(f)
    assert g.objcon == 2.5 and g.Aeq.shape[0] + g.Aineq.shape[0] >= 14
    g.lb, g.ub = _bounded(g)
    r0 = linprog(g.f, A_ub=g.Aineq, b_ub=g.bineq, A_eq=g.Aeq if g.Aeq.shape[0] else None,
                 b_eq=g.beq if g.Aeq.shape[0] else None, bounds=list(zip(g.lb, g.ub)), method="highs")
    assert r0.status == 0
    s = mps.to_standard_form(g)
    assert s.A.shape == (g.Aeq.shape[0] + g.Aineq.shape[0] + g.f.size, g.f.size + g.Aineq.shape[0] + g.f.size)
    r1 = linprog(s.c, A_eq=s.A, b_eq=s.b, bounds=(0, None), method="highs")
    assert r1.status == 0
    assert abs((r0.fun + g.objcon) - (r1.fun + s.objcon)) <= 1e-7 * (1 + abs(r0.fun))
    x = s.recover(r1.x)
    assert np.all(x >= g.lb - 1e-7) and np.all(x <= g.ub + 1e-7)
    assert np.all(g.Aineq @ x <= g.bineq + 1e-6) and np.allclose(g.Aeq @ x, g.beq, atol=1e-6)


def test_unbounded_below_defaults_of_preprocess_m(tmp_path):
    """preprocess.m:35-37 in MATLAB arithmetic: (lb > -inf) .* lb = 0 * -Inf = NaN -> -1e6, then -1e8 is added: a free
    variable is shifted by -1.01e8; finite lower bounds are kept."""
    f = str(tmp_path / "u.mps")
    open(f, "w").write("NAME U\nROWS\n N C\n E R0\nCOLUMNS\n X0 C 1.0 R0 1.0\n X1 C 1.0 R0 1.0\nRHS\n RHS R0 3.0\n"
                       "BOUNDS\n MI B X0\n LO B X1 2.0\nENDATA\n")
    s = mps.to_standard_form(mps.read_mps(f))
    assert s.lb_shift.tolist() == [-1.01e8, 2.0]
    assert s.b.tolist() == [3.0 + 1.01e8 - 2.0] and s.objcon == -1.01e8 + 2.0
    assert s.A.shape == (1, 2)


def test_write_read_roundtrip_and_empty_rows(tmp_path):
    p = problems.random_lp(30, 90, 3, seed=5)
    f = str(tmp_path / "s.mps")
    mps.write_mps(f, p.csc(), p.b, p.c, objcon=1.25)
    s = mps.load_standard_form(f)
    assert (abs(s.A - p.csc())).max() == 0 and np.array_equal(s.b, p.b) and np.array_equal(s.c, p.c)
    assert s.objcon == 1.25 and s.n_orig == p.n
    A2 = sp.vstack([p.csc(), sp.csr_matrix((1, p.n))]).tocsc()
    s2 = mps.drop_empty_rows(mps.StandardLP(A2, np.append(p.b, 0.0), p.c, 0.0, np.zeros(p.n), p.n))
    assert s2.A.shape[0] == p.m
    with pytest.raises(ValueError):
        mps.drop_empty_rows(mps.StandardLP(A2, np.append(p.b, 1.0), p.c, 0.0, np.zeros(p.n), p.n))


def test_cli_argument_forms():
    from abip_b200 import cli
    a = cli.parse_args(["x.mps", "3600", "100000", "10000000", "0", "1e-10", "1e-4", "5", "1", "out/x"])
    assert (a.mps, a.time_limit, a.max_ipm_iters, a.max_admm_iters, a.tol, a.out) == ("x.mps", 3600.0, 100000, 10000000,
                                                                                    1e-4, "out/x")
    b = cli.parse_args(["x.mps", "--tol", "1e-6", "--out", "y"])
    assert b.tol == 1e-6 and b.out == "y" and b.time_limit is None


@pytest.mark.gpu
def test_cli_end_to_end_on_gpu(tmp_path):
    rng = np.random.default_rng(7)
    f = str(tmp_path / "g.mps")
    write_general_mps(f, rng, m=20, n=40)
    g = mps.read_mps(f)
    lb, ub = _bounded(g)
    # box the free directions in the file as well, so that the LP has a finite optimum
    txt = open(f).read().replace("ENDATA\n", "")
    for j in range(g.f.size):
        if not np.isfinite(g.lb[j]):
            txt += f" LO BND X{j} -50.0\n"
        if not np.isfinite(g.ub[j]):
            txt += f" UP BND X{j} 50.0\n"
    open(f, "w").write(txt + "ENDATA\n")
    g = mps.read_mps(f)
    r0 = linprog(g.f, A_ub=g.Aineq, b_ub=g.bineq, A_eq=g.Aeq if g.Aeq.shape[0] else None,
                 b_eq=g.beq if g.Aeq.shape[0] else None, bounds=list(zip(g.lb, g.ub)), method="highs")
    assert r0.status == 0
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "abip_b200.cli", f, "600", "500", "1000000", "0", "1e-10", "1e-5", "5", "1", out]
    pr = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert pr.returncode == 0, pr.stdout[-2000:] + pr.stderr[-2000:]
    res = json.load(open(out + ".json"))
    assert res["status"] == "Solved"
    for key in ("pres", "dres", "time", "status", "pobj", "dobj", "admm_iter", "ipm_iter",          # analyze_abip.py:10-28
                "InnerIter", "OuterIter", "PResABIP", "DResABIP", "ABIPTime", "PObj", "DObj", "CGIter"):  # :33-60
        assert key in res
    assert res["CGIter"] > 0 and res["InnerIter"] == res["admm_iter"]
    assert abs(res["pobj"] - (r0.fun + g.objcon)) <= 2e-3 * (1 + abs(r0.fun + g.objcon))
    x = np.loadtxt(out + ".sol")
    assert x.size == g.f.size
    assert np.all(g.Aineq @ x <= g.bineq + 1e-2 * (1 + np.abs(g.bineq)))


@pytest.mark.parametrize("hdr", ["OBJSENSE MAX", "OBJSENSE MAXIMIZE", "OBJSENSE\n    MAX", "OBJSENSE_MAX"])
def test_objsense_max_is_rejected_in_every_spelling(tmp_path, hdr):
    """a maximisation problem must not be silently minimised (mpsread rejects it as well)"""
    f = str(tmp_path / "o.mps")
    open(f, "w").write("NAME O\n" + hdr + "\nROWS\n N C\n E R0\nCOLUMNS\n X0 C 1.0 R0 1.0\nRHS\n RHS R0 3.0\nENDATA\n")
    with pytest.raises(ValueError):
        mps.read_mps(f)
