"""The mu rules of csrc/lp_logic.h (shared by the host loop and the device-resident outer loop of the batch engine, k_batch
kind BATCH_SOLVE) against the oracle's restatement of src/abip.c:753-992 (rules) and :2251-2277 (selection), on random states
that reach every branch of the table.  Host-only: the header is compiled with g++ (tests/tools/mu_logic_check.cpp)."""
import os
import shutil
import subprocess
import types

import numpy as np
import pytest

from oracle import lp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_mu_rules_match_the_oracle(tmp_path):
    exe = str(tmp_path / "mu_logic_check")
    subprocess.run(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "tools", "mu_logic_check.cpp"), "-o", exe], check=True)
    rng = np.random.default_rng(7)
    cases, lines = [], []
    for t in range(600):
        eps = 10.0 ** rng.integers(-6, -2)
        mu = eps * 10.0 ** rng.uniform(-3.5, 4.5)
        err = eps * rng.choice([0.5, 1.5, 2.5, 3.5, 5.0, 8.0, 20.0])
        res = err * np.array([1.0, rng.uniform(0, 1), rng.uniform(0, 1)])[rng.permutation(3)]
        sp = rng.choice([0.01, 0.15, 0.3, 0.6])
        spr = rng.choice([0.05, 0.25, 0.45])
        hybrid = int(rng.integers(0, 2))
        dyn = rng.choice([0.0, -0.5, -1.0, 0.3])
        second = rng.choice([0.0, 0.5])
        n = int(rng.integers(5, 50))
        xs = rng.uniform(1e-3, 2.0, size=n + 1) * mu
        c = dict(mu=mu, sigma=rng.choice([0.3, 0.5, 0.8]), gamma=rng.choice([2.0, 3.0]), dyn=dyn, fc=int(rng.integers(0, 2)),
                 dc=int(rng.integers(0, 2)), eps=eps, sp=sp, spr=spr, second=second, dynx=0.8, thresh=1000.0, hybrid=hybrid, n=n,
                 res=res, xs=xs)
        cases.append(c)
        lines.append("%.17g %.17g %.17g %.17g %d %d %.17g %.17g %.17g %.17g %.17g %.17g %d %d %.17g %.17g %.17g %.17g %.17g" % (
            c["mu"], c["sigma"], c["gamma"], c["dyn"], c["fc"], c["dc"], eps, sp, spr, second, 0.8, 1000.0, hybrid, n + 1,
            res[0], res[1], res[2], float(xs.min()), float(xs.sum())))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(cases)
    seen_rules = set()
    for c, ln in zip(cases, out):
        rule, rc, mu, sigma, gamma, dyn, fc, dc, stopper = ln.split()
        seen_rules.add(int(rule))
        st = types.SimpleNamespace(eps=c["eps"], sparsity_ratio=c["spr"], hybrid_mu=c["hybrid"], dynamic_sigma=c["dyn"],
                                   dynamic_sigma_second=c["second"], dynamic_x=c["dynx"], hybrid_thresh=c["thresh"], avg_criterion=0)
        m = 3
        u = np.concatenate([np.zeros(m), np.sqrt(c["xs"])])
        w = types.SimpleNamespace(stgs=st, mu=c["mu"], sigma=c["sigma"], gamma=c["gamma"], sp=c["sp"], final_check=c["fc"],
                                  double_check=c["dc"], m=m, n=c["n"], u=u, v=u.copy(), u_avgcon=u, v_avgcon=u)
        r = types.SimpleNamespace(res_pri=c["res"][0], res_dual=c["res"][1], rel_gap=c["res"][2])
        O.update_mu(w, r)
        assert int(rc) == 0
        assert abs(float(mu) - w.mu) <= 1e-14 * abs(w.mu), (c, ln)
        assert abs(float(sigma) - w.sigma) <= 1e-15 and abs(float(gamma) - w.gamma) <= 1e-15, (c, ln)
        assert float(dyn) == st.dynamic_sigma and int(fc) == w.final_check and int(dc) == w.double_check, (c, ln)
        spmin = min(c["sp"], c["spr"])
        want = round(w.mu ** -0.35) if spmin > 0.5 else round(w.mu ** -1.0) if spmin > 0.2 else 1000000
        assert abs(int(stopper) - want) <= 1
    assert seen_rules == {0, 1, 2, 3}
