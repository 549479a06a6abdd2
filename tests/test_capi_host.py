"""CPU tests of the C-ABI library: it loads, exports every symbol include/abip_gpu.h declares, its host-side
functions (validate / copy / equilibration) match the oracle, and compute entry points fail loudly -- never fall
back -- when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from abip_b200 import _capi, api, problems
from oracle import lp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    hdr = open(os.path.join(ROOT, "include", "abip_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(abip(?:gpu)?_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.DECLARED_SYMBOLS), declared ^ set(_capi.DECLARED_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_struct_layouts_match_reference_abi():
    # -DDLONG layout of the shipped reference build: all fields 8 bytes (include/abip.h:36-105)
    assert C.sizeof(_capi.ABIPSettings) == 31 * 8
    assert C.sizeof(_capi.ABIPInfo) == 32 + 12 * 8
    assert C.sizeof(_capi.ABIPMatrix) == 5 * 8
    assert C.sizeof(_capi.ABIPData) == 7 * 8
    st = _capi.default_settings()
    assert (st.max_ipm_iters, st.max_admm_iters, st.alpha, st.rho_y, st.adaptive_lookback) == (500, 1000000, 1.8, 1e-3, 20)
    assert (st.dynamic_sigma, st.hybrid_thresh, st.dynamic_sigma_second, st.restart_thresh) == (-1.0, 1000.0, 0.5, 100000)


@pytest.mark.parametrize("kw", [dict(), dict(origin_rescale=1), dict(qp_rescale=1, pc_ruiz_rescale=0), dict(scale=2.0),
                                dict(ruiz_iter=3)])
def test_normalize_A_matches_oracle_and_roundtrips(kw):
    L = _capi.lib()
    p = problems.random_lp(200, 700, 4, seed=3)
    H = api.CscHolder((p.m, p.n, p.Ap.copy(), p.Ai.copy(), p.Ax.copy()))
    st = _capi.default_settings(verbose=0, **kw)
    sc = _capi.ABIPScaling()
    L.abip_normalize_A(C.byref(H.c), C.byref(st), C.byref(sc))
    D = np.ctypeslib.as_array(sc.D, shape=(p.m,)).copy()
    E = np.ctypeslib.as_array(sc.E, shape=(p.n,)).copy()
    As, Do, Eo, mr, mc = O.normalize_A(p.csc(), O.Settings(**kw))
    assert np.allclose(H.Ax, As.data, rtol=1e-13, atol=0)
    assert np.allclose(D, Do, rtol=1e-13) and np.allclose(E, Eo, rtol=1e-13)
    assert abs(sc.mean_norm_row_A - mr) < 1e-12 and abs(sc.mean_norm_col_A - mc) < 1e-12
    L.abip_un_normalize_A(C.byref(H.c), C.byref(st), C.byref(sc))
    assert np.allclose(H.Ax, p.Ax, rtol=1e-13)


def test_validate_and_copy():
    L = _capi.lib()
    p = problems.random_lp(20, 50, 3, seed=1)
    H = api.CscHolder(p.csc())
    assert L.abip_validate_lin_sys(C.byref(H.c)) == 0
    dst = C.POINTER(_capi.ABIPMatrix)()
    assert L.abip_copy_A_matrix(C.byref(dst), C.byref(H.c)) == 1
    assert dst.contents.m == p.m and dst.contents.n == p.n
    assert np.array_equal(np.ctypeslib.as_array(dst.contents.x, shape=(p.nnz,)), H.Ax)
    L.abip_free_A_matrix(dst)
    bad = api.CscHolder((p.m, p.n, p.Ap.copy(), p.Ai.copy(), p.Ax.copy()))
    bad.Ai[0] = p.m + 3   # row index out of range
    assert L.abip_validate_lin_sys(C.byref(bad.c)) == -1
    bad2 = api.CscHolder((p.m, p.n, p.Ap.copy(), p.Ai.copy(), p.Ax.copy()))
    bad2.Ap[3] = bad2.Ap[4] + 1  # decreasing column pointers
    assert L.abip_validate_lin_sys(C.byref(bad2.c)) == -1


def test_entry_rejects_invalid_input_like_reference():
    # validate(): m > n is rejected (src/abip.c:1661-1665) -> status ABIP_FAILED, NaN outputs; no GPU needed
    import scipy.sparse as sp
    A = sp.random(8, 5, density=0.6, random_state=1, format="csc")
    x, y, s, info = api.lp_solve(A, np.ones(8), np.ones(5), dict(verbose=0))
    assert info["status_val"] == -4 and np.isnan(x).all() and np.isnan(y).all()
    with pytest.raises(ValueError):
        api.lp_solve(A.T.tocsc(), np.ones(5), np.ones(8), dict(verbose=0, pcg=0))  # direct path: out of scope
    with pytest.raises(ValueError):
        api.abip({"A": A, "b": np.ones(8), "c": np.ones(5)}, {"s": 3})


@pytest.mark.skipif(_have_gpu(), reason="checks behaviour without a CUDA device")
def test_no_cpu_fallback_without_gpu():
    p = problems.random_lp(20, 50, 3, seed=1)
    x, y, s, info = api.lp_solve(p.csc(), p.b, p.c, dict(verbose=0))
    assert info["status_val"] == -4 and np.isnan(x).all()      # ABIP_FAILED, like a failed init in the reference
    with pytest.raises(RuntimeError):
        api.LinSysPlugin(p.csc())
    with pytest.raises(RuntimeError):
        api.LpEngine(p.csc())


@pytest.mark.parametrize("deal", [0, 1])
@pytest.mark.parametrize("case", ["short", "mixed", "long", "empty_rows", "tiny"])
def test_spmv_plan_invariants(case, deal):
    """abipgpu_plan_debug (host-only): the SpMV plan of csrc/spmv_host.h covers every row exactly once -- short rows
    in chunks within the nonzero / row limits, long rows as consecutive equal pieces with consecutive scratch slots --
    and deals the chunks to the warps in CTA-contiguous groups."""
    import ctypes as C
    import numpy as np
    from abip_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(3)
    lens = {"short": rng.integers(1, 9, 5000), "mixed": np.where(rng.random(3000) < 0.02, rng.integers(300, 2000, 3000),
                                                                 rng.integers(0, 40, 3000)),
            "long": rng.integers(253, 5000, 40), "empty_rows": np.where(rng.random(2000) < 0.5, 0, 7),
            "tiny": np.array([3, 1, 2])}[case]
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nrows, nnz, ctas = len(lens), int(ptr[-1]), 5
    cap = 4 * (nnz // 8 + nrows + 64)
    chunks = np.zeros(4 * cap, dtype=np.int32)
    info = np.zeros(6, dtype=np.int32)
    wc = np.zeros(ctas * 64 + 1, dtype=np.int32)
    L.abipgpu_plan_debug.restype = C.c_long
    L.abipgpu_plan_debug.argtypes = [C.c_long, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    n = L.abipgpu_plan_debug(nrows, ptr.ctypes.data, ctas, deal, chunks.ctypes.data, cap, wc.ctypes.data, info.ctypes.data)
    assert n > 0
    warps, lim_nnz, lim_rows, n_long, n_pieces, _ = info.tolist()
    ch = chunks[:4 * n].reshape(n, 4)
    W = ctas * warps
    wc = wc[:W + 1]
    assert wc[0] == 0 and wc[-1] == n and np.all(np.diff(wc) >= 0)
    row_seen = np.zeros(nrows, dtype=np.int64)
    nnz_seen = np.zeros(max(nnz, 1), dtype=np.int64)
    pieces = {}
    for row0, nnz0, nr, cnt in ch.tolist():
        assert 0 <= cnt <= lim_nnz and 0 <= nnz0 and nnz0 + cnt <= nnz
        nnz_seen[nnz0:nnz0 + cnt] += 1
        if nr > 0:                                   # chunk of whole rows
            assert nr <= lim_rows and ptr[row0] == nnz0 and ptr[row0 + nr] == nnz0 + cnt
            assert np.all(lens[row0:row0 + nr] <= lim_nnz)
            row_seen[row0:row0 + nr] += 1
        else:                                        # piece of a long row, scratch slot -nr - 1
            assert lens[row0] > lim_nnz and ptr[row0] <= nnz0 and nnz0 + cnt <= ptr[row0 + 1] and cnt > 0
            pieces.setdefault(row0, []).append((-nr - 1, nnz0, cnt))
    assert np.all(row_seen[lens <= lim_nnz] == 1) and np.all(row_seen[lens > lim_nnz] == 0)
    assert np.all(nnz_seen[:nnz] == 1)               # every nonzero in exactly one chunk
    assert len(pieces) == n_long == int(np.sum(lens > lim_nnz))
    slots = sorted(s for ps in pieces.values() for s, _, _ in ps)
    assert slots == list(range(n_pieces))
    for row, ps in pieces.items():
        ps.sort()
        assert [s for s, _, _ in ps] == list(range(ps[0][0], ps[0][0] + len(ps)))        # consecutive slots
        assert ps[0][1] == ptr[row] and sum(c for _, _, c in ps) == lens[row]
        assert all(a[1] + a[2] == b[1] for a, b in zip(ps, ps[1:]))                        # in order, contiguous
    # the pieces of a long row stay inside one CTA (they are combined after a CTA barrier)
    cta_of_chunk = np.searchsorted(wc[::warps], np.arange(n), side="right") - 1
    for row, ps in pieces.items():
        idx = [i for i, c in enumerate(ch.tolist()) if c[2] <= 0 and c[0] == row]
        assert len(set(cta_of_chunk[idx].tolist())) == 1
    if not deal and case != "tiny":                  # contiguous mode: CTA b owns a contiguous range of rows
        first_rows = [ch[wc[b * warps]:wc[(b + 1) * warps], 0] for b in range(ctas) if wc[(b + 1) * warps] > wc[b * warps]]
        for a, b in zip(first_rows, first_rows[1:]):
            assert a.max() < b.min()


def test_spmv_plan_with_measured_costs():
    """abipgpu_plan_debug_cost: the CTA row ranges follow the per-row costs handed in (the path of the measured balance,
    lp_engine.cu tune_balance): with the first half of the rows four times as expensive, the CTAs split the COST evenly,
    every row is still covered exactly once, and NULL costs reproduce the structural plan."""
    import ctypes as C
    import numpy as np
    from abip_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 12, 8000)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nrows, nnz, ctas = len(lens), int(ptr[-1]), 8
    cap = 4 * (nnz // 8 + nrows + 64)
    L.abipgpu_plan_debug_cost.restype = C.c_long
    L.abipgpu_plan_debug_cost.argtypes = [C.c_long, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p,
                                          C.c_void_p]
    L.abipgpu_plan_debug.restype = C.c_long
    L.abipgpu_plan_debug.argtypes = [C.c_long, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]

    def plan(cost):
        chunks = np.zeros(4 * cap, dtype=np.int32)
        info = np.zeros(6, dtype=np.int32)
        wc = np.zeros(ctas * 64 + 1, dtype=np.int32)
        n = L.abipgpu_plan_debug_cost(nrows, ptr.ctypes.data, ctas, 0, cost.ctypes.data if cost is not None else None,
                                      chunks.ctypes.data, cap, wc.ctypes.data, info.ctypes.data)
        assert n > 0
        warps = int(info[0])
        return chunks[:4 * n].reshape(n, 4).copy(), wc[:ctas * warps + 1].copy(), warps

    cost = (lens + 2.0) * np.where(np.arange(nrows) < nrows // 2, 4.0, 1.0)
    ch, wc, warps = plan(cost)
    seen = np.zeros(nrows, dtype=np.int64)
    for row0, nnz0, nr, cnt in ch.tolist():
        assert nr > 0 and ptr[row0] == nnz0 and ptr[row0 + nr] == nnz0 + cnt
        seen[row0:row0 + nr] += 1
    assert np.all(seen == 1)
    per_cta = []
    for b in range(ctas):
        rows = [(r0, nr) for r0, _, nr, _ in ch[wc[b * warps]:wc[(b + 1) * warps]].tolist()]
        lo, hi = min(r for r, _ in rows), max(r + k for r, k in rows)
        assert sum(k for _, k in rows) == hi - lo            # contiguous range of rows
        per_cta.append(cost[lo:hi].sum())
    per_cta = np.array(per_cta)
    assert per_cta.max() / per_cta.min() < 1.02              # equal cost per CTA (row granularity)
    # the expensive half is spread over 4/5 of the CTAs
    first_half_ctas = sum(1 for b in range(ctas) if ch[wc[b * warps], 0] < nrows // 2)
    assert first_half_ctas in (6, 7)
    # NULL costs: identical to the structural plan
    ch0, wc0, _ = plan(None)
    chunks = np.zeros(4 * cap, dtype=np.int32)
    wcb = np.zeros(ctas * 64 + 1, dtype=np.int32)
    info = np.zeros(6, dtype=np.int32)
    n = L.abipgpu_plan_debug(nrows, ptr.ctypes.data, ctas, 0, chunks.ctypes.data, cap, wcb.ctypes.data, info.ctypes.data)
    assert n == len(ch0) and np.array_equal(chunks[:4 * n].reshape(n, 4), ch0) and np.array_equal(wcb[:len(wc0)], wc0)
