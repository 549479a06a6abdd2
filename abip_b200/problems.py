"""Deterministic synthetic problem generators for the BASELINE.json configs.

All arrays are FP64 / int64 CSC with sorted row indices, no duplicate entries, no empty rows or
columns.  LPs are feasible and bounded by construction: draw complementary
x0, s0 >= 0 and y0 ~ N(0,1), then b = A x0, c = A' y0 + s0.

The reference ships no generator for its LP path (its benchmarks read MPS files,
scripts/bench-lp/README.md); the Lasso-like recipe for the QCP config follows
scripts/bench-qcp/get_lasso_simu_data.m:1-15 in spirit (random design, sparse ground truth).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np
import scipy.sparse as sp


@dataclass
class LPProblem:
    """Standard-form LP  min c'x  s.t. Ax = b, x >= 0, with A in CSC (reference amatrix.h:10-17)."""
    m: int
    n: int
    Ap: np.ndarray  # int64 [n+1]
    Ai: np.ndarray  # int64 [nnz]
    Ax: np.ndarray  # float64 [nnz]
    b: np.ndarray
    c: np.ndarray
    name: str = "lp"
    meta: dict = field(default_factory=dict)

    @property
    def nnz(self) -> int:
        return int(self.Ap[-1])

    def csc(self) -> sp.csc_matrix:
        return sp.csc_matrix((self.Ax, self.Ai, self.Ap), shape=(self.m, self.n))


def _finish_lp(A: sp.csc_matrix, rng: np.random.Generator, name: str, meta: dict) -> LPProblem:
    A = A.tocsc()
    A.sum_duplicates()
    A.sort_indices()
    m, n = A.shape
    # complementary primal/dual pair: first half of a random permutation basic, rest nonbasic
    perm = rng.permutation(n)
    x0 = np.zeros(n)
    s0 = np.zeros(n)
    half = n // 2
    x0[perm[:half]] = rng.uniform(0.1, 1.1, size=half)
    s0[perm[half:]] = rng.uniform(0.1, 1.1, size=n - half)
    y0 = rng.standard_normal(m)
    b = A @ x0
    c = A.T @ y0 + s0
    return LPProblem(m=m, n=n, Ap=A.indptr.astype(np.int64), Ai=A.indices.astype(np.int64),
                     Ax=A.data.astype(np.float64), b=np.asarray(b, dtype=np.float64),
                     c=np.asarray(c, dtype=np.float64), name=name, meta=meta)


def _patch_empty_rows(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray, m: int, n: int,
                      rng: np.random.Generator):
    """Give every empty row one entry (in a random column) so diag(AA') > 0 (indirect.c:68-72)."""
    present = np.zeros(m, dtype=bool)
    present[rows] = True
    empty = np.flatnonzero(~present)
    if empty.size:
        rows = np.concatenate([rows, empty])
        cols = np.concatenate([cols, rng.integers(0, n, size=empty.size)])
        vals = np.concatenate([vals, rng.standard_normal(empty.size)])
    return rows, cols, vals


def random_lp(m: int, n: int, nnz_per_col: int = 5, seed: int = 0, name: str | None = None) -> LPProblem:
    """cfg 1 family: k nonzeros per column at uniformly random distinct rows, values N(0,1)."""
    rng = np.random.default_rng(seed)
    k = min(nnz_per_col, m)
    # distinct rows per column: sample with a random offset + distinct strides is biased; do it exactly
    # in chunks with argpartition of random keys when m is small, else rejection on duplicates.
    if m <= 4096:
        keys = rng.random((n, m))
        rows = np.argpartition(keys, k - 1, axis=1)[:, :k].reshape(-1)
    else:
        rows = rng.integers(0, m, size=(n, k))
        for _ in range(50):
            srt = np.sort(rows, axis=1)
            dup = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
            if not dup.any():
                break
            rows[dup] = rng.integers(0, m, size=(int(dup.sum()), k))
        rows = rows.reshape(-1)
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    vals = rng.standard_normal(n * k)
    rows, cols, vals = _patch_empty_rows(rows.astype(np.int64), cols, vals, m, n, rng)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(m, n)).tocsc()
    return _finish_lp(A, rng, name or f"random_lp_m{m}_n{n}", {"family": "random", "seed": seed,
                                                               "nnz_per_col": k})


def mcf_lp(commodities: int = 20, nodes: int = 7500, arcs: int = 47600, side_rows: int = 2400,
           side_row_nnz: int = 875, seed: int = 0, name: str | None = None) -> LPProblem:
    """cfg 2 / cfg 4 family: multicommodity-flow structure.

    Columns: flow x[k][e] for K commodities x E arcs, then E capacity slacks.
    Rows: K*V flow-conservation rows (+1 tail / -1 head), E capacity rows (sum_k x[k][e] + slack_e),
    R random dense-ish side rows over the flow columns (values N(0,1)).
    Defaults give m = 199,600 + 400 = 200,000?  (K*V + E + R = 150,000 + 47,600 + 2,400), n = K*E + E
    = 999,600, nnz ~ 3*K*E + E + R*side_row_nnz ~ 5.0M.
    """
    rng = np.random.default_rng(seed)
    K, V, E, R = commodities, nodes, arcs, side_rows
    tail = rng.integers(0, V, size=E)
    head = (tail + 1 + rng.integers(0, V - 1, size=E)) % V  # head != tail
    # make sure every node is touched: thread a Hamiltonian cycle through the first V arcs
    if E >= V:
        order = rng.permutation(V)
        tail[:V] = order
        head[:V] = np.roll(order, -1)
    n_flow = K * E
    n = n_flow + E
    m = K * V + E + R
    e_idx = np.tile(np.arange(E, dtype=np.int64), K)
    k_idx = np.repeat(np.arange(K, dtype=np.int64), E)
    col_flow = k_idx * E + e_idx
    rows = [k_idx * V + tail[e_idx], k_idx * V + head[e_idx], K * V + e_idx,
            K * V + np.arange(E, dtype=np.int64)]
    cols = [col_flow, col_flow, col_flow, n_flow + np.arange(E, dtype=np.int64)]
    vals = [np.ones(n_flow), -np.ones(n_flow), np.ones(n_flow), np.ones(E)]
    if R > 0:
        w = min(side_row_nnz, n_flow)
        sc = rng.integers(0, n_flow, size=(R, w))
        rows.append(np.repeat(K * V + E + np.arange(R, dtype=np.int64), w))
        cols.append(sc.reshape(-1))
        vals.append(rng.standard_normal(R * w))
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    vals = np.concatenate(vals)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(m, n)).tocsc()  # duplicates (rare) are summed
    A.eliminate_zeros()
    return _finish_lp(A, rng, name or f"mcf_lp_K{K}_V{V}_E{E}_R{R}",
                      {"family": "mcf", "seed": seed, "K": K, "V": V, "E": E, "R": R,
                       "side_row_nnz": side_row_nnz})


def cfg1(seed: int = 1) -> LPProblem:
    """BASELINE.json configs[0]: m=1,000 n=5,000 density 0.5%."""
    return random_lp(1000, 5000, 5, seed=seed, name="cfg1_lp_m1k_n5k")


def cfg2(seed: int = 2, scale: float = 1.0) -> LPProblem:
    """BASELINE.json configs[1]: m=200k n=1M nnz=5M multicommodity-flow structure (scale<1 shrinks V,E,R)."""
    V = max(8, int(round(7500 * scale)))
    E = max(V + 8, int(round(47600 * scale)))
    R = max(1, int(round(2400 * scale)))
    w = max(4, min(int(round(875 * min(1.0, scale * 4))), 20 * E))
    p = mcf_lp(20, V, E, R, w, seed=seed, name=f"cfg2_mcf_lp_scale{scale:g}")
    return p


def cfg4(seed: int = 4, scale: float = 1.0) -> LPProblem:
    """BASELINE.json configs[3]: m=2M n=10M nnz=60M (same family as cfg2, 6 nnz/col)."""
    V = max(8, int(round(75000 * scale)))
    E = max(V + 8, int(round(476000 * scale)))
    R = max(1, int(round(24000 * scale)))
    return mcf_lp(20, V, E, R, 1310, seed=seed, name=f"cfg4_mcf_lp_scale{scale:g}")


def cfg5_batch(count: int = 4096, base_seed: int = 5000, m: int = 500, n: int = 2000,
               nnz_per_col: int = 5):
    """BASELINE.json configs[4]: independent small LPs (density 1% -> 5 nnz per column)."""
    return [random_lp(m, n, nnz_per_col, seed=base_seed + i, name=f"cfg5_lp_{i}") for i in range(count)]


# ---------------------------------------------------------------------------------------------------------
# ABIP-QCP:  min 1/2 x'Qx + c'x  s.t. Ax = b, x in K,  K = SOC^q x RSOC^rq x free x zero x R+   (column order
# q -> rq -> f -> z -> l, reference README.md:121 and src/abip-qcp/include/abip.h:63-76)
# ---------------------------------------------------------------------------------------------------------
@dataclass
class QCPProblem:
    m: int
    n: int
    A: sp.csc_matrix
    Q: sp.csc_matrix | None
    b: np.ndarray
    c: np.ndarray
    K: dict
    name: str = "qcp"


def toy_qcp() -> QCPProblem:
    """The explicit 2 x 8 QCP of the reference's only test (test/test_abip_install.m:32-43)."""
    A = sp.csc_matrix(np.array([[1, 2, 3, 4, 5, 6, 7, 8], [0, 1, 2, 1, 2, 3, 1, 2]], dtype=np.float64))
    return QCPProblem(2, 8, A, sp.identity(8, format="csc", dtype=np.float64), np.array([4.0, 3.0]),
                      np.array([1.0, 0, 2, 1, 4, 2, 3, 0]), {"q": [3], "rq": [3], "f": 1, "l": 1}, "toy_qcp")


def _cone_interior_point(K: dict, n: int, rng: np.random.Generator):
    """A point strictly inside K (primal) / K* (dual; all cones here are self-dual except free<->zero)."""
    x = np.zeros(n)
    pos = 0
    for d in K.get("q", []):
        if d > 0:
            t = rng.standard_normal(d - 1) if d > 1 else np.zeros(0)
            x[pos] = np.linalg.norm(t) + rng.uniform(0.5, 1.5)
            x[pos + 1:pos + d] = t
            pos += d
    for d in K.get("rq", []):
        t = rng.standard_normal(d - 2)
        a = rng.uniform(0.5, 1.5)
        bb = (t @ t) / (2 * a) + rng.uniform(0.5, 1.5)
        x[pos], x[pos + 1] = a, bb
        x[pos + 2:pos + d] = t
        pos += d
    pos += K.get("f", 0) + K.get("z", 0)
    ll = K.get("l", 0)
    x[pos:pos + ll] = rng.uniform(0.1, 1.1, size=ll)
    return x


def random_qcp(m: int, n_soc: int, soc_dim: int, n_rsoc: int = 0, rsoc_dim: int = 4, n_free: int = 0, n_lin: int = 0,
               nnz_per_col: int = 4, q_offdiag_per_col: int = 2, seed: int = 0, with_q: bool = True,
               name: str | None = None) -> QCPProblem:
    """cfg 3 family: SOCP/QCP with `n_soc` second-order cones (+ optional rotated cones, free and linear blocks),
    random sparse A, Q = B'B-like sparse PSD (diagonal + symmetric off-diagonals, diagonally dominant).
    Feasible by construction: x0 in int K, b = A x0; s0 in int K*, y0 random, c = A'y0 + s0 - Q x0."""
    rng = np.random.default_rng(seed)
    K = {}
    if n_soc:
        K["q"] = [soc_dim] * n_soc
    if n_rsoc:
        K["rq"] = [rsoc_dim] * n_rsoc
    if n_free:
        K["f"] = n_free
    if n_lin:
        K["l"] = n_lin
    n = n_soc * soc_dim + n_rsoc * rsoc_dim + n_free + n_lin
    k = min(nnz_per_col, m)
    rows = rng.integers(0, m, size=(n, k))
    cols = np.repeat(np.arange(n), k)
    vals = rng.standard_normal(n * k)
    r, c_, v = _patch_empty_rows(rows.reshape(-1).astype(np.int64), cols.astype(np.int64), vals, m, n, rng)
    A = sp.coo_matrix((v, (r, c_)), shape=(m, n)).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    Q = None
    if with_q:
        qo = q_offdiag_per_col
        ri = rng.integers(0, n, size=n * qo)
        ci = np.repeat(np.arange(n), qo)
        vv = 0.1 * rng.standard_normal(n * qo)
        off = sp.coo_matrix((vv, (ri, ci)), shape=(n, n)).tocsr()
        off = off + off.T
        off.setdiag(0)
        dom = np.asarray(abs(off).sum(axis=1)).ravel()
        Q = (off + sp.diags(dom + rng.uniform(0.05, 1.0, size=n))).tocsc()
        Q.sort_indices()
    x0 = _cone_interior_point(K, n, rng)
    pos_f = n_soc * soc_dim + n_rsoc * rsoc_dim
    x0[pos_f:pos_f + n_free] = rng.standard_normal(n_free)
    s0 = _cone_interior_point(K, n, rng)
    s0[pos_f:pos_f + n_free] = 0.0   # dual of the free cone is {0}
    y0 = rng.standard_normal(m)
    b = A @ x0
    c = A.T @ y0 + s0 - (Q @ x0 if Q is not None else 0.0)
    return QCPProblem(m, n, A, Q, np.asarray(b), np.asarray(c), K, name or f"random_qcp_m{m}_n{n}")


def cfg3(seed: int = 3, scale: float = 1.0) -> QCPProblem:
    """BASELINE.json configs[2]: SOCP with 10k second-order cones of dimension 50 (n = 500k), nnz(A) = 10M
    (20 per column), sparse PSD Q.  m = 100k."""
    n_soc = max(2, int(round(10000 * scale)))
    m = max(4, int(round(100000 * scale)))
    return random_qcp(m, n_soc, 50, nnz_per_col=20, q_offdiag_per_col=2, seed=seed, with_q=True,
                      name=f"cfg3_socp_scale{scale:g}")


# ---------------------------------------------------------------------------------------------------------
# Lasso instances (abip_b200/lasso.py): seeded, shared by tests/golden/make_golden_lasso.py and the tests; the recipe
# follows the reference's scripts/bench-qcp/get_lasso_simu_data.m (Gaussian design, sparse ground truth, noisy response,
# lambda a fraction of |X'y|_inf)
# ---------------------------------------------------------------------------------------------------------


def _lasso_case(m, n, density, k, seed, frac):
    def make():
        rng = np.random.default_rng(seed)
        if density < 1.0:
            X = sp.random(m, n, density=density, random_state=seed, format="csc")
            X.data = rng.standard_normal(X.nnz)
        else:
            X = sp.csc_matrix(rng.standard_normal((m, n)))
        w0 = np.zeros(n)
        w0[rng.choice(n, k, replace=False)] = rng.standard_normal(k)
        y = X @ w0 + 0.01 * rng.standard_normal(m)
        lam = frac * float(np.max(np.abs(X.T @ y)))
        return X, y, lam
    return make


LASSO_CASES = {
    "sparse_wide": _lasso_case(60, 150, 0.3, 8, 11, 0.1),
    "dense_tall": _lasso_case(120, 40, 1.0, 6, 12, 0.05),
    "sparse_wide_more_features": _lasso_case(80, 200, 0.1, 10, 13, 0.1),
}


# Soft-margin SVM instances (abip_b200/svm.py): seeded Gaussian features, labels from a noisy linear rule


def _svm_case(m, n, density, seed, C, noise):
    def make():
        rng = np.random.default_rng(seed)
        if density < 1.0:
            X = sp.random(m, n, density=density, random_state=seed, format="csc")
            X.data = rng.standard_normal(X.nnz)
        else:
            X = sp.csc_matrix(rng.standard_normal((m, n)))
        w0 = rng.standard_normal(n)
        y = np.sign(X @ w0 + noise * rng.standard_normal(m))
        y[y == 0] = 1.0
        return X, y, float(C)
    return make


SVM_CASES = {
    "dense_tall": _svm_case(120, 20, 1.0, 21, 1.0, 0.3),
    "sparse_tall": _svm_case(200, 40, 0.3, 22, 0.5, 0.5),
    "dense_wide": _svm_case(40, 60, 1.0, 23, 2.0, 0.2),
}
