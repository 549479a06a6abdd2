"""MPS reader and standard-form preprocessing in front of the ABIP-LP path.

The reference's benchmark pipeline is MATLAB: `mpsread` followed by scripts/bench-lp/preprocess.m:1-80, which turns a
general LP  min f'x  s.t. Aeq x = beq, Aineq x <= bineq, lb <= x <= ub  into ABIP's standard form
min c'x  s.t. A x = b, x >= 0.  `read_mps` restates what mpsread returns (E rows -> Aeq, L rows -> Aineq, G rows
negated into Aineq, ranged rows -> two inequalities, N row -> objective, BOUNDS section), `to_standard_form` restates
preprocess.m line by line:

    lb: finite kept, NaN -> -1e6, -inf -> -1.01e8 (MATLAB: 0 * -Inf = NaN -> -1e6, then -1e8 added)   (preprocess.m:35-37)
    x' = x - lb >= 0; one slack per inequality; one row  x'_j + t_j = ub_j - lb_j  per finite ub   (:40-57)
    A = [Aeq 0 0; Aineq I 0; D 0 I],  b = [beq - Aeq lb; bineq - Aineq lb; ub - lb],  c = [f; 0]
    objcon = f'lb with the SHIFTED lb (+ the objective constant of the file).  Deliberate deviation: preprocess.m:75
             uses prob.lb, which is -Inf (NaN for f_j = 0) as soon as the problem has a free variable      (:75-80)

Host-side Python (the reference's own preprocessing is a script); the GPU engine only ever sees the standard form.
"""
from __future__ import annotations

import dataclasses
import gzip

import numpy as np
import scipy.sparse as sp


@dataclasses.dataclass
class GeneralLP:
    """What MATLAB's mpsread returns (the fields preprocess.m uses)."""
    name: str
    f: np.ndarray
    Aeq: sp.csr_matrix
    beq: np.ndarray
    Aineq: sp.csr_matrix
    bineq: np.ndarray
    lb: np.ndarray
    ub: np.ndarray
    objcon: float
    col_names: list
    n_integer: int = 0       # integrality markers are read and ignored (LP relaxation)


@dataclasses.dataclass
class StandardLP:
    A: sp.csc_matrix
    b: np.ndarray
    c: np.ndarray
    objcon: float
    lb_shift: np.ndarray     # x_original = x_std[:n_orig] + lb_shift
    n_orig: int
    name: str = ""

    def recover(self, x_std: np.ndarray) -> np.ndarray:
        return np.asarray(x_std)[: self.n_orig] + self.lb_shift


def _open(path):
    return gzip.open(path, "rt") if str(path).endswith(".gz") else open(path, "r")


def read_mps(path) -> GeneralLP:
    """Free- or fixed-format MPS (names without blanks).  Negative RHS on the objective row is the objective
    constant with the sign convention of the format (obj = f'x - rhs_N)."""
    name = ""
    row_type: dict[str, str] = {}
    row_order: list[str] = []
    obj_row = None
    cols: dict[str, int] = {}
    col_names: list[str] = []
    entries: list[tuple[str, int, float]] = []     # (row, col, value)
    rhs: dict[str, float] = {}
    ranges: dict[str, float] = {}
    bounds: list[tuple[str, int, float]] = []
    section = None
    integer = False
    int_cols = set()
    with _open(path) as fh:
        for raw in fh:
            if not raw.strip() or raw.lstrip().startswith("*"):
                continue
            if not raw[0].isspace():                         # section header
                tok = raw.split()
                section = tok[0].upper()
                if section == "NAME":
                    name = tok[1] if len(tok) > 1 else ""
                elif section == "ENDATA":
                    break
                elif section.startswith("OBJSENSE"):
                    # "OBJSENSE MAX" / "OBJSENSE MAXIMIZE" on the header line, or extensions such as OBJSENSE_MAX
                    rest = (section[len("OBJSENSE"):] + " " + " ".join(tok[1:])).upper()
                    if "MAX" in rest:
                        raise ValueError("maximisation problems are not supported (mpsread rejects them as well)")
                    section = "OBJSENSE"
                continue
            tok = raw.split()
            if section == "OBJSENSE":
                if tok[0].upper().startswith("MAX"):
                    raise ValueError("maximisation problems are not supported (mpsread rejects them as well)")
                continue
            if section == "ROWS":
                t, r = tok[0].upper(), tok[1]
                if t == "N":
                    if obj_row is None:
                        obj_row = r
                    row_type[r] = "N"
                else:
                    row_type[r] = t
                    row_order.append(r)
            elif section == "COLUMNS":
                if len(tok) >= 3 and tok[1].upper() == "'MARKER'":
                    integer = "INTORG" in tok[2].upper()
                    continue
                cname = tok[0]
                if cname not in cols:
                    cols[cname] = len(col_names)
                    col_names.append(cname)
                    if integer:
                        int_cols.add(cols[cname])
                for k in range(1, len(tok) - 1, 2):
                    entries.append((tok[k], cols[cname], float(tok[k + 1])))
            elif section == "RHS":
                start = 1 if len(tok) % 2 == 1 else 0       # optional RHS-set name
                for k in range(start, len(tok) - 1, 2):
                    rhs[tok[k]] = float(tok[k + 1])
            elif section == "RANGES":
                start = 1 if len(tok) % 2 == 1 else 0
                for k in range(start, len(tok) - 1, 2):
                    ranges[tok[k]] = float(tok[k + 1])
            elif section == "BOUNDS":
                t = tok[0].upper()
                if t in ("FR", "MI", "PL", "BV"):
                    cname = tok[2] if len(tok) >= 3 else tok[1]
                    val = 0.0
                else:
                    cname, val = (tok[2], float(tok[3])) if len(tok) >= 4 else (tok[1], float(tok[2]))
                if cname not in cols:
                    raise ValueError(f"BOUNDS refers to unknown column {cname}")
                bounds.append((t, cols[cname], val))
    n = len(col_names)
    f = np.zeros(n)
    ridx = {r: i for i, r in enumerate(row_order)}
    m_all = len(row_order)
    rr, cc, vv = [], [], []
    for r, j, v in entries:
        if r == obj_row:
            f[j] += v
        elif r in ridx:
            rr.append(ridx[r]); cc.append(j); vv.append(v)
        elif row_type.get(r) == "N":
            pass                                                # further free rows are ignored, like mpsread
        else:
            raise ValueError(f"COLUMNS refers to unknown row {r}")
    Aall = sp.csr_matrix((vv, (rr, cc)), shape=(m_all, n))
    ball = np.array([rhs.get(r, 0.0) for r in row_order])
    objcon = -rhs.get(obj_row, 0.0) if obj_row is not None else 0.0
    eq_rows, in_rows, in_sign, in_rhs = [], [], [], []
    for i, r in enumerate(row_order):
        t = row_type[r]
        if r in ranges:
            R = ranges[r]
            if t == "L":
                lo, hi = ball[i] - abs(R), ball[i]
            elif t == "G":
                lo, hi = ball[i], ball[i] + abs(R)
            else:                                               # E row: sign of R picks the side
                lo, hi = (ball[i], ball[i] + abs(R)) if R >= 0 else (ball[i] - abs(R), ball[i])
            in_rows += [i, i]; in_sign += [1.0, -1.0]; in_rhs += [hi, -lo]
        elif t == "E":
            eq_rows.append(i)
        elif t == "L":
            in_rows.append(i); in_sign.append(1.0); in_rhs.append(ball[i])
        elif t == "G":
            in_rows.append(i); in_sign.append(-1.0); in_rhs.append(-ball[i])
    Aeq = Aall[eq_rows, :] if eq_rows else sp.csr_matrix((0, n))
    beq = ball[eq_rows] if eq_rows else np.zeros(0)
    if in_rows:
        Aineq = sp.diags(in_sign) @ Aall[in_rows, :]
        bineq = np.array(in_rhs)
    else:
        Aineq, bineq = sp.csr_matrix((0, n)), np.zeros(0)
    lb = np.zeros(n)
    ub = np.full(n, np.inf)
    for t, j, v in bounds:
        if t == "UP":
            ub[j] = v
            if v < 0 and lb[j] == 0:
                lb[j] = -np.inf                                 # the usual MPS convention
        elif t == "LO":
            lb[j] = v
        elif t == "FX":
            lb[j] = ub[j] = v
        elif t == "FR":
            lb[j], ub[j] = -np.inf, np.inf
        elif t == "MI":
            lb[j] = -np.inf
        elif t == "PL":
            ub[j] = np.inf
        elif t == "BV":
            lb[j], ub[j] = 0.0, 1.0
        elif t == "LI":
            lb[j] = v
        elif t == "UI":
            ub[j] = v
        else:
            raise ValueError(f"unknown bound type {t}")
    return GeneralLP(name, f, sp.csr_matrix(Aeq), beq, sp.csr_matrix(Aineq), bineq, lb, ub, objcon, col_names,
                     len(int_cols))


def to_standard_form(g: GeneralLP) -> StandardLP:
    """scripts/bench-lp/preprocess.m:20-80."""
    n = g.f.size
    m1, m2 = g.Aeq.shape[0], g.Aineq.shape[0]
    # preprocess.m:35-37 in MATLAB arithmetic: (lb > -inf) .* lb is 0 * -Inf = NaN for a free variable, NaN becomes -1e6,
    # then -1e8 is added: a free variable is shifted by -1.01e8 (and a NaN bound in the input by -1e6)
    with np.errstate(invalid="ignore"):
        lb = (g.lb > -np.inf) * g.lb
    lb = np.where(np.isnan(lb), -1e6, lb)                        # :36
    lb = lb + (g.lb == -np.inf) * (-1e8)                         # :37
    idxub = g.ub < np.inf
    m3 = int(idxub.sum())
    D = sp.identity(n, format="csr")[np.flatnonzero(idxub), :]
    brhs = g.ub[idxub] - lb[idxub]
    A = _stack(g.Aeq, g.Aineq, D, m1, m2, m3, n)
    b = np.concatenate([g.beq - g.Aeq @ lb, g.bineq - g.Aineq @ lb, brhs])
    c = np.concatenate([g.f, np.zeros(m2 + m3)])
    return StandardLP(A, b, c, float(g.f @ lb) + g.objcon, lb, n, g.name)


def _stack(Aeq, Aineq, D, m1, m2, m3, n) -> sp.csc_matrix:
    blocks = []
    if m1:
        blocks.append(sp.hstack([Aeq, sp.csr_matrix((m1, m2 + m3))], format="csr"))
    if m2:
        blocks.append(sp.hstack([Aineq, sp.identity(m2), sp.csr_matrix((m2, m3))], format="csr"))
    if m3:
        blocks.append(sp.hstack([D, sp.csr_matrix((m3, m2)), sp.identity(m3)], format="csr"))
    if not blocks:
        return sp.csc_matrix((0, n + m2 + m3))
    A = sp.vstack(blocks, format="csc")
    A.sort_indices()
    return A


def drop_empty_rows(s: StandardLP) -> StandardLP:
    """Rows without entries make diag(AA') singular (the reference's preconditioner divides by it): a zero row with
    zero right-hand side is dropped, one with a nonzero right-hand side is infeasible."""
    A = s.A.tocsr()
    nnz_row = np.diff(A.indptr)
    empty = nnz_row == 0
    if not empty.any():
        return s
    if np.any(np.abs(s.b[empty]) > 0):
        raise ValueError("infeasible: empty row with nonzero right-hand side")
    keep = np.flatnonzero(~empty)
    return dataclasses.replace(s, A=A[keep, :].tocsc(), b=s.b[keep])


def write_mps(path, A, b, c, name="ABIPLP", objcon=0.0):
    """Standard-form LP (A x = b, x >= 0) as free-format MPS; used by the tests and the CLI round trip."""
    A = sp.csc_matrix(A)
    m, n = A.shape
    with open(path, "w") as fh:
        fh.write(f"NAME {name}\nROWS\n N COST\n")
        for i in range(m):
            fh.write(f" E R{i}\n")
        fh.write("COLUMNS\n")
        for j in range(n):
            if c[j] != 0 or A.indptr[j] == A.indptr[j + 1]:
                fh.write(f" X{j} COST {float(c[j])!r}\n")
            for k in range(A.indptr[j], A.indptr[j + 1]):
                fh.write(f" X{j} R{A.indices[k]} {float(A.data[k])!r}\n")
        fh.write("RHS\n")
        if objcon:
            fh.write(f" RHS COST {float(-objcon)!r}\n")
        for i in range(m):
            if b[i] != 0:
                fh.write(f" RHS R{i} {float(b[i])!r}\n")
        fh.write("ENDATA\n")


def load_standard_form(path) -> StandardLP:
    return drop_empty_rows(to_standard_form(read_mps(path)))


__all__ = ["GeneralLP", "StandardLP", "read_mps", "to_standard_form", "drop_empty_rows", "write_mps",
           "load_standard_form"]
