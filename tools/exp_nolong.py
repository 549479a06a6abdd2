"""GPU experiment: pass times of cfg2 with and without the rows longer than one chunk (the 2,400 random side constraints).
Prints the per-CTA section times of the PCG loop measured by k_tune (ABIP_GPU_TUNE_VERBOSE=2)."""
import os
import sys

import numpy as np

os.environ["ABIP_GPU_TUNE"] = "1"
os.environ["ABIP_GPU_TUNE_ROUNDS"] = "0"
os.environ["ABIP_GPU_TUNE_VERBOSE"] = "2"
os.environ.setdefault("ABIP_GPU_TUNE_REPS", "9")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abip_b200 import LpSolver, problems  # noqa: E402

p = problems.cfg2()
A = p.csc()
for label, mat in (("full", A), ("no long rows", None)):
    if mat is None:
        R = A.tocsr()
        keep = np.diff(R.indptr) <= 252
        mat = R[keep].tocsc()
        mat.sort_indices()
    print("==== %s: m=%d n=%d nnz=%d" % (label, mat.shape[0], mat.shape[1], mat.nnz), flush=True)
    sys.stderr.write("==== %s\n" % label)
    sys.stderr.flush()
    s = LpSolver(mat, dict(tol=1e-4, verbose=1))
    s.close()
