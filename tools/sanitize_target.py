"""Target of the compute-sanitizer tests (tests/test_sanitizer_gpu.py): one solve of a small LP with short and long rows
through the single-problem engine and a small lock-step batch; prints a hash of the solutions."""
import hashlib
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import numpy as np  # noqa: E402
from abip_b200 import problems, lp_solve, lp_solve_batch  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "single"
h = hashlib.sha256()
if mode == "single":
    p = problems.mcf_lp(4, 40, 200, 6, 300, seed=6)
    x, y, s, info = lp_solve(p.csc(), p.b, p.c, dict(tol=1e-3, verbose=0))
    assert info["status_val"] == 1, info
    h.update(np.ascontiguousarray(x).tobytes())
else:
    probs = [problems.random_lp(40, 120, 3, seed=80 + i) for i in range(4)]
    for x, y, s, info in lp_solve_batch(probs, dict(tol=1e-3, verbose=0), concurrency=4, ctas_per_problem=1):
        assert info["status_val"] == 1, info
        h.update(np.ascontiguousarray(x).tobytes())
print("OK", h.hexdigest()[:16])
