"""GPU experiment: device-resident outer loop of the batch engine (default) against the host outer loop
(ABIP_GPU_BATCH_HOST_OUTER=1), and the knobs that matter once a problem is one launch."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abip_b200 import lp_solve_batch, problems  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 512
probs = [problems.random_lp(500, 2000, 5, seed=5000 + i) for i in range(count)]
par = dict(tol=1e-4, verbose=0)
lp_solve_batch(probs[:32], par, concurrency=32)  # context + module load
ref = None


def run(conc, **env):
    global ref
    for k in ("ABIP_GPU_BATCH_HOST_OUTER",):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    best, res = 0.0, None
    for _ in range(2):
        t = time.perf_counter()
        res = lp_solve_batch(probs, par, concurrency=conc)
        dt = time.perf_counter() - t
        best = max(best, count / dt)
    sig = [(r[3]["status"], r[3]["admm_iter"], round(r[3]["pobj"], 9)) for r in res]
    if ref is None:
        ref = sig
    same = sum(a == b for a, b in zip(sig, ref))
    print("conc %4d %-60s %7.1f LP/s  solved %d  identical to first run %d/%d" % (
        conc, " ".join(f"{k[9:]}={v}" for k, v in env.items()), best, sum(s[0] == "Solved" for s in sig), same, count), flush=True)


run(296, ABIP_GPU_BATCH_HOST_OUTER=1)
run(296)
for g in (1, 6, 12, 32):
    run(296, ABIP_GPU_BATCH_SETUP_GATE=g)
os.environ["ABIP_GPU_BATCH_SETUP_GATE"] = "3"
for c in (148, 200, 360, 444, 512):
    run(min(c, count))
for w_, s_ in ((0, 6), (50, 6), (200, 2), (200, 12), (1000, 6)):
    run(296, ABIP_GPU_BATCH_WAIT_US=w_, ABIP_GPU_BATCH_SLOTS=s_)
