"""GPU experiment: knobs of the batch executor (accumulation window, launcher slots, problems in flight, set-up gate) at cfg5."""
import itertools
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abip_b200 import lp_solve_batch, problems  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 512
probs = [problems.random_lp(500, 2000, 5, seed=5000 + i) for i in range(count)]
par = dict(tol=1e-4, verbose=0)
lp_solve_batch(probs[:32], par, concurrency=32)  # context + module load


def run(conc, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    best = 0.0
    for _ in range(2):
        t = time.perf_counter()
        res = lp_solve_batch(probs, par, concurrency=conc)
        dt = time.perf_counter() - t
        best = max(best, count / dt)
    ok = sum(r[3]["status"] == "Solved" for r in res)
    print("conc %4d %-70s %7.1f LP/s (best of 2)  solved %d" % (conc, " ".join(f"{k[9:]}={v}" for k, v in env.items()), best, ok), flush=True)


base = dict(ABIP_GPU_BATCH_WAIT_US=200, ABIP_GPU_BATCH_SLOTS=6, ABIP_GPU_BATCH_SETUP_GATE=3)
run(296, **base)
for w in (0, 50, 100, 400, 800):
    run(296, **{**base, "ABIP_GPU_BATCH_WAIT_US": w})
for s in (2, 3, 4, 8, 12, 16):
    run(296, **{**base, "ABIP_GPU_BATCH_SLOTS": s})
for c in (148, 200, 240, 360, 444, 512):
    run(min(c, count), **base)
for w, s, c in ((100, 8, 360), (100, 12, 444), (400, 4, 296), (50, 12, 296), (100, 8, 240)):
    run(min(c, count), **{**base, "ABIP_GPU_BATCH_WAIT_US": w, "ABIP_GPU_BATCH_SLOTS": s})
