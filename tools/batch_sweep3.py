"""GPU experiment: problems in flight / set-up gate with the device-resident outer loop, median of 3 runs."""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abip_b200 import lp_solve_batch, problems  # noqa: E402

counts = [int(x) for x in sys.argv[1:]] or [512]
allp = [problems.random_lp(500, 2000, 5, seed=5000 + i) for i in range(max(counts))]
par = dict(tol=1e-4, verbose=0)
lp_solve_batch(allp[:32], par, concurrency=32)
for count in counts:
    probs = allp[:count]
    for gate in (3, 8, 32):
        os.environ["ABIP_GPU_BATCH_SETUP_GATE"] = str(gate)
        for conc in (148, 176, 200, 240, 296):
            v = []
            for _ in range(3):
                t = time.perf_counter()
                res = lp_solve_batch(probs, par, concurrency=min(conc, count))
                v.append(count / (time.perf_counter() - t))
            print("count %5d gate %2d conc %3d: median %7.1f  min %7.1f  max %7.1f LP/s  solved %d" % (
                count, gate, conc, statistics.median(v), min(v), max(v), sum(r[3]["status"] == "Solved" for r in res)), flush=True)
