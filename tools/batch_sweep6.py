"""GPU experiment: ungated finish / in-flight count after the polling change (median of 3, statistics on the last run)."""
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
    for env in (dict(), dict(ABIP_GPU_BATCH_FINISH_GATE=0), dict(ABIP_GPU_BATCH_FINISH_GATE=0, ABIP_GPU_BATCH_SETUP_GATE=4),
                dict(ABIP_GPU_BATCH_SLOTS=24)):
        for k in ("ABIP_GPU_BATCH_SLOTS", "ABIP_GPU_BATCH_FINISH_GATE", "ABIP_GPU_BATCH_SETUP_GATE"):
            os.environ.pop(k, None)
        for k, v in env.items():
            os.environ[k] = str(v)
        for conc in (160, 192, 224, 296):
            v = []
            for rep in range(3):
                if rep == 2 and conc == 192:
                    os.environ["ABIP_GPU_BATCH_VERBOSE"] = "1"
                t = time.perf_counter()
                res = lp_solve_batch(probs, par, concurrency=min(conc, count))
                v.append(count / (time.perf_counter() - t))
                os.environ.pop("ABIP_GPU_BATCH_VERBOSE", None)
            print("count %5d %-36s conc %3d: median %7.1f  min %7.1f  max %7.1f LP/s  solved %d" % (
                count, " ".join(f"{k[15:]}={v_}" for k, v_ in env.items()), conc, statistics.median(v), min(v), max(v),
                sum(r[3]["status"] == "Solved" for r in res)), flush=True)
