#!/usr/bin/env python
"""bench.py -- ADMM iterations/sec and time-to-1e-4 of the ABIP hot path (BASELINE.json metric), all five configs.

  python bench.py --gpus N --steps K --warmup W                       # our arm (GPU engine)
  python bench.py --impl reference --gpus N --steps K --warmup W      # the reference's own C solver on the host cores

Headline (`metric`, `value`, `e2e`, `roofline`): BASELINE.json configs[1], "cfg2" -- one complete solve to eps = 1e-4 of
the synthetic multicommodity-flow LP m=200k, n=1M, nnz=5M per step.  `value` times abip_gpu_solve with A, b, c resident in
HBM (CUDA events); `e2e` times the reference-facing entry abip_gpu_main (init + solve + finish) from HOST buffers, i.e.
CSR build, ordering, H2D copies, device equilibration, solve and the D2H of x, y, s.  For N > 1 every rank solves the
SAME instance (replicas; weak scaling, no data-path collective).

Further blocks of the same JSON line (BASELINE.json configs[2..4]; each measured with CUDA events / device-synchronised
host clocks, max over ranks):
  cfg3_qcp      N = 1: ABIP-QCP, SOCP with 10k second-order cones, n=500k, nnz(A)=10M + Q          (it/s, time, roofline)
  cfg4          m=2M n=10M nnz=60M.  N = 1: single-GPU solve; N > 1: ONE instance sharded over the N GPUs (column blocks,
                in-kernel NVLink peer-memory exchange), with set-up time and a parity block against the single-GPU solve
  cfg5_batch    512 independent LPs (m=500, n=2000) per GPU = 4,096 on 8 GPUs, sharded one problem set per GPU (LP/s)
  sharded_single_instance (N > 1): cfg2 as ONE instance over the N GPUs + parity against the replica solve
The reference arm runs a bounded sample of the cfg2 solve with all host threads (OMP_NUM_THREADS set explicitly).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("ABIP_BENCH_SCALE", "1.0")),
                    help="shrink every config (debug only; the reported configs are scale=1)")
    ap.add_argument("--eps", type=float, default=1e-4)
    ap.add_argument("--cpu-sample-iters", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only", default=os.environ.get("ABIP_BENCH_ONLY", ""),
                    help="comma list of extra blocks to run (cfg3,cfg4,cfg5,sharded); default: all")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes of one k_bb_round launch, from the committed `ncu --set full` capture of the shipped build (profiles/):
    a constant taken from that capture, NOT a measurement of this run."""
    path = os.path.join(ROOT, "profiles", "r02_k_bb_round_ncu.txt")
    if not os.path.exists(path):
        return None
    rd = wr = None
    for ln in open(path):
        if ln.startswith("dram__bytes_read.sum"):
            v = float(ln.split("=")[1]); rd = v * (1e9 if "Gbyte" in ln else 1e6)
        if ln.startswith("dram__bytes_write.sum"):
            v = float(ln.split("=")[1]); wr = v * (1e9 if "Gbyte" in ln else 1e6)
    if rd is None or wr is None:
        return None
    return {"dram_bytes_per_launch": rd + wr,
            "note": "from profiles/r02_k_bb_round_ncu.txt: ncu --set full of ONE k_bb_round launch of the shipped build "
                    "(two solves, 47 CG iterations, ~10.4 GB algorithmic); not re-measured in this run"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


L2_NOTE = ("GPU arm: 256 MiB buffer written between timed steps (L2 flush), matrices (120 MB) ~ L2 size; "
           "reference (CPU) arm: not applicable")


def bench_config(name):
    """identical in both arms (the driver compares the `config` objects)"""
    return {"workload": name, "l2": L2_NOTE}


def workload(args):
    from abip_b200 import problems
    p = problems.cfg2(seed=2, scale=args.scale)
    name = (f"cfg2: ABIP-LP synthetic multicommodity-flow LP m={p.m} n={p.n} nnz={p.nnz}, eps={args.eps:g}"
            + ("" if args.scale == 1.0 else f" (DEBUG scale={args.scale})"))
    return p, name


# ---------------------------------------------------------------------------------------------------------------
# reference arm / CPU baselines (the only places that execute oracle/)
# ---------------------------------------------------------------------------------------------------------------
def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def _ref_kind():
    """Prefer the OpenMP build of the reference (only the SpMV loop is parallel, linsys/common.c:620-622).  torchrun
    exports OMP_NUM_THREADS=1: the thread count is set EXPLICITLY here (environment before the library is loaded, and
    omp_set_num_threads afterwards)."""
    cores = _host_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import ref_lp
    if ref_lp.available("indirect_omp"):
        try:
            import ctypes
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(cores))
        except Exception:  # noqa: BLE001
            pass
        return "indirect_omp", cores, "reference indirect.c + OpenMP SpMV"
    if ref_lp.available("indirect"):
        return "indirect", 1, "reference indirect.c, single thread as shipped"
    return None, 0, "oracle/_ref not built"


def _quiet(fn):
    """The reference's C code prints progress lines to stdout: keep stdout to the one JSON line of the contract."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        return fn()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def _ref_sample(args, p, which, iters):
    """Bounded sample of the cfg2 solve on the host cores: the first `iters` ADMM iterations with the Barzilai-Borwein
    searches that run between them.  The complete solve (196 ADMM iterations) is pinned by
    tests/golden/lp_golden_large.json["cfg2_full"] (451 s on 6 threads of the build container = 0.43 it/s)."""
    from oracle import ref_lp
    return _quiet(lambda: ref_lp.solve(p, which=which, eps=args.eps, max_admm_iters=iters + 1))


def _full_solve_fixture():
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "lp_golden_large.json")))["cfg2_full"]
        return {"admm_iter": g["admm_iter"], "solve_time_s": g["solve_time_ms"] / 1e3, "threads": g.get("threads"),
                "iter_per_s": g["admm_iter"] / (g["solve_time_ms"] / 1e3),
                "note": "complete reference solve of this workload, run once in the build container (tests/golden/make_golden_large.py)"}
    except Exception:  # noqa: BLE001
        return None


def run_reference(args):
    """Reference arm: the reference's own CPU implementation (oracle/_ref, compiled unmodified from its sources) with all
    the host threads it can use.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    which, cores, desc = _ref_kind()
    if which is None:
        print(json.dumps({"impl": "reference", "unavailable": desc}))
        return
    p, name = workload(args)
    its_total, ms_total, per_step = 0, 0.0, []
    for step in range(args.warmup + args.steps):
        r = _ref_sample(args, p, which, 1 if step < args.warmup else args.cpu_sample_iters)
        if step >= args.warmup:
            its_total += r["admm_iter"]
            ms_total += r["solve_time_ms"]
            per_step.append(r["solve_time_ms"])
    value = its_total / (ms_total / 1e3)
    sample = (f"first {args.cpu_sample_iters} ADMM iterations (incl. BB searches) of the cfg2 solve per step "
              f"(warm-up steps: 1 iteration); {desc}; {cores} thread(s)")
    line = {"impl": "reference", "metric": "ADMM iters/sec", "value": value, "unit": "iter/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(per_step)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(name),
            "cpu_baseline": {"value": value, "unit": "iter/s", "cores": cores, "kind": "reference", "sample": sample,
                             "full_solve_fixture": _full_solve_fixture()},
            "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline(args, p):
    which, cores, desc = _ref_kind()
    if which is None:
        return {"value": None, "unit": "iter/s", "cores": 0, "kind": "reference", "sample": desc}
    try:
        r = _ref_sample(args, p, which, args.cpu_sample_iters)
        return {"value": r["admm_iter"] / (r["solve_time_ms"] / 1e3), "unit": "iter/s", "cores": cores,
                "kind": "reference",
                "sample": f"first {args.cpu_sample_iters} ADMM iterations (incl. BB searches) of the same cfg2 solve; "
                          f"{desc}; {r['solve_time_ms'] / 1e3:.1f} s of CPU work",
                "full_solve_fixture": _full_solve_fixture()}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "unit": "iter/s", "cores": cores, "kind": "reference", "sample": f"failed: {ex}"}


def _cfg5_ref_one(seed):
    from abip_b200 import problems
    from oracle import ref_lp
    p = problems.random_lp(500, 2000, 5, seed=seed, name="cfg5")
    r = ref_lp.solve(p, which="indirect", eps=1e-4)
    return r["admm_iter"]


def cfg5_cpu_baseline(count):
    """the reference looped over the problems, one process per core"""
    import multiprocessing as mp
    cores = _host_cores()
    n = max(cores, min(count, 4 * cores))
    try:
        def run_pool():  # fork INSIDE the redirection: the children inherit stderr as their fd 1
            ctx = mp.get_context("fork")
            with ctx.Pool(cores) as pool:
                t0 = time.perf_counter()
                pool.map(_cfg5_ref_one, [5000 + i for i in range(n)])
                return time.perf_counter() - t0
        dt = _quiet(run_pool)
        return {"value": n / dt, "unit": "LP/s", "cores": cores, "kind": "reference",
                "sample": f"{n} of the cfg5 problems, reference ABIP(main) looped, one process per core"}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "unit": "LP/s", "cores": cores, "kind": "reference", "sample": f"failed: {ex}"}


def cfg3_cpu_baseline(args):
    """reference direct path (QDLDL; its PCG path does not run) at the largest scale the factorisation
    survives in seconds"""
    try:
        from abip_b200 import problems
        from oracle import ref_qcp
        sc = 0.01 * min(1.0, args.scale)
        q = problems.cfg3(scale=sc)
        t0 = time.perf_counter()
        r = _quiet(lambda: ref_qcp.solve(q, eps_p=args.eps, eps_d=args.eps, eps_g=args.eps))
        dt = time.perf_counter() - t0
        st = r.get("solve_time_ms", dt * 1e3) / 1e3
        return {"value": r["admm_iter"] / max(st, 1e-9), "unit": "iter/s", "cores": 1, "kind": "reference",
                "sample": f"cfg3 at scale {sc:g} (n={q.n}, nnz(A)={q.A.nnz}): reference abip() with linsys_solver=1 (QDLDL), "
                          f"{r['admm_iter']} ADMM iterations, set-up {r.get('setup_time_ms', 0) / 1e3:.1f} s + solve {st:.2f} s; "
                          "the full size does not factor in reasonable time"}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "unit": "iter/s", "cores": 1, "kind": "reference", "sample": f"failed: {ex}"}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from abip_b200 import LpSolver, lp_solve, problems

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    os.environ["ABIP_GPU_DEVICE"] = str(local)
    dev = torch.device("cuda", local)
    only = set(x for x in args.only.split(",") if x) or {"cfg3", "cfg4", "cfg5", "sharded"}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item()

    def allsum(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t[0].item()

    p, name = workload(args)
    A = p.csc()
    params = dict(tol=args.eps, verbose=0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- cfg2, device-resident -------------------------------------------------------------------------------
    solver = LpSolver(A, params)
    for _ in range(args.warmup):
        solver.solve(p.b, p.c)
        flush.fill_(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    its = 0
    event_ms, per_step = 0.0, []
    agg = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        x, y, s, info = solver.solve(p.b, p.c)
        its += info["admm_iter"]
        event_ms += info["stats"]["solve_event_ms"]
        per_step.append(info["stats"]["solve_event_ms"])
        for k, v in info["stats"].items():
            agg[k] = agg.get(k, 0) + v
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    last = info
    x_single, pobj_single, its_single = x.copy(), info["pobj"], info["admm_iter"]
    solver.close()

    # ---- cfg2, end to end through the reference-facing entry with host buffers -------------------------------
    del flush
    torch.cuda.empty_cache()
    barrier()
    e2e_its, e2e_s, h2d, d2h = 0, 0.0, 0.0, 0.0
    lp_solve(A, p.b, p.c, params)  # untimed warm-up of this code path
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        t1 = time.perf_counter()
        x, y, s, inf2 = lp_solve(A, p.b, p.c, params, want_stats=True)
        e2e_s += time.perf_counter() - t1
        e2e_its += inf2["admm_iter"]
        h2d = inf2["stats"]["h2d_bytes"] + 12.0 * p.nnz * 2 + 4.0 * (p.m + p.n + 2)  # + CSR(A), CSR(A') uploads (init)
        d2h = inf2["stats"]["d2h_bytes"]
    barrier()
    event_ms_max, its_all = allmax(event_ms), allsum(its)
    e2e_max, e2e_its_all = allmax(e2e_s), allsum(e2e_its)

    extra = {}

    def guarded(key, fn):
        try:
            extra[key] = fn()
        except Exception as ex:  # noqa: BLE001
            extra[key] = {"error": f"{type(ex).__name__}: {ex}"}
        barrier()

    # ---- N > 1: cfg2 as ONE instance sharded over the GPUs ----------------------------------------------------
    def block_sharded(pp, label, x_ref, pobj_ref, its_ref):
        from abip_b200.dist import LpSolverDist
        barrier()
        t_s = time.perf_counter()
        ds = LpSolverDist(pp.csc(), params)
        torch.cuda.synchronize()
        setup_s = allmax(time.perf_counter() - t_s)
        ds.solve(pp.b, pp.c)  # warm-up
        barrier()
        sh_ms, sh_its = 0.0, 0
        nst = max(1, min(args.steps, 3))
        for _ in range(nst):
            xs, ys, ss, si = ds.solve(pp.b, pp.c)
            sh_ms += si["stats"]["solve_event_ms"]
            sh_its += si["admm_iter"]
        barrier()
        ds.close()
        sh_ms = allmax(sh_ms)
        used = si.get("gpus_used", world)
        out = {"workload": f"{label} as ONE instance: {world} GPUs offered, {used} used (column blocks of A; abip_b200/dist.py refuses to "
                           f"shard below {4_000_000:,} nonzeros per GPU or over more than 4 GPUs, see profiles/r02_multi_gpu.md)",
               "value": sh_its / (sh_ms / 1e3), "unit": "iter/s", "scaling": "strong",
               "time_to_1e-4_s": sh_ms / nst / 1e3, "setup_s": setup_s, "status": si["status"],
               "admm_iter_per_solve": sh_its / nst, "gpus_used": si.get("gpus_used", world),
               "collective": "in-kernel peer-memory sum of the m-vector A_g x_g + scalar blocks (no NCCL in the solve)"}
        if x_ref is not None:
            out["parity"] = {"against": "single-GPU engine, same instance, same run",
                             "admm_iter_equal": bool(si["admm_iter"] == its_ref),
                             "admm_iter": [int(si["admm_iter"]), int(its_ref)],
                             "pobj_rel": float(abs(si["pobj"] - pobj_ref) / (abs(pobj_ref) + 1e-300)),
                             "x_rel": float(np.max(np.abs(xs - x_ref)) / (np.max(np.abs(x_ref)) + 1e-300))}
        return out

    if world > 1 and "sharded" in only:
        guarded("sharded_single_instance", lambda: block_sharded(p, "cfg2", x_single, pobj_single, its_single))

    # ---- cfg4: m=2M n=10M nnz=60M ------------------------------------------------------------------------------
    def block_cfg4():
        p4 = problems.cfg4(seed=4, scale=args.scale)
        A4 = p4.csc()
        t_s = time.perf_counter()
        s4 = LpSolver(A4, params)
        torch.cuda.synchronize()
        setup1 = time.perf_counter() - t_s
        x4, y4, ss4, i4 = s4.solve(p4.b, p4.c)
        ms1, it1 = i4["stats"]["solve_event_ms"], i4["admm_iter"]
        achieved = i4["stats"]["alg_bytes"] / (ms1 * 1e-3) / 1e9
        s4.close()
        out = {"workload": f"cfg4: ABIP-LP synthetic multicommodity-flow LP m={p4.m} n={p4.n} nnz={p4.nnz}, eps={args.eps:g}",
               "single_gpu": {"value": it1 / (ms1 / 1e3), "unit": "iter/s", "time_to_1e-4_s": ms1 / 1e3, "setup_s": setup1,
                              "status": i4["status"], "admm_iter": it1, "pres": i4["pres"], "dres": i4["dres"], "gap": i4["gap"],
                              "achieved_GBps": achieved, "roofline_frac": achieved / peaks()[0]}}
        if world > 1:
            sh = block_sharded(p4, "cfg4", x4, i4["pobj"], it1)
            sh["speedup_vs_single_gpu"] = (ms1 / 1e3) / sh["time_to_1e-4_s"]
            out["sharded"] = sh
        return out

    if "cfg4" in only:
        guarded("cfg4", block_cfg4)

    # ---- cfg5: 512 LPs per GPU ---------------------------------------------------------------------------------
    def block_cfg5():
        from abip_b200 import lp_solve_batch
        per_gpu = max(8, int(round(512 * min(1.0, args.scale))))
        probs = [problems.random_lp(500, 2000, 5, seed=5000 + rank * per_gpu + i, name=f"cfg5_lp_{rank * per_gpu + i}")
                 for i in range(per_gpu)]
        lp_solve_batch(probs[:192], dict(tol=args.eps, verbose=0), concurrency=min(per_gpu, 192))  # warm-up: one wave
        barrier()
        t_b = time.perf_counter()
        res = lp_solve_batch(probs, dict(tol=args.eps, verbose=0), concurrency=min(per_gpu, 192))
        torch.cuda.synchronize()
        dt = allmax(time.perf_counter() - t_b)
        solved = allsum(sum(1 for r in res if r[3]["status"] == "Solved"))
        out = {"workload": f"cfg5: {per_gpu} independent LPs m=500 n=2000 per GPU ({per_gpu * world} in total), eps={args.eps:g}",
               "value": per_gpu * world / dt, "unit": "LP/s", "scaling": "weak", "wall_s": dt, "solved": int(solved),
               "engine_wall_s": allmax(res[0][3]["batch_wall_s"]),  # inside abip_gpu_batch_main (without the Python marshalling)
               "total": per_gpu * world, "timing": "host clock around lp_solve_batch = abip_gpu_batch_main + marshalling (set-up, H2D, solves, D2H) of ONE batch, max over ranks; warm-up: one wave of 192 problems",
               "in_flight": min(per_gpu, 192)}
        try:  # parity of a slice against the reference's own results (fixture: first 64 problems)
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg5_golden.json")))
            if rank == 0:
                nchk = min(len(gold), per_gpu)
                bad = sum(1 for g, r in zip(gold[:nchk], res[:nchk])
                          if r[3]["status"] != g["status"] or abs(r[3]["admm_iter"] - g["admm_iter"]) > max(2, 0.05 * g["admm_iter"])
                          or abs(float(probs[gold.index(g)].c @ r[0]) - g["pobj"]) > 1e-6 * abs(g["pobj"]) + 2e-8)
                out["parity"] = {"against": "tests/golden/cfg5_golden.json (compiled reference)", "checked": nchk, "mismatches": bad}
        except Exception as ex:  # noqa: BLE001
            out["parity"] = {"error": str(ex)}
        if rank == 0 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cfg5_cpu_baseline(per_gpu)
        return out

    if "cfg5" in only:
        guarded("cfg5_batch", block_cfg5)

    # ---- cfg3: ABIP-QCP (N = 1) --------------------------------------------------------------------------------
    def block_cfg3():
        from abip_b200.qcp import qcp_solve_raw
        q = problems.cfg3(seed=3, scale=args.scale)
        kw = dict(eps_p=args.eps, eps_d=args.eps, eps_g=args.eps, verbose=0)
        qcp_solve_raw(q.A, q.Q, q.b, q.c, q.K, **kw)  # warm-up
        tq = time.perf_counter()
        xq, yq, sq, iq = qcp_solve_raw(q.A, q.Q, q.b, q.c, q.K, **kw)
        e2e_q = time.perf_counter() - tq
        class _V:  # counters of the last solve (abip_qcp_gpu_last_counters, read by qcp_solve_raw)
            def __init__(self, v):
                self.value = v
        n_it, n_cg, n_in, kms = _V(iq["n_iter"]), _V(iq["n_cg"]), _V(iq["n_inner"]), _V(iq["kernel_ms"])
        V, I = 8.0, 4.0
        nnzA, nnzQ = q.A.nnz, (q.Q.nnz if q.Q is not None else 0)
        B_A = nnzA * (V + I) + (q.m + 1) * I + q.n * V + q.m * V
        B_AT = nnzA * (V + I) + (q.n + 1) * I + q.m * V + q.n * V
        B_Q = nnzQ * (V + I) + (q.n + 1) * I + 2 * q.n * V
        # per outer (Schur) CG iteration: A' pass + A pass + 12 m-vector passes; per inner (H^-1) CG iteration: Q pass +
        # 12 n-vector passes; per ADMM iteration: one more (A, A', Q) triple for the convergence check
        alg = n_cg.value * (B_A + B_AT + 12 * q.m * V) + n_in.value * (B_Q + 12 * q.n * V) + n_it.value * (B_A + B_AT + B_Q + 30 * (q.m + q.n) * V)
        solve_s = iq["solve_time_ms"] / 1e3
        ach = alg / max(kms.value * 1e-3, 1e-9) / 1e9
        Qx = q.Q @ xq if q.Q is not None else 0.0
        pres = float(np.max(np.abs(q.A @ xq - q.b)) / (1 + max(np.max(np.abs(q.A @ xq)), np.max(np.abs(q.b)))))
        out = {"workload": f"cfg3: ABIP-QCP SOCP, {len(q.K.get('q', []))} second-order cones, n={q.n}, m={q.m}, nnz(A)={nnzA}, nnz(Q)={nnzQ}, eps={args.eps:g}",
               "value": iq["admm_iter"] / solve_s, "unit": "iter/s", "time_to_1e-4_s": solve_s, "setup_s": iq["setup_time_ms"] / 1e3,
               "e2e_s": e2e_q, "status": iq["status"], "admm_iter": iq["admm_iter"], "ipm_iter": iq["ipm_iter"],
               "outer_cg_iters": n_cg.value, "inner_cg_iters": n_in.value, "pres_recomputed": pres,
               "roofline": {"bound": "hbm", "kernel": "k_qcp_iter", "achieved": ach, "peak": peaks()[0], "unit": "GB/s",
                            "frac": ach / peaks()[0], "kernel_ms": kms.value,
                            "traffic": None, "note": "algorithmic bytes from the engine's CG counters (bench.py block_cfg3)"}}
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cfg3_cpu_baseline(args)
        return out

    if world == 1 and "cfg3" in only:
        guarded("cfg3_qcp", block_cfg3)

    if rank == 0:
        peak, peak_src = peaks()
        admm_ms, bb_ms = agg["admm_kernel_ms"], agg["bb_kernel_ms"]
        ach_admm = agg["alg_bytes_admm"] / (admm_ms * 1e-3) / 1e9 if admm_ms > 0 else None
        ach_bb = agg["alg_bytes_bb"] / (bb_ms * 1e-3) / 1e9 if bb_ms > 0 else None
        line = {
            "metric": "ADMM iters/sec", "value": its_all / (event_ms_max / 1e3), "unit": "iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": event_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(name),
            "run": {"per_gpu": "the same instance on every GPU (replicas, no collective)" if world > 1 else "single instance",
                    "status": last["status"], "admm_iter_per_solve": its / args.steps,
                    "ipm_iter": last["ipm_iter"], "pres": last["pres"], "dres": last["dres"], "gap": last["gap"],
                    "pobj": last["pobj"]},
            "time_to_1e-4_s": event_ms_max / args.steps / 1e3,
            "host_wall_s_per_step": wall / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_its_all / e2e_max, "unit": "iter/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "time_to_1e-4_s": e2e_max / n_e2e,
                    "note": "abip_gpu_main from host buffers: CSR build + ordering + H2D + device equilibration + solve + D2H of x, y, s"},
            "gpu_launches": int(agg["n_kernel_launches"]),
            # dominant kernel: k_bb_round (one Barzilai-Borwein lookback round = two solves), ~82 % of the step
            "roofline": {"bound": "hbm", "kernel": "k_bb_round", "achieved": ach_bb, "peak": peak, "unit": "GB/s",
                         "frac": (ach_bb / peak) if ach_bb else None, "traffic": None, "peak_source": peak_src,
                         "launches": int(agg["n_bb_launch"]), "avg_launch_ms": bb_ms / max(1, agg["n_bb_launch"]),
                         "k_admm_iter": {"achieved": ach_admm, "frac": (ach_admm / peak) if ach_admm else None,
                                         "launches": int(agg["n_admm_launch"]),
                                         "avg_launch_ms": admm_ms / max(1, agg["n_admm_launch"])},
                         "share_of_step": {"k_admm_iter": admm_ms / event_ms, "k_bb_round": bb_ms / event_ms}},
            "counters": {"cg_iters": int(agg["n_cg_iters"]), "solves": int(agg["n_solves"]),
                         "spmv_A": int(agg["n_spmv_A"]), "spmv_AT": int(agg["n_spmv_AT"])},
        }
        line.update(extra)
        tr = ncu_traffic()
        if tr is not None:
            line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
            line["roofline"]["traffic_note"] = tr["note"]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, p)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
