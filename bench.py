#!/usr/bin/env python
"""bench.py -- ADMM iterations/sec and time-to-1e-4 of the ABIP-LP hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W              # our arm (GPU engine)
  python bench.py --impl reference --gpus N --steps K --warmup W   # reference's own C solver on the host cores

A "step" is one complete solve to eps = 1e-4 of one synthetic LP of BASELINE.json configs[1]
(m=200k, n=1M, nnz=5M, multicommodity-flow structure).  `value` times abip_gpu_solve with A, b, c already
resident in HBM (abip_gpu_init done before the timed region); `e2e` times the reference-facing entry
abip_gpu_main (init + solve + finish) from HOST buffers, i.e. including equilibration, CSR build, all H2D
copies and the D2H of x, y, s.  For N > 1 each rank solves its own instance (independent LPs sharded one per
GPU, no data-path collective): weak scaling.
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
                    help="shrink cfg2 (debug only; the reported config is scale=1)")
    ap.add_argument("--eps", type=float, default=1e-4)
    ap.add_argument("--cpu-sample-iters", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes of one k_admm_iter launch from the committed ncu --set full capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r01_k_admm_iter_ncu_v2.txt")
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
            "note": "ncu --set full capture of one k_admm_iter launch with 26 CG iterations (~5.0 GB algorithmic), "
                    "profiles/r01_k_admm_iter_ncu_v2.txt"}


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


def workload(args, rank):
    from abip_b200 import problems
    p = problems.cfg2(seed=2 + rank, scale=args.scale)
    name = (f"cfg2: ABIP-LP synthetic multicommodity-flow LP m={p.m} n={p.n} nnz={p.nnz}, eps={args.eps:g}"
            + ("" if args.scale == 1.0 else f" (DEBUG scale={args.scale})"))
    return p, name


def _ref_kind():
    """Prefer the OpenMP build of the reference (only the SpMV loop is parallel, linsys/common.c:620-622)."""
    from oracle import ref_lp
    cores = os.cpu_count() or 1
    if ref_lp.available("indirect_omp"):
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        return "indirect_omp", int(os.environ["OMP_NUM_THREADS"]), "reference indirect.c + OpenMP SpMV"
    if ref_lp.available("indirect"):
        return "indirect", 1, "reference indirect.c, single thread as shipped"
    return None, 0, "oracle/_ref not built"


def _ref_sample(args, p, which):
    """Bounded sample of the cfg2 solve on the host cores: the first `cpu_sample_iters` ADMM iterations, with the
    Barzilai-Borwein searches that run between them (early iterations use the loosest CG tolerance, so this
    over-states the CPU's average iteration rate)."""
    from oracle import ref_lp
    # the reference's C code prints progress lines to stdout ("Done the pc rescaling!"): keep stdout to the one JSON
    # line of the contract by pointing fd 1 at stderr while it runs
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        return ref_lp.solve(p, which=which, eps=args.eps, max_admm_iters=args.cpu_sample_iters + 1)
    finally:
        os.dup2(saved, 1)
        os.close(saved)


def run_reference(args):
    """Reference arm: the reference's own CPU implementation (oracle/_ref, compiled unmodified from its sources)
    with all the host threads it can use.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    which, cores, desc = _ref_kind()
    if which is None:
        print(json.dumps({"impl": "reference", "unavailable": desc}))
        return
    p, name = workload(args, 0)
    its_total, ms_total, per_step = 0, 0.0, []
    for step in range(args.warmup + args.steps):
        r = _ref_sample(args, p, which)
        if step >= args.warmup:
            its_total += r["admm_iter"]
            ms_total += r["solve_time_ms"]
            per_step.append(r["solve_time_ms"])
    value = its_total / (ms_total / 1e3)
    sample = (f"first {args.cpu_sample_iters} ADMM iterations (incl. BB searches) of the cfg2 solve per step; {desc}; "
              f"{cores} thread(s)")
    line = {"impl": "reference", "metric": "ADMM iters/sec", "value": value, "unit": "iter/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(per_step)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name},
            "cpu_baseline": {"value": value, "unit": "iter/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline(args, p):
    which, cores, desc = _ref_kind()
    if which is None:
        return {"value": None, "unit": "iter/s", "cores": 0, "kind": "reference", "sample": desc}
    try:
        r = _ref_sample(args, p, which)
        return {"value": r["admm_iter"] / (r["solve_time_ms"] / 1e3), "unit": "iter/s", "cores": cores,
                "kind": "reference",
                "sample": f"first {args.cpu_sample_iters} ADMM iterations (incl. BB searches) of the same cfg2 solve; "
                          f"{desc}; {r['solve_time_ms'] / 1e3:.1f} s of CPU work"}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "unit": "iter/s", "cores": cores, "kind": "reference", "sample": f"failed: {ex}"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from abip_b200 import LpSolver, lp_solve

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    os.environ["ABIP_GPU_DEVICE"] = str(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    p, name = workload(args, rank)
    A = p.csc()
    params = dict(tol=args.eps, verbose=0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    solver = LpSolver(A, params)
    for _ in range(args.warmup):
        solver.solve(p.b, p.c)
        flush.fill_(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    its = 0
    solve_ms, event_ms, per_step = 0.0, 0.0, []
    agg = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        x, y, s, info = solver.solve(p.b, p.c)
        its += info["admm_iter"]
        solve_ms += info["solve_time_ms"]
        event_ms += info["stats"]["solve_event_ms"]
        per_step.append(info["stats"]["solve_event_ms"])
        for k, v in info["stats"].items():
            agg[k] = agg.get(k, 0) + v
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    last = info
    solver.close()

    # end-to-end through the reference-facing entry with host buffers (init + solve + finish per step)
    del flush
    torch.cuda.empty_cache()
    barrier()
    e2e_its, e2e_s, h2d, d2h = 0, 0.0, 0.0, 0.0
    lp_solve(A, p.b, p.c, params)  # untimed warm-up of this code path
    for _ in range(max(1, min(args.steps, 2))):
        t1 = time.perf_counter()
        x, y, s, inf2 = lp_solve(A, p.b, p.c, params, want_stats=True)
        e2e_s += time.perf_counter() - t1
        e2e_its += inf2["admm_iter"]
        h2d = inf2["stats"]["h2d_bytes"] + 12.0 * p.nnz * 2 + 4.0 * (p.m + p.n + 2)  # + CSR(A), CSR(A') uploads
        d2h = inf2["stats"]["d2h_bytes"]
    barrier()

    # N > 1: the same workload as ONE instance sharded over the GPUs (column blocks of A, in-kernel NVLink
    # peer-memory all-reduce; abip_b200/dist.py) -- strong scaling, reported beside the weak-scaling value
    sharded = None
    if world > 1:
        from abip_b200.dist import LpSolverDist
        from abip_b200 import problems as _pb
        p0 = _pb.cfg2(seed=2, scale=args.scale)
        ds = LpSolverDist(p0.csc(), params)
        ds.solve(p0.b, p0.c)  # warm-up
        barrier()
        sh_ms, sh_its = 0.0, 0
        for _ in range(args.steps):
            xs, ys, ss, si = ds.solve(p0.b, p0.c)
            sh_ms += si["stats"]["solve_event_ms"]
            sh_its += si["admm_iter"]
        barrier()
        ds.close()
        tsh = torch.tensor([sh_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tsh, op=dist.ReduceOp.MAX)
        sharded = {"workload": "cfg2 as ONE instance, column blocks of A over %d GPUs" % world,
                   "value": sh_its / (tsh[0].item() / 1e3), "unit": "iter/s", "scaling": "strong",
                   "time_to_1e-4_s": tsh[0].item() / args.steps / 1e3, "status": si["status"],
                   "admm_iter_per_solve": sh_its / args.steps,
                   "collective": "in-kernel peer-memory sum of the m-vector A_g x_g + scalar blocks (no NCCL in the solve)"}

    # max over ranks of the device-timed region; whole-job iterations
    t = torch.tensor([event_ms, float(its), e2e_s, float(e2e_its)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        event_ms_max, its_all, e2e_max, e2e_its_all = tmax[0].item(), tsum[1].item(), tmax[2].item(), tsum[3].item()
    else:
        event_ms_max, its_all, e2e_max, e2e_its_all = event_ms, float(its), e2e_s, float(e2e_its)

    if rank == 0:
        peak, peak_src = peaks()
        admm_ms, bb_ms = agg["admm_kernel_ms"], agg["bb_kernel_ms"]
        # dominant kernel: the persistent ADMM-iteration kernel (k_admm_iter); k_bb_round reported beside it
        ach = agg["alg_bytes_admm"] / (admm_ms * 1e-3) / 1e9 if admm_ms > 0 else None
        ach_bb = agg["alg_bytes_bb"] / (bb_ms * 1e-3) / 1e9 if bb_ms > 0 else None
        line = {
            "metric": "ADMM iters/sec", "value": its_all / (event_ms_max / 1e3), "unit": "iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": event_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "per_gpu": "one independent instance per GPU" if world > 1 else "single instance",
                       "l2": "256 MiB buffer written between timed steps (L2 flush); matrices (120 MB) ~ L2 size",
                       "status": last["status"], "admm_iter_per_solve": its / args.steps,
                       "ipm_iter": last["ipm_iter"], "pres": last["pres"], "dres": last["dres"], "gap": last["gap"]},
            "time_to_1e-4_s": event_ms_max / args.steps / 1e3,
            "host_wall_s_per_step": wall / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_its_all / e2e_max, "unit": "iter/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "time_to_1e-4_s": e2e_max / max(1, min(args.steps, 2)),
                    "note": "abip_gpu_main from host buffers: CSR build + H2D + device equilibration + solve + D2H of x, y, s"},
            "gpu_launches": int(agg["n_kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_admm_iter", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": (ach / peak) if ach else None, "traffic": None, "peak_source": peak_src,
                         "launches": int(agg["n_admm_launch"]), "avg_launch_ms": admm_ms / max(1, agg["n_admm_launch"]),
                         "k_bb_round": {"achieved": ach_bb, "frac": (ach_bb / peak) if ach_bb else None,
                                        "launches": int(agg["n_bb_launch"]),
                                        "avg_launch_ms": bb_ms / max(1, agg["n_bb_launch"])},
                         "share_of_step": {"k_admm_iter": admm_ms / event_ms, "k_bb_round": bb_ms / event_ms}},
            "counters": {"cg_iters": int(agg["n_cg_iters"]), "solves": int(agg["n_solves"]),
                         "spmv_A": int(agg["n_spmv_A"]), "spmv_AT": int(agg["n_spmv_AT"])},
        }
        if sharded is not None:
            line["sharded_single_instance"] = sharded
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
