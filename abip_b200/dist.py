"""Multi-GPU ABIP-LP: one process per GPU (torchrun), torch.distributed for the plumbing (exchange of the CUDA-IPC
handles, assembly of the solution shards); the data path runs inside the persistent kernels over NVLink peer memory
(csrc/lp_device.cuh: comm_sum_vec / comm_sum_scalars) -- no NCCL call inside the solve.

Partition: rank r owns a contiguous block of COLUMNS of A balanced by nonzeros, i.e. a row block of the stored
CSR(A'); m-space vectors are replicated, n-space vectors sharded."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _capi
from .api import CscHolder, _lp_settings, _fp


def column_partition(n: int, Ap: np.ndarray, world: int, rank: int):
    """[c0, c0 + nl) owned by `rank` (same rule as the C side: abip_gpu_column_partition)."""
    L = _capi.lib()
    Ap = np.ascontiguousarray(Ap, dtype=np.int64)
    c0, nl = C.c_long(), C.c_long()
    L.abip_gpu_column_partition(int(n), Ap.ctypes.data_as(C.POINTER(C.c_long)), int(world), int(rank), C.byref(c0),
                                C.byref(nl))
    return int(c0.value), int(nl.value)


def exchange_handles(handle: bytes, group=None) -> bytes:
    """all-gather of the 64-byte IPC handles, rank order (works with gloo on CPU tensors and nccl on CUDA tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return b"".join(bytes(t.cpu().tolist()) for t in out)


def assemble_shards(local: np.ndarray, group=None) -> np.ndarray:
    """sum of the zero-padded shards = full vector, on every rank"""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


# Where sharding is refused (measured, profiles/r02_multi_gpu.md):
#   * below MIN_NNZ_PER_GPU nonzeros per GPU the exchange of the m-vector (two cross-GPU flag round trips + NVLink
#     transfers per operator application) costs as much as the local SpMV it saves: cfg2 (5.0 M nonzeros) runs at
#     0.98x - 1.02x on 2 GPUs and 0.74x / 0.70x on 4 / 8 (SCALE_r01), so it stays on one GPU;
#   * beyond MAX_SHARD_GPUS the in-kernel reduce-scatter + all-gather stops paying: cfg4 (60.5 M nonzeros) 1.70x on 2 GPUs,
#     2.37x on 4, but 1.24x on 8.
# A solver asked to use more GPUs than that shards over the first `effective_world` ranks only; the others hold no engine
# and contribute zeros to the assembly of x and s.
MIN_NNZ_PER_GPU = 4_000_000
MAX_SHARD_GPUS = 4


def effective_world(nnz: int, world: int, min_nnz_per_gpu: int | None = None) -> int:
    lim = MIN_NNZ_PER_GPU if min_nnz_per_gpu is None else min_nnz_per_gpu
    return max(1, min(world, MAX_SHARD_GPUS if min_nnz_per_gpu is None else world, nnz // max(1, lim)))


class LpSolverDist:
    """Collective counterpart of LpSolver: every rank constructs it with the FULL problem and calls solve()."""

    def __init__(self, A, params: dict | None = None, group=None, min_nnz_per_gpu: int | None = None, **raw_settings):
        import torch.distributed as dist
        self.L = _capi.lib()
        self.outer_group = group
        self.outer_rank, self.outer_world = dist.get_rank(group), dist.get_world_size(group)
        self.p, self.st = _lp_settings(params, **raw_settings)
        self.H = CscHolder(A)
        self.world = effective_world(self.H.nnz, self.outer_world, min_nnz_per_gpu)
        self.w = None
        # every rank of the outer group must take part in new_group()
        if self.world < self.outer_world:
            ranks = list(range(self.world)) if group is None else None
            if ranks is None:
                raise ValueError("a custom group with fewer effective ranks is not supported")
            group = dist.new_group(ranks=ranks)
        self.group = group
        self.active = self.outer_rank < self.world
        self.rank = self.outer_rank
        if not self.active:
            self.setup_time_ms = 0.0
            return
        self.info = _capi.ABIPInfo()
        self._b = np.zeros(self.H.m)
        self._c = np.zeros(self.H.n)
        self.d = _capi.ABIPData(self.H.m, self.H.n, C.pointer(self.H.c), _fp(self._b), _fp(self._c),
                                float(self.H.nnz) / (float(self.H.m) * float(self.H.n)), C.pointer(self.st))
        self.w = self.L.abip_gpu_init_dist(C.byref(self.d), C.byref(self.info), self.rank, self.world)
        if not self.w:
            raise RuntimeError("abip_gpu_init_dist failed")
        if self.world > 1:
            buf = (C.c_ubyte * 64)()
            if self.L.abip_gpu_comm_export(self.w, C.cast(buf, C.c_void_p)) != 0:
                raise RuntimeError("abip_gpu_comm_export failed")
            allh = exchange_handles(bytes(buf), group)
            hb = (C.c_ubyte * len(allh)).from_buffer_copy(allh)
            if self.L.abip_gpu_comm_connect(self.w, C.cast(hb, C.c_void_p)) != 0:
                raise RuntimeError("abip_gpu_comm_connect failed")
            dist.barrier(group)  # every rank has mapped every buffer before the first kernel spins on a flag
        self.setup_time_ms = self.info.setup_time
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [C.c_void_p]

    def _solve_idle(self):
        """rank outside the effective group: zero shards + the result record of rank 0"""
        import torch.distributed as dist
        x = assemble_shards(np.zeros(self.H.n), self.outer_group)
        s = assemble_shards(np.zeros(self.H.n), self.outer_group)
        box = [None]
        dist.broadcast_object_list(box, src=0, group=self.outer_group)
        res, y = box[0]
        return x, y, s, res

    def solve(self, b, c):
        if not self.active:
            return self._solve_idle()
        self._b[:] = b
        self._c[:] = c
        sol = _capi.ABIPSolution()
        info = _capi.ABIPInfo()
        self.L.abip_gpu_solve(self.w, C.byref(self.d), C.byref(sol), C.byref(info))
        stats = _capi.ABIPGpuStats()
        self.L.abip_gpu_get_stats(self.w, C.byref(stats))
        out = {}
        for name, ln in (("x", self.H.n), ("y", self.H.m), ("s", self.H.n)):
            ptr = getattr(sol, name)
            out[name] = np.ctypeslib.as_array(ptr, shape=(ln,)).copy() if ptr else np.full(ln, np.nan)
            if ptr:
                self._libc.free(C.cast(ptr, C.c_void_p))
        x = assemble_shards(out["x"], self.outer_group)
        s = assemble_shards(out["s"], self.outer_group)
        res = dict(status=info.status.decode(), status_val=int(info.status_val), ipm_iter=int(info.ipm_iter),
                   admm_iter=int(info.admm_iter), pres=info.res_pri, dres=info.res_dual, gap=info.rel_gap,
                   pobj=info.pobj, dobj=info.dobj, solve_time_ms=info.solve_time, gpus_used=self.world,
                   stats={f: getattr(stats, f) for f, _ in stats._fields_})
        if self.world < self.outer_world:
            import torch.distributed as dist
            dist.broadcast_object_list([(res, out["y"])], src=0, group=self.outer_group)
        return x, out["y"], s, res

    def close(self):
        if self.w:
            self.L.abip_gpu_finish(self.w)
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_indices(count: int, world: int, rank: int) -> list:
    """Problems of a batch owned by `rank`: interleaved, so that every GPU sees the same mix of problem sizes."""
    return list(range(rank, count, world))


def lp_solve_batch_sharded(problems, params: dict | None = None, concurrency: int = 192, ctas_per_problem: int = 1,
                           group=None, gather: bool = True, solve_fn=None, **raw_settings):
    """BASELINE.json configs[4]: a batch of independent LPs sharded one problem set per GPU (one process per GPU,
    no data-path collective).  Every rank calls it with the FULL list; rank r solves problems r, r + world, ... with
    `lp_solve_batch` on its own GPU.  gather=True: the results are all-gathered (torch.distributed objects) and every
    rank returns the full list in the original order; gather=False: {index: result} of the local shard.
    solve_fn(problems, params, concurrency, ctas_per_problem) replaces lp_solve_batch (CPU tests of the plumbing)."""
    import torch.distributed as dist
    from .api import lp_solve_batch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    problems = list(problems)
    mine = shard_indices(len(problems), world, rank)
    fn = solve_fn or (lambda ps, prm, cc, ct: lp_solve_batch(ps, prm, concurrency=cc, ctas_per_problem=ct,
                                                            **raw_settings))
    local = fn([problems[i] for i in mine], params, min(concurrency, max(len(mine), 1)), ctas_per_problem) if mine else []
    res = dict(zip(mine, local))
    if not gather:
        return res
    parts = [None] * world
    dist.all_gather_object(parts, res, group=group)
    merged = {}
    for part in parts:
        merged.update(part)
    return [merged[i] for i in range(len(problems))]
