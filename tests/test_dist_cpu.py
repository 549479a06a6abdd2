"""CPU tests (gloo, world_size 2) of the host-side multi-GPU plumbing in abip_b200/dist.py: column partition,
exchange of the 64-byte IPC handles, assembly of the solution shards.  The data path itself (in-kernel NVLink
peer-memory all-reduce) is covered on GPUs by tests/test_dist_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from abip_b200 import problems
from abip_b200.dist import assemble_shards, column_partition, exchange_handles, lp_solve_batch_sharded, shard_indices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = problems.mcf_lp(4, 40, 200, 6, 300, seed=6)
        c0, nl = column_partition(p.n, p.Ap, world, rank)
        # 1. handles travel in rank order and unmodified
        mine = bytes([(rank * 37 + i) % 256 for i in range(64)])
        allh = exchange_handles(mine)
        assert len(allh) == 64 * world
        for r in range(world):
            assert allh[64 * r:64 * (r + 1)] == bytes([(r * 37 + i) % 256 for i in range(64)])
        # 2. every rank contributes its shard, zeros elsewhere; the sum is the full vector on every rank
        full = np.arange(p.n, dtype=np.float64) * 0.5 + 1.0
        shard = np.zeros(p.n)
        shard[c0:c0 + nl] = full[c0:c0 + nl]
        got = assemble_shards(shard)
        assert np.array_equal(got, full)
        # 3. the partitions tile [0, n) and are balanced by nonzeros
        t = torch.tensor([c0, nl, int(p.Ap[c0 + nl] - p.Ap[c0])], dtype=torch.int64)
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        pos = 0
        for q in parts:
            assert int(q[0]) == pos and int(q[1]) > 0
            pos += int(q[1])
            assert abs(int(q[2]) - p.nnz / world) <= 0.1 * p.nnz / world + 64
        assert pos == p.n
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2_plumbing(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_column_partition_properties(world):
    p = problems.random_lp(50, 400, 5, seed=8)
    pos = 0
    for r in range(world):
        c0, nl = column_partition(p.n, p.Ap, world, r)
        assert c0 == pos and nl >= 0
        pos += nl
    assert pos == p.n


def _batch_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        probs = [("problem", i) for i in range(7)]
        seen = []

        def fake_solve(ps, params, conc, ctas):     # stands in for lp_solve_batch (no GPU here)
            seen.extend(ps)
            assert conc <= max(len(ps), 1) and ctas == 1
            return [(p[1] * 10, rank) for p in ps]

        out = lp_solve_batch_sharded(probs, dict(tol=1e-4), solve_fn=fake_solve)
        assert [p[1] for p in seen] == shard_indices(7, world, rank)
        assert [o[0] for o in out] == [10 * i for i in range(7)]           # original order on every rank
        assert [o[1] for o in out] == [i % world for i in range(7)]       # problem i was solved by rank i % world
        local = lp_solve_batch_sharded(probs, None, solve_fn=fake_solve, gather=False)
        assert sorted(local) == shard_indices(7, world, rank)
        open(os.path.join(out_dir, f"bok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2_batch_sharding(tmp_path):
    """configs[4]: independent LPs sharded one problem set per GPU -- no data-path collective, results gathered."""
    world = 2
    mp.spawn(_batch_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"bok{r}") for r in range(world))
    assert shard_indices(5, 8, 6) == [] and shard_indices(10, 4, 1) == [1, 5, 9]
