"""CPU, world_size 2 over gloo: the sharding arithmetic and the single gather of per-candidate rows
(clairs_to_b200/dist.py) that the multi-GPU path uses with NCCL."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clairs_to_b200.dist import exchange_sizes, gather_rows, gather_rows_padded, shard_bounds, shard_sizes


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_total * 16, dtype=torch.float32).reshape(n_total, 8, 2)
    lo, hi = shard_bounds(n_total, world, rank)
    got = gather_rows(full[lo:hi].clone(), n_total)
    # uneven shards whose lengths the ranks only know locally (bench.py configs 2-4: batches of a strong-scaled shard)
    cut = n_total // 3 if rank == 0 else None
    lo2, hi2 = (0, n_total // 3) if rank == 0 else (n_total // 3, n_total)
    sizes = exchange_sizes(hi2 - lo2, torch.device("cpu"))
    got2 = gather_rows_padded(full[lo2:hi2].clone(), sizes)
    got3 = gather_rows_padded(full[lo2:hi2].clone())
    if rank == 0:
        q.put(bool(torch.equal(got, full)) and bool(torch.equal(got2, full)) and bool(torch.equal(got3, full)) and sizes == [n_total // 3, n_total - n_total // 3])
    else:
        assert got is None and got2 is None and got3 is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [9, 64])
def test_gather_rows_world2(n_total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True
