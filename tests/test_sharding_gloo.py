"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, 'solve' their shard with the CPU oracle
(standing in for the per-rank GPU solve, which needs no communication) and all-gather the log rows; the
gathered result must equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch
from tum_control_b200.sharding import all_gather_logs, pack_log_rows, shard_bounds, unpack_log_rows

N, TP, B = 12, 0.96, 11          # odd batch: shards of unequal size


def _solve(x0, yref):
    import oracle
    O = oracle.Oracle(OcpConfig(N=N, Tp=TP))
    X = np.repeat(x0[:, None, :], N + 1, axis=1).copy()
    U = np.zeros((x0.shape[0], N, 2))
    _, stats = O.rti_batch(x0, yref, X, U, nthreads=1)
    dbg = np.stack([stats[:, 0], np.zeros(len(x0)), np.ones(len(x0)), stats[:, 1], stats[:, 2]], axis=1)
    return pack_log_rows(X, U, dbg)


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x0, yref = synthetic_batch("lvms", B, N, TP, seed=7)
    lo, hi = shard_bounds(B, rank, world)
    rows = torch.from_numpy(_solve(x0[lo:hi], yref[lo:hi]))
    full = all_gather_logs(rows, B)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for batch in (1, 7, 8, 4096, 16384):
        for world in (1, 2, 4, 8):
            edges = [shard_bounds(batch, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == batch
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_two_rank_gather_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    x0, yref = synthetic_batch("lvms", B, N, TP, seed=7)
    ref = _solve(x0, yref)
    got = np.load(out)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    X, U, dbg = unpack_log_rows(got, N)
    assert X.shape == (B, N + 1, 8) and U.shape == (B, N, 2) and np.all(dbg[:, 4] == 0)
