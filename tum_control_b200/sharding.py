"""Multi-GPU plumbing: the batch of independent problem instances is sharded statically over the ranks of
one node (one process per GPU); there is no collective on the solve path.  The only exchange is an optional
all-gather of the solved trajectories for logging (the role of ``np.savez(full_logs.npz)`` in the reference,
``Utils/Logging_Plotting.py:282``).  ``torch.distributed`` is plumbing here: NCCL over NVLink on the GPU box,
gloo in the CPU tests."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: rank r of G gets instances [r*B/G, (r+1)*B/G) (the first B % G ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} not in [0, {world})")
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_log_rows(X: np.ndarray, U: np.ndarray, debug: np.ndarray) -> np.ndarray:
    """[B, (N+1)*8 + N*2 + 5] rows = [x_0..x_N, u_0..u_{N-1}, cost, time_tot, sqp_iter, qp_iter, status]."""
    B = X.shape[0]
    return np.concatenate([X.reshape(B, -1), U.reshape(B, -1), debug.reshape(B, 5)], axis=1)


def unpack_log_rows(rows: np.ndarray, N: int):
    B = rows.shape[0]
    nx = (N + 1) * 8
    return rows[:, :nx].reshape(B, N + 1, 8), rows[:, nx:nx + 2 * N].reshape(B, N, 2), rows[:, nx + 2 * N:]


def all_gather_logs(local_rows, batch: int):
    """All-gather the per-rank log rows (a torch tensor on the rank's device) into the full [batch, row] tensor
    in instance order.  Shards may differ in size by one row, so the gather is padded to the largest shard."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    out = torch.empty((world * maxn, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * maxn: r * maxn + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
