"""``BatchPipeline``: cold-start RTI solves of independent batches with several batches in flight.

The reference's callers that evaluate many independent problems (the Bayesian-optimisation objective,
``BO_WMPC/bayesian_optimization.py:389``; the vectorised RL environments, ``RL_WMPC/environment.py:290``) hand the solver
one batch after the other.  A single batch of 4096 instances is 1.73 waves of the QP kernel's resident warps: the SMs drain
behind its last instances.  This class keeps ``depth`` batches in flight instead -- every batch has its own handles (own
CUDA streams) and its own pinned host buffers, ``submit`` only enqueues (pinned H2D of ``x0`` / ``yref``, the RTI step,
asynchronous D2H of the iterate and the statistics), ``result`` waits for one batch -- so the tail of one batch is filled by
the next (1.65 M -> 2.10 M solves/s end to end at B = 4096, N = 30).  Each batch is further split over
``split`` handles so that the copies of one part overlap the kernels of the others.

Results are those of ``BatchedOcpSolver`` on the same inputs, bit for bit (``tests/test_gpu_api.py``).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import _lib
from .config import OcpConfig
from .solver import BatchedOcpSolver


class BatchPipeline:
    def __init__(self, ocp: OcpConfig, batch: int, device: int = 0, depth: int = 2, split: int = 2):
        if depth < 1 or split < 1 or batch < split:
            raise ValueError("BatchPipeline: need depth >= 1 and 1 <= split <= batch")
        self.ocp, self.batch, self.depth, self.split = ocp, int(batch), int(depth), int(split)
        N = self.N = int(ocp.N)
        B = self.batch
        self.parts: List[Tuple[int, int]] = [(i * B // split, (i + 1) * B // split) for i in range(split)]
        self._slots = []
        for _ in range(depth):
            handles = [BatchedOcpSolver(ocp, batch=hi - lo, device=device, squeeze=False) for lo, hi in self.parts]
            bufs = {"x0": _lib.PinnedArray((B, 8)), "yref": _lib.PinnedArray((B, N + 1, 6)),
                    "x": _lib.PinnedArray((B, N + 1, 8)), "u": _lib.PinnedArray((B, N, 2)), "stats": _lib.PinnedArray((B, 5))}
            self._slots.append((handles, bufs))
        self._submitted = 0          # tickets handed out
        self._collected = 0          # tickets whose results were waited for

    @property
    def handles(self) -> List[BatchedOcpSolver]:
        return [h for hs, _ in self._slots for h in hs]

    def inputs(self, ticket: int):
        """Pinned input buffers ``(x0 [B, 8], yref [B, N+1, 6])`` of the slot that ticket ``ticket`` uses (ticket modulo
        ``depth``); with ``ticket = self.next_ticket`` they are what the next ``submit()`` without arguments sends."""
        b = self._slots[ticket % self.depth][1]
        return b["x0"].array, b["yref"].array

    @property
    def next_ticket(self) -> int:
        return self._submitted

    @property
    def in_flight(self) -> int:
        return self._submitted - self._collected

    @property
    def oldest(self) -> int:
        """Ticket of the oldest batch in flight (the one ``result`` accepts next)."""
        return self._collected

    def submit(self, x0=None, yref=None) -> int:
        """Enqueue one batch (x = x0 on every stage, u = 0: a freshly reset solver) and return its ticket; never waits for the
        device.  With ``depth`` batches already in flight the slot is still in use: RuntimeError -- collect the oldest ticket
        first (``solve_many`` does the bookkeeping).  Without arguments the slot's pinned input buffers are sent as they are
        (``inputs(next_ticket)`` to fill them in place)."""
        t = self._submitted
        if t - self._collected >= self.depth:
            raise RuntimeError(f"BatchPipeline: {self.depth} batches in flight, collect ticket {self._collected} first")
        handles, b = self._slots[t % self.depth]
        if x0 is not None:
            b["x0"].array[...] = x0
        if yref is not None:
            b["yref"].array[...] = yref
        for Sh, (lo, hi) in zip(handles, self.parts):
            Sh.set_x0_ptr(b["x0"].array[lo:hi].ctypes.data)
            Sh.set_yref_all_ptr(b["yref"].array[lo:hi].ctypes.data)
            Sh.init_iterate_from_x0()
            Sh.solve_async()
            Sh.get_results_async_ptr(b["x"].array[lo:hi].ctypes.data, b["u"].array[lo:hi].ctypes.data,
                                     b["stats"].array[lo:hi].ctypes.data)
        self._submitted += 1
        return t

    def result(self, ticket: int):
        """Wait for batch ``ticket`` (the oldest one in flight) -> ``(x [B, N+1, 8], u [B, N, 2], stats [B, 5])``, views of
        the slot's pinned buffers (valid until the slot is submitted again, ``depth`` tickets later).  ``stats`` rows are
        the reference's ``simSolverDebug`` columns: cost, time_tot, sqp_iter, qp_iter, status."""
        if ticket != self._collected or ticket >= self._submitted:
            raise RuntimeError(f"BatchPipeline: results are collected in order, next is ticket {self._collected}")
        handles, b = self._slots[ticket % self.depth]
        for Sh in handles:
            Sh.synchronize()
        self._collected += 1
        return b["x"].array, b["u"].array, b["stats"].array

    def drain(self):
        out = None
        while self._collected < self._submitted:
            out = self.result(self._collected)
        return out

    def solve_many(self, batches):
        """Generator over ``(x, u, stats)`` of an iterable of ``(x0, yref)`` batches, ``depth`` of them in flight.  The arrays
        are views of pinned buffers: copy what must outlive the next ``depth`` batches."""
        for x0, yref in batches:
            if self._submitted - self._collected >= self.depth:
                yield self.result(self._collected)
            self.submit(x0, yref)
        while self._collected < self._submitted:
            yield self.result(self._collected)

    def close(self):
        for hs, _ in self._slots:
            for h in hs:
                h.close()
        self._slots = []
