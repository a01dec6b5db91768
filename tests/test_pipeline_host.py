"""Host logic of tum_control_b200.pipeline.BatchPipeline (tickets, slots, in-flight limit) with the device handles and the
pinned buffers replaced by stand-ins: runs without a GPU.  The real thing is compared with BatchedOcpSolver bit for bit in
tests/test_gpu_api.py."""
import numpy as np
import pytest

from tum_control_b200 import pipeline as pl
from tum_control_b200.config import OcpConfig


class _Buf:
    def __init__(self, shape):
        self.array = np.zeros(shape)


class _FakeSolver:
    """'Solves' by copying x0 into every stage of x when synchronised; records the call order."""
    log = []

    def __init__(self, ocp, batch, device, squeeze):
        self.batch, self.N, self.pending = batch, ocp.N, None
        self.id = len(_FakeSolver.log)

    def _view(self, ptr, shape):
        import ctypes
        n = int(np.prod(shape))
        return np.frombuffer((ctypes.c_double * n).from_address(ptr), dtype=np.float64, count=n).reshape(shape)

    def set_x0_ptr(self, ptr): self.x0 = self._view(ptr, (self.batch, 8)).copy()
    def set_yref_all_ptr(self, ptr): self.yref = self._view(ptr, (self.batch, self.N + 1, 6)).copy()
    def init_iterate_from_x0(self): pass
    def solve_async(self): _FakeSolver.log.append(("solve", id(self)))
    def get_results_async_ptr(self, X, U, D): self.pending = (X, U, D)

    def synchronize(self):
        if self.pending:
            X, U, D = self.pending
            self._view(X, (self.batch, self.N + 1, 8))[...] = self.x0[:, None, :]
            self._view(D, (self.batch, 5))[:, 0] = self.yref[:, 0, 0]
            self.pending = None

    def close(self): pass


@pytest.fixture
def pipe(monkeypatch):
    monkeypatch.setattr(pl, "BatchedOcpSolver", _FakeSolver)
    monkeypatch.setattr(pl._lib, "PinnedArray", _Buf)
    return pl.BatchPipeline(OcpConfig(N=4, Tp=0.32), batch=6, depth=2, split=2)


def _batch(seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((6, 8)), rng.standard_normal((6, 5, 6))


def test_parts_cover_the_batch_and_every_slot_has_its_own_handles(pipe):
    assert pipe.parts == [(0, 3), (3, 6)] and len(pipe.handles) == 4 and len({id(h) for h in pipe.handles}) == 4


def test_results_come_back_in_ticket_order_from_the_right_slot(pipe):
    b = [_batch(s) for s in range(5)]
    out = [(x.copy(), d.copy()) for x, u, d in pipe.solve_many(b)]
    assert len(out) == 5 and pipe.in_flight == 0 and pipe.next_ticket == 5
    for (x0, yref), (x, d) in zip(b, out):
        assert np.array_equal(x, np.repeat(x0[:, None, :], 5, axis=1)) and np.array_equal(d[:, 0], yref[:, 0, 0])


def test_in_flight_limit_and_collection_order_are_enforced(pipe):
    t0 = pipe.submit(*_batch(0)); t1 = pipe.submit(*_batch(1))
    assert (t0, t1, pipe.in_flight, pipe.oldest) == (0, 1, 2, 0)
    with pytest.raises(RuntimeError):
        pipe.submit(*_batch(2))
    with pytest.raises(RuntimeError):
        pipe.result(t1)
    pipe.result(t0)
    t2 = pipe.submit()                      # in place: the slot keeps the inputs of ticket 0
    assert t2 == 2 and pipe.in_flight == 2
    x0_slot, _ = pipe.inputs(t2)
    assert np.array_equal(x0_slot, _batch(0)[0])
    assert pipe.drain() is not None and pipe.in_flight == 0 and pipe.drain() is None


def test_bad_shapes_are_rejected():
    with pytest.raises(ValueError):
        pl.BatchPipeline(OcpConfig(N=4, Tp=0.32), batch=1, depth=2, split=2)
