"""Edge cases of the batched solve through the C-ABI: failure isolation, ragged and extreme sizes, vacuous rows and
grossly infeasible starts -- each against the CPU oracle on the same inputs."""
import numpy as np
import pytest

import oracle
from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch

pytestmark = pytest.mark.gpu

TOL = 1e-6


def _solver(ocp, B):
    from tum_control_b200.solver import BatchedOcpSolver
    return BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def _oracle_batch(ocp, x0, yref, steps):
    O = oracle.Oracle(ocp)
    B, N = x0.shape[0], ocp.N
    X = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); U = np.zeros((B, N, 2))
    stats = None
    for _ in range(steps):
        _, stats = O.rti_batch(x0, yref, X, U, nthreads=8)
    return X, U, stats


def test_nan_instance_fails_alone_and_keeps_its_iterate():
    """A NaN in one instance's x0 gives acados' status 4 (QP failure) for that instance only; its iterate is left
    untouched (the reference keeps pred_X of the last success, NMPC_class.py:195-201); neighbours in the same block
    of four are unaffected."""
    N, Tp, B = 30, 2.4, 37
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("monteblanco", B, N, Tp, seed=11)
    Xo, Uo, _ = _oracle_batch(ocp, x0, yref, 1)
    bad = [5, 18]
    x0n = x0.copy(); x0n[5, 3] = np.nan
    yrefn = yref.copy(); yrefn[18, 7, 1] = np.inf
    S = _solver(ocp, B)
    S.set_x0(np.where(np.isnan(x0n), 0.0, x0n)); S.init_iterate_from_x0()      # finite initial iterate
    X_before, U_before = S.get_iterate()
    S.set_x0(x0n); S.set_yref_all(yrefn)
    status = S.solve()
    assert status[bad].tolist() == [4, 4] and np.all(np.delete(status, bad) == 0)
    X, U = S.get_iterate()
    assert np.array_equal(X[bad], X_before[bad]) and np.array_equal(U[bad], U_before[bad])
    good = np.delete(np.arange(B), bad)
    assert _rel(X[good], Xo[good]) <= TOL and _rel(U[good], Uo[good]) <= TOL
    # the instance recovers once its inputs are finite again (main.py:59-61 re-initialises the solver)
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    assert np.all(S.solve() == 0)
    X, U = S.get_iterate()
    assert _rel(X, Xo) <= TOL and _rel(U, Uo) <= TOL


@pytest.mark.parametrize("N,B", [(2, 3), (3, 1), (5, 129), (128, 5)])
def test_extreme_horizons_and_ragged_batches(N, B):
    """Shortest and longest supported horizons (B200NMPC_MAX_N = 128), batch sizes that do not fill a block."""
    Tp = 0.08 * N
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("modena", B, N, Tp, seed=N)
    Xo, Uo, so = _oracle_batch(ocp, x0, yref, 2)
    S = _solver(ocp, B)
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    for _ in range(2):
        assert np.all(S.solve() == 0)
    X, U = S.get_iterate()
    ok = so[:, 3] <= 1e-9
    assert ok.mean() >= 0.8
    assert _rel(X[ok], Xo[ok]) <= TOL and _rel(U[ok], Uo[ok]) <= TOL


def test_rows_without_penalty_are_vacuous_and_huge_violations_are_absorbed_by_the_slacks():
    """zl = zu = Zl = Zu = 0 on a row removes it (no barrier, no slack cost); a start far outside every limit
    (steering angle at 3x its bound, 15 m/s^2 of acceleration) is handled by the L1/L2 slacks like in acados."""
    N, Tp, B = 30, 2.4, 16
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("lvms", B, N, Tp, seed=4)
    x0[:, 6] = 3 * 0.610865 * np.sign(np.arange(B) % 2 - 0.5)
    x0[:, 7] = np.linspace(-15, 15, B)
    S = _solver(ocp, B)
    O = oracle.Oracle(ocp)
    for k in range(1, N + 1):                       # drop the h row on every second stage
        if k % 2 == 0:
            n = 2 if k == N else 3
            for f in ("zl", "zu", "Zl", "Zu"):
                v = np.full(n, ocp.L1_pen if f[0] == "z" else ocp.L2_pen); v[-1] = 0.0
                S.cost_set(k, f, v)
            O.zl[k, 2] = O.zu[k, 2] = O.Zl[k, 2] = O.Zu[k, 2] = 0.0
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    Xo = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); Uo = np.zeros((B, N, 2))
    for _ in range(3):
        assert np.all(S.solve() == 0)
        _, so = O.rti_batch(x0, yref, Xo, Uo, nthreads=8)
    X, U = S.get_iterate()
    assert _rel(X, Xo) <= TOL and _rel(U, Uo) <= TOL
    su = S.get(1, "su"); sl = S.get(1, "sl")
    assert max(su.max(), sl.max()) > 0.1            # the steering-angle bound of stage 1 is violated: slack in use
    assert np.all(S.get(2, "su")[:, 2] == 0.0)      # vacuous row: no slack reported


def test_batch_beyond_32bit_element_offsets():
    """Maximum sizes: 602 112 instances put the stage records (B x 31 x 120 doubles = 2.24e9 elements, 17.9 GB) and the
    side arrays past 2^31 elements, so any 32-bit offset in a kernel or in the C-ABI copies would corrupt the high
    instances.  The batch is a 4096-instance block tiled 147 times: instances are independent, so every tile must
    reproduce tile 0 bit for bit -- iterate, cost, iteration count, status -- and tile 0 must equal the same block
    solved alone."""
    N, Tp, blk, tiles = 30, 2.4, 4096, 147
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("monteblanco", blk, N, Tp, seed=3)
    S1 = _solver(ocp, blk)
    S1.set_x0(x0); S1.set_yref_all(yref); S1.init_iterate_from_x0()
    S1.solve()
    X1, U1 = (a.copy() for a in S1.get_iterate())
    cost1, it1 = S1.get_cost().copy(), S1.get_stats("qp_iter").copy()
    S1.close()
    B = blk * tiles
    assert B * (N + 1) * 120 > 2 ** 31
    S = _solver(ocp, B)
    S.set_x0(np.tile(x0, (tiles, 1))); S.set_yref_all(np.tile(yref, (tiles, 1, 1))); S.init_iterate_from_x0()
    status = S.solve()
    X, U = S.get_iterate()
    cost, it = S.get_cost(), S.get_stats("qp_iter")
    assert not status.any()
    X, U = X.reshape(tiles, blk, N + 1, 8), U.reshape(tiles, blk, N, 2)
    assert np.array_equal(X[0], X1) and np.array_equal(U[0], U1)
    assert np.array_equal(cost[:blk], cost1) and np.array_equal(it[:blk], it1)
    for t in (1, 73, tiles - 1):
        assert np.array_equal(X[t], X1) and np.array_equal(U[t], U1), f"tile {t}"
    assert np.array_equal(cost.reshape(tiles, blk), np.broadcast_to(cost1, (tiles, blk)))
    assert np.array_equal(it.reshape(tiles, blk), np.broadcast_to(it1, (tiles, blk)))
    # a stage-wise getter and the QP data of a high instance
    assert np.array_equal(S.get(N, "x").reshape(tiles, blk, 8)[tiles - 1], X1[:, N])
    S.close()
