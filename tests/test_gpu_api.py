"""The AcadosOcpSolver-shaped surface: per-stage set/get round trips, broadcast vs per-instance values,
runtime-mutable stage data and the error behaviour (ValueError on misuse, like acados_template)."""
import numpy as np
import pytest

import oracle
from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch

pytestmark = pytest.mark.gpu


def _solver(ocp, B, **kw):
    from tum_control_b200.solver import BatchedOcpSolver
    return BatchedOcpSolver(ocp, batch=B, device=0, **kw)


def test_set_get_roundtrip_and_reset():
    N, B = 12, 5
    S = _solver(OcpConfig(N=N, Tp=0.96), B, squeeze=False)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((B, N + 1, 8)); U = rng.standard_normal((B, N, 2))
    for k in range(N + 1):
        S.set(k, "x", X[:, k])
    for k in range(N):
        S.set(k, "u", U[:, k])
    for k in (0, 3, N):
        assert np.array_equal(S.get(k, "x"), X[:, k])
    assert np.array_equal(S.get(N - 1, "u"), U[:, N - 1])
    Xg, Ug = S.get_iterate()
    assert np.array_equal(Xg, X) and np.array_equal(Ug, U)
    S.set(2, "x", np.arange(8.0))                      # broadcast form
    assert np.array_equal(S.get(2, "x"), np.tile(np.arange(8.0), (B, 1)))
    S.reset()
    Xg, Ug = S.get_iterate()
    assert not Xg.any() and not Ug.any()


def test_errors_like_acados_template():
    N = 10
    S = _solver(OcpConfig(N=N, Tp=0.8), 2, squeeze=False)
    with pytest.raises(ValueError):
        S.set(N + 1, "x", np.zeros(8))
    with pytest.raises(ValueError):
        S.set(N, "u", np.zeros(2))
    with pytest.raises(ValueError):
        S.set(0, "nonsense", np.zeros(8))
    with pytest.raises(ValueError):
        S.set(0, "x", np.zeros(7))
    with pytest.raises(ValueError):
        S.constraints_set(0, "uh", np.array([1.0]))      # nh_0 = 0
    with pytest.raises(ValueError):
        S.cost_set(1, "W", np.ones((6, 6)))              # non-diagonal W
    with pytest.raises(ValueError):
        S.get(0, "lam")
    with pytest.raises(ValueError):
        S.get_stats("nope")


def test_batch_of_one_is_a_drop_in_for_the_reference_calls():
    """The exact call sequence of NMPC_class.py:111-112,169-206 with un-batched arrays."""
    N, Tp = 30, 2.4
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("monteblanco", 1, N, Tp, seed=11)
    x0, yref = x0[0], yref[0]
    S = _solver(ocp, 1)
    S.constraints_set(0, "lbx", x0); S.constraints_set(0, "ubx", x0)
    for i in range(N + 1):
        S.set(i, "x", x0)
    for j in range(N):
        S.set(j, "yref", yref[j])
    S.set(N, "yref", yref[N, :4])
    status = S.solve()
    assert status == 0
    u0 = S.get(0, "u")
    pred_X = np.stack([S.get(j, "x") for j in range(N)])
    stats = [S.get_cost(), S.get_stats("time_tot"), S.get_stats("sqp_iter"), np.max(S.get_stats("qp_iter")), status]
    O = oracle.Oracle(ocp)
    Xo, Uo = O.init_guess(x0)
    st, so = O.rti_step(x0, yref, Xo, Uo)
    assert u0.shape == (2,) and pred_X.shape == (N, 8)
    assert np.abs(u0 - Uo[0]).max() <= 1e-6 and np.abs(pred_X - Xo[:N]).max() <= 1e-6 * np.abs(Xo).max()
    assert abs(stats[0] - so[0]) <= 1e-8 * max(1, so[0]) and stats[2] == 1 and stats[1] > 0


def test_runtime_mutable_stage_data_matches_oracle():
    """cost_set / constraints_set per stage (broadcast and per instance) change the solution exactly as in
    the oracle: weights without the 0.01 factor, tighter bounds on a few stages, per-instance uh."""
    N, Tp, B = 20, 1.6, 32
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("lvms", B, N, Tp, seed=2)
    rng = np.random.default_rng(1)
    S = _solver(ocp, B, squeeze=False)
    O = [oracle.Oracle(ocp) for _ in range(B)]
    p = np.array([4.7, 1.2, 11.9, 5.9, 276.6, 630.2, 1555.4])           # F.csv row 0
    W = np.diag([p[0], p[0], p[1], p[2], p[3], p[4]])
    for k in range(N):
        S.cost_set(k, "W", W)
    S.cost_set(N, "W", W[:4, :4])
    for k in range(N + 1):
        n = 1 if k == 0 else (2 if k == N else 3)
        for f, v in (("zl", p[5]), ("zu", p[5]), ("Zl", p[6]), ("Zu", p[6])):
            S.cost_set(k, f, np.full(n, v))
    uh = rng.uniform(0.6, 1.0, (B, 1))
    for k in range(1, N + 1):
        S.constraints_set(k, "uh", uh)
    S.constraints_set(3, "ubx", np.array([0.01])); S.constraints_set(3, "lbx", np.array([-0.01]))
    S.constraints_set(0, "ubu", np.array([0.05]))
    for b in range(B):
        O[b].update_cost_function_weights(p)
        O[b].ub[1:, 2] = uh[b, 0]
        O[b].ub[3, 1] = 0.01; O[b].lb[3, 1] = -0.01
        O[b].ub[0, 0] = 0.05
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    Xo = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); Uo = np.zeros((B, N, 2))
    for step in range(3):
        assert np.all(S.solve() == 0)
        for b in range(B):
            st, _ = O[b].rti_step(x0[b], yref[b], Xo[b], Uo[b])
            assert st == 0
        Xg, Ug = S.get_iterate()
        assert np.max(np.abs(Xg - Xo) / np.maximum(1, np.abs(Xo))) <= 1e-6
        assert np.max(np.abs(Ug - Uo) / np.maximum(1, np.abs(Uo))) <= 1e-6
    # slack values of the last QP are exposed like acados' get(stage, "sl"/"su")
    assert S.get(1, "su").shape == (B, 3) and S.get(0, "sl").shape == (B, 1) and S.get(N, "su").shape == (B, 2)


def test_async_results_into_pinned_buffers_equal_the_synchronous_getters():
    """b200nmpc_get_results_async: stream-ordered D2H into pinned host memory, complete after synchronize()."""
    from tum_control_b200 import _lib
    N, Tp, B = 20, 1.6, 24
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("modena", B, N, Tp, seed=9)
    S = _solver(ocp, B, squeeze=False)
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    hX = _lib.PinnedArray((B, N + 1, 8)); hU = _lib.PinnedArray((B, N, 2)); hD = _lib.PinnedArray((B, 5))
    S.solve_async()
    S.get_results_async_ptr(hX.array.ctypes.data, hU.array.ctypes.data, hD.array.ctypes.data)
    S.synchronize()
    X, U = S.get_iterate()
    dbg = S.get_solver_debug()
    assert np.array_equal(hX.array, X) and np.array_equal(hU.array, U)
    assert np.array_equal(hD.array[:, [0, 2, 3, 4]], dbg[:, [0, 2, 3, 4]]) and np.all(hD.array[:, 1] == 0.0)
