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


def test_per_stage_calls_without_a_sync_and_the_host_mirror_of_the_iterate():
    """The reference's own call pattern: 2N + 4 per-stage set calls, one solve, N + 1 per-stage gets.  The set calls go
    through the pinned host ring and do not wait for the device (the caller's buffer may be reused at once -- it IS reused
    here); the gets of a small batch are served from one read-back of the whole iterate, which every call that changes the
    iterate must invalidate."""
    N, Tp, B = 30, 2.4, 7
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("monteblanco", B, N, Tp, seed=4)
    ref = _solver(ocp, B, squeeze=False)
    ref.set_x0(x0); ref.set_yref_all(yref); ref.init_iterate_from_x0()
    S = _solver(ocp, B, squeeze=False)
    buf = np.empty((B, 8)); ybuf = np.empty((B, 6))
    buf[...] = x0
    S.constraints_set(0, "lbx", buf); S.constraints_set(0, "ubx", buf)
    for j in range(N + 1):                              # one scratch buffer reused for every stage
        buf[...] = x0; S.set(j, "x", buf)
        buf[...] = np.nan
    for j in range(N):
        S.set(j, "u", np.zeros(2))
        ybuf[...] = yref[:, j]; S.set(j, "yref", ybuf)
        ybuf[...] = np.nan
    S.set(N, "yref", yref[:, N, :4].copy())
    for rounds in range(3):                             # > ring capacity in total: exercises the wrap
        for j in range(1, N):
            S.constraints_set(j, "uh", np.array([1.0])); S.cost_set(j, "W", np.asarray(ocp.W))
    assert np.array_equal(S.solve(), ref.solve())
    Xr, Ur = ref.get_iterate()
    assert np.array_equal(np.stack([S.get(j, "x") for j in range(N + 1)], axis=1), Xr)
    assert np.array_equal(np.stack([S.get(j, "u") for j in range(N)], axis=1), Ur)
    # invalidation of the host mirror: set, solve, reset, init_iterate_from_x0
    S.set(3, "x", np.arange(8.0))
    assert np.array_equal(S.get(3, "x"), np.tile(np.arange(8.0), (B, 1))) and np.array_equal(S.get(2, "x"), Xr[:, 2])
    S.set(3, "x", Xr[:, 3])
    assert np.array_equal(S.solve(), ref.solve())
    Xr2, Ur2 = ref.get_iterate()
    assert np.array_equal(S.get(5, "x"), Xr2[:, 5]) and np.array_equal(S.get(0, "u"), Ur2[:, 0]) and not np.array_equal(Xr2, Xr)
    S.init_iterate_from_x0()
    assert np.array_equal(S.get(N, "x"), x0)
    S.reset()
    assert not S.get(1, "x").any()


def test_options_bounds_getters_and_snmpc_qp_in_guard():
    N, Tp, B = 30, 2.4, 300
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("lvms", B, N, Tp, seed=6)

    def run(hint):
        S = _solver(ocp, B, squeeze=False)
        S.options_set("launch_order_hint", hint)
        out = []
        for _ in range(3):                              # the 2nd and 3rd solve start in the order the previous one suggests
            S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
            st = S.solve()
            out.append((st.copy(),) + S.get_iterate())
        return S, out

    S1, a = run(1)
    S0, b = run(0)
    for (sa, Xa, Ua), (sb, Xb, Ub) in zip(a, b):        # scheduling only: results are bit-identical
        assert np.array_equal(sa, sb) and np.array_equal(Xa, Xb) and np.array_equal(Ua, Ub)
    assert all(np.array_equal(a[0][1], x[1]) for x in a[1:])
    with pytest.raises(ValueError):
        S1.options_set("no_such_option", 1)
    # bounds in force (shared values broadcast; per-instance after a per-instance set)
    assert np.all(S1.constraints_get(3, "ubx") == ocp.vehicle.delta_f_max) and np.all(S1.constraints_get(N, "lh") == 0.0)
    ub = np.linspace(0.5, 0.9, B)[:, None]
    S1.constraints_set(4, "uh", ub)
    assert np.array_equal(S1.constraints_get(4, "uh"), ub) and np.all(S1.constraints_get(5, "uh") == 1.0)
    with pytest.raises(ValueError):
        S1.constraints_get(0, "uh")
    # SNMPC: no single (A, B, b) inside the propagation horizon
    Ss = _solver(OcpConfig(N=N, Tp=Tp, variant="snmpc", rk_steps=1, uph=5), 4, squeeze=False)
    Ss.set_x0(x0[:4]); Ss.set_yref_all(yref[:4]); Ss.init_iterate_from_x0()
    assert np.all(Ss.solve() == 0)
    with pytest.raises(ValueError):
        Ss.get_from_qp_in(2, "A")
    assert Ss.get_from_qp_in(5, "A").shape == (4, 8, 8) and Ss.get_from_qp_in(7, "b").shape == (4, 8)


def test_weight_rows_of_a_batch_that_equals_the_row_length():
    """cost_set('W') with batch == 6: a (6, 6) array of per-instance DIAGONALS (positive rows) must not be mistaken for one
    weight matrix; a diagonal (6, 6) matrix still is one."""
    N, Tp, B = 10, 0.8, 6
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("monteblanco", B, N, Tp, seed=8)
    rows = np.asarray(ocp.W) * np.linspace(0.5, 2.0, B)[:, None]
    Sa = _solver(ocp, B, squeeze=False); Sb = _solver(ocp, B, squeeze=False)
    for S in (Sa, Sb):
        S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    for k in range(N):
        Sa.cost_set(k, "W", rows)                                   # (6, 6): rows of diagonals
        Sb.cost_set(k, "W", rows[:, None, :] * np.eye(6))           # (B, 6, 6): explicit matrices
    assert np.array_equal(Sa.solve(), Sb.solve())
    assert np.array_equal(Sa.get_iterate()[0], Sb.get_iterate()[0])
    Sc = _solver(ocp, B, squeeze=False); Sd = _solver(ocp, B, squeeze=False)
    for S in (Sc, Sd):
        S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    for k in range(N):
        Sc.cost_set(k, "W", np.diag(2.0 * np.asarray(ocp.W)))       # one diagonal matrix for every instance
        Sd.cost_set(k, "W", 2.0 * np.asarray(ocp.W))
    Sc.solve(); Sd.solve()
    assert np.array_equal(Sc.get_iterate()[0], Sd.get_iterate()[0]) and not np.array_equal(Sc.get_iterate()[0], Sa.get_iterate()[0])


def test_batch_pipeline_with_batches_in_flight_equals_one_solver_bit_for_bit():
    """tum_control_b200.pipeline.BatchPipeline: a stream of independent batches, two in flight, each split over two
    handles -- every batch's (x, u, stats) equals what one BatchedOcpSolver gives for that batch alone."""
    from tum_control_b200.pipeline import BatchPipeline
    N, Tp, B = 30, 2.4, 96
    ocp = OcpConfig(N=N, Tp=Tp)
    batches = [synthetic_batch(track, B, N, Tp, seed=s) for s, track in enumerate(("monteblanco", "modena", "lvms", "monteblanco", "modena"))]
    pipe = BatchPipeline(ocp, batch=B, device=0, depth=2, split=2)
    got = [(x.copy(), u.copy(), d.copy()) for x, u, d in pipe.solve_many(batches)]
    assert len(got) == len(batches) and pipe.in_flight == 0
    S = _solver(ocp, B, squeeze=False)
    for (x0, yref), (x, u, d) in zip(batches, got):
        S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
        assert np.all(S.solve() == 0)
        X, U = S.get_iterate()
        dbg = S.get_solver_debug()
        assert np.array_equal(x, X) and np.array_equal(u, U)
        assert np.array_equal(d[:, [0, 2, 3, 4]], dbg[:, [0, 2, 3, 4]])
    # explicit tickets: results are collected in order, at most `depth` batches in flight
    t0 = pipe.submit(*batches[0]); t1 = pipe.submit(*batches[1])
    with pytest.raises(RuntimeError):
        pipe.submit(*batches[2])
    with pytest.raises(RuntimeError):
        pipe.result(t1)
    assert np.array_equal(pipe.result(t0)[0], got[0][0])
    assert np.array_equal(pipe.result(t1)[0], got[1][0])
    pipe.close()
