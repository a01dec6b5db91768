"""Finite-difference and structural checks of the oracle's hand-derived derivatives, the plant
replay against the golden logs (plant transitions are pinned to ~1e-16) and the constants folded
into the reference's committed acados_ocp_SNMPC.json."""
import os

import numpy as np
import pytest

import oracle
from tum_control_b200.config import OcpConfig
from replay_util import load_log


@pytest.fixture(scope="module")
def O():
    return oracle.Oracle(OcpConfig(N=38, Tp=3.04))


def _states(n, seed=0):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 8))
    x[:, 0:2] = rng.uniform(-500, 500, (n, 2)); x[:, 2] = rng.uniform(-7, 7, n)
    x[:, 3] = rng.uniform(5, 37.5, n); x[:, 4] = rng.normal(0, 0.5, n); x[:, 5] = rng.normal(0, 0.2, n)
    x[:, 6] = rng.normal(0, 0.05, n); x[:, 7] = rng.uniform(-3.5, 2.5, n)
    u = np.stack([rng.uniform(-8, 6, n), rng.uniform(-0.3, 0.3, n)], axis=1)
    return x, u


def test_jacobian_finite_difference_and_sparsity(O):
    x, u = _states(20)
    eps = 1e-6
    for xi, ui in zip(x, u):
        f, Jx, Ju = O.f_jac(xi, ui)
        np.testing.assert_array_equal(f, O.f(xi, ui))
        J = np.zeros((8, 10))
        for i in range(10):
            d = np.zeros(10); d[i] = eps
            J[:, i] = (O.f(xi + d[:8], ui + d[8:]) - O.f(xi - d[:8], ui - d[8:])) / (2 * eps)
        assert np.abs(J[:, :8] - Jx).max() <= 5e-7 * max(1.0, np.abs(Jx).max())
        assert np.abs(J[:, 8:] - Ju).max() <= 1e-9
        assert np.count_nonzero(Jx) <= 22                      # 22 structural non-zeros


def test_sensitivity_structure_and_fd(O):
    x, u = _states(8, seed=1)
    eps = 1e-6
    for xi, ui in zip(x, u):
        xn, A, B = O.integrate(xi, ui)
        # columns 0,1 = e0,e1; rows 6,7 = e6',e7'; B rows 6,7 = [0,dt],[dt,0]
        np.testing.assert_array_equal(A[:, 0], np.eye(8)[0]); np.testing.assert_array_equal(A[:, 1], np.eye(8)[1])
        np.testing.assert_array_equal(A[6], np.eye(8)[6]); np.testing.assert_array_equal(A[7], np.eye(8)[7])
        np.testing.assert_allclose(B[6], [0, O.M.dt], atol=1e-15); np.testing.assert_allclose(B[7], [O.M.dt, 0], atol=1e-15)
        Jn = np.zeros((8, 10))
        for i in range(10):
            d = np.zeros(10); d[i] = eps
            Jn[:, i] = (O.integrate(xi + d[:8], ui + d[8:])[0] - O.integrate(xi - d[:8], ui - d[8:])[0]) / (2 * eps)
        assert np.abs(Jn[:, :8] - A).max() <= 1e-6 and np.abs(Jn[:, 8:] - B).max() <= 1e-6


def test_h_gradient_and_ggv_regions(O):
    x, _ = _states(30, seed=2)
    x[:10, 3] = np.linspace(10.5, 12.5, 10)                   # the only sloped ggv interval (11.11..12)
    eps = 1e-6
    for xi in x:
        h, g = O.h(xi)
        gn = np.array([(O.h(xi + eps * e)[0] - O.h(xi - eps * e)[0]) / (2 * eps) for e in np.eye(8)])
        # skip points that straddle a knot or a = 0 inside the FD stencil
        if min(abs(xi[3] - 11.11), abs(xi[3] - 12.0)) < 1e-5 or abs(xi[7]) < 1e-5:
            continue
        assert np.abs(g - gn).max() <= 1e-6 * max(1.0, np.abs(g).max())
        assert np.count_nonzero(g) <= 3


def test_low_speed_guard(O):
    """if_else(vlong > 0.001, ..., 0): slip angles and their derivatives vanish (pred_model...:148-149)."""
    x = np.array([0, 0, 0.3, 0.0005, 0.2, 0.1, 0.05, 0.5]); u = np.array([0.1, 0.05])
    f, Jx, _ = O.f_jac(x, u)
    assert np.all(np.isfinite(f)) and np.all(np.isfinite(Jx))
    assert f[5] == pytest.approx(1.484 * (-(0.009 + 0.002 * 3.6 * np.hypot(x[3], x[4]) / 100
                                          + 0.0003 * (3.6 * np.hypot(x[3], x[4]) / 100) ** 4)
                                          * 2520 * 1.644 * 9.81 / 3.128) * np.sin(0.05) / 13600, rel=1e-12)


def test_plant_replay_matches_logged_transitions(O):
    """CiLX[i+1] = RK4(4 finite elements, Ts = 0.02) of the 7-state plant with inputs
    [MPC_SimX[i+1,7], simU[i,1]] (SimulationMode_main_class.py:113-118)."""
    log = load_log("monteblanco_0")
    C = log["CiLX"].copy(); C[:, 2] = np.unwrap(C[:, 2])
    worst = 0.0
    for i in list(range(0, 300)) + list(range(2000, 2200)) + list(range(5200, 5400)):
        xn = O.plant_step(C[i], [log["MPC_SimX"][i + 1, 7], log["simU"][i, 1]])
        worst = max(worst, np.abs(xn - C[i + 1]).max())
    assert worst <= 1e-11


def test_constants_folded_in_reference_snmpc_json(O, golden_dir):
    """acados_ocp_SNMPC.json folds Fz_r=11728.3, Fz_f=12992.9, Fmax_f=21309.9, Fmax_r=22129,
    0.5 rho S Cd = 0.621687, 1/m = 3.96825e-4, 1/Iz = 7.35294e-5."""
    M = O.M
    Fz_f = M.m * M.lr * M.g / (M.lf + M.lr); Fz_r = M.m * M.lf * M.g / (M.lf + M.lr)
    assert Fz_r == pytest.approx(11728.3, abs=0.05) and Fz_f == pytest.approx(12992.9, abs=0.05)
    assert Fz_f * np.sqrt(1 + M.Cf ** 2) == pytest.approx(21309.9, abs=0.05)
    assert Fz_r * np.sqrt(1 + M.Cr ** 2) == pytest.approx(22129.0, abs=0.5)
    assert 0.5 * M.ro * M.S * M.Cd == pytest.approx(0.621687, abs=1e-6)
    g = np.load(os.path.join(golden_dir, "snmpc_json_golden.npz"))
    N, nx, nu, np_, nh, nh0, ns, ns0, nse, ny, nye = g["dims"]
    assert (N, nx, nu, nh, nh0, ns, ns0, nse, ny, nye) == (38, 88, 2, 1, 0, 3, 1, 2, 6, 4)
    assert g["lh"][0] == 0 and g["uh"][0] == 1
    assert g["lbu"][0] == -0.322 and g["ubx"][0] == 0.610865


def test_batch_driver_equals_single_instance_calls():
    from tum_control_b200.planner import synthetic_batch
    N, Tp, B = 30, 2.4, 24
    ocp = OcpConfig(N=N, Tp=Tp)
    Ob = oracle.Oracle(ocp)
    x0, yref = synthetic_batch("monteblanco", B, N, Tp, seed=4)
    Xb = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); Ub = np.zeros((B, N, 2))
    Xs, Us = Xb.copy(), Ub.copy()
    for _ in range(2):
        bad, stats = Ob.rti_batch(x0, yref, Xb, Ub, nthreads=4)
        assert bad == 0
        for b in range(B):
            st, s1 = Ob.rti_step(x0[b], yref[b], Xs[b], Us[b])
            assert st == 0 and s1[0] == stats[b, 0]
    assert np.array_equal(Xb, Xs) and np.array_equal(Ub, Us)
    assert stats[:, 3].max() <= 1e-9
