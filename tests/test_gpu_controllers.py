"""The batched controller classes (host mirrors of NMPC_class / Reduced_Robustified_NMPC_class) in a short
closed loop, against the same loop driven through the CPU oracle, and the fused R2NMPC tightening kernel
against the oracle's restatement of Reduced_Robustified_NMPC_class.py:287-365."""
import numpy as np
import pytest

import oracle
from tum_control_b200.config import OcpConfig, load_reftraj
from tum_control_b200.planner import PlannerEmulator, initial_state

pytestmark = pytest.mark.gpu


def _windows(planner, poses, N, Tp):
    ws = [planner(p, N + 1, Tp)[1] for p in poses]
    return {k: np.stack([w[k] for w in ws]) for k in ("pos_x", "pos_y", "ref_yaw", "ref_v")}


def _yref(win, N):
    B = win["pos_x"].shape[0]
    y = np.zeros((B, N + 1, 6))
    for i, k in enumerate(("pos_x", "pos_y", "ref_yaw", "ref_v")):
        y[:, :, i] = win[k]
    return y


def test_nominal_controller_closed_loop_matches_oracle():
    from tum_control_b200.controllers import Nonlinear_Model_Predictive_Controller, ocp_from_reference_params
    ref = load_reftraj("monteblanco")
    planner = PlannerEmulator(ref)
    B = 6
    sim = {"Tp": 2.4, "Ts_MPC": 0.08}
    ocp = ocp_from_reference_params(sim_main_params=sim)
    N, Tp = ocp.N, ocp.Tp
    assert N == 30
    X0 = np.stack([initial_state(ref, i) for i in (0, 150, 400, 700, 950, 1100)])
    mpc = Nonlinear_Model_Predictive_Controller(X0, sim_main_params=sim)
    O = [oracle.Oracle(ocp) for _ in range(B)]
    it = [O[b].init_guess(X0[b]) for b in range(B)]
    x = X0.copy()
    for step in range(6):
        poses = [[x[b, 0], x[b, 1], x[b, 2], np.hypot(x[b, 3], x[b, 4])] for b in range(B)]
        win = _windows(planner, poses, N, Tp)
        u0, pred_X, stats = mpc.solve(win)
        assert u0.shape == (B, 2) and pred_X.shape == (B, N, 8) and stats.shape == (B, 5)
        yref = _yref(win, N)
        for b in range(B):
            st, so = O[b].rti_step(x[b], yref[b], it[b][0], it[b][1])
            assert st == 0 and stats[b, 4] == 0 and stats[b, 2] == 1
            assert np.abs(u0[b] - it[b][1][0]).max() <= 1e-6
            assert np.max(np.abs(pred_X[b] - it[b][0][:N]) / np.maximum(1, np.abs(it[b][0][:N]))) <= 1e-6
            assert abs(stats[b, 0] - so[0]) <= 1e-8 * max(1.0, so[0])
        x = np.stack([it[b][0][1] for b in range(B)])          # MPC-sim mode: next state = predicted x_1
        mpc.set_initial_state(x)
    # reset(x0) restores the constructor state: x_k = x0, u = 0
    mpc.reset(X0)
    Xg, Ug = mpc.acados_solver.get_iterate()
    assert np.array_equal(Xg, np.repeat(X0[:, None, :], N + 1, axis=1)) and not Ug.any()


def test_r2_tightening_matches_oracle_and_reduces_to_nominal():
    from tum_control_b200.controllers import Reduced_Robustified_NMPC, Nonlinear_Model_Predictive_Controller
    ref = load_reftraj("modena")
    planner = PlannerEmulator(ref)
    sim = {"Tp": 2.4, "Ts_MPC": 0.08}
    idx = (10, 200, 420, 640, 880)
    B = len(idx)
    X0 = np.stack([initial_state(ref, i) for i in idx])
    X0[:, 3] *= 0.97; X0[:, 5] += 0.02
    r2 = Reduced_Robustified_NMPC(X0, sim_main_params=sim)
    ocp = r2.ocp
    N, Tp = ocp.N, ocp.Tp
    poses = [[x[0], x[1], x[2], np.hypot(x[3], x[4])] for x in X0]
    win = _windows(planner, poses, N, Tp)
    yref = _yref(win, N)
    O = [oracle.Oracle(ocp) for _ in range(B)]
    it = [O[b].init_guess(X0[b]) for b in range(B)]
    qp = oracle.QpView()
    for step in range(3):
        u0, pred_X, stats = r2.solve(win)
        for b in range(B):
            st, _ = O[b].rti_step(X0[b], yref[b], it[b][0], it[b][1], qp=qp)
            assert st == 0
            A = qp.get("A", N).reshape(N, 8, 8).copy()             # pre-step linearisation of the solve just done
            O[b].r2_tighten(A, it[b][0])                            # post-step iterate
            assert np.abs(u0[b] - it[b][1][0]).max() <= 1e-6, (step, b)
            assert np.max(np.abs(pred_X[b] - it[b][0][:N]) / np.maximum(1, np.abs(it[b][0][:N]))) <= 1e-6
    # the tightened bounds differ from the nominal ones on stages 1..N-1 and are what the oracle computed
    assert all(O[b].ub[2, 2] < 1.0 and O[b].ub[N - 1, 1] < ocp.vehicle.delta_f_max for b in range(B))
    # zero uncertainty => back-offs vanish (up to sqrt(2.5e-11)) => the R2 controller is the nominal controller
    z = {"stds": [0.0] * 8}
    r2z = Reduced_Robustified_NMPC(X0, mpc_params=z, sim_main_params=sim)
    nom = Nonlinear_Model_Predictive_Controller(X0, sim_main_params=sim)
    for _ in range(3):
        ua, Xa, _ = r2z.solve(win)
        ub, Xb, _ = nom.solve(win)
    assert np.abs(ua - ub).max() <= 1e-4 and np.abs(Xa - Xb).max() <= 1e-3
