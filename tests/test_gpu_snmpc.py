"""GPU parity of the Stochastic NMPC variant: the CUDA path (head condensing + Riccati tail, through the C-ABI)
against the dense full-condensing CPU oracle (oracle/snmpc_oracle.c), which solves the same QP by a different
method.  SNMPC parity is UNPINNED by reference data: the oracle is the reference here, and it
is itself cross-checked against the golden-pinned Riccati oracle in tests/test_oracle_snmpc.py.

Tolerance: the linearised lh = 0 rows make the QP degenerate, so an interior-point iterate is O(sqrt(mu)) from the
optimum; both sides run to ipm_tol = 1e-13 and are compared at 1e-6 * max(1, |.|) (SURVEY section 8c contract).
"""
import numpy as np
import pytest

import oracle
from tum_control_b200 import pce
from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch

pytestmark = pytest.mark.gpu

N, TP = 30, 2.4
TOL = 1e-6


def _solver(ocp, B):
    from tum_control_b200.solver import BatchedOcpSolver
    return BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def _oracle_steps(O, x0, yref, steps):
    """Per-instance oracle iterates after `steps` RTI steps from the reference's initial guess."""
    B = x0.shape[0]
    Xs, XSs, Us, stats = [], [], [], []
    for b in range(B):
        X, XS, U = O.init_guess(x0[b])
        st = None
        for _ in range(steps):
            code, st = O.rti_step(x0[b], yref[b], X, XS, U)
            assert code == 0
        Xs.append(X); XSs.append(XS); Us.append(U); stats.append(st)
    return np.array(Xs), np.array(XSs), np.array(Us), stats


@pytest.mark.parametrize("ns,uph", [(10, 5), (13, 5), (10, 3)])
def test_head_linearisation_matches_dense_oracle(ns, uph):
    """The condensed head (nominal PCE-mean maps G_k, c_k, chance-constraint rows and values) equals the dense
    oracle's condensed multiple-shooting maps restricted to the head inputs."""
    B = 6
    ocp = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1, n_samples=ns, uph=uph, ipm_tol=1e-13)
    O = oracle.SnmpcOracle(ocp)
    x0, yref = synthetic_batch("monteblanco", B, N, TP, seed=5)
    X, XS, U, _ = _oracle_steps(O, x0, yref, 2)
    S = _solver(ocp, B)
    S.set_x0(x0); S.set_yref_all(yref); S.set_iterate(X, U)
    XSg = np.zeros((B, uph + 1, ns, 8)); XSg[:, :uph] = XS[:, :uph, :ns]
    S.set_samples(XSg)
    assert np.all(S.solve() == 0)
    head = S.snmpc_head()
    nh = 2 * uph
    for b in range(B):
        Xo, XSo, Uo = X[b].copy(), oracle._c(XS[b].copy()), U[b].copy()
        st, stats, G, rows, v0 = O.rti_step(x0[b], yref[b], Xo, XSo, Uo, debug=True)
        assert st == 0
        G = G.reshape(N + 1, 8, 2 * N)
        for k in range(1, uph + 1):
            assert np.abs(G[k][:, nh:]).max() == 0.0
            assert np.abs(head["G"][b, k - 1] - G[k][:, :nh]).max() <= 1e-10 * max(1.0, np.abs(G[k]).max())
        for k in range(1, uph):
            assert np.abs(head["rows"][b, k] - rows[3 * k + 2, :nh]).max() <= 1e-10 * max(1.0, np.abs(rows[3 * k + 2]).max())
            assert abs(head["v0"][b, k] - v0[3 * k + 2]) <= 1e-10
            assert abs(head["c"][b, k - 1, 6] + X[b, k, 6] - v0[3 * k + 1]) <= 1e-10        # bx row value
        # and the step itself
        Xg, Ug = S.get_iterate()
        assert _rel(Xg[b], Xo) <= TOL and _rel(Ug[b], Uo) <= TOL
        XSn = S.get_samples()
        assert _rel(XSn[b, :uph], XSo[:uph, :ns]) <= TOL


@pytest.mark.parametrize("track,ns,uph,B", [("monteblanco", 10, 5, 24), ("monteblanco", 13, 5, 16), ("lvms", 10, 2, 12)])
def test_rti_steps_match_dense_oracle(track, ns, uph, B):
    """Cold start + 3 warm RTI steps; nominal iterate, sample iterates, cost against the dense oracle."""
    ocp = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1, n_samples=ns, uph=uph, ipm_tol=1e-13)
    O = oracle.SnmpcOracle(ocp)
    x0, yref = synthetic_batch(track, B, N, TP, seed=1)
    S = _solver(ocp, B)
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    Xo = []; XSo = []; Uo = []
    for b in range(B):
        a, s, u = O.init_guess(x0[b]); Xo.append(a); XSo.append(s); Uo.append(u)
    n_ipm = 0
    for step in range(4):
        status = S.solve()
        assert np.all(status == 0)
        Xg, Ug = S.get_iterate(); XSg = S.get_samples(); dbg = S.get_solver_debug()
        for b in range(B):
            code, st = O.rti_step(x0[b], yref[b], Xo[b], XSo[b], Uo[b])
            assert code == 0
            assert _rel(Xg[b], Xo[b]) <= TOL, (step, b, _rel(Xg[b], Xo[b]))
            assert _rel(Ug[b], Uo[b]) <= TOL, (step, b, _rel(Ug[b], Uo[b]))
            assert _rel(XSg[b, :uph], XSo[b][:uph, :ns]) <= TOL, (step, b)
            assert abs(dbg[b, 0] - st[0]) <= 1e-7 * max(1.0, abs(st[0]))
            n_ipm += st[1] > 0
        assert S.get_stats("kkt").max() <= 1e-9
    assert n_ipm > 0


def test_no_propagation_horizon_equals_riccati_oracle():
    """UPH = 0: every stop flag is 1, the SNMPC OCP is the nominal OCP with a 1-step integrator and |v| in the cost
    and h; the CUDA path then runs the plain Riccati kernel and must agree with the Riccati oracle."""
    B = 32
    ocp = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1, uph=0)
    O = oracle.Oracle(ocp)
    x0, yref = synthetic_batch("modena", B, N, TP, seed=2)
    S = _solver(ocp, B)
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    Xo = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); Uo = np.zeros((B, N, 2))
    for _ in range(3):
        assert np.all(S.solve() == 0)
        O.rti_batch(x0, yref, Xo, Uo, nthreads=8)
        Xg, Ug = S.get_iterate()
        assert _rel(Xg, Xo) <= TOL and _rel(Ug, Uo) <= TOL


def test_zero_spread_samples_reduce_to_no_propagation():
    """All sample offsets zero: every sample equals the nominal state, the PCE mean is the RK4 step and the variance
    term vanishes, so UPH = 5 (head path) must equal UPH = 0 (plain Riccati path)."""
    B = 16
    _, A, _ = pce.samples_and_A(10, OcpConfig().stds)
    o5 = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1, uph=5, pce_A=A, pce_w_off=np.zeros((10, 8)), ipm_tol=1e-13)
    o0 = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1, uph=0, ipm_tol=1e-13)
    x0, yref = synthetic_batch("lvms", B, N, TP, seed=4)
    S5, S0 = _solver(o5, B), _solver(o0, B)
    for S in (S5, S0):
        S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    for _ in range(3):
        assert np.all(S5.solve() == 0) and np.all(S0.solve() == 0)
    X5, U5 = S5.get_iterate(); X0, U0 = S0.get_iterate()
    assert _rel(X5, X0) <= TOL and _rel(U5, U0) <= TOL
    XS = S5.get_samples()
    assert np.abs(XS[:, :5] - X5[:, :5, None, :]).max() <= 1e-9      # samples collapse onto the nominal trajectory


def test_acados_surface_of_the_snmpc_handle():
    """The calls SNMPC_class.py makes: constraints_set(0,'lbx',x0_samples.flatten()), set(j,'x',...), set(j,'p',...),
    get(j,'x')[0:8]; shapes carry the augmented state."""
    ocp = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1)
    A, w_off = ocp.pce_matrices()
    x0, yref = synthetic_batch("monteblanco", 1, N, TP, seed=0)
    xs = np.concatenate([x0[0][None], x0[0][None] + w_off]).flatten()
    S = _solver(ocp, 1); S.squeeze = True
    S.constraints_set(0, "lbx", xs); S.constraints_set(0, "ubx", xs)
    stop = np.zeros(N + 1); stop[ocp.uph:] = 1
    for j in range(N + 1):
        S.set(j, "p", np.concatenate((A.flatten(), [ocp.gamma], [stop[j]])))
        S.set(j, "x", xs)
        S.set(j, "yref", yref[0, j, :6] if j < N else yref[0, j, :4])
    assert S.uph == ocp.uph
    assert S.solve() == 0
    x1 = S.get(1, "x")
    assert x1.shape == (88,)
    # same through the bulk calls
    S2 = _solver(ocp, 1)
    S2.set_x0(x0); S2.set_yref_all(yref); S2.init_iterate_from_x0()
    assert np.all(S2.solve() == 0)
    X2, U2 = S2.get_iterate()
    assert np.abs(x1[:8] - X2[0, 1]).max() <= 1e-12
    assert np.abs(S.get(0, "u") - U2[0, 0]).max() <= 1e-12
    assert np.abs(x1[8:].reshape(10, 8) - S2.get_samples()[0, 1]).max() <= 1e-12
    # frozen samples beyond the propagation horizon
    assert np.array_equal(S.get(N, "x")[8:], S.get(ocp.uph, "x")[8:])
    # a stop-flag pattern that is not a prefix is refused at solve time
    S.set(ocp.uph + 2, "p", np.concatenate((A.flatten(), [ocp.gamma], [0.0])))
    from tum_control_b200.solver import SolverError
    with pytest.raises(SolverError):
        S.solve()


def test_stochastic_controller_mirror():
    from tum_control_b200.controllers import Stochastic_Nonlinear_Model_Predictive_Controller as SNMPC
    B = 8
    ocp = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1, ipm_tol=1e-13)
    x0, yref = synthetic_batch("monteblanco", B, N, TP, seed=7)
    ctl = SNMPC(x0, ocp=ocp)
    ref = {"pos_x": yref[:, :, 0], "pos_y": yref[:, :, 1], "ref_yaw": yref[:, :, 2], "ref_v": yref[:, :, 3]}
    u0, pred_X, stats = ctl.solve(ref)
    assert u0.shape == (B, 2) and pred_X.shape == (B, N, 8) and stats.shape == (B, 5)
    O = oracle.SnmpcOracle(ocp)
    for b in range(B):
        X, XS, U = O.init_guess(x0[b])
        code, st = O.rti_step(x0[b], yref[b], X, XS, U)
        assert code == 0
        assert _rel(u0[b], U[0]) <= TOL and _rel(pred_X[b], X[:N]) <= TOL
    # set_initial_state moves the whole sample matrix
    ctl.set_initial_state(pred_X[:, 1])
    u0b, _, _ = ctl.solve(ref)
    assert np.all(np.isfinite(u0b))


def test_device_x0_sample_matrix_equals_the_reference_json(golden_dir):
    """lbx_0 of the committed acados_ocp_SNMPC.json is the reference's own x0 sample matrix (compute_x0dist with the
    chaospy Hammersley samples, stds 1.1 / 0.2 / 0.05): the matrix the device derives from x0 must equal it."""
    import os
    g = np.load(os.path.join(golden_dir, "snmpc_json_golden.npz"))
    lbx0 = g["lbx_0"].reshape(11, 8)
    ocp = OcpConfig(N=38, Tp=3.04, variant="snmpc", rk_steps=1, n_samples=10, stds=np.array([0, 0, 0, 1.1, 0.2, 0.05, 0, 0]))
    S = _solver(ocp, 2)
    S.set_x0(np.stack([lbx0[0], lbx0[0] + 1.0]))
    S.init_iterate_from_x0()
    x = S.get(0, "x")
    assert x.shape == (2, 88)
    np.testing.assert_allclose(x[0].reshape(11, 8), lbx0, rtol=0, atol=1e-14)
    np.testing.assert_allclose(x[1].reshape(11, 8), lbx0 + 1.0, rtol=0, atol=1e-14)
    # and every later stage starts from the same matrix (SNMPC_class.py:127-130)
    np.testing.assert_array_equal(S.get(3, "x"), x)
