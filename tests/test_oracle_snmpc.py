"""Stochastic NMPC: host PCE set-up against the reference's committed acados_ocp_SNMPC.json, the SNMPC MODEL of the
oracle (augmented discrete dynamics, chance constraint, cost output) against the folded expressions the reference
committed in that JSON (evaluated numerically: to round-off with full-precision constants, to 1e-5 with the literals as
printed), and the dense (full condensing) SNMPC oracle through its reductions and finite differences.  The reference
commits no SNMPC *solver outputs*; the QP on top of the pinned model is the same unique-optimum convex QP class that the
nominal oracle is pinned on, and the dense and Riccati solvers are checked against each other."""
import os

import numpy as np
import pytest

import oracle
from tum_control_b200 import pce
from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch

N, TP = 30, 2.4


def test_hammersley_samples_match_reference_json(golden_dir):
    """lbx_0 of acados_ocp_SNMPC.json is the 11x8 x0 sample matrix made by chaospy's Hammersley rule with stds
    (1.1, 0.2, 0.05) on (vlong, vlat, yawrate)."""
    g = np.load(os.path.join(golden_dir, "snmpc_json_golden.npz"))
    lbx0 = g["lbx_0"].reshape(11, 8)
    std = [0, 0, 0, 1.1, 0.2, 0.05, 0, 0]
    w, A, alphas = pce.samples_and_A(10, std)
    np.testing.assert_allclose(lbx0[0] + pce.sample_offsets(w, std), lbx0[1:], rtol=0, atol=1e-15)
    np.testing.assert_array_equal(g["ubx_0"], g["lbx_0"])
    # alpha order of alphaGeneration(3, 2): constant term first
    assert alphas.tolist() == [[0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0], [0, 0, 2], [0, 1, 1], [0, 2, 0], [1, 0, 1],
                               [1, 1, 0], [2, 0, 0]]
    assert A.shape == (10, 10)
    assert abs(A[0].sum() - 1.0) < 1e-12 and np.abs(A[1:].sum(axis=1)).max() < 1e-12   # mean row sums to one
    # gamma = 0.8 => kappa = 0.5, the factor folded into the JSON's con_h_expr
    assert np.sqrt((1 - 0.8) / 0.8) == pytest.approx(0.5)


def test_dense_oracle_equals_riccati_oracle_without_samples():
    """n_s = 0, UPH = 0: the SNMPC OCP is the nominal OCP with a 1-step integrator and |v| in cost and h.  The
    dense full-condensing oracle and the Riccati oracle are independent solvers of the same QP.  The QP is
    degenerate where the linearised lh = 0 row is weakly active (SURVEY B-3), so an interior-point iterate is only
    O(sqrt(complementarity)) from the optimum: both solvers are run to ipm_tol = 1e-13 and compared at 1e-6."""
    x0, yref = synthetic_batch("monteblanco", 6, N, TP, seed=3)
    ocp = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1, ipm_tol=1e-13)
    OR = oracle.Oracle(ocp)
    OD = oracle.SnmpcOracle(ocp, A=np.zeros((0, 0)), H=0)
    n_ipm = 0
    for b in range(6):
        X1, U1 = OR.init_guess(x0[b]); X2, XS, U2 = OD.init_guess(x0[b])
        for _ in range(3):
            st1, s1 = OR.rti_step(x0[b], yref[b], X1, U1)
            st2, s2 = OD.rti_step(x0[b], yref[b], X2, XS, U2)
            assert st1 == 0 and st2 == 0
            assert np.abs(X1 - X2).max() <= 1e-6 and np.abs(U1 - U2).max() <= 1e-6
            assert abs(s1[0] - s2[0]) <= 1e-8 * max(1.0, s1[0]) and abs(s1[1] - s2[1]) <= 3
            n_ipm += s1[1] > 0
    assert n_ipm > 0


def test_zero_spread_samples_reduce_to_the_nominal_block():
    """All sample offsets zero => every sample equals the nominal state, the PCE mean is the RK4 step and the
    variance term vanishes: UPH = 5 with 10 samples must equal UPH = 0 without samples."""
    x0, yref = synthetic_batch("lvms", 3, N, TP, seed=1)
    ocp = OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1)
    _, A, _ = pce.samples_and_A(10, ocp.stds)
    O0 = oracle.SnmpcOracle(ocp, A=np.zeros((0, 0)), H=0)
    Oz = oracle.SnmpcOracle(ocp, A=A, w_off=np.zeros((10, 8)), H=5)
    for b in range(3):
        Xa, XSa, Ua = O0.init_guess(x0[b]); Xb, XSb, Ub = Oz.init_guess(x0[b])
        for _ in range(3):
            assert O0.rti_step(x0[b], yref[b], Xa, XSa, Ua)[0] == 0
            assert Oz.rti_step(x0[b], yref[b], Xb, XSb, Ub)[0] == 0
        assert np.abs(Xa - Xb).max() <= 1e-6 and np.abs(Ua - Ub).max() <= 1e-6
        assert np.abs(XSb[:5, :10] - Xb[:5, None, :]).max() <= 1e-9


def test_condensed_linearisation_by_finite_differences():
    """At a dynamically consistent iterate the condensed multiple-shooting maps equal the derivatives of the
    single-shooting roll-out of the augmented discrete model (pred_model_dynamic_disc.py:187-220) and of the
    chance constraint (SNMPC_acados_settings.py:131-133,184)."""
    x0, yref = synthetic_batch("monteblanco", 2, N, TP, seed=3)
    b = 1
    OS = oracle.SnmpcOracle(OcpConfig(N=N, Tp=TP, variant="snmpc", rk_steps=1))
    assert OS.ns == 10 and OS.H == 5
    rng = np.random.default_rng(0)
    U = np.stack([rng.normal(0, 0.5, N), rng.normal(0, 0.02, N)], axis=1)

    def rollout(U):
        xa = np.concatenate([x0[b]] + [x0[b] + OS.w_off[i] for i in range(OS.ns)])
        X = np.zeros((N + 1, 8)); XS = np.zeros((OS.H, OS.ns, 8)); hs = np.zeros(N + 1)
        for k in range(N + 1):
            X[k] = xa[:8]
            if k < OS.H:
                XS[k] = xa[8:].reshape(OS.ns, 8)
            hs[k] = OS.h_tilde(xa, 0 if k < OS.H else 1)
            if k < N:
                xa = OS.disc(xa, U[k], 0 if k < OS.H else 1)
        return X, XS, hs

    X, XS, hs = rollout(U)
    st, stats, G, rows, v0 = OS.rti_step(x0[b], yref[b], X.copy(), oracle._c(XS.copy()), U.copy(), debug=True)
    assert st == 0
    assert np.abs(v0[2::3][1:] - hs[1:]).max() <= 1e-12          # row values at the linearisation point
    # chance constraint is a back-off: larger than the nominal h on the propagation horizon
    assert all(hs[k] > OS.h(X[k])[0] for k in range(1, OS.H))
    eps = 1e-6
    Gfd = np.zeros_like(G); Rfd = np.zeros((N + 1, 2 * N))
    for a in range(2 * N):
        Up = U.copy().ravel(); Um = U.copy().ravel(); Up[a] += eps; Um[a] -= eps
        Xp, _, hp = rollout(Up.reshape(N, 2)); Xm, _, hm = rollout(Um.reshape(N, 2))
        Gfd[:, a] = ((Xp - Xm) / (2 * eps)).ravel(); Rfd[:, a] = (hp - hm) / (2 * eps)
    assert np.abs(G - Gfd).max() <= 1e-6 * np.abs(G).max()
    assert np.abs(rows[2::3][1:] - Rfd[1:]).max() <= 1e-7


# ------------------------------------------------------------------------------------------------------
# Numeric pin against the reference's own folded model (acados_ocp_SNMPC.json: disc_dyn_expr :775,
# con_h_expr :745, cost_y_expr :772).  The fixture holds those printed expressions evaluated by
# tests/golden/make_snmpc_model_golden.py (parser: tests/golden/casadi_print_eval.py) on 160 seeded augmented
# states, half with stop_flag = 1: once with the literals as printed (6 significant digits) and once with every
# printed literal that is the rounding of a committed model constant replaced by that constant.
# ------------------------------------------------------------------------------------------------------
def _json_oracle(g):
    dt = float(g["dt"][0])
    ocp = OcpConfig(N=38, Tp=38 * dt, variant="snmpc", rk_steps=1)       # the JSON's horizon: N = 38, tf = 3.04
    assert abs(ocp.dt - 0.08) < 1e-15
    return oracle.SnmpcOracle(ocp, A=g["A"], w_off=np.zeros((10, 8)), H=5)


def test_discrete_model_equals_the_reference_json_expression(golden_dir):
    g = np.load(os.path.join(golden_dir, "snmpc_model_golden.npz"))
    OS = _json_oracle(g)
    n = len(g["xa"])
    assert n >= 100 and set(g["stop"]) == {0.0, 1.0}
    worst_exact = worst_print = 0.0
    for c in range(n):
        out = OS.disc(g["xa"][c], g["u"][c], int(g["stop"][c]))
        scale = np.maximum(1.0, np.abs(g["disc_exact"][c]))
        worst_exact = max(worst_exact, float((np.abs(out - g["disc_exact"][c]) / scale).max()))
        worst_print = max(worst_print, float((np.abs(out - g["disc_print"][c]) / scale).max()))
    assert worst_exact <= 1e-12, worst_exact        # structure + full-precision constants: round-off only
    assert worst_print <= 1e-5, worst_print         # the literals as printed: the JSON's print precision


def test_chance_constraint_equals_the_reference_json_expression(golden_dir):
    g = np.load(os.path.join(golden_dir, "snmpc_model_golden.npz"))
    OS = _json_oracle(g)
    for key in ("h", "he"):       # path and terminal expression are the same function of (x, p)
        for c in range(len(g["xa"])):
            v = OS.h_tilde(g["xa"][c], int(g["stop"][c]))
            assert abs(v - g[key + "_exact"][c]) <= 1e-12 * max(1.0, abs(v))
            assert abs(v - g[key + "_print"][c]) <= 1e-5 * max(1.0, abs(v))
    # kappa = sqrt((1 - gamma) / gamma) = 0.5 is the factor printed in con_h_expr
    assert abs(OS.S.kappa - 0.5) < 1e-15
    # the propagated constraint differs from the nominal one on the stop_flag = 0 cases (the pin is not vacuous)
    d = [abs(OS.h_tilde(g["xa"][c], 0) - OS.h_tilde(g["xa"][c], 1)) for c in range(0, 40, 2)]
    assert min(d) > 1e-6


def test_cost_output_equals_the_reference_json_expression(golden_dir):
    g = np.load(os.path.join(golden_dir, "snmpc_model_golden.npz"))
    OS = _json_oracle(g)
    for c in range(len(g["xa"])):
        y = OS.cost_output(g["xa"][c][:8], g["u"][c]); ye = OS.cost_output(g["xa"][c][:8])
        np.testing.assert_allclose(y, g["y_exact"][c], rtol=0, atol=1e-12)
        np.testing.assert_allclose(ye, g["ye_exact"][c], rtol=0, atol=1e-12)
        np.testing.assert_allclose(y, g["y_print"][c], rtol=0, atol=2e-5)      # 6.28319 vs 2 pi
    assert (g["y_exact"][:, 2] >= 0).all() and (g["y_exact"][:, 2] < 2 * np.pi).all()
    assert (g["xa"][:, 2] < 0).any() and (g["xa"][:, 2] > 2 * np.pi).any()      # both wrap branches are exercised
