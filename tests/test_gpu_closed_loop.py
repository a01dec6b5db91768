"""The closed loop on the device (planner -> solve -> plant -> state estimation -> logs)
against (i) vectors produced by the reference's own PlannerEmulator, (ii) the reference's committed closed-loop logs,
run FREE (not teacher-forced: every quantity of step i+1 comes from the device's own step i), and (iii) the same
loop assembled on the host from the oracle pieces."""
import os

import numpy as np
import pytest

import oracle
from tum_control_b200.closed_loop import ClosedLoopSim
from tum_control_b200.config import OcpConfig, load_reftraj
from tum_control_b200.planner import PlannerEmulator, StateEstimation, initial_state
from replay_util import golden_ocp, load_log

pytestmark = pytest.mark.gpu


def _solver(ocp, B):
    from tum_control_b200.solver import BatchedOcpSolver
    return BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)


def _wrap_diff(a, b):
    d = np.abs(a - b)
    return np.minimum(d, np.abs(d - 2 * np.pi))


@pytest.mark.parametrize("track", ["monteblanco", "lvms", "modena"])
def test_device_planner_matches_reference_planner(track, golden_dir):
    """tests/golden/planner_golden.npz holds outputs of the reference's own PlannerEmulator (make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "planner_golden.npz"))
    pose, win, closest, cfg = (g[f"{track}_{k}"] for k in ("pose", "win", "closest", "cfg"))
    for N, Tp in sorted({(int(n), float(t)) for n, t in cfg}):
        sel = np.nonzero((cfg[:, 0] == N) & (cfg[:, 1] == Tp))[0]
        S = _solver(OcpConfig(N=N, Tp=Tp), len(sel))
        sim = ClosedLoopSim(S, track, log_capacity=1)
        ci = sim.plan(pose[sel])
        assert np.array_equal(ci, closest[sel].astype(np.int32))
        yref = S.get_yref_all()
        np.testing.assert_allclose(yref[:, :, [0, 1, 2, 3]], win[sel, : N + 1], rtol=0, atol=1e-12)
        assert np.all(yref[:, :, 4:] == 0.0)


def _host_closed_loop(ocp, track, X0, nsteps, params=None, snmpc=False, w_deriv=None, w_est=None):
    """main.py:48-78 on the host with the oracle as the solver (one vehicle).  w_deriv / w_est: [7] or [nsteps, 7]
    additive disturbances; as in the reference (SimulationMode_main_class.py:119-143) the disturbed state only feeds
    the state estimate, the undisturbed one continues as the plant state."""
    O = oracle.SnmpcOracle(ocp) if snmpc else oracle.Oracle(ocp)
    if params is not None:
        O.update_cost_function_weights(params)
    Op = oracle.Oracle(OcpConfig(N=ocp.N, Tp=ocp.Tp))
    planner = PlannerEmulator(load_reftraj(track))
    se = StateEstimation()
    x0 = np.array(X0, float)
    cil = x0[:7].copy()
    pose = [cil[0], cil[1], cil[2], np.hypot(cil[3], cil[4])]
    it = O.init_guess(x0)
    CiLX, MPCX, simU, REF, DBG = [cil.copy()], [x0.copy()], [], [], []
    for i in range(nsteps):
        yref = planner.yref(pose, ocp.N, ocp.Tp)
        st, stats = O.rti_step(x0, yref, *it)
        assert st == 0
        X, U = it[0], it[-1]
        a_in = X[1, 7]
        u_sim = np.array([a_in, U[0, 1]])
        wd = None if w_deriv is None else np.asarray(w_deriv, float).reshape(-1, 7)[min(i, np.asarray(w_deriv).reshape(-1, 7).shape[0] - 1)]
        we = np.zeros(7) if w_est is None else np.asarray(w_est, float).reshape(-1, 7)[min(i, np.asarray(w_est).reshape(-1, 7).shape[0] - 1)]
        dist = cil_next = Op.plant_step(cil, u_sim)
        if wd is not None:
            dist = Op.plant_step(cil, u_sim, w7=wd)
        cil = cil_next
        pose = [cil[0], cil[1], cil[2], np.hypot(cil[3], cil[4])]
        x0 = se(np.append(dist + we, a_in))
        CiLX.append(cil.copy()); MPCX.append(X[1].copy()); simU.append(U[0].copy()); REF.append(yref[0, :4].copy())
        DBG.append([stats[0], stats[1]])
    return np.array(CiLX), np.array(MPCX), np.array(simU), np.array(REF), np.array(DBG)


def test_free_running_closed_loop_reproduces_reference_logs():
    """Two committed laps of the reference (different weight sets), the first 700 steps each, both vehicles in one
    batch with per-instance weights.  Free running: errors are not reset by the log, so the bound is looser than the
    teacher-forced oracle gate; observed max |du0| ~1e-6, plant state ~1e-6."""
    keys = ["monteblanco_0", "monteblanco_17"]
    logs = [load_log(k) for k in keys]
    n = 700
    S = _solver(golden_ocp(), len(keys))
    S.update_cost_function_weights(np.stack([lg["params"] for lg in logs]))
    sim = ClosedLoopSim(S, "monteblanco", log_capacity=n)
    X0 = np.stack([lg["MPC_SimX"][0] for lg in logs])
    sim.reset(X0)
    sim.run(n)
    out = sim.logs()
    assert np.all(out["simSolverDebug"][:, :, 4] == 0)
    for b, lg in enumerate(logs):
        eu = np.abs(out["simU"][b] - lg["simU"][:n]).max(axis=1)
        cil = out["CiLX"][b].copy(); ref = lg["CiLX"][: n + 1]
        ex = np.abs(cil - ref); ex[:, 2] = _wrap_diff(np.mod(cil[:, 2], 2 * np.pi), ref[:, 2])
        eref = np.abs(out["simREF"][b] - lg["simREF"][:n]).max()
        assert np.median(eu) <= 1e-6 and eu.max() <= 5e-4, (keys[b], np.median(eu), eu.max())
        assert ex.max() <= 1e-3, (keys[b], ex.max())
        assert eref <= 1e-9
        ec = np.abs(out["simSolverDebug"][b, :, 0] - lg["dbg"][:n, 0]) / np.maximum(1.0, np.abs(lg["dbg"][:n, 0]))
        assert np.median(ec) <= 1e-5


@pytest.mark.parametrize("track,N,Tp", [("lvms", 30, 2.4), ("modena", 20, 1.6)])
def test_device_loop_equals_host_loop_with_oracle(track, N, Tp):
    B, nsteps = 6, 120
    ocp = OcpConfig(N=N, Tp=Tp)
    ref = load_reftraj(track)
    starts = np.linspace(0, len(ref["pos_x"]) - 1, B, endpoint=False).astype(int)
    X0 = np.stack([initial_state(ref, int(i)) for i in starts])
    S = _solver(ocp, B)
    sim = ClosedLoopSim(S, track, log_capacity=nsteps)
    sim.reset(X0)
    sec = sim.run(nsteps)
    assert sec > 0 and sim.steps_done == nsteps
    out = sim.logs()
    for b in range(B):
        CiLX, MPCX, simU, REF, DBG = _host_closed_loop(ocp, track, X0[b], nsteps)
        assert np.abs(out["CiLX"][b] - CiLX).max() <= 1e-6
        assert np.abs(out["simU"][b] - simU).max() <= 1e-6
        assert np.abs(out["MPC_SimX"][b] - MPCX).max() <= 1e-6
        assert np.abs(out["simREF"][b] - REF).max() <= 1e-9
        assert np.abs(out["simSolverDebug"][b, :, 0] - DBG[:, 0]).max() <= 1e-6 * max(1.0, np.abs(DBG[:, 0]).max())


def test_snmpc_device_loop_equals_host_loop_with_dense_oracle():
    B, nsteps, N, Tp = 3, 40, 30, 2.4
    ocp = OcpConfig(N=N, Tp=Tp, variant="snmpc", rk_steps=1, ipm_tol=1e-13)
    ref = load_reftraj("monteblanco")
    X0 = np.stack([initial_state(ref, i) for i in (40, 500, 900)])
    S = _solver(ocp, B)
    sim = ClosedLoopSim(S, "monteblanco", log_capacity=nsteps)
    sim.reset(X0)
    sim.run(nsteps)
    out = sim.logs()
    assert np.all(out["simSolverDebug"][:, :, 4] == 0)
    for b in range(B):
        CiLX, MPCX, simU, REF, DBG = _host_closed_loop(ocp, "monteblanco", X0[b], nsteps, snmpc=True)
        assert np.abs(out["CiLX"][b] - CiLX).max() <= 1e-5
        assert np.abs(out["simU"][b] - simU).max() <= 1e-5


def test_derivative_disturbance_and_log_file(tmp_path):
    """Additive state-derivative disturbance enters the plant like simulator_step_disturbed
    (VehicleSimulator.py:79-94); save_logs writes the reference's full_logs.npz keys with yaw wrapped."""
    N, Tp, nsteps = 30, 2.4, 60
    ocp = OcpConfig(N=N, Tp=Tp)
    ref = load_reftraj("monteblanco")
    X0 = initial_state(ref, 1100)[None]
    w = np.array([0.05, -0.05, 0.002, 0.3, 0.1, 0.01, 0.0])
    S = _solver(ocp, 1)
    sim = ClosedLoopSim(S, "monteblanco", log_capacity=nsteps)
    sim.reset(X0)
    sim.set_disturbances(deriv=w[None])
    sim.run(nsteps)
    out = sim.logs()
    CiLX, MPCX, simU, REF, DBG = _host_closed_loop(ocp, "monteblanco", X0[0], nsteps, w_deriv=w)
    assert np.abs(out["CiLX"][0] - CiLX).max() <= 1e-5 and np.abs(out["simU"][0] - simU).max() <= 1e-5
    p = str(tmp_path / "full_logs.npz")
    sim.save_logs(p, 0)
    d = np.load(p)
    for k in ("MPC_SimX", "CiLX", "simU", "simREF", "simSolverDebug", "sim_disturbance_derivatives",
              "sim_disturbance_state_estimation", "a_lat", "dev_lat", "dev_long", "dev_vel", "dev_yaw", "t"):
        assert k in d.files
    assert d["CiLX"].shape == (nsteps + 1, 7) and d["simU"].shape == (nsteps, 2) and d["simSolverDebug"].shape == (nsteps, 5)
    assert np.all((d["CiLX"][:, 2] >= 0) & (d["CiLX"][:, 2] < 2 * np.pi))


def test_generated_disturbances_statistics_playback_and_host_replay():
    """Device-side disturbance generators (generate_disturbances, Utils/MPC_sim_utils.py:54-84) with the w_* bounds of
    sim_main_params.yaml:54-79: distribution checks, determinism from the seed per vehicle, playback, and the same
    closed loop on the host with the logged rows."""
    N, Tp, B, nsteps = 30, 2.4, 48, 150
    ocp = OcpConfig(N=N, Tp=Tp)
    ref = load_reftraj("monteblanco")
    starts = (np.arange(B) * len(ref["pos_x"])) // B
    X0 = np.stack([initial_state(ref, int(i)) for i in starts])
    w_dot = np.array([0.8, 0.8, 0.1, 1.1, 0.1, 0.05, 0.1])
    w_est = np.array([0.15, 0.15, 0.01, 0.8, 0.35, 0.05, 0.005])

    def run(batch_sel, seed):
        S = _solver(ocp, len(batch_sel))
        sim = ClosedLoopSim(S, "monteblanco", log_capacity=nsteps)
        sim.configure_disturbances("uniform", w_dot, "gaussian", w_est, seed=seed)
        sim.reset(X0[batch_sel])
        sim.run(nsteps)
        return sim.logs()

    lg = run(np.arange(B), seed=7)
    d, e = lg["sim_disturbance_derivatives"], lg["sim_disturbance_state_estimation"]
    # uniform = sampleFromEllipsoid(0, diag(w)): inside the ellipsoid with semi-axes sqrt(w_i); radius^7 uniform
    rho = np.sqrt(np.sum(d ** 2 / w_dot, axis=-1)).ravel()
    assert rho.max() <= 1.0 + 1e-12
    assert abs(np.mean(rho ** 7) - 0.5) < 0.02 and abs(np.mean(d)) < 0.01
    # gaussian: N(0, w_i^2)
    assert np.all(np.abs(e.std(axis=(0, 1)) / w_est - 1.0) < 0.05) and np.all(np.abs(e.mean(axis=(0, 1))) < 0.05 * w_est)
    # the state estimate that reaches the controller is the disturbed plant state plus the estimation noise
    assert np.abs(lg["DisturbedX"][:, 1:] - lg["CiLX"][:, 1:]).max() > 1e-3
    # counter-based: same seed -> same rows for the same vehicle index, whatever else is in the batch
    lg2 = run(np.arange(B), seed=7)
    assert np.array_equal(lg2["CiLX"], lg["CiLX"]) and np.array_equal(lg2["sim_disturbance_derivatives"], d)
    lg3 = run(np.arange(B), seed=8)
    assert not np.array_equal(lg3["sim_disturbance_derivatives"], d)
    # playback of the logged rows reproduces the run bit for bit
    S = _solver(ocp, B)
    sim = ClosedLoopSim(S, "monteblanco", log_capacity=nsteps)
    sim.set_playback(d, e)
    sim.reset(X0)
    sim.run(nsteps)
    lp = sim.logs()
    assert np.array_equal(lp["CiLX"], lg["CiLX"]) and np.array_equal(lp["simU"], lg["simU"])
    # and the host loop with the oracle, fed the same rows, agrees
    for b in (0, 17):
        CiLX, MPCX, simU, REF, DBG = _host_closed_loop(ocp, "monteblanco", X0[b], 60, w_deriv=d[b], w_est=e[b])
        assert np.abs(lg["CiLX"][b, :61] - CiLX).max() <= 1e-5 and np.abs(lg["simU"][b, :60] - simU).max() <= 1e-5


def test_log_ranges_are_slices_of_the_full_logs():
    """b200nmpc_sim_get_logs_range: the chunked callers' read-back equals the same rows of the full history,
    for ranges after every run(), an interior range, the empty range; a range past the steps done is refused."""
    S = _solver(OcpConfig(N=20, Tp=1.6), 5)
    sim = ClosedLoopSim(S, "lvms", log_capacity=40)
    sim.reset(sim.initial_states([0, 100, 200, 300, 400]))
    chunks = []
    for k in (7, 1, 12):
        n0 = sim.steps_done
        sim.run(k)
        chunks.append((n0, sim.logs_range(n0)))
    full = sim.logs()
    assert full["simU"].shape[1] == 20
    for n0, lg in chunks + [(3, sim.logs_range(3, 9)), (20, sim.logs_range(20)), (0, sim.logs_range(0))]:
        c = lg["simU"].shape[1]
        for key in ("simU", "simREF", "ref_idx"):
            assert np.array_equal(lg[key], full[key][:, n0:n0 + c]), key
        for key in ("MPC_SimX", "CiLX"):
            assert np.array_equal(lg[key], full[key][:, n0:n0 + c + 1]), key
        assert np.array_equal(lg["simSolverDebug"][:, :, [0, 2, 3, 4]], full["simSolverDebug"][:, n0:n0 + c][:, :, [0, 2, 3, 4]])
    part = sim.logs_range(4, 6, keys=("CiLX", "ref_idx"))
    assert set(part) == {"CiLX", "ref_idx"} and np.array_equal(part["CiLX"], full["CiLX"][:, 4:11])
    assert np.array_equal(part["ref_idx"], full["ref_idx"][:, 4:10])
    with pytest.raises(ValueError):
        sim.logs_range(0, 1, keys=("nope",))
    with pytest.raises(ValueError):
        sim.logs_range(15, 6)
    with pytest.raises(ValueError):
        sim.logs_range(-1, 2)


@pytest.mark.parametrize("variant", ["nominal", "snmpc"])
def test_reset_of_single_instances(variant):
    """b200nmpc_sim_reset_instances: one vehicle restarts at step 10 (its RL episode ended) while the batch goes on.  Its
    following rows equal, bit for bit, the first rows of a fresh run from the same state; the other vehicles equal an
    uninterrupted run (instances are independent)."""
    kw = dict(variant="snmpc", rk_steps=1, n_samples=10, uph=3) if variant == "snmpc" else {}
    ocp = OcpConfig(N=20, Tp=1.6, **kw)
    S = _solver(ocp, 3)
    sim = ClosedLoopSim(S, "monteblanco", log_capacity=40)
    x_a = sim.initial_states([0, 100, 300])
    sim.reset(x_a)
    sim.run(20)
    straight = sim.logs()
    sim.reset(x_a)
    sim.run(10)
    x_new = sim.initial_states([0, 200, 0])[1]
    sim.reset_instances([1], x_new[None, :])
    sim.run(10)
    mixed = sim.logs()
    for b in (0, 2):
        for key in ("CiLX", "MPC_SimX", "simU", "simREF", "ref_idx"):
            assert np.array_equal(mixed[key][b], straight[key][b]), (b, key)
    sim.reset(sim.initial_states([200, 200, 200]))
    sim.run(10)
    fresh = sim.logs()
    assert np.array_equal(mixed["CiLX"][1, 10:21], fresh["CiLX"][1, :11])
    assert np.array_equal(mixed["MPC_SimX"][1, 10:21], fresh["MPC_SimX"][1, :11])
    for key in ("simU", "simREF", "ref_idx"):
        assert np.array_equal(mixed[key][1, 10:20], fresh[key][1, :10]), key
    assert np.array_equal(mixed["simSolverDebug"][1, 10:20, 3], fresh["simSolverDebug"][1, :10, 3])      # same iteration counts
    # a boolean mask and one row for all
    sim.reset_instances(np.array([True, False, True]), x_new)
    assert sim.steps_done == 10
    with pytest.raises(ValueError):
        sim.reset_instances([0], np.zeros((2, 8)))
