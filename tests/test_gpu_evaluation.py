"""Batched closed-loop evaluation API of the BO / RL callers against the reference's
committed laps: the segment objectives computed from a golden log must equal those of the device loop."""
import numpy as np
import pytest

from tum_control_b200.closed_loop import deviations
from tum_control_b200.config import load_reftraj
from tum_control_b200.evaluation import BatchedSegmentEvaluator, BatchedWMPCEnvironment, root_mean_square
from replay_util import golden_ocp, load_log

pytestmark = pytest.mark.gpu


def _golden_objectives(log, ref, end):
    """objective_function.py:153,176-184 evaluated on a committed log: steps until the closest index equals `end`."""
    hit = np.nonzero((log["simREF"][:, 0] == ref["pos_x"][end]) & (log["simREF"][:, 1] == ref["pos_y"][end]))[0]
    k = int(hit[0]) + 1
    lat, vel = deviations(log["CiLX"][:k], log["simREF"][:k])
    return k, np.array([-np.max(np.abs(lat)), -root_mean_square(vel)])


@pytest.mark.parametrize("device_stop", [True, False])
def test_segment_objectives_match_reference_laps(device_stop):
    keys = ["monteblanco_0", "monteblanco_17"]
    logs = [load_log(k) for k in keys]
    ref = load_reftraj("monteblanco")
    ends = [120, 200]
    ev = BatchedSegmentEvaluator("monteblanco", batch=2, ocp=golden_ocp(), max_steps=1200, chunk=200)
    out = ev.evaluate(np.stack([lg["params"] for lg in logs]), start_idx=[0, 0], end_idx=ends, device_stop=device_stop)
    assert out["feasible"].all() and out["done"].all()
    for b, lg in enumerate(logs):
        k, obj = _golden_objectives(lg, ref, ends[b])
        assert out["steps"][b] == k
        np.testing.assert_allclose(out["objectives"][b], obj, rtol=0, atol=2e-5)


def test_crash_predicate_marks_pair_unfeasible():
    lg = load_log("monteblanco_0")
    ev = BatchedSegmentEvaluator("monteblanco", batch=2, ocp=golden_ocp(), max_steps=400, chunk=100)
    # instance 0: generous limits; instance 1 shares the run but the limit below is violated within the first metres
    out = ev.evaluate(lg["params"], start_idx=[0, 300], end_idx=[60, 360], max_lat_dev=0.02)
    assert not out["feasible"].all()
    bad = ~out["feasible"]
    assert np.isnan(out["objectives"][bad]).all()
    assert np.isfinite(out["objectives"][~bad]).all()


def test_device_stop_conditions_equal_host_predicates():
    """b200nmpc_sim_set_stop: the end / crash predicates evaluated on the device after every step (finished pairs drop
    out of the batch) give the same steps, flags and objectives as the host predicates on the full-batch run; a mixed
    batch: pairs that reach their end, pairs that crash on a tight lateral limit, pairs cut off by max_steps."""
    B = 24
    rng = np.random.default_rng(5)
    base = load_log("monteblanco_0")["params"]
    params = base * np.exp(rng.uniform(-0.8, 0.8, size=(B, 7)))
    starts = rng.integers(0, 1000, size=B)
    ends = starts + rng.integers(15, 90, size=B)
    ev = BatchedSegmentEvaluator("monteblanco", batch=B, ocp=golden_ocp(), max_steps=260, chunk=65)
    dev = ev.evaluate(params, starts, ends, max_lat_dev=0.12, max_a_comb=1.02, device_stop=True)
    batch_steps_dev = ev.sim.steps_done
    st, dn, cr = ev.sim.stop_state()
    host = ev.evaluate(params, starts, ends, max_lat_dev=0.12, max_a_comb=1.02, device_stop=False)
    assert np.array_equal(dev["steps"], host["steps"])
    assert np.array_equal(dev["done"], host["done"]) and np.array_equal(dev["feasible"], host["feasible"])
    assert np.array_equal(np.isnan(dev["objectives"]), np.isnan(host["objectives"]))
    ok = dev["feasible"]
    assert np.array_equal(dev["objectives"][ok], host["objectives"][ok])       # same device arithmetic up to the stop
    assert np.array_equal(np.isnan(dev["lat_dev"]), np.isnan(host["lat_dev"][:, :dev["lat_dev"].shape[1]]))
    # the batch holds all three outcomes, and the stop state mirrors them
    assert dev["done"].any() and (~dev["feasible"]).any() and (dev["steps"] == 260).any()
    assert np.array_equal(st >= 0, dev["done"] | ~dev["feasible"]) and np.array_equal(cr, ~dev["feasible"])
    assert np.array_equal(st[st >= 0], dev["steps"][st >= 0]) and batch_steps_dev == 260


@pytest.mark.parametrize("variant", ["nominal", "snmpc"])
def test_stopped_instance_is_frozen(variant):
    """A stopped instance keeps its plant state, iterate (SNMPC: sample iterates too) and logs while the rest of the
    batch goes on."""
    from tum_control_b200.closed_loop import ClosedLoopSim
    from tum_control_b200.config import OcpConfig
    from tum_control_b200.solver import BatchedOcpSolver
    ocp = golden_ocp() if variant == "nominal" else OcpConfig(N=20, Tp=1.6, variant="snmpc", rk_steps=1, n_samples=10, uph=3)
    S = BatchedOcpSolver(ocp, batch=3, device=0, squeeze=False)
    sim = ClosedLoopSim(S, "monteblanco", log_capacity=80)
    sim.set_stop_conditions(end_idx=[5, 2000, 9])          # 2000: never reached
    sim.reset(sim.initial_states([0, 50, 0]))
    sim.run(40)
    st, dn, cr = sim.stop_state()
    assert st[0] > 0 and st[2] > st[0] and st[1] == -1 and dn[0] and dn[2] and not dn[1] and not cr.any()
    lg40, x40, u40 = sim.logs(), S.get_iterate()[0].copy(), S.get_iterate()[1].copy()
    xs40 = S.get_samples().copy() if variant == "snmpc" else None
    sim.run(40)
    lg80, x80, u80 = sim.logs(), *S.get_iterate()
    if variant == "snmpc":
        xs80 = S.get_samples()
        assert np.array_equal(xs40[0], xs80[0]) and np.array_equal(xs40[2], xs80[2]) and not np.array_equal(xs40[1], xs80[1])
    for b in (0, 2):
        k = int(st[b])
        assert lg40["ref_idx"][b, k - 1] == (5, 0, 9)[b]
        assert np.array_equal(x40[b], x80[b]) and np.array_equal(u40[b], u80[b])
        for key in ("CiLX", "MPC_SimX"):
            assert np.array_equal(lg80[key][b, :k + 1], lg40[key][b, :k + 1])
            assert np.array_equal(lg80[key][b, k + 1:41], lg40[key][b, k + 1:41])      # rows after the stop: untouched
        assert np.array_equal(lg80["simU"][b, :k], lg40["simU"][b, :k])
    assert not np.array_equal(x40[1], x80[1])
    # without stop conditions the same batch runs on: instance 1 is bit-identical in both runs (instances are independent)
    sim.set_stop_conditions(enable=False)
    sim.reset(sim.initial_states([0, 50, 0]))
    sim.run(80)
    free = sim.logs()
    assert np.array_equal(free["CiLX"][1], lg80["CiLX"][1]) and np.array_equal(free["simU"][1], lg80["simU"][1])
    assert np.array_equal(free["CiLX"][0, :int(st[0]) + 1], lg80["CiLX"][0, :int(st[0]) + 1])
    assert (sim.stop_state()[0] == -1).all()


def test_wmpc_environment_step():
    """RLEnvironment.step semantics (environment.py:112-189): 20 closed-loop steps per action, reward from the RMS
    deviations of those steps (reward.py:17-38)."""
    lg = load_log("monteblanco_0")
    sets = np.stack([lg["params"], load_log("monteblanco_17")["params"]])
    env = BatchedWMPCEnvironment("monteblanco", batch=4, parameter_sets=sets, ocp=golden_ocp())
    env.reset([0, 0, 100, 400])
    r1, trunc, info = env.step([0, 1, 0, 1])
    assert r1.shape == (4,) and np.all((r1 > 0) & (r1 <= 1)) and not trunc.any()
    # environment 0 runs the committed lap 0 with its own weight set: reward from the golden log
    lat, vel = deviations(lg["CiLX"][:20], lg["simREF"][:20])
    m = np.array([root_mean_square(lat), root_mean_square(vel)])
    nor = np.clip(m / np.array([0.4, 1.0]), 0, 1)
    assert abs(r1[0] - np.exp(-np.sum(nor ** 2 / (2 * np.array([0.1, 0.5]))))) <= 1e-6
    # environment 1 ends its episode and restarts where environment 0 started: its next step repeats environment 0's first
    env.reset_envs([1], [0])
    r2, _, info = env.step([1, 0, 0, 1])
    assert r2[1] == r1[0] and np.all((r2 > 0) & (r2 <= 1))
    r2, _, _ = env.step([1, 1, 1, 1])
    v, lat_dev, vel_dev = env.observation_states()
    assert v.shape == (4,) and np.all(np.isfinite(r2))


def test_wmpc_environment_runs_past_the_log_capacity():
    """The device log is a linear array: an environment stepped beyond `capacity` closed-loop steps rebases the log at the
    current state and must return exactly what an environment with a large log returns; a raw run past the capacity is
    an error instead of silently unlogged steps (the stop predicates read the log row of the step)."""
    lg = load_log("monteblanco_0")
    sets = np.stack([lg["params"], load_log("monteblanco_17")["params"]])
    big = BatchedWMPCEnvironment("monteblanco", batch=3, parameter_sets=sets, ocp=golden_ocp())
    small = BatchedWMPCEnvironment("monteblanco", batch=3, parameter_sets=sets, ocp=golden_ocp(), capacity=50)
    for env in (big, small):
        env.reset([0, 100, 400])
    for i in range(6):                               # 120 closed-loop steps > capacity 50: two rebases
        a = [i % 2, (i + 1) % 2, 0]
        rb, tb, ib = big.step(a)
        rs, ts, isd = small.step(a)
        assert np.array_equal(rb, rs) and np.array_equal(tb, ts) and np.array_equal(ib["ref_idx"], isd["ref_idx"])
        if i == 3:                                   # a single environment restarts in both
            big.reset_envs([2], [250]); small.reset_envs([2], [250])
    assert all(np.array_equal(a, b) for a, b in zip(big.observation_states(), small.observation_states()))
    assert small.sim.steps_done <= 50 and big.sim.steps_done == 120
    with pytest.raises(ValueError, match="log capacity"):
        small.sim.run(51)
    t = small.sim.logs()["simSolverDebug"][:, :, 1]
    assert np.all(t > 0)                             # time_tot stays meaningful
