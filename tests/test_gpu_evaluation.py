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


def test_segment_objectives_match_reference_laps():
    keys = ["monteblanco_0", "monteblanco_17"]
    logs = [load_log(k) for k in keys]
    ref = load_reftraj("monteblanco")
    ends = [120, 200]
    ev = BatchedSegmentEvaluator("monteblanco", batch=2, ocp=golden_ocp(), max_steps=1200, chunk=200)
    out = ev.evaluate(np.stack([lg["params"] for lg in logs]), start_idx=[0, 0], end_idx=ends)
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
    r2, _, _ = env.step([1, 1, 1, 1])
    v, lat_dev, vel_dev = env.observation_states()
    assert v.shape == (4,) and np.all(np.isfinite(r2))
