"""Host-side input producers against vectors produced by the reference's OWN functions
(tests/golden/make_golden.py imports Utils/MPC_sim_utils.PlannerEmulator and
Utils/SimulationMode_main_class.moving_average_filter)."""
import os

import numpy as np
import pytest

from tum_control_b200.config import load_reftraj
from tum_control_b200.planner import PlannerEmulator, StateEstimation, synthetic_batch
from replay_util import load_log, replay_inputs


@pytest.mark.parametrize("track", ["monteblanco", "lvms", "modena"])
def test_planner_window_matches_reference(track, golden_dir):
    g = np.load(os.path.join(golden_dir, "planner_golden.npz"))
    pl = PlannerEmulator(load_reftraj(track))
    for pose, win, c, (N, Tp) in zip(g[track + "_pose"], g[track + "_win"], g[track + "_closest"], g[track + "_cfg"]):
        N = int(N)
        ci, w = pl(pose, N + 1, float(Tp))
        assert ci == int(c)
        got = np.stack([w[k] for k in ("pos_x", "pos_y", "ref_yaw", "ref_v")], axis=1)
        np.testing.assert_allclose(got, win[: N + 1], rtol=0, atol=1e-12)


def test_moving_average_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "planner_golden.npz"))
    data, out = g["maf_data"], g["maf_out"]
    for wi, w in enumerate(range(1, 6)):
        se = StateEstimation(windows=[w])
        for n in range(1, 16):
            y = se([data[n - 1]])
            assert abs(y[0] - out[n - 1, wi]) <= 1e-15


def test_replay_reference_point_matches_log():
    """simREF[i] is the first planner point of step i (Logging_Plotting.py:140-143)."""
    log = load_log("lvms_3")
    for i, x0, yref in replay_inputs(log, 300):
        np.testing.assert_allclose(yref[0, :4], log["simREF"][i], rtol=0, atol=1e-9)


def test_synthetic_batch_is_deterministic():
    x0a, ya = synthetic_batch("monteblanco", 16, 30, 2.4, seed=0)
    x0b, yb = synthetic_batch("monteblanco", 16, 30, 2.4, seed=0)
    assert x0a.shape == (16, 8) and ya.shape == (16, 31, 6)
    assert np.array_equal(x0a, x0b) and np.array_equal(ya, yb)
    assert np.all(ya[:, :, 4:] == 0)
