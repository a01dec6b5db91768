"""Host-side pieces of the batched evaluation / closed-loop API that need no GPU: metric formulas against the
reference's expressions, retyped here from Utils/MPC_sim_utils.py:102-112, Utils/Logging_Plotting.py:164-179 and
Learning_To_Adapt/SafeRL_WMPC/RL_WMPC/reward.py:17-38."""
import numpy as np

from tum_control_b200.closed_loop import deviations, lon_lat_deviations
from tum_control_b200.config import VehicleModel
from tum_control_b200.evaluation import _interp_lin, combined_acceleration, root_mean_square


def test_lon_lat_deviations_rotate_into_the_vehicle_frame():
    rng = np.random.default_rng(0)
    yaw, x, y, rx, ry = rng.normal(size=(5, 50))
    dlon, dlat = lon_lat_deviations(yaw, x, y, rx, ry)
    rotcos, rotsin = np.cos(-yaw), np.sin(-yaw)                       # MPC_sim_utils.py:108-111
    np.testing.assert_allclose(dlon, rotcos * (rx - x) - rotsin * (ry - y), rtol=0, atol=1e-15)
    np.testing.assert_allclose(dlat, rotsin * (rx - x) + rotcos * (ry - y), rtol=0, atol=1e-15)
    # a reference point straight ahead has no lateral deviation; one to the left has a positive one
    assert abs(lon_lat_deviations(0.3, 0.0, 0.0, np.cos(0.3), np.sin(0.3))[1]) < 1e-15
    assert lon_lat_deviations(0.0, 0.0, 0.0, 0.0, 1.0)[1] == 1.0
    CiLX = np.zeros((3, 7)); CiLX[:, 3] = [10.0, 11.0, 12.0]
    REF = np.zeros((3, 4)); REF[:, 3] = 10.5; REF[:, 1] = 0.2
    lat, vel = deviations(CiLX, REF)
    np.testing.assert_allclose(lat, 0.2); np.testing.assert_allclose(vel, [-0.5, 0.5, 1.5])


def test_combined_acceleration_as_logged_by_the_reference():
    veh = VehicleModel()
    CiLX = np.zeros((4, 7)); CiLX[:, 3] = [5.0, 11.5, 20.0, 30.0]; CiLX[:, 5] = [0.1, 0.2, -0.1, 0.05]
    a = np.array([1.0, 2.0, -2.0, -3.5])
    got = combined_acceleration(CiLX, a, veh)
    for j in range(4):
        v = CiLX[j, 3]
        alat = v * CiLX[j, 5]                                         # sim_model_dynamic_stm_pacejka.py:223
        alat_lim = _interp_lin(veh.ggv_v, veh.ggv_ay, v)
        if a[j] > 0:                                                  # Logging_Plotting.py:170-176
            alon_nor = a[j] / _interp_lin(veh.ggv_v, veh.ggv_ax, v)
        else:
            alon_nor = abs(a[j]) / (-3.5)
        assert abs(got[j] - np.sqrt(alon_nor ** 2 + (alat / alat_lim) ** 2)) < 1e-15
    # braking at the limit with no lateral acceleration uses the whole budget
    assert abs(combined_acceleration(np.array([0, 0, 0, 20.0, 0, 0, 0]), np.array(-3.5), veh) - 1.0) < 1e-15


def test_linear_interpolant_extrapolates_like_casadi():
    xs, ys = np.array([0.0, 1.0, 3.0]), np.array([0.0, 2.0, 2.0])
    np.testing.assert_allclose(_interp_lin(xs, ys, np.array([-1.0, 0.5, 1.0, 2.0, 4.0])), [-2.0, 1.0, 2.0, 2.0, 2.0])
    assert root_mean_square(np.array([3.0, 4.0])) == np.sqrt(12.5)


def test_segment_objectives_match_the_per_instance_loop():
    """Batched `_get_objectives` (objective_function.py:176-184): max |lat_dev| and rms(vel_dev) over the first steps[b]
    columns; columns after the last step are masked, a NaN inside the range propagates."""
    from tum_control_b200.evaluation import segment_objectives
    rng = np.random.default_rng(0)
    lat, vel = rng.normal(size=(6, 25)), rng.normal(size=(6, 25))
    steps = np.array([25, 1, 7, 12, 3, 20])
    lat[4, 1] = np.nan                      # inside the range of instance 4
    lat[2, 9] = np.nan                      # after the last step of instance 2: must not matter
    obj, lat_m, vel_m = segment_objectives(lat, vel, steps)
    for b, k in enumerate(steps):
        ref = [-np.max(np.abs(lat[b, :k])), -root_mean_square(vel[b, :k])]
        np.testing.assert_allclose(obj[b], ref, rtol=1e-14, equal_nan=True)
        assert np.isnan(lat_m[b, k:]).all() and np.isnan(vel_m[b, k:]).all()
        assert np.array_equal(vel_m[b, :k], vel[b, :k])
    assert np.isnan(obj[4, 0]) and np.isfinite(obj[2]).all()


def test_wmpc_rewards_match_the_per_environment_loop():
    """Batched reward of RL_WMPC/reward.py:17-38 over the first length[b] steps of every environment."""
    from tum_control_b200.evaluation import wmpc_rewards
    rng = np.random.default_rng(1)
    lat, vel = 0.3 * rng.normal(size=(5, 20)), 0.8 * rng.normal(size=(5, 20))
    length = np.array([20, 20, 4, 1, 13])
    lims, sig = np.array([[0.0, 0.4], [0.0, 1.0]]), np.array([0.1, 0.5])
    r = wmpc_rewards(lat, vel, length, lims, sig)
    for b, k in enumerate(length):
        m = np.array([root_mean_square(lat[b, :k]), root_mean_square(vel[b, :k])])
        nor = np.clip((m - lims[:, 0]) / (lims[:, 1] - lims[:, 0]), 0.0, 1.0)
        assert abs(r[b] - np.exp(-np.sum(nor ** 2 / (2.0 * sig)))) <= 1e-15
    assert np.all((r > 0) & (r <= 1))
