"""Batched closed-loop evaluation for the reference's weight-adaptation callers.

The reference evaluates one (parameter set, track segment) pair per Python closed loop -- eight of them at a time
in a ``multiprocessing.Pool`` for Bayesian optimisation (``BO_WMPC/objective_function.py:25-200``,
``bayesian_optimization.py:372-407``) and sixteen forked environments for PPO (``RL_WMPC/environment.py:112-240``).
Here every pair is one instance of a batched on-device closed loop (:mod:`tum_control_b200.closed_loop`): per-instance
weight sets through ``update_cost_function_weights([B, 7])``, per-instance start / end indices, and the crash
predicates and metrics of the reference evaluated on the device-written logs.

Semantics kept from the reference:
  * ``objective_function``: a segment ends when the planner's closest index equals ``segment['end']`` (:153); it
    crashes when ``lat_devs[step-1] > max_lat_dev`` (signed, :188) or ``acomb > max_a_comb`` (:189) with ``acomb``
    as logged by ``Logger.logging_step`` (``Logging_Plotting.py:164-179``); objectives are
    ``[-max |lat_dev|, -rms(vel_dev)]`` over the steps simulated (:176-184); an unfeasible pair yields NaN (:171-172).
  * ``RLEnvironment.step``: ``n_mpc_steps`` closed-loop steps under the selected weight set, reward
    ``exp(-sum(clip(normalize(rms metrics))^2 / (2 sigma)))`` (``RL_WMPC/reward.py:17-38``), truncation when
    ``lat_devs[step-1] > max_lat_dev`` (:239-240).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from .closed_loop import ClosedLoopSim, deviations
from .config import OcpConfig
from .solver import BatchedOcpSolver


def combined_acceleration(CiLX: np.ndarray, a_lon: np.ndarray, vehicle) -> np.ndarray:
    """``acomb`` of ``Logger.logging_step`` (``Logging_Plotting.py:164-179``): lateral acceleration ``v_lon * yaw_rate``
    and the MPC's longitudinal acceleration, each normalised by the g-g-v limit at the current speed."""
    v = CiLX[..., 3]
    alat = v * CiLX[..., 5]
    ay_lim = _interp_lin(vehicle.ggv_v, vehicle.ggv_ay, v)
    ax_lim = _interp_lin(vehicle.ggv_v, vehicle.ggv_ax, v)
    # braking: the reference divides |a| by acc_min = -3.5 (sign kept: the quotient is squared)
    alon_nor = np.where(a_lon > 0, a_lon / ax_lim, np.abs(a_lon) / (-vehicle.ax_brake))
    return np.sqrt(alon_nor ** 2 + (alat / ay_lim) ** 2)


def _interp_lin(xs, ys, x):
    """casadi.interpolant('linear'): piecewise linear with linear extrapolation (``NMPC_class.py:332-333``)."""
    xs, ys = np.asarray(xs, float), np.asarray(ys, float)
    i = np.clip(np.searchsorted(xs, x, side="right") - 1, 0, len(xs) - 2)
    return ys[i] + (ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]) * (x - xs[i])


def root_mean_square(a, axis=-1):
    return np.sqrt(np.mean(np.square(a), axis=axis))


def segment_objectives(lat, vel, steps):
    """``_get_objectives`` of ``objective_function.py:176-184`` for a batch: ``[-max |lat_dev|, -rms(vel_dev)]`` over the
    first ``steps[b]`` columns of every row.  Returns ``(objectives [B, 2], lat, vel)`` with the columns after an
    instance's last step set to NaN; a NaN inside the simulated range propagates like ``np.max`` / ``np.mean`` would."""
    lat, vel, steps = np.asarray(lat, float), np.asarray(vel, float), np.asarray(steps, int)
    after = np.arange(lat.shape[1])[None, :] >= steps[:, None]
    obj = np.stack([-np.max(np.where(after, 0.0, np.abs(lat)), axis=1),
                    -np.sqrt(np.sum(np.where(after, 0.0, np.square(vel)), axis=1) / np.maximum(steps, 1))], axis=1)
    return obj, np.where(after, np.nan, lat), np.where(after, np.nan, vel)


def wmpc_rewards(lat, vel, length, lims, sigmas):
    """Reward of ``RL_WMPC/reward.py:17-38`` for a batch: ``exp(-sum(clip(normalize(m))^2 / (2 sigma)))`` with
    ``m = [rms(lat_dev), rms(vel_dev)]`` over the first ``length[b]`` steps of every row (the reference stops the inner
    loop at the crashing step)."""
    lat, vel, length = np.asarray(lat, float), np.asarray(vel, float), np.asarray(length, int)
    lims, sigmas = np.asarray(lims, float), np.asarray(sigmas, float)
    after = np.arange(lat.shape[1])[None, :] >= length[:, None]
    n = np.maximum(length, 1)
    metrics = np.stack([np.sqrt(np.sum(np.where(after, 0.0, np.square(lat)), axis=1) / n),
                        np.sqrt(np.sum(np.where(after, 0.0, np.square(vel)), axis=1) / n)], axis=1)
    nor = np.clip((metrics - lims[:, 0]) / (lims[:, 1] - lims[:, 0]), 0.0, 1.0)
    return np.exp(-np.sum(nor ** 2 / (2.0 * sigmas), axis=1))


class BatchedSegmentEvaluator:
    """``objective_function`` for ``B`` (parameter set, segment) pairs at once."""

    def __init__(self, track: str, batch: int, ocp: Optional[OcpConfig] = None, device: int = 0, max_steps: int = 3000,
                 chunk: int = 100):
        self.ocp = ocp if ocp is not None else OcpConfig(N=38, Tp=3.04)     # sim_main_params.yaml: Tp 3.04, Ts_MPC 0.08
        self.batch, self.max_steps, self.chunk = int(batch), int(max_steps), int(chunk)
        self.solver = BatchedOcpSolver(self.ocp, batch=self.batch, device=device, squeeze=False)
        self.sim = ClosedLoopSim(self.solver, track, log_capacity=self.max_steps)

    def evaluate(self, params, start_idx: Sequence[int], end_idx: Sequence[int], max_lat_dev: float = 2.0,
                 max_a_comb: float = 1.02, device_stop: bool = True) -> Dict[str, np.ndarray]:
        """params ``[B, 7]`` or ``[7]`` = ``[q_xy, q_yaw, q_vel, r_jerk, r_steer, L1, L2]``.  Returns ``objectives [B, 2]``
        (NaN rows where unfeasible), ``feasible [B]``, ``steps [B]`` (closed-loop steps until end / crash), ``done [B]``,
        ``lat_dev`` / ``vel_dev [B, n]`` (NaN after an instance's last step).

        ``device_stop=True``: the end / crash predicates run on the device after every step and a finished pair drops
        out of the batch (``b200nmpc_sim_set_stop``); the host only polls the stop state between chunks.
        ``device_stop=False``: the whole batch runs until the last pair has finished and the predicates are evaluated
        here on the ranged log read-backs (the two paths agree: ``tests/test_gpu_evaluation.py``)."""
        B = self.batch
        end = np.broadcast_to(np.asarray(end_idx, int), (B,))
        self.solver.update_cost_function_weights(np.asarray(params, float))
        self.sim.set_stop_conditions(end, max_lat_dev, max_a_comb, enable=device_stop)
        self.sim.reset(self.sim.initial_states(start_idx))
        stop = np.full(B, -1)
        crashed = np.zeros(B, bool)
        done = np.zeros(B, bool)
        if device_stop:
            while self.sim.steps_done < self.max_steps and np.any(stop < 0):
                self.sim.run(min(self.chunk, self.max_steps - self.sim.steps_done))
                stop, done, crashed = self.sim.stop_state()
            lg = self.sim.logs_range(0, keys=("CiLX", "simREF"))
            lat, vel = deviations(lg["CiLX"][:, :-1], lg["simREF"])
        else:
            lat_all, vel_all = [], []
            while self.sim.steps_done < self.max_steps and np.any(stop < 0):
                n0 = self.sim.steps_done
                self.sim.run(min(self.chunk, self.max_steps - n0))
                lg = self.sim.logs_range(n0)                 # only the rows this chunk wrote
                lat, vel = deviations(lg["CiLX"][:, :-1], lg["simREF"])
                acomb = combined_acceleration(lg["CiLX"][:, :-1], lg["MPC_SimX"][:, 1:, 7], self.ocp.vehicle)
                lat_all.append(lat); vel_all.append(vel)
                hit_end = lg["ref_idx"] == end[:, None]
                crash = (lat > max_lat_dev) | (acomb > max_a_comb)
                event = hit_end | crash
                first = event.argmax(axis=1)
                for b in np.nonzero((stop < 0) & event.any(axis=1))[0]:
                    i = int(first[b])
                    stop[b] = n0 + i + 1                     # steps simulated, including the one that ended the run
                    crashed[b] = bool(crash[b, i])
                    done[b] = bool(hit_end[b, i])
            lat, vel = np.concatenate(lat_all, axis=1), np.concatenate(vel_all, axis=1)
        steps = np.where(stop < 0, self.sim.steps_done, stop)
        obj, lat, vel = segment_objectives(lat, vel, steps)
        feasible = ~crashed
        obj[~feasible] = np.nan
        return {"objectives": obj, "feasible": feasible, "steps": steps, "done": done, "lat_dev": lat, "vel_dev": vel}


class BatchedWMPCEnvironment:
    """``RLEnvironment.step`` for ``B`` environments at once: every environment selects one of the parameter sets
    (``_parameters/F.csv`` rows), the batch advances ``n_mpc_steps`` closed-loop steps, rewards and truncation flags
    come back as arrays (defaults: ``_config/rl_config.yaml:6-7,33-35``).  The observation generator of the RL stack is out of scope; :meth:`observation_states`
    returns the quantities it consumes (``Logger.get_observation_states``)."""

    def __init__(self, track: str, batch: int, parameter_sets, ocp: Optional[OcpConfig] = None, device: int = 0,
                 n_mpc_steps: int = 20, max_lat_dev: float = 2.0, sigmas=(0.1, 0.5), normalization_lims=((0.0, 0.4), (0.0, 1.0)),
                 capacity: int = 6000):
        self.ocp = ocp if ocp is not None else OcpConfig(N=38, Tp=3.04)
        self.batch = int(batch)
        self.parameter_sets = np.asarray(parameter_sets, float)
        self.n_mpc_steps, self.max_lat_dev = int(n_mpc_steps), float(max_lat_dev)
        self.sigmas, self.lims = np.asarray(sigmas, float), np.asarray(normalization_lims, float)
        if int(capacity) < self.n_mpc_steps:
            raise ValueError(f"capacity {capacity} is smaller than one environment step ({self.n_mpc_steps} MPC steps)")
        self.solver = BatchedOcpSolver(self.ocp, batch=self.batch, device=device, squeeze=False)
        self.sim = ClosedLoopSim(self.solver, track, log_capacity=capacity)

    def reset(self, start_idx=0):
        self.sim.reset(self.sim.initial_states(start_idx))

    def reset_envs(self, envs, start_idx):
        """Restart single environments (their episode ended) while the others keep their state: ``envs`` is a list of
        environment indices, ``start_idx`` one reference index per listed environment (or one for all)."""
        envs = np.asarray(envs, int).reshape(-1)
        starts = np.broadcast_to(np.asarray(start_idx, int), envs.shape)
        full = np.zeros(self.batch, int)
        full[envs] = starts
        self.sim.reset_instances(envs, self.sim.initial_states(full)[envs])

    def step(self, actions):
        """actions: int ``[B]`` indices into the parameter sets.  Returns (reward [B], truncated [B], info)."""
        a = np.broadcast_to(np.asarray(actions, int), (self.batch,))
        self.solver.update_cost_function_weights(self.parameter_sets[a])
        if self.sim.steps_done + self.n_mpc_steps > self.sim.capacity:
            self.sim.rebase_logs()          # the device log is linear: restart it at the current state, nothing else changes
        n0 = self.sim.steps_done
        self.sim.run(self.n_mpc_steps)
        lg = self.sim.logs_range(n0)
        lat, vel = deviations(lg["CiLX"][:, :-1], lg["simREF"])
        crash = lat > self.max_lat_dev
        truncated = crash.any(axis=1)
        # the reference stops the inner loop at the crashing step: metrics over the steps up to and including it
        length = np.where(truncated, crash.argmax(axis=1) + 1, lat.shape[1])
        reward = wmpc_rewards(lat, vel, length, self.lims, self.sigmas)
        self._last = (lg["CiLX"][:, -2, 3], lat[:, -1], vel[:, -1])
        return reward, truncated, {"ref_idx": lg["ref_idx"][:, -1], "steps": length}

    def observation_states(self):
        """(v, lat_dev, vel_dev) of the last logged step (``Logging_Plotting.py`` ``get_observation_states``)."""
        return self._last
