"""Batched closed-loop simulation on the device: the loop of ``main.py:48-78`` (planner emulator -> solve ->
plant step -> state estimation -> logging) for ``B`` independent vehicles, advancing without host round trips.

Mirrors, with a leading batch axis, what the reference spreads over
  * ``Utils/SimulationMode_main_class.py`` (``MPC_Sim``: ``X0_MPC`` :44-66, ``sim_step`` :106-150, ``StateEstimation`` :152-156),
  * ``Utils/MPC_sim_utils.py:137-194`` (``PlannerEmulator``),
  * ``Vehicle_Simulator/VehicleSimulator.py:60-94`` (RK4 plant, 4 finite elements over ``Ts``),
  * ``Utils/Logging_Plotting.py:124-145,276-282,330-372`` (log arrays and the ``full_logs.npz`` schema).
The arithmetic runs in ``libb200nmpc.so`` (``b200nmpc_sim_*``, ``csrc/k_sim.cu``); there is no host fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from .config import SIM_MAIN_PARAMS, STATE_ESTIMATION_WINDOWS, load_reftraj
from .planner import initial_state
from .solver import BatchedOcpSolver, _check


class ClosedLoopSim:
    def __init__(self, solver: BatchedOcpSolver, track: str, Ts: Optional[float] = None, log_capacity: int = 6000,
                 plant_substeps: int = 4, windows=STATE_ESTIMATION_WINDOWS):
        self.solver = solver
        self.batch, self.N = solver.batch, solver.N
        self.track = track
        self.reftraj = load_reftraj(track)
        self.Ts = float(SIM_MAIN_PARAMS["Ts"] if Ts is None else Ts)
        self.Tp = float(solver.ocp.Tp)
        self._L = _lib.lib()
        cfg = _lib.B200SimConfig()
        cfg.Ts, cfg.Tp, cfg.plant_substeps, cfg.log_capacity = self.Ts, self.Tp, int(plant_substeps), int(log_capacity)
        cfg.n_ref = len(self.reftraj["pos_x"])
        for i, w in enumerate(windows):
            cfg.windows[i] = int(w)
        self.capacity = int(log_capacity)
        arrs = [np.ascontiguousarray(self.reftraj[k], np.float64) for k in ("pos_x", "pos_y", "ref_yaw", "ref_v")]
        h = C.c_void_p()
        _check(self._L.b200nmpc_sim_create(solver._h, C.byref(cfg), *[a.ctypes.data for a in arrs], C.byref(h)))
        self._h = h
        solver._children.add(self)          # the solver closes its closed-loop objects before it destroys the handle

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.solver, "_h", None):
                self._L.b200nmpc_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up
    def initial_states(self, start_idx) -> np.ndarray:
        """``X0_MPC`` (``SimulationMode_main_class.py:44-66``) at the given reference indices: ``[B, 8]``."""
        idx = np.broadcast_to(np.asarray(start_idx, int), (self.batch,))
        return np.stack([initial_state(self.reftraj, int(i)) for i in idx])

    def reset(self, X0_MPC):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(X0_MPC, np.float64), (self.batch, 8)))
        _check(self._L.b200nmpc_sim_reset(self._h, a.ctypes.data))

    def reset_instances(self, which, X0_MPC):
        """Reset single vehicles while the others keep running (``RLEnvironment.reset`` of one environment of a
        vectorised run): ``which`` is a boolean mask ``[B]`` or a list of instance indices, ``X0_MPC`` is ``[B, 8]``,
        ``[len(which), 8]`` (one row per listed instance) or ``[8]``.  Plant, pose, filter, stop state and the
        controller (``MPC.reset(X0_MPC)``: x0 and the initial guess) restart; the log rows continue at the current step."""
        w = np.asarray(which)
        mask = np.zeros(self.batch, np.int32)
        if w.dtype == bool:
            if w.shape != (self.batch,):
                raise ValueError(f"reset_instances: mask must have shape ({self.batch},)")
            mask[w] = 1
        else:
            mask[w.astype(int).reshape(-1)] = 1
        x = np.asarray(X0_MPC, np.float64)
        full = np.zeros((self.batch, 8))
        if x.shape == (self.batch, 8):
            full[:] = x
        elif x.ndim == 2 and x.shape == (int(mask.sum()), 8) and w.dtype != bool:
            full[w.astype(int).reshape(-1)] = x
        elif x.shape == (8,):
            full[mask != 0] = x
        else:
            raise ValueError(f"reset_instances: X0_MPC has shape {x.shape}")
        _check(self._L.b200nmpc_sim_reset_instances(self._h, mask.ctypes.data, full.ctypes.data))

    def set_disturbances(self, deriv=None, est=None):
        """Additive disturbances for the following steps (``Utils/MPC_sim_utils.py:54-99``), ``[B, 7]`` each."""
        d = None if deriv is None else np.ascontiguousarray(np.broadcast_to(np.asarray(deriv, np.float64), (self.batch, 7)))
        e = None if est is None else np.ascontiguousarray(np.broadcast_to(np.asarray(est, np.float64), (self.batch, 7)))
        _check(self._L.b200nmpc_sim_set_disturbances(self._h, None if d is None else d.ctypes.data,
                                                     None if e is None else e.ctypes.data))

    def set_stop_conditions(self, end_idx=None, max_lat_dev: float = float("inf"), max_a_comb: float = float("inf"),
                            enable: bool = True):
        """Per-instance stop conditions evaluated on the device after every step (``objective_function.py:153,186-200``):
        closest reference index ``== end_idx[b]`` (done) and ``lat_dev > max_lat_dev or acomb > max_a_comb`` (crash).
        A stopped instance keeps its state and logs and is skipped by all kernels of the following steps.
        :meth:`reset` clears the stop state; ``enable=False`` turns the predicates off."""
        e = None if end_idx is None else np.ascontiguousarray(np.broadcast_to(np.asarray(end_idx, np.int32), (self.batch,)))
        _check(self._L.b200nmpc_sim_set_stop(self._h, int(bool(enable)), None if e is None else e.ctypes.data,
                                             float(max_lat_dev), float(max_a_comb)))

    def stop_state(self):
        """``(stop_step [B], done [B], crashed [B])``: ``stop_step`` is -1 while the instance runs, else the number of
        steps it simulated including the one that stopped it."""
        st, fl = np.empty(self.batch, np.int32), np.empty(self.batch, np.int32)
        _check(self._L.b200nmpc_sim_get_stop(self._h, st.ctypes.data, fl.ctypes.data))
        return st, (fl & 1).astype(bool), (fl & 2).astype(bool)

    DISTURBANCE_TYPES = {None: 0, "off": 0, "uniform": 1, "gaussian": 2, "absolute": 3}

    def configure_disturbances(self, type_derivatives=None, bounds_derivatives=None, type_state_estimation=None,
                               bounds_state_estimation=None, seed: int = 0):
        """Disturbances generated on the device with the reference's three types (``generate_disturbances``,
        ``Utils/MPC_sim_utils.py:54-84``); ``bounds_*`` are the ``w_*`` rows of ``sim_main_params.yaml:54-79``
        (``[7]``: posx, posy, yaw, vlong, vlat, yawrate, delta_f).  Reproducible per vehicle from ``seed``."""
        t0, t1 = self.DISTURBANCE_TYPES[type_derivatives], self.DISTURBANCE_TYPES[type_state_estimation]
        b0 = np.ascontiguousarray(np.zeros(7) if bounds_derivatives is None else np.asarray(bounds_derivatives, np.float64))
        b1 = np.ascontiguousarray(np.zeros(7) if bounds_state_estimation is None else np.asarray(bounds_state_estimation, np.float64))
        _check(self._L.b200nmpc_sim_config_disturbances(self._h, t0, b0.ctypes.data, t1, b1.ctypes.data, int(seed)))

    def set_playback(self, deriv=None, est=None):
        """``disturbance_playback``: replay the ``sim_disturbance_*`` rows of a previous run, ``[B, n, 7]`` each."""
        n = (deriv if deriv is not None else est).shape[-2]
        d = None if deriv is None else np.ascontiguousarray(np.broadcast_to(np.asarray(deriv, np.float64), (self.batch, n, 7)))
        e = None if est is None else np.ascontiguousarray(np.broadcast_to(np.asarray(est, np.float64), (self.batch, n, 7)))
        _check(self._L.b200nmpc_sim_set_playback(self._h, None if d is None else d.ctypes.data,
                                                 None if e is None else e.ctypes.data, int(n)))

    # ---- run
    def run(self, nsteps: int, sync: bool = True) -> float:
        """Advance every vehicle by ``nsteps`` closed-loop steps; returns the device time in seconds."""
        sec = C.c_double(0.0)
        _check(self._L.b200nmpc_sim_run(self._h, int(nsteps), C.byref(sec) if sync else None))
        return float(sec.value)

    def rebase_logs(self):
        """Restart the log at the current state (row 0 = state after the last step, ``steps_done`` = 0) without touching
        vehicles, controller or filter: lets a chunk-wise caller run longer than ``log_capacity`` steps."""
        _check(self._L.b200nmpc_sim_rebase_logs(self._h))

    @property
    def steps_done(self) -> int:
        return int(self._L.b200nmpc_sim_steps_done(self._h))

    def plan(self, pose):
        """Planner only: fills the solver's yref from ``pose [B, 4]``; returns the closest reference indices."""
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(pose, np.float64), (self.batch, 4)))
        idx = np.zeros(self.batch, dtype=np.int32)
        _check(self._L.b200nmpc_sim_plan(self._h, a.ctypes.data, idx.ctypes.data))
        return idx

    # ---- logs
    def logs(self) -> Dict[str, np.ndarray]:
        """Batched log arrays in the reference's names (``Logging_Plotting.py:34-41``), each with a leading batch
        axis: ``MPC_SimX [B, n+1, 8]``, ``CiLX [B, n+1, 7]``, ``simU [B, n, 2]``, ``simREF [B, n, 4]``,
        ``simSolverDebug [B, n, 5] = [cost, time_tot, sqp_iter, qp_iter, status]``."""
        n, B = min(self.steps_done, self.capacity), self.batch
        out = {"MPC_SimX": np.empty((B, n + 1, 8)), "CiLX": np.empty((B, n + 1, 7)), "simU": np.empty((B, n, 2)),
               "simREF": np.empty((B, n, 4)), "simSolverDebug": np.empty((B, n, 5))}
        _check(self._L.b200nmpc_sim_get_logs(self._h, *[out[k].ctypes.data for k in
                                                         ("MPC_SimX", "CiLX", "simU", "simREF", "simSolverDebug")]))
        idx = np.zeros((B, n), dtype=np.int32)
        _check(self._L.b200nmpc_sim_get_ref_idx(self._h, idx.ctypes.data))
        out["ref_idx"] = idx          # current_ref_idx of every step (not a Logger array; the BO / RL callers need it)
        out["DisturbedX"] = np.empty((B, n + 1, 7))
        out["sim_disturbance_derivatives"] = np.zeros((B, n, 7))
        out["sim_disturbance_state_estimation"] = np.zeros((B, n, 7))
        _check(self._L.b200nmpc_sim_get_disturbance_logs(self._h, out["DisturbedX"].ctypes.data,
                                                         out["sim_disturbance_derivatives"].ctypes.data,
                                                         out["sim_disturbance_state_estimation"].ctypes.data))
        return out

    LOG_KEYS = ("MPC_SimX", "CiLX", "simU", "simREF", "simSolverDebug", "ref_idx")

    def logs_range(self, first: int, count: Optional[int] = None, keys=None) -> Dict[str, np.ndarray]:
        """The solver logs of the closed-loop steps ``[first, first + count)`` only (default: up to the last step done):
        ``MPC_SimX`` / ``CiLX`` hold ``count + 1`` rows starting at row ``first``, ``simU`` / ``simREF`` /
        ``simSolverDebug`` / ``ref_idx`` hold ``count`` rows; ``keys`` selects a subset of the arrays.  What a chunked
        caller reads after every :meth:`run` instead of the whole history (``objective_function.py:120-189`` inspects
        only the newest step)."""
        n, B = min(self.steps_done, self.capacity), self.batch
        first = int(first)
        count = n - first if count is None else int(count)
        if first < 0 or count < 0 or first + count > n:
            raise ValueError(f"log range [{first}, {first + count}) outside the {n} steps done")
        keys = self.LOG_KEYS if keys is None else tuple(keys)
        if not set(keys) <= set(self.LOG_KEYS):
            raise ValueError(f"logs_range: unknown key in {keys!r}")
        shapes = {"MPC_SimX": (B, count + 1, 8), "CiLX": (B, count + 1, 7), "simU": (B, count, 2), "simREF": (B, count, 4),
                  "simSolverDebug": (B, count, 5), "ref_idx": (B, count)}
        out = {k: (np.zeros(shapes[k], dtype=np.int32) if k == "ref_idx" else np.empty(shapes[k])) for k in keys}
        _check(self._L.b200nmpc_sim_get_logs_range(self._h, first, count,
                                                   *[out[k].ctypes.data if k in out else None for k in self.LOG_KEYS]))
        return out

    def save_logs(self, path: str, instance: int = 0, logs: Optional[Dict[str, np.ndarray]] = None):
        """``full_logs.npz`` of one vehicle with the reference's keys (``Logging_Plotting.py:330-372``): yaw columns
        wrapped to [0, 2pi) on save (:341-343), deviations of ``CiLX[1:]`` against ``simREF`` (:339-345), lateral
        acceleration ``v_lon * yaw_rate`` (``sim_model_dynamic_stm_pacejka.py:231-237``), ``t = linspace(0, T, n)``."""
        from .planner import postprocess_yaw
        lg = self.logs() if logs is None else logs
        b = int(instance)
        MPC_SimX, CiLX = lg["MPC_SimX"][b].copy(), lg["CiLX"][b].copy()
        simU, simREF, dbg = lg["simU"][b], lg["simREF"][b], lg["simSolverDebug"][b]
        n = simU.shape[0]
        dev_vel = np.abs(CiLX[1:, 3] - simREF[:, 3])
        CiLX[:, 2] = postprocess_yaw(CiLX[:, 2]); MPC_SimX[:, 2] = postprocess_yaw(MPC_SimX[:, 2])
        dev_yaw = np.abs(CiLX[1:, 2] - simREF[:, 2])
        dev_long, dev_lat = lon_lat_deviations(CiLX[1:, 2], CiLX[1:, 0], CiLX[1:, 1], simREF[:, 0], simREF[:, 1])
        np.savez(path, MPC_SimX=MPC_SimX, CiLX=CiLX, simU=simU, simREF=simREF, simSolverDebug=dbg,
                 sim_disturbance_derivatives=lg["sim_disturbance_derivatives"][b],
                 sim_disturbance_state_estimation=lg["sim_disturbance_state_estimation"][b],
                 a_lat=CiLX[:, 3] * CiLX[:, 5], dev_lat=dev_lat, dev_long=dev_long, dev_vel=dev_vel, dev_yaw=dev_yaw,
                 t=np.linspace(0.0, n * self.Ts, n))


def lon_lat_deviations(ego_yaw, ego_x, ego_y, ref_x, ref_y):
    """``LonLatDeviations`` (``Utils/MPC_sim_utils.py:102-112``): the offset to the reference point rotated into the
    vehicle frame."""
    c, s = np.cos(-ego_yaw), np.sin(-ego_yaw)
    return c * (ref_x - ego_x) - s * (ref_y - ego_y), s * (ref_x - ego_x) + c * (ref_y - ego_y)


def deviations(CiLX: np.ndarray, simREF: np.ndarray):
    """Per-step lateral and velocity deviation as ``Logger.logging_step`` records them (``Logging_Plotting.py:147-159``):
    state of step i against the reference point of step i."""
    _, lat = lon_lat_deviations(CiLX[..., 2], CiLX[..., 0], CiLX[..., 1], simREF[..., 0], simREF[..., 1])
    return lat, CiLX[..., 3] - simREF[..., 3]
