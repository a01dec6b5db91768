"""Host-side producers of the solver inputs: planner window, state-estimation filter and the
synthetic (x0, yref) batches used by the benchmarks.

These mirror the *callers* of the hot path in the reference (``Utils/MPC_sim_utils.py:126-194``,
``Utils/SimulationMode_main_class.py:10-19,44-66,152-156``); they are needed to regenerate the
golden inputs and to draw benchmark batches, and are numpy-only host logic.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, Optional, Tuple

import numpy as np

from .config import STATE_ESTIMATION_WINDOWS, SIM_MAIN_PARAMS, load_reftraj

TWO_PI = 2.0 * np.pi


def postprocess_yaw(yaw):
    """``fmod`` to (-2pi, 2pi) then shift negatives: ``Utils/MPC_sim_utils.py:126-135``."""
    y = np.fmod(yaw, TWO_PI)
    return np.where(y < 0, y + TWO_PI, y) if isinstance(y, np.ndarray) else (y + TWO_PI if y < 0 else y)


class PlannerEmulator:
    """Reference-window extraction of ``PlannerEmulator(ref_traj_set, pose, N+1, Tp, loop_circuit=True)``
    (``Utils/MPC_sim_utils.py:137-194``): nearest point, march forward until the accumulated
    time ``sum(|dxy| / ref_v)`` exceeds Tp, then resample every key to N+1 points uniformly in
    index space (yaw unwrapped first when the window straddles 0/2pi)."""

    def __init__(self, reftraj: Dict[str, np.ndarray]):
        self.px = np.asarray(reftraj["pos_x"], float)
        self.py = np.asarray(reftraj["pos_y"], float)
        self.v = np.asarray(reftraj["ref_v"], float)
        self.yaw = np.asarray(reftraj["ref_yaw"], float)
        n = self.n = len(self.px)
        nxt = (np.arange(n) + 1) % n
        # time to travel from point i to its successor, divided by the successor's speed (:170)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.seg_dt = np.hypot(self.px[nxt] - self.px, self.py[nxt] - self.py) / self.v[nxt]

    def closest(self, x: float, y: float) -> int:
        return int(np.argmin(np.hypot(self.px - x, self.py - y)))

    def window_indices(self, c: int, Tp: float) -> np.ndarray:
        idx = [c]
        T = 0.0
        cur = c
        while T <= Tp:
            T += self.seg_dt[cur]
            cur = cur + 1 if cur + 1 < self.n else 0
            idx.append(cur)
        return np.asarray(idx)

    def __call__(self, pose, n_pts: int, Tp: float) -> Tuple[int, Dict[str, np.ndarray]]:
        c = self.closest(pose[0], pose[1])
        idx = self.window_indices(c, Tp)
        out = {}
        L = len(idx)
        for key, arr in (("pos_x", self.px), ("pos_y", self.py), ("ref_v", self.v), ("ref_yaw", self.yaw)):
            vals = arr[idx]
            if L != n_pts:
                xq = np.linspace(0, L - 1, n_pts)
                xp = np.arange(L)
                res = np.interp(xq, xp, vals)
                if key == "ref_yaw" and (np.abs(np.diff(vals)) > np.deg2rad(250)).any():
                    res = np.mod(np.interp(xq, xp, np.unwrap(vals, period=TWO_PI)), TWO_PI)
                vals = res
            out[key] = vals
        return c, out

    def yref(self, pose, N: int, Tp: float) -> np.ndarray:
        """[N+1, 6] rows ``[x_ref, y_ref, yaw_ref, v_ref, 0, 0]`` as set by
        ``NMPC_class.py:169-180`` (the terminal row only uses its first four entries)."""
        _, w = self(pose, N + 1, Tp)
        y = np.zeros((N + 1, 6))
        y[:, 0] = w["pos_x"]; y[:, 1] = w["pos_y"]; y[:, 2] = w["ref_yaw"]; y[:, 3] = w["ref_v"]
        return y


class StateEstimation:
    """Per-state truncated moving average (``SimulationMode_main_class.py:10-19,82-86,152-156``)."""

    def __init__(self, windows=STATE_ESTIMATION_WINDOWS, buffer_size: int = 15):
        self.windows = np.asarray(windows, int)
        self.buf = [deque(maxlen=buffer_size) for _ in self.windows]

    def __call__(self, x):
        x = np.array(x, dtype=float)
        for i, w in enumerate(self.windows):
            self.buf[i].append(x[i])
            d = np.asarray(self.buf[i])
            x[i] = np.mean(d[-w:]) if len(d) >= w else np.mean(d)
        return x


def initial_state(reftraj: Dict[str, np.ndarray], idx: int) -> np.ndarray:
    """``X0_MPC`` of ``SimulationMode_main_class.py:44-66``."""
    return np.array([reftraj["pos_x"][idx], reftraj["pos_y"][idx], postprocess_yaw(float(reftraj["ref_yaw"][idx])),
                     reftraj["ref_v"][idx], 0.0, 0.0, 0.0, 0.0])


def synthetic_batch(track: str, B: int, N: int, Tp: float, seed: int = 0,
                    sigma: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic benchmark batch of SURVEY section 8(d): start indices ~ U{0..n-1}, nominal state on the
    reference, gaussian perturbation with the state-estimation stds of
    ``Config/EDGAR/sim_main_params.yaml:54-60``, planner window per instance.
    Returns x0 [B, 8] and yref [B, N+1, 6]."""
    ref = load_reftraj(track)
    planner = PlannerEmulator(ref)
    rng = np.random.default_rng(seed)
    if sigma is None:
        s = SIM_MAIN_PARAMS
        sigma = np.array([s["w_posx"], s["w_posy"], s["w_yaw"], s["w_vlong"], s["w_vlat"], s["w_yawrate"],
                          s["w_delta_f"], 0.0])
    idx = rng.integers(0, planner.n, size=B)
    noise = rng.standard_normal((B, 8)) * sigma
    x0 = np.zeros((B, 8))
    yref = np.zeros((B, N + 1, 6))
    for b in range(B):
        x = initial_state(ref, int(idx[b])) + noise[b]
        x0[b] = x
        yref[b] = planner.yref([x[0], x[1], x[2], np.hypot(x[3], x[4])], N, Tp)
    return x0, yref
