"""Teacher-forced replay of the reference's committed closed-loop logs (SURVEY section 8c recipe).

For step i of a log: x0 is what the reference fed to acados (moving-average filter of the logged
plant state + the MPC's own predicted acceleration), yref is the planner window at the logged pose,
the warm start is the SOLVER-UNDER-TEST's own previous solution (acados does not shift), and the
outputs compared are u0 = simU[i], x_1 = MPC_SimX[i+1] and the cost simSolverDebug[i, 0].
"""
import os

import numpy as np

from tum_control_b200.config import OcpConfig, load_reftraj
from tum_control_b200.planner import PlannerEmulator, StateEstimation

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_GOLD, TP_GOLD = 38, 3.04          # every committed artefact uses Tp=3.04, Ts_MPC=0.08 (SURVEY B-11)


def load_log(key):
    d = np.load(os.path.join(GOLDEN, "baseline_F.npz"))
    track, k = key.rsplit("_", 1)
    log = {n: d[f"{key}_{n}"] for n in ("CiLX", "MPC_SimX", "simU", "dbg", "simREF")}
    log["params"] = d["F_params"][int(k)]
    log["track"] = track
    return log


def replay_inputs(log, nsteps=None):
    """Yield (i, x0, yref) exactly as the reference produced them (teacher forcing)."""
    CiLX = log["CiLX"].copy()
    CiLX[:, 2] = np.unwrap(CiLX[:, 2])          # logs wrap yaw on save (Logging_Plotting.py:341-343)
    MX = log["MPC_SimX"]
    n = len(log["simU"]) if nsteps is None else min(nsteps, len(log["simU"]))
    planner = PlannerEmulator(load_reftraj(log["track"]))
    se = StateEstimation()
    x0 = MX[0].copy()
    x0[2] = CiLX[0, 2]
    for i in range(n):
        pose = [CiLX[i, 0], CiLX[i, 1], CiLX[i, 2], np.hypot(CiLX[i, 3], CiLX[i, 4])]
        yield i, x0, planner.yref(pose, N_GOLD, TP_GOLD)
        x0 = se(np.append(CiLX[i + 1], MX[i + 1, 7]))


def wrap_err_x1(x1, ref):
    x1 = np.array(x1, float)
    x1[2] = np.fmod(x1[2], 2 * np.pi)
    if x1[2] < 0:
        x1[2] += 2 * np.pi
    e = np.abs(x1 - ref)
    e[2] = min(e[2], abs(e[2] - 2 * np.pi))
    return e.max()


def golden_ocp(**kw):
    return OcpConfig(N=N_GOLD, Tp=TP_GOLD, **kw)
