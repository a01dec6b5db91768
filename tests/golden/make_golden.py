#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the reference tree.

Runs ONLY in the build container (needs /root/reference); the GPU box and the tests read the
committed .npz files.  Nothing here is imported by the product.

Fixtures
--------
baseline_F.npz
    Down-selected closed-loop logs of the reference's nominal NMPC
    (Learning_To_Adapt/SafeRL_WMPC/_baseline/F/{track}/{k}.npz, written by
    Utils/Logging_Plotting.py:357-372) together with the weight rows of
    Learning_To_Adapt/SafeRL_WMPC/_parameters/F.csv.  Per log: CiLX, MPC_SimX, simU and the
    (cost, qp_iter, status) columns of simSolverDebug, truncated to `steps`.
planner_golden.npz
    Outputs of the reference's own PlannerEmulator / moving_average_filter, obtained by importing
    Utils/MPC_sim_utils.py and Utils/SimulationMode_main_class.py with matplotlib/casadi stubbed
    (both modules import them at the top but the two functions do not use them).
snmpc_json_golden.npz
    Numbers decoded from acados_ocp_SNMPC.json (dims, lbx_0 sample matrix, folded constants).
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

LOGS = [  # (track, k, steps)  None = full lap
    ("monteblanco", 0, None),
    ("monteblanco", 17, 1500),
    ("lvms", 21, 1500),
    ("lvms", 3, 800),
]


def make_baseline():
    F = np.loadtxt(f"{REF}/Learning_To_Adapt/SafeRL_WMPC/_parameters/F.csv", delimiter=",")
    out = {"F_params": F}
    for track, k, steps in LOGS:
        d = np.load(f"{REF}/Learning_To_Adapt/SafeRL_WMPC/_baseline/F/{track}/{k}.npz")
        n = steps if steps is not None else d["simU"].shape[0]
        key = f"{track}_{k}"
        out[key + "_CiLX"] = d["CiLX"][: n + 1]
        out[key + "_MPC_SimX"] = d["MPC_SimX"][: n + 1]
        out[key + "_simU"] = d["simU"][:n]
        out[key + "_dbg"] = d["simSolverDebug"][:n][:, [0, 3, 4]]
        out[key + "_simREF"] = d["simREF"][:n]
    np.savez_compressed(os.path.join(OUT, "baseline_F.npz"), **out)
    print("baseline_F.npz", os.path.getsize(os.path.join(OUT, "baseline_F.npz")) // 1024, "KiB")


def _stub_modules():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "casadi",
                 "matplotlib.patches", "matplotlib.animation", "imageio", "pylab"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.setdefault("__path__", [])
            sys.modules[name] = m
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def make_planner():
    _stub_modules()
    sys.path.insert(0, REF)
    from Utils.MPC_sim_utils import PlannerEmulator  # the reference's own function
    from Utils.SimulationMode_main_class import moving_average_filter

    rng = np.random.default_rng(123)
    out = {}
    for track in ("monteblanco", "lvms", "modena"):
        with open(f"{REF}/Trajectories/reftraj_{track}_edgar.json") as f:
            ref = json.load(f)
        n = len(ref["pos_x"])
        # cover the wrap (last indices), the yaw 0/2pi crossing and random poses with offsets
        idxs = np.unique(np.concatenate([rng.integers(0, n, 40), np.arange(n - 6, n), np.arange(0, 4)]))
        poses, wins, cidx, cfg = [], [], [], []
        for (N, Tp) in ((38, 3.04), (30, 2.4), (80, 6.4)):
            for i in idxs:
                pose = [ref["pos_x"][i] + rng.normal(0, 0.5), ref["pos_y"][i] + rng.normal(0, 0.5),
                        ref["ref_yaw"][i], ref["ref_v"][i]]
                c, w = PlannerEmulator(ref, pose, N + 1, Tp, True)
                win = np.zeros((81, 4))
                win[: N + 1] = np.stack([np.asarray(w[k]) for k in ("pos_x", "pos_y", "ref_yaw", "ref_v")], axis=1)
                poses.append(pose); wins.append(win); cidx.append(c); cfg.append([N, Tp])
        out[track + "_pose"] = np.asarray(poses)
        out[track + "_win"] = np.asarray(wins)
        out[track + "_closest"] = np.asarray(cidx)
        out[track + "_cfg"] = np.asarray(cfg)
    # moving average filter
    data = rng.standard_normal(15)
    out["maf_data"] = data
    out["maf_out"] = np.array([[moving_average_filter(data[: n], w) for w in range(1, 6)] for n in range(1, 16)])
    np.savez_compressed(os.path.join(OUT, "planner_golden.npz"), **out)
    print("planner_golden.npz", os.path.getsize(os.path.join(OUT, "planner_golden.npz")) // 1024, "KiB")


def make_snmpc_json():
    with open(f"{REF}/acados_ocp_SNMPC.json") as f:
        J = json.load(f)
    dims = J["dims"]
    out = {
        "dims": np.array([dims[k] for k in ("N", "nx", "nu", "np", "nh", "nh_0", "ns", "ns_0", "ns_e", "ny", "ny_e")]),
        "lbx_0": np.asarray(J["constraints"]["lbx_0"], float),
        "ubx_0": np.asarray(J["constraints"]["ubx_0"], float),
        "W": np.asarray(J["cost"]["W"], float),
        "W_e": np.asarray(J["cost"]["W_e"], float),
        "zl": np.asarray(J["cost"]["zl"], float), "Zl": np.asarray(J["cost"]["Zl"], float),
        "lh": np.asarray(J["constraints"]["lh"], float), "uh": np.asarray(J["constraints"]["uh"], float),
        "lbx": np.asarray(J["constraints"]["lbx"], float), "ubx": np.asarray(J["constraints"]["ubx"], float),
        "lbu": np.asarray(J["constraints"]["lbu"], float), "ubu": np.asarray(J["constraints"]["ubu"], float),
        "p0": np.asarray(J["parameter_values"], float),
        "tf": np.array([J["solver_options"]["tf"]]),
    }
    np.savez_compressed(os.path.join(OUT, "snmpc_json_golden.npz"), **out)
    print("snmpc_json_golden.npz", {k: v.shape for k, v in out.items()})


def make_trajectories():
    """--trajectories: the reduced reference trajectories shipped as package data."""
    dst = os.path.join(os.path.dirname(os.path.dirname(OUT)), "tum_control_b200", "data")
    for t in ("monteblanco", "lvms", "modena"):
        with open(f"{REF}/Trajectories/reftraj_{t}_edgar.json") as f:
            d = json.load(f)
        with open(f"{dst}/reftraj_{t}_edgar.json", "w") as f:
            json.dump({k: d[k] for k in ("pos_x", "pos_y", "ref_v", "ref_yaw")}, f, separators=(",", ":"))


if __name__ == "__main__":
    if "--trajectories" in sys.argv:
        make_trajectories()
    make_baseline()
    make_planner()
    make_snmpc_json()
