"""Problem configuration for the batched NMPC hot path.

The reference reads four YAML files and one CSV under ``Config/EDGAR/`` at start-up
(``main.py:25-28``, ``NMPC_class.py:44-87``).  The values the hot path needs are re-typed here with
the reference location of each, so that nothing at run time depends on ``/root/reference``; the
same dictionaries can be overridden from the reference's own YAML files with
:func:`load_reference_config` when they are available.

All quantities are float64 (the reference is numpy + acados ``double``).
"""
from __future__ import annotations

import dataclasses
import json
import os
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np

NX, NU, NY, NYE, NC = 8, 2, 6, 4, 3
MAX_GGV = 16
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# Config/EDGAR/veh_params_pred.yaml:3-25  (prediction model)
VEH_PARAMS_PRED: Dict[str, float] = dict(
    lf=1.484, lr=1.644, m=2520.0, Iz=13600.0, ro=1.225, S=2.9, Cd=0.35,
    veh_length=4.973, veh_width=1.941,
    jerk_min=-8.0, jerk_max=6.0, lat_acc_min=-5.886, lat_acc_max=5.886,
    acc_min=-3.5, acc_max=2.5, delta_f_min=-0.610865, delta_f_max=0.610865,
    delta_f_dot_min=-0.322, delta_f_dot_max=0.322,
)
# Config/EDGAR/veh_params_sim.yaml (plant) -- identical numbers in the committed config
VEH_PARAMS_SIM: Dict[str, float] = dict(VEH_PARAMS_PRED)
# Config/EDGAR/pacejka_params.yaml:3-13
PACEJKA_PARAMS: Dict[str, float] = dict(
    Bf=10.0, Cf=1.3, Df=15591.427, Ef=0.97, Br=10.0, Cr=1.6, Dr=24629.523, Er=0.97, mu=1.0489,
)
# Config/EDGAR/ggv.csv  (vel_max_mps, ax_max_mps2, ax_min_mps2, ay_max_mps2)
GGV_TABLE = np.array([
    [0.0, 3.0, -4.5, 5.886], [4.0, 3.0, -4.5, 5.886], [8.0, 3.0, -4.5, 5.886],
    [11.11, 3.0, -4.5, 5.886], [12.0, 2.5, -3.5, 5.886], [20.0, 2.5, -3.5, 5.886],
    [24.0, 2.5, -3.5, 5.886], [28.0, 2.5, -3.5, 5.886], [32.0, 2.5, -3.5, 5.886],
    [37.5, 2.5, -3.5, 5.886],
])
# Config/EDGAR/MPC_params.yaml:10-47
MPC_PARAMS: Dict[str, object] = dict(
    s_lon=1.0, s_lat=1.0, s_yaw=1.0, s_vel=1.0, s_jerk=1.0, s_steering_rate=1.0,
    q_lon=2.8, q_lat=2.8, q_yaw=0.4, q_vel=0.2, r_jerk=38.1, r_steering_rate=101.4,
    L1_pen=106.7, L2_pen=9.9, combined_acc_limits=2,
    stds=[0.0, 0.0, 0.0, 0.8, 0.35, 0.035, 0.0, 0.0],
    uncertainty_propagation_horizon=5, n_samples=10, gamma=0.8, expansion_degree=2,
)
# Config/EDGAR/sim_main_params.yaml:39-60
SIM_MAIN_PARAMS: Dict[str, object] = dict(
    Ts=0.02, T=100.0, Tp=3.04, Ts_MPC=0.08,
    w_posx=0.15, w_posy=0.15, w_yaw=0.01, w_vlong=0.8, w_vlat=0.35, w_yawrate=0.05, w_delta_f=0.005,
)
# Utils/SimulationMode_main_class.py:86  moving-average windows of StateEstimation
STATE_ESTIMATION_WINDOWS = np.array([1, 1, 4, 2, 2, 3, 4, 2])


@dataclass
class VehicleModel:
    """Constants of the dynamic single-track Pacejka model
    (``Prediction_Models/pred_model_dynamic_stm_pacejka.py:34-67``)."""
    lf: float = 1.484
    lr: float = 1.644
    m: float = 2520.0
    Iz: float = 13600.0
    ro: float = 1.225
    S: float = 2.9
    Cd: float = 0.35
    Bf: float = 10.0
    Cf: float = 1.3
    Df: float = 15591.427
    Ef: float = 0.97
    Br: float = 10.0
    Cr: float = 1.6
    Dr: float = 24629.523
    Er: float = 0.97
    g: float = 9.81        # pred_model_dynamic_stm_pacejka.py:38
    fr0: float = 0.009     # :44
    fr1: float = 0.002     # :45
    fr4: float = 0.0003    # :46
    ax_brake: float = 3.5  # -acc_min, NMPC_STM_acados_settings.py:73
    delta_f_max: float = 0.610865
    delta_f_dot_max: float = 0.322
    ggv_v: np.ndarray = field(default_factory=lambda: GGV_TABLE[:, 0].copy())
    ggv_ax: np.ndarray = field(default_factory=lambda: GGV_TABLE[:, 1].copy())
    ggv_ay: np.ndarray = field(default_factory=lambda: GGV_TABLE[:, 3].copy())

    @classmethod
    def from_dicts(cls, veh: Dict[str, float], tire: Dict[str, float], ggv: np.ndarray) -> "VehicleModel":
        return cls(
            lf=veh["lf"], lr=veh["lr"], m=veh["m"], Iz=veh["Iz"], ro=veh["ro"], S=veh["S"], Cd=veh["Cd"],
            Bf=tire["Bf"], Cf=tire["Cf"], Df=tire["Df"], Ef=tire["Ef"],
            Br=tire["Br"], Cr=tire["Cr"], Dr=tire["Dr"], Er=tire["Er"],
            ax_brake=-veh["acc_min"], delta_f_max=veh["delta_f_max"], delta_f_dot_max=veh["delta_f_dot_max"],
            ggv_v=np.asarray(ggv[:, 0], float), ggv_ax=np.asarray(ggv[:, 1], float),
            ggv_ay=np.asarray(ggv[:, 3], float),
        )


@dataclass
class OcpConfig:
    """Everything ``acados_settings(...)`` bakes into the solver
    (``NMPC_STM_acados_settings.py:16-245``) plus the QP-solve knobs of this implementation."""
    N: int = 30
    Tp: float = 2.4
    variant: str = "nominal"            # "nominal" | "r2" | "snmpc"
    vehicle: VehicleModel = field(default_factory=VehicleModel)
    rk_steps: int = 3                   # sim_method_num_steps, :240 (4 stages fixed, :239)
    # W = 0.01 * blkdiag(Q, R), W_e = 0.01 * Qe   (:55-56, NMPC_class.py:76-78)
    W: np.ndarray = field(default_factory=lambda: 0.01 * np.array([2.8, 2.8, 0.4, 0.2, 38.1, 101.4]))
    W_e: np.ndarray = field(default_factory=lambda: 0.01 * np.array([2.8, 2.8, 0.4, 0.2]))
    L1_pen: float = 106.7
    L2_pen: float = 9.9
    ipm_max_iter: int = 50              # qp_solver_iter_max, :232
    # KKT tolerance of the QP solve (complementarity, side residuals, stationarity).  1e-9 is the bound the parity contract
    # states and 10x tighter than the reference's own HPIPM default (1e-8: qp_solver_tol_* are unset in
    # NMPC_STM_acados_settings.py:226-241).  Round 1 ran 1e-11: 1.1 more interior-point iterations per solve and, replayed
    # against the reference's committed logs, the same agreement to the last digit shown (median |du0| 1.6e-8, max 4.6e-6 on
    # monteblanco/0; the differences are the reference's tolerance, not ours) -- DESIGN.md section 2.
    ipm_tol: float = 1e-9
    ipm_reg: float = 1e-9               # dual proximal regularisation of the IPM Newton system
    ipm_eta_max: float = 0.99           # cap of the fraction-to-the-boundary parameter max(0.995, 1 - gap)
    ipm_s0: float = 0.1                 # initial slack of the interior start: s = ipm_s0 + violation (1.0 costs 1.4 more iterations)
    ipm_corr_mode: int = 0              # Mehrotra second-order term scaled by 1 (0), a_aff (1), a_aff^2 (2)
    qp_no_shortcut: bool = False
    # SNMPC / R2 (MPC_params.yaml:33-47)
    stds: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 0.0, 0.8, 0.35, 0.035, 0.0, 0.0]))
    uph: int = 5
    n_samples: int = 10
    gamma: float = 0.8
    expansion_degree: int = 2
    pce_A: Optional[np.ndarray] = None       # override of the PCE matrix / sample offsets (see pce_matrices)
    pce_w_off: Optional[np.ndarray] = None
    r2_sigma0: Optional[np.ndarray] = None   # override of the R2NMPC ellipsoid data (see r2_matrices): diag(Sigma_0) [8]
    r2_wd: Optional[np.ndarray] = None       # ... and diag(B W_d B') on states 2..5 [4]

    @property
    def dt(self) -> float:
        return self.Tp / self.N

    @property
    def cost_vel_abs(self) -> bool:
        # SNMPC tracks sqrt(vl^2+vt^2) (SNMPC_acados_settings.py:62-65); nominal/R2 track v_lon
        return self.variant == "snmpc"

    def pce_matrices(self):
        """SNMPC polynomial-chaos data (``SNMPC_class.py:87-99``): ``A = (Phi'Phi)^-1 Phi'`` [L, n_samples] for the
        Hammersley collocation samples, and the x0 sample offsets ``std (.) w_i`` [n_samples, 8]
        (``stochastic_mpc_utils.py:56-91``).  ``pce_A`` / ``pce_w_off`` override the computed values (tests)."""
        if self.pce_A is not None:
            return np.asarray(self.pce_A, float), np.asarray(self.pce_w_off, float)
        from . import pce
        w, A, _ = pce.samples_and_A(int(self.n_samples), self.stds, int(self.expansion_degree))
        return A, pce.sample_offsets(w, self.stds)

    def r2_matrices(self):
        """R2NMPC ellipsoid data (``Reduced_Robustified_NMPC_class.py:104-148``): diagonal of
        ``Sigma_0 = (0.5 * diag(1e-5, 1e-5, w_yaw, w_vlong, w_vlat, w_yawrate, 1e-5, 1e-5))**2`` and of
        ``B W_d B'`` on states 2..5 with ``W_d = Ts_MPC * diag(stds[2:6])**2``."""
        s = np.asarray(self.stds, float)
        sig0 = (0.5 * np.array([1e-5, 1e-5, s[2], s[3], s[4], s[5], 1e-5, 1e-5])) ** 2
        wd = self.dt * s[2:6] ** 2
        if self.r2_sigma0 is not None:
            sig0 = np.asarray(self.r2_sigma0, float).reshape(8)
        if self.r2_wd is not None:
            wd = np.asarray(self.r2_wd, float).reshape(4)
        return sig0, wd

    def stage_bounds(self):
        """Default per-stage bounds in slack order [bu(omega_f), bx(delta_f), h]:
        ``NMPC_STM_acados_settings.py:107-139``.  Returns (lb, ub) with shape [N+1, 3]."""
        v = self.vehicle
        lb = np.tile(np.array([-v.delta_f_dot_max, -v.delta_f_max, 0.0]), (self.N + 1, 1))
        ub = np.tile(np.array([v.delta_f_dot_max, v.delta_f_max, 1.0]), (self.N + 1, 1))
        return lb, ub

    def stage_weights(self):
        """[N+1, 6] diagonal NLS weights; the terminal row carries W_e in its first 4 entries."""
        W = np.tile(np.asarray(self.W, float), (self.N + 1, 1))
        W[self.N, :4] = self.W_e
        W[self.N, 4:] = 0.0
        return W

    def stage_penalties(self):
        """(zl, zu, Zl, Zu), each [N+1, 3] (``NMPC_STM_acados_settings.py:164-224``)."""
        z = np.full((self.N + 1, NC), float(self.L1_pen))
        Z = np.full((self.N + 1, NC), float(self.L2_pen))
        return z.copy(), z.copy(), Z.copy(), Z.copy()


def horizon_from_sim_params(Tp: float, Ts_MPC: float) -> int:
    """``N = int(Tp / Ts_MPC)`` (``NMPC_class.py:49``)."""
    return int(Tp / Ts_MPC)


def load_reference_config(config_path: str):
    """Parse the reference's own ``Config/EDGAR`` files (same parsing as ``main.py:25-28`` and
    ``NMPC_class.py:44-45,322-335``).  Returns (veh_pred, veh_sim, tire, ggv, mpc_params, sim_params)."""
    import yaml

    def y(name):
        with open(os.path.join(config_path, name), "r") as f:
            return yaml.load(f, Loader=yaml.FullLoader)

    sim = y("sim_main_params.yaml")
    mpc = y("MPC_params.yaml")
    base = os.path.dirname(os.path.normpath(config_path))
    veh_pred = y(os.path.basename(sim["veh_params_file_MPC"]))
    veh_sim = y(os.path.basename(sim["veh_params_file_simulator"]))
    tire_raw = y(os.path.basename(sim["tire_params_file_MPC"]))
    tire = dict(tire_raw["tire_params"]["front"]); tire.update(tire_raw["tire_params"]["rear"]); tire["mu"] = tire_raw["mu"]
    ggv = np.genfromtxt(os.path.join(base, mpc["lookuptable_gg_limits"]), delimiter=",", skip_header=1)
    return veh_pred, veh_sim, tire, ggv, mpc, sim


def load_reftraj(name_or_path: str) -> Dict[str, np.ndarray]:
    """Load a reference trajectory (``SimulationMode_main_class.py:37-40``).  ``name_or_path`` is a
    track name ("monteblanco" | "lvms" | "modena") resolved in the package data, or a JSON path."""
    path = name_or_path
    if not os.path.exists(path):
        path = os.path.join(DATA_DIR, f"reftraj_{name_or_path}_edgar.json")
    with open(path, "r") as f:
        d = json.load(f)
    return {k: np.asarray(v, dtype=np.float64) for k, v in d.items() if k in ("pos_x", "pos_y", "ref_v", "ref_yaw")}
