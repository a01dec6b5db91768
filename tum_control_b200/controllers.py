"""Host-side mirrors of the reference's controller classes (the L3 interface of SURVEY section 1), batched.

Reference classes and the calls they make on ``self.acados_solver``:
  * ``Model_Predictive_Controller/Nominal_NMPC/NMPC_class.py:35-317``      ``Nonlinear_Model_Predictive_Controller``
  * ``Model_Predictive_Controller/Reduced_Robustified_NMPC/Reduced_Robustified_NMPC_class.py:37-440``
Same public methods (``solve``, ``set_initial_state``, ``reset``, ``reintialize_solver``,
``update_cost_function_weights``), same return convention ``(u0, pred_X, stats[5])``; every array gains a
leading batch axis (squeezed away for ``batch == 1``).  What the reference does in a Python loop of
``2N + 4`` ctypes calls per solve is three bulk calls here.

  * ``Model_Predictive_Controller/Stochastic_NMPC/SNMPC_class.py:38-330``       ``Stochastic_Nonlinear_Model_Predictive_Controller``

The WMPC branch (``NMPC_class.py:120-161,209-239``, needs stable_baselines3) is out of scope (SURVEY section 2.1 #2;
DESIGN.md).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .config import MPC_PARAMS, SIM_MAIN_PARAMS, OcpConfig, VehicleModel, horizon_from_sim_params
from .solver import BatchedOcpSolver


def ocp_from_reference_params(mpc_params: Optional[Dict] = None, sim_main_params: Optional[Dict] = None,
                              vehicle: Optional[VehicleModel] = None, variant: str = "nominal", **overrides) -> OcpConfig:
    """Build the OCP exactly as ``NMPC_class.__init__`` does from the YAML dictionaries
    (``NMPC_class.py:44-107``): ``N = int(Tp / Ts_MPC)``, ``Q = diag(q_lon/s_lon^2, ...)``,
    ``W = 0.01 * blkdiag(Q, R)``, ``W_e = 0.01 * Q`` (``NMPC_STM_acados_settings.py:55-56``)."""
    p = dict(MPC_PARAMS); p.update(mpc_params or {})
    s = dict(SIM_MAIN_PARAMS); s.update(sim_main_params or {})
    N = horizon_from_sim_params(s["Tp"], s["Ts_MPC"])
    Q = np.array([p["q_lon"] / p["s_lon"] ** 2, p["q_lat"] / p["s_lat"] ** 2, p["q_yaw"] / p["s_yaw"] ** 2,
                  p["q_vel"] / p["s_vel"] ** 2])
    R = np.array([p["r_jerk"] / p["s_jerk"] ** 2, p["r_steering_rate"] / p["s_steering_rate"] ** 2])
    if int(p.get("combined_acc_limits", 2)) != 2:
        raise NotImplementedError("only the circle-shaped combined acceleration limit (combined_acc_limits: 2) is "
                                  "consistent in the reference (SURVEY B-2) and built here")
    kw = dict(N=N, Tp=float(s["Tp"]), variant=variant, W=0.01 * np.concatenate([Q, R]), W_e=0.01 * Q,
              L1_pen=float(p["L1_pen"]), L2_pen=float(p["L2_pen"]), stds=np.asarray(p["stds"], float),
              uph=int(p["uncertainty_propagation_horizon"]), n_samples=int(p["n_samples"]), gamma=float(p["gamma"]))
    if vehicle is not None:
        kw["vehicle"] = vehicle
    kw.update(overrides)
    return OcpConfig(**kw)


class Nonlinear_Model_Predictive_Controller:
    """Batched drop-in for ``NMPC_class.Nonlinear_Model_Predictive_Controller``."""

    variant = "nominal"

    def __init__(self, X0_MPC, mpc_params: Optional[Dict] = None, sim_main_params: Optional[Dict] = None,
                 vehicle: Optional[VehicleModel] = None, device: int = 0, ocp: Optional[OcpConfig] = None):
        X0 = np.atleast_2d(np.asarray(X0_MPC, float))
        self.batch = X0.shape[0]
        self.ocp = ocp if ocp is not None else ocp_from_reference_params(mpc_params, sim_main_params, vehicle,
                                                                          variant=self.variant)
        self.N, self.Tp, self.nx = self.ocp.N, self.ocp.Tp, 8
        self.device = device
        self.acados_solver = BatchedOcpSolver(self.ocp, batch=self.batch, device=device, squeeze=False)
        self.x0 = X0.copy()
        # NMPC_class.py:111-112 + acados' initial guess x_k = x0
        self.acados_solver.set_x0(self.x0)
        self.acados_solver.init_iterate_from_x0()
        self.stats = np.zeros((self.batch, 5))
        self.pred_X = np.zeros((self.batch, self.N, self.nx))

    # ---- helpers
    def _sq(self, a):
        return a[0] if self.batch == 1 else a

    def _yref(self, current_ref_traj: Dict) -> np.ndarray:
        """``[x_ref, y_ref, yaw_ref, v_ref, 0, 0]`` per stage (``NMPC_class.py:169-180``)."""
        y = np.zeros((self.batch, self.N + 1, 6))
        for i, key in enumerate(("pos_x", "pos_y", "ref_yaw", "ref_v")):
            v = np.asarray(current_ref_traj[key], float)
            y[:, :, i] = np.broadcast_to(v.reshape((-1, self.N + 1)) if v.ndim > 1 else v, (self.batch, self.N + 1))
        return y

    # ---- reference API
    def solve(self, current_ref_traj: Dict) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        S = self.acados_solver
        S.set_yref_all(self._yref(current_ref_traj))
        status = S.solve()
        X, U = S.get_iterate()
        u0 = U[:, 0, :].copy()
        ok = status == 0                               # pred_X only refreshed on success (NMPC_class.py:195-201)
        self.pred_X[ok] = X[ok, : self.N, :]
        self.stats = S.get_solver_debug()
        return self._sq(u0), self._sq(self.pred_X), self._sq(self.stats)

    def set_initial_state(self, x0):
        self.x0 = np.broadcast_to(np.asarray(x0, float), (self.batch, 8)).copy()
        self.acados_solver.set_x0(self.x0)

    def reset(self, x0):
        """``acados_solver.reset()`` then ``set(i, 'x', x0)`` for all stages (``NMPC_class.py:250-254``)."""
        self.acados_solver.reset()
        self.set_initial_state(x0)
        self.acados_solver.init_iterate_from_x0()

    def reintialize_solver(self, X0_MPC, **_ignored):
        """The reference rebuilds the AcadosOcpSolver after a failure (``main.py:59-61``); here the handle is
        kept and its iterate re-initialised, which is what a fresh solver amounts to."""
        self.reset(X0_MPC)

    def update_cost_function_weights(self, params):
        """``[q_xy, q_yaw, q_vel, r_jerk, r_steer, L1, L2]`` (``NMPC_class.py:269-317``): ``[7]`` for all
        instances or ``[B, 7]`` for one weight set per instance (the BO / RL callers)."""
        if hasattr(params, "numpy"):
            params = params.numpy()
        self.acados_solver.update_cost_function_weights(np.asarray(params, float))


class Reduced_Robustified_NMPC(Nonlinear_Model_Predictive_Controller):
    """Batched drop-in for ``Reduced_Robustified_NMPC_class``: the nominal OCP whose ``lbx/ubx/uh`` of stages
    1..N-1 are tightened after every successful solve by the ellipsoid back-offs
    (``Reduced_Robustified_NMPC_class.py:287-365``).  The reference does this in a host loop of CasADi calls
    and adds its wall time to ``stats[1]``; here it is one fused kernel behind the QP solve."""

    variant = "r2"


class Stochastic_Nonlinear_Model_Predictive_Controller(Nonlinear_Model_Predictive_Controller):
    """Batched drop-in for ``SNMPC_class.Stochastic_Nonlinear_Model_Predictive_Controller``: the OCP over the
    augmented state of ``1 + n_samples`` vehicle copies with the PCE chance constraint
    (``SNMPC_acados_settings.py:19-330``).  The reference re-sends ``p = [A, gamma, stop_flag]`` to every stage on
    every solve (``SNMPC_class.py:185-196``); the values never change, so they are sent once here.  ``pred_X`` takes
    the nominal block ``get(j, "x")[0:8]`` (``:214``)."""

    variant = "snmpc"

    def __init__(self, X0_MPC, mpc_params: Optional[Dict] = None, sim_main_params: Optional[Dict] = None,
                 vehicle: Optional[VehicleModel] = None, device: int = 0, ocp: Optional[OcpConfig] = None):
        if ocp is None:
            ocp = ocp_from_reference_params(mpc_params, sim_main_params, vehicle, variant="snmpc", rk_steps=1)
        super().__init__(X0_MPC, device=device, ocp=ocp)
        S = self.acados_solver
        self.A, self.w_off = ocp.pce_matrices()
        self.n_samples = int(ocp.n_samples)
        self.stop_flags = np.zeros(self.N + 1)
        self.stop_flags[int(ocp.uph):] = 1                        # SNMPC_class.py:96-98
        self.risk_parameter = np.array([ocp.gamma])
        for i in range(self.N + 1):                               # :122-125
            S.set(i, "p", np.concatenate((self.A.flatten(), self.risk_parameter, self.stop_flags[i].reshape(1))))

    def x0_samples(self, x0) -> np.ndarray:
        """``compute_x0dist`` (``stochastic_mpc_utils.py:78-91``): ``[B, 1+ns, 8]``."""
        x0 = np.broadcast_to(np.asarray(x0, float), (self.batch, 8))
        return np.concatenate([x0[:, None, :], x0[:, None, :] + self.w_off[None, :, :]], axis=1)

    def set_initial_state(self, x0):
        self.x0 = np.broadcast_to(np.asarray(x0, float), (self.batch, 8)).copy()
        self.acados_solver.set_x0_samples(self.x0_samples(self.x0))      # :259-264
