"""ctypes loader of ``libb200nmpc.so`` (the C-ABI of ``include/b200nmpc.h``).

There is deliberately NO fallback: if the shared library is missing or fails to load, importing
callers get a ``RuntimeError`` that says how to build it.  The library itself refuses to create a
handle without an sm_100 device (``B200NMPC_E_NODEVICE``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200NMPC_LIB", os.path.join(HERE, "libb200nmpc.so"))   # override: A/B testing of kernel builds
CSRC = os.path.join(HERE, "csrc")

MAX_GGV = 16
MAX_NS = 16
PCE_L = 10
MAX_UPH = 8

# field / stat ids of include/b200nmpc.h
F_X, F_U, F_YREF, F_P = 0, 1, 2, 3
F_LBX, F_UBX, F_LBU, F_UBU, F_LH, F_UH = 10, 11, 12, 13, 14, 15
F_W, F_ZL_LIN, F_ZU_LIN, F_ZL_QUAD, F_ZU_QUAD = 20, 21, 22, 23, 24
F_SL, F_SU = 30, 31
F_QP_A, F_QP_B, F_QP_b = 40, 41, 42
S_TIME_TOT, S_SQP_ITER, S_QP_ITER, S_KKT, S_TIME_BATCH, S_TIME_LINEARIZE, S_TIME_QP = 0, 1, 2, 3, 4, 5, 6
VARIANTS = {"nominal": 0, "r2": 1, "snmpc": 2}
E_NODEVICE = -6


class B200Config(C.Structure):
    """Mirror of ``struct b200nmpc_config``."""
    _fields_ = (
        [(n, C.c_double) for n in ("lf", "lr", "m", "Iz", "ro", "S", "Cd",
                                   "Bf", "Cf", "Df", "Ef", "Br", "Cr", "Dr", "Er",
                                   "g", "fr0", "fr1", "fr4", "ax_brake")]
        + [("ggv_v", C.c_double * MAX_GGV), ("ggv_ax", C.c_double * MAX_GGV), ("ggv_ay", C.c_double * MAX_GGV),
           ("dt", C.c_double), ("W", C.c_double * 6), ("W_e", C.c_double * 4),
           ("zl", C.c_double), ("zu", C.c_double), ("Zl", C.c_double), ("Zu", C.c_double),
           ("lbu", C.c_double), ("ubu", C.c_double), ("lbx", C.c_double), ("ubx", C.c_double),
           ("lh", C.c_double), ("uh", C.c_double),
           ("ipm_tol", C.c_double), ("ipm_reg", C.c_double), ("ipm_eta_max", C.c_double), ("ipm_s0", C.c_double),
           ("r2_sigma0_diag", C.c_double * 8), ("r2_wd", C.c_double * 4), ("sn_x0_std", C.c_double * 8),
           ("sn_gamma", C.c_double), ("sn_A", C.c_double * (PCE_L * MAX_NS)), ("sn_w_off", C.c_double * (MAX_NS * 8)),
           ("n_ggv", C.c_int), ("N", C.c_int), ("rk_steps", C.c_int), ("variant", C.c_int),
           ("ipm_max_iter", C.c_int), ("qp_no_shortcut", C.c_int), ("uph", C.c_int), ("n_samples", C.c_int),
           ("sn_L", C.c_int), ("ipm_corr_mode", C.c_int)]
    )


class B200SimConfig(C.Structure):
    """Mirror of ``struct b200nmpc_sim_config``."""
    _fields_ = [("Ts", C.c_double), ("Tp", C.c_double), ("plant_substeps", C.c_int), ("log_capacity", C.c_int),
                ("n_ref", C.c_int), ("windows", C.c_int * 8)]


EXPORTS = [
    "b200nmpc_create", "b200nmpc_destroy", "b200nmpc_last_error", "b200nmpc_version", "b200nmpc_sizeof_config",
    "b200nmpc_batch", "b200nmpc_horizon", "b200nmpc_stream", "b200nmpc_set", "b200nmpc_constraints_set",
    "b200nmpc_cost_set", "b200nmpc_solve", "b200nmpc_get_status", "b200nmpc_get", "b200nmpc_get_cost",
    "b200nmpc_get_stats", "b200nmpc_reset", "b200nmpc_set_x0", "b200nmpc_set_yref_all", "b200nmpc_set_iterate",
    "b200nmpc_init_iterate_from_x0", "b200nmpc_get_iterate", "b200nmpc_get_solver_debug",
    "b200nmpc_update_cost_function_weights", "b200nmpc_set_x0_dev", "b200nmpc_set_yref_all_dev",
    "b200nmpc_get_iterate_dev", "b200nmpc_get_solver_debug_dev", "b200nmpc_synchronize", "b200nmpc_get_results_async",
    "b200nmpc_alloc_host", "b200nmpc_free_host", "b200nmpc_timer_start", "b200nmpc_timer_stop", "b200nmpc_launch_count", "b200nmpc_r2_tighten",
    "b200nmpc_set_x0_samples", "b200nmpc_set_samples", "b200nmpc_get_samples", "b200nmpc_snmpc_uph", "b200nmpc_snmpc_get_head",
    "b200nmpc_get_yref_all", "b200nmpc_sim_create", "b200nmpc_sim_destroy", "b200nmpc_sim_reset", "b200nmpc_sim_run", "b200nmpc_sim_steps_done", "b200nmpc_sim_rebase_logs", "b200nmpc_options_set",
    "b200nmpc_sim_get_logs", "b200nmpc_sim_get_logs_range", "b200nmpc_sim_get_ref_idx", "b200nmpc_sim_set_stop", "b200nmpc_sim_get_stop", "b200nmpc_sim_reset_instances", "b200nmpc_sim_set_disturbances", "b200nmpc_sim_plan",
    "b200nmpc_sim_config_disturbances", "b200nmpc_sim_set_playback", "b200nmpc_sim_get_disturbance_logs",
]


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", CSRC] + ([] if verbose else ["-s"]), check=True)
    return LIB_PATH


_lib = None


def lib():
    """Load the shared library (once).  Raises RuntimeError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {CSRC}`.  tum_control_b200 has no CPU or PyTorch fallback.")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:
        raise RuntimeError(f"cannot load {LIB_PATH}: {e}.  tum_control_b200 has no CPU or PyTorch fallback.") from e
    vp, ip, dp = C.c_void_p, C.c_int, C.POINTER(C.c_double)
    L.b200nmpc_create.argtypes = [C.POINTER(B200Config), ip, ip, C.POINTER(vp)]
    L.b200nmpc_destroy.argtypes = [vp]
    L.b200nmpc_last_error.restype = C.c_char_p
    L.b200nmpc_batch.argtypes = [vp]
    L.b200nmpc_horizon.argtypes = [vp]
    L.b200nmpc_stream.argtypes = [vp]
    L.b200nmpc_stream.restype = vp
    for name in ("b200nmpc_set", "b200nmpc_constraints_set", "b200nmpc_cost_set"):
        getattr(L, name).argtypes = [vp, ip, ip, vp, ip]
    L.b200nmpc_solve.argtypes = [vp, vp]
    L.b200nmpc_get_status.argtypes = [vp, vp]
    L.b200nmpc_get.argtypes = [vp, ip, ip, vp]
    L.b200nmpc_get_cost.argtypes = [vp, vp]
    L.b200nmpc_get_stats.argtypes = [vp, ip, vp]
    L.b200nmpc_reset.argtypes = [vp]
    L.b200nmpc_set_x0.argtypes = [vp, vp]
    L.b200nmpc_set_yref_all.argtypes = [vp, vp]
    L.b200nmpc_get_yref_all.argtypes = [vp, vp]
    L.b200nmpc_set_iterate.argtypes = [vp, vp, vp]
    L.b200nmpc_init_iterate_from_x0.argtypes = [vp]
    L.b200nmpc_get_iterate.argtypes = [vp, vp, vp]
    L.b200nmpc_get_solver_debug.argtypes = [vp, vp]
    L.b200nmpc_update_cost_function_weights.argtypes = [vp, vp, ip]
    L.b200nmpc_set_x0_dev.argtypes = [vp, vp]
    L.b200nmpc_set_yref_all_dev.argtypes = [vp, vp]
    L.b200nmpc_get_iterate_dev.argtypes = [vp, vp, vp]
    L.b200nmpc_get_results_async.argtypes = [vp, vp, vp, vp]
    L.b200nmpc_get_solver_debug_dev.argtypes = [vp, vp]
    L.b200nmpc_synchronize.argtypes = [vp]
    L.b200nmpc_alloc_host.argtypes = [C.c_size_t]
    L.b200nmpc_alloc_host.restype = vp
    L.b200nmpc_free_host.argtypes = [vp]
    L.b200nmpc_timer_start.argtypes = [vp]
    L.b200nmpc_timer_stop.argtypes = [vp, vp]
    L.b200nmpc_launch_count.argtypes = [vp]
    L.b200nmpc_launch_count.restype = C.c_longlong
    L.b200nmpc_r2_tighten.argtypes = [vp]
    L.b200nmpc_set_x0_samples.argtypes = [vp, vp]
    L.b200nmpc_set_samples.argtypes = [vp, vp]
    L.b200nmpc_get_samples.argtypes = [vp, vp]
    L.b200nmpc_snmpc_uph.argtypes = [vp]
    L.b200nmpc_snmpc_get_head.argtypes = [vp, vp, C.POINTER(ip)]
    L.b200nmpc_sim_create.argtypes = [vp, C.POINTER(B200SimConfig), vp, vp, vp, vp, C.POINTER(vp)]
    L.b200nmpc_sim_destroy.argtypes = [vp]
    L.b200nmpc_sim_reset.argtypes = [vp, vp]
    L.b200nmpc_sim_run.argtypes = [vp, ip, vp]
    L.b200nmpc_sim_steps_done.argtypes = [vp]
    L.b200nmpc_sim_rebase_logs.argtypes = [vp]
    L.b200nmpc_options_set.argtypes = [vp, C.c_char_p, C.c_double]
    L.b200nmpc_sim_get_logs.argtypes = [vp, vp, vp, vp, vp, vp]
    L.b200nmpc_sim_get_ref_idx.argtypes = [vp, vp]
    L.b200nmpc_sim_get_logs_range.argtypes = [vp, ip, ip, vp, vp, vp, vp, vp, vp]
    L.b200nmpc_sim_set_stop.argtypes = [vp, ip, vp, C.c_double, C.c_double]
    L.b200nmpc_sim_get_stop.argtypes = [vp, vp, vp]
    L.b200nmpc_sim_reset_instances.argtypes = [vp, vp, vp]
    L.b200nmpc_sim_set_disturbances.argtypes = [vp, vp, vp]
    L.b200nmpc_sim_plan.argtypes = [vp, vp, vp]
    L.b200nmpc_sim_config_disturbances.argtypes = [vp, ip, vp, ip, vp, C.c_ulonglong]
    L.b200nmpc_sim_set_playback.argtypes = [vp, vp, vp, ip]
    L.b200nmpc_sim_get_disturbance_logs.argtypes = [vp, vp, vp, vp]
    if L.b200nmpc_sizeof_config() != C.sizeof(B200Config):
        raise RuntimeError("b200nmpc_config layout mismatch between include/b200nmpc.h and tum_control_b200/_lib.py")
    _lib = L
    return L


def last_error() -> str:
    return lib().b200nmpc_last_error().decode()


def make_config(ocp) -> B200Config:
    """``tum_control_b200.config.OcpConfig`` -> ``b200nmpc_config``."""
    v = ocp.vehicle
    c = B200Config()
    for n in ("lf", "lr", "m", "Iz", "ro", "S", "Cd", "Bf", "Cf", "Df", "Ef", "Br", "Cr", "Dr", "Er",
              "g", "fr0", "fr1", "fr4", "ax_brake"):
        setattr(c, n, float(getattr(v, n)))
    n = len(v.ggv_v)
    if n > MAX_GGV:
        raise ValueError(f"ggv table has {n} rows, max {MAX_GGV}")
    for i in range(n):
        c.ggv_v[i], c.ggv_ax[i], c.ggv_ay[i] = float(v.ggv_v[i]), float(v.ggv_ax[i]), float(v.ggv_ay[i])
    c.n_ggv = n
    c.N, c.dt, c.rk_steps = int(ocp.N), float(ocp.dt), int(ocp.rk_steps)
    for i in range(6):
        c.W[i] = float(ocp.W[i])
    for i in range(4):
        c.W_e[i] = float(ocp.W_e[i])
    c.zl = c.zu = float(ocp.L1_pen)
    c.Zl = c.Zu = float(ocp.L2_pen)
    c.lbu, c.ubu = -float(v.delta_f_dot_max), float(v.delta_f_dot_max)
    c.lbx, c.ubx = -float(v.delta_f_max), float(v.delta_f_max)
    c.lh, c.uh = 0.0, 1.0
    c.ipm_tol, c.ipm_reg = float(ocp.ipm_tol), float(ocp.ipm_reg)
    c.ipm_eta_max, c.ipm_corr_mode, c.ipm_s0 = float(ocp.ipm_eta_max), int(ocp.ipm_corr_mode), float(ocp.ipm_s0)
    c.ipm_max_iter, c.qp_no_shortcut = int(ocp.ipm_max_iter), int(bool(ocp.qp_no_shortcut))
    c.variant = VARIANTS[ocp.variant]
    c.uph, c.n_samples = int(ocp.uph), int(ocp.n_samples)
    sig0, wd = ocp.r2_matrices()
    for i in range(8):
        c.r2_sigma0_diag[i] = float(sig0[i])
        c.sn_x0_std[i] = float(ocp.stds[i])
    for i in range(4):
        c.r2_wd[i] = float(wd[i])
    c.sn_gamma = float(ocp.gamma)
    if ocp.variant == "snmpc":
        # host-side PCE set-up (SNMPC_class.py:87-99): Hammersley samples, Hermite least-squares matrix, x0 offsets
        A, w_off = ocp.pce_matrices()
        L, ns = A.shape
        if ns != int(ocp.n_samples) or ns > MAX_NS or L > PCE_L:
            raise ValueError(f"SNMPC: PCE matrix {A.shape} does not fit n_samples={ocp.n_samples} (max {MAX_NS}) / L <= {PCE_L}")
        c.sn_L = L
        for j in range(L):
            for i in range(ns):
                c.sn_A[j * MAX_NS + i] = float(A[j, i])
        for i in range(ns):
            for k in range(8):
                c.sn_w_off[i * 8 + k] = float(w_off[i, k])
    return c


class PinnedArray:
    """A float64 numpy array backed by page-locked host memory from ``b200nmpc_alloc_host``."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape))
        self._ptr = lib().b200nmpc_alloc_host(max(n, 1) * 8)
        if not self._ptr:
            raise MemoryError("b200nmpc_alloc_host failed")
        buf = (C.c_double * max(n, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(self.shape)

    def __del__(self):
        try:
            if getattr(self, "_ptr", None):
                self.array = None
                lib().b200nmpc_free_host(self._ptr)
                self._ptr = None
        except Exception:
            pass
