"""ctypes binding of the CPU ORACLE (test infrastructure, NOT the product).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  See ``nmpc_oracle.h``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("ORC_LIB", os.path.join(_HERE, "_build", "liboracle.so"))

MAX_GGV = 16
MAX_N = 128
NC = 3


class OrcModel(C.Structure):
    _fields_ = (
        [(n, C.c_double) for n in ("lf", "lr", "m", "Iz", "ro", "S", "Cd",
                                   "Bf", "Cf", "Df", "Ef", "Br", "Cr", "Dr", "Er",
                                   "g", "fr0", "fr1", "fr4", "ax_brake")]
        + [("ggv_v", C.c_double * MAX_GGV), ("ggv_ax", C.c_double * MAX_GGV), ("ggv_ay", C.c_double * MAX_GGV),
           ("dt", C.c_double), ("ipm_tol", C.c_double), ("ipm_reg", C.c_double), ("ipm_eta_max", C.c_double), ("ipm_s0", C.c_double),
           ("n_ggv", C.c_int), ("N", C.c_int), ("rk_steps", C.c_int), ("cost_vel_abs", C.c_int),
           ("ipm_max_iter", C.c_int), ("qp_no_shortcut", C.c_int), ("ipm_corr_mode", C.c_int)]
    )


MAX_NS, MAX_L = 16, 10


class OrcSnmpc(C.Structure):
    _fields_ = [("A", C.c_double * (MAX_L * MAX_NS)), ("w_off", C.c_double * (MAX_NS * 8)), ("kappa", C.c_double),
                ("ns", C.c_int), ("L", C.c_int), ("H", C.c_int), ("pad_", C.c_int)]


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc, a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("nmpc_oracle.c", "snmpc_oracle.c", "nmpc_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if "ORC_LIB" in os.environ:
        return _LIB_PATH
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        mp = C.POINTER(OrcModel)
        L.orc_f.argtypes = [mp, _dp, _dp, _dp]
        L.orc_f_jac.argtypes = [mp, _dp, _dp, _dp, _dp, _dp]
        L.orc_integrate.argtypes = [mp, _dp, _dp, _dp, _dp, _dp]
        L.orc_flow.argtypes = [mp, _dp, _dp, C.c_double, C.c_int, _dp]
        L.orc_h.argtypes = [mp, _dp, _dp]
        L.orc_h.restype = C.c_double
        L.orc_cost.argtypes = [mp] + [_dp] * 10
        L.orc_cost_output.argtypes = [mp, _dp, C.c_void_p, _dp]
        L.orc_cost_output.restype = None
        L.orc_cost.restype = C.c_double
        L.orc_rti_step.argtypes = [mp] + [_dp] * 12 + [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_rti_step.restype = C.c_int
        L.orc_rti_batch.argtypes = [mp, C.c_int, C.c_int] + [_dp] * 12
        L.orc_rti_batch.restype = C.c_int
        L.orc_linearize.argtypes = [mp] + [_dp] * 11 + [C.c_void_p]
        L.orc_qp_solve.argtypes = [mp, C.c_void_p]
        L.orc_r2_tighten.argtypes = [mp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, _dp, _dp, _dp]
        L.orc_plant_step.argtypes = [mp, _dp, _dp, C.c_void_p, C.c_double, C.c_int, _dp]
        sp = C.POINTER(OrcSnmpc)
        L.orc_snmpc_rti_step.argtypes = [mp, sp] + [_dp] * 13 + [C.c_void_p] * 5
        L.orc_snmpc_rti_step.restype = C.c_int
        L.orc_snmpc_disc.argtypes = [mp, sp, _dp, _dp, C.c_int, _dp]
        L.orc_snmpc_h.argtypes = [mp, sp, _dp, C.c_int]
        L.orc_snmpc_h.restype = C.c_double
        L.orc_sizeof_snmpc.restype = C.c_int
        assert L.orc_sizeof_snmpc() == C.sizeof(OrcSnmpc), "OrcSnmpc layout mismatch"
        L.orc_sizeof_qp.restype = C.c_int
        L.orc_sizeof_model.restype = C.c_int
        assert L.orc_sizeof_model() == C.sizeof(OrcModel), "OrcModel layout mismatch"
        _lib = L
    return _lib


def make_model(ocp, vehicle=None) -> OrcModel:
    """Build the oracle's model struct from a ``tum_control_b200.config.OcpConfig``."""
    v = vehicle if vehicle is not None else ocp.vehicle
    M = OrcModel()
    for n in ("lf", "lr", "m", "Iz", "ro", "S", "Cd", "Bf", "Cf", "Df", "Ef", "Br", "Cr", "Dr", "Er",
              "g", "fr0", "fr1", "fr4", "ax_brake"):
        setattr(M, n, float(getattr(v, n)))
    n = len(v.ggv_v)
    assert n <= MAX_GGV
    for i in range(n):
        M.ggv_v[i], M.ggv_ax[i], M.ggv_ay[i] = float(v.ggv_v[i]), float(v.ggv_ax[i]), float(v.ggv_ay[i])
    M.n_ggv = n
    M.N = int(ocp.N)
    M.dt = float(ocp.dt)
    M.rk_steps = int(ocp.rk_steps)
    M.cost_vel_abs = int(bool(ocp.cost_vel_abs))
    M.ipm_max_iter = int(ocp.ipm_max_iter)
    M.ipm_tol = float(ocp.ipm_tol)
    M.ipm_reg = float(ocp.ipm_reg)
    M.ipm_eta_max = float(ocp.ipm_eta_max)
    M.ipm_s0 = float(ocp.ipm_s0)
    M.ipm_corr_mode = int(ocp.ipm_corr_mode)
    M.qp_no_shortcut = int(bool(ocp.qp_no_shortcut))
    return M


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class QpView:
    """Field access into an ``orc_qp`` blob (layout mirrors nmpc_oracle.h)."""

    _layout = [("N", "i", 1, 1)]

    def __init__(self):
        self.buf = np.zeros(lib().orc_sizeof_qp() // 8 + 1, dtype=np.float64)
        n1 = MAX_N + 1
        off = 1  # int N padded to 8 bytes
        self._f = {}
        for name, rows, cols in [
            ("A", MAX_N, 64), ("B", MAX_N, 16), ("b", MAX_N, 8), ("Q", n1, 64), ("R", MAX_N, 4),
            ("q", n1, 8), ("r", MAX_N, 2), ("cval", n1, 3), ("gh", n1, 8), ("lo", n1, 3), ("up", n1, 3),
            ("zl", n1, 3), ("zu", n1, 3), ("Zl", n1, 3), ("Zu", n1, 3),
        ]:
            self._f[name] = (off, rows, cols); off += rows * cols
        self._act_off = off  # int act[n1][3] -> n1*3 ints
        nint = n1 * 3
        off += (nint + 1) // 2
        for name, rows, cols in [("dx0", 1, 8), ("dx", n1, 8), ("du", MAX_N, 2), ("sl", n1, 3), ("su", n1, 3),
                                 ("lam_l", n1, 3), ("lam_u", n1, 3), ("pi", n1, 8)]:
            self._f[name] = (off, rows, cols); off += rows * cols
        self._tail = off  # int iters, int status ; double kkt

    @property
    def ptr(self):
        return self.buf.ctypes.data

    def get(self, name, N=None):
        off, rows, cols = self._f[name]
        a = self.buf[off:off + rows * cols].reshape(rows, cols)
        return a if N is None else a[:N]

    @property
    def iters(self):
        return int(self.buf[self._tail:self._tail + 1].view(np.int32)[0])

    @property
    def status(self):
        return int(self.buf[self._tail:self._tail + 1].view(np.int32)[1])

    @property
    def kkt(self):
        return float(self.buf[self._tail + 1])


class Oracle:
    """Single-instance and batched RTI steps with reference-default stage data."""

    def __init__(self, ocp, vehicle=None):
        self.ocp = ocp
        self.M = make_model(ocp, vehicle)
        self.N = ocp.N
        self.W = _c(ocp.stage_weights())
        lb, ub = ocp.stage_bounds()
        self.lb, self.ub = _c(lb), _c(ub)
        zl, zu, Zl, Zu = ocp.stage_penalties()
        self.zl, self.zu, self.Zl, self.Zu = _c(zl), _c(zu), _c(Zl), _c(Zu)

    # -- reference: update_cost_function_weights (NMPC_class.py:269-317): W without the 0.01 factor
    def update_cost_function_weights(self, p):
        W = np.tile(np.array([p[0], p[0], p[1], p[2], p[3], p[4]], float), (self.N + 1, 1))
        W[self.N, 4:] = 0.0
        self.W = _c(W)
        self.zl = _c(np.full((self.N + 1, 3), float(p[5]))); self.zu = self.zl.copy()
        self.Zl = _c(np.full((self.N + 1, 3), float(p[6]))); self.Zu = self.Zl.copy()

    def init_guess(self, x0):
        """Constructor / reset() initial guess: x_k = x0 for all k, u = 0."""
        X = np.tile(np.asarray(x0, float), (self.N + 1, 1))
        U = np.zeros((self.N, 2))
        return _c(X), _c(U)

    def rti_step(self, x0, yref, X, U, want_slacks=False, qp: QpView = None):
        stats = np.zeros(4)
        sl = np.zeros((self.N + 1, 3)); su = np.zeros((self.N + 1, 3))
        st = lib().orc_rti_step(C.byref(self.M), _c(x0), _c(yref), self.W, self.lb, self.ub, self.zl, self.zu,
                                self.Zl, self.Zu, X, U, stats, sl.ctypes.data, su.ctypes.data,
                                qp.ptr if qp is not None else None)
        return (st, stats, sl, su) if want_slacks else (st, stats)

    def linearize(self, x0, yref, X, U, qp: QpView):
        lib().orc_linearize(C.byref(self.M), _c(x0), _c(X), _c(U), _c(yref), self.W, self.lb, self.ub,
                            self.zl, self.zu, self.Zl, self.Zu, qp.ptr)

    def qp_solve(self, qp: QpView):
        lib().orc_qp_solve(C.byref(self.M), qp.ptr)

    def rti_batch(self, x0, yref, X, U, nthreads=1):
        B = x0.shape[0]
        stats = np.zeros((B, 4))
        bad = lib().orc_rti_batch(C.byref(self.M), B, nthreads, _c(x0), _c(yref), self.W, self.lb, self.ub,
                                  self.zl, self.zu, self.Zl, self.Zu, X, U, stats)
        return bad, stats

    def integrate(self, x, u):
        xn = np.zeros(8); A = np.zeros((8, 8)); Bm = np.zeros((8, 2))
        lib().orc_integrate(C.byref(self.M), _c(x), _c(u), xn, A, Bm)
        return xn, A, Bm

    def f(self, x, u):
        xd = np.zeros(8)
        lib().orc_f(C.byref(self.M), _c(x), _c(u), xd)
        return xd

    def f_jac(self, x, u):
        xd = np.zeros(8); Jx = np.zeros((8, 8)); Ju = np.zeros((8, 2))
        lib().orc_f_jac(C.byref(self.M), _c(x), _c(u), xd, Jx, Ju)
        return xd, Jx, Ju

    def h(self, x):
        g = np.zeros(8)
        return lib().orc_h(C.byref(self.M), _c(x), g), g

    def cost_output(self, x, u=None):
        """NLS output y(x, u) of a path stage (6) or, with u None, of the terminal stage (4)."""
        y = np.zeros(6 if u is not None else 4)
        lib().orc_cost_output(C.byref(self.M), _c(x), None if u is None else _c(u).ctypes.data, y)
        return y

    def plant_step(self, x7, u2, Ts=0.02, nsub=4, w7=None):
        out = np.zeros(7)
        w = _c(w7) if w7 is not None else None
        lib().orc_plant_step(C.byref(self.M), _c(x7), _c(u2), w.ctypes.data if w is not None else None,
                             float(Ts), int(nsub), out)
        return out

    def r2_tighten(self, A, X, uph=None):
        """Reduced_Robustified_NMPC_class.py:287-365 on one instance: A [N,8,8] (pre-step linearisation), X [N+1,8]
        (post-step iterate).  Updates self.lb / self.ub in place (bx rows and the upper h bound)."""
        sig0, wd = self.ocp.r2_matrices()
        N = self.N
        lbx = self.lb[:, 1].copy(); ubx = self.ub[:, 1].copy(); uh = self.ub[:, 2].copy()
        lib().orc_r2_tighten(C.byref(self.M), _c(A.reshape(N, 64)), _c(X), _c(np.diag(sig0)), _c(wd),
                             int(self.ocp.uph if uph is None else uph), float(self.ocp.vehicle.delta_f_max), lbx, ubx, uh)
        self.lb[:, 1] = lbx; self.ub[:, 1] = ubx; self.ub[:, 2] = uh


class SnmpcOracle(Oracle):
    """Stochastic NMPC RTI step through the dense (full condensing) oracle of snmpc_oracle.c."""

    def __init__(self, ocp, A=None, w_off=None, H=None):
        super().__init__(ocp)
        from tum_control_b200 import pce
        S = OrcSnmpc()
        if A is None:
            w, A, _ = pce.samples_and_A(ocp.n_samples, ocp.stds, ocp.expansion_degree)
            w_off = pce.sample_offsets(w, ocp.stds)
        A = np.asarray(A, float)
        self.ns = A.shape[1] if A.size else 0
        self.H = int(ocp.uph if H is None else H)
        S.ns, S.L, S.H = self.ns, (A.shape[0] if A.size else 0), self.H
        for j in range(S.L):
            for i in range(self.ns):
                S.A[j * self.ns + i] = float(A[j, i])
        if self.ns:
            wo = np.asarray(w_off, float)
            for i in range(self.ns):
                for c in range(8):
                    S.w_off[i * 8 + c] = float(wo[i, c])
        S.kappa = float(np.sqrt((1.0 - ocp.gamma) / ocp.gamma))
        self.S, self.A, self.w_off = S, A, (np.asarray(w_off, float) if self.ns else np.zeros((0, 8)))

    def init_guess(self, x0):
        """SNMPC_class.py:127-130: every stage starts at the x0 sample matrix, u = 0."""
        X, U = super().init_guess(x0)
        XS = np.zeros((max(self.H, 1), max(self.ns, 1), 8))
        for k in range(self.H):
            XS[k, :self.ns] = np.asarray(x0, float) + self.w_off
        return X, _c(XS), U

    def rti_step(self, x0, yref, X, XS, U, debug=False):
        stats = np.zeros(4)
        N = self.N
        G = np.zeros(((N + 1) * 8, 2 * N)) if debug else None
        rows = np.zeros((3 * (N + 1), 2 * N)) if debug else None
        v0 = np.zeros(3 * (N + 1)) if debug else None
        st = lib().orc_snmpc_rti_step(C.byref(self.M), C.byref(self.S), _c(x0), _c(yref), self.W, self.lb, self.ub,
                                      self.zl, self.zu, self.Zl, self.Zu, X, XS, U, stats, None, None,
                                      G.ctypes.data if debug else None, rows.ctypes.data if debug else None,
                                      v0.ctypes.data if debug else None)
        return (st, stats, G, rows, v0) if debug else (st, stats)

    def disc(self, xa, u, stop):
        out = np.zeros_like(np.asarray(xa, float))
        lib().orc_snmpc_disc(C.byref(self.M), C.byref(self.S), _c(xa), _c(u), int(stop), out)
        return out

    def h_tilde(self, xa, stop):
        return lib().orc_snmpc_h(C.byref(self.M), C.byref(self.S), _c(xa), int(stop))
