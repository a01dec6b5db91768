"""``BatchedOcpSolver``: the AcadosOcpSolver-shaped object the reference controllers hold as
``self.acados_solver`` (created at ``NMPC_STM_acados_settings.py:243``), with a leading batch axis on
every array and the arithmetic in hand-written sm_100a CUDA behind ``libb200nmpc.so``.

Method names, stage/field strings and error behaviour follow the subset of acados_template's
``AcadosOcpSolver`` that the reference actually calls:
``solve, set, get, constraints_set, cost_set, get_cost, get_stats, get_from_qp_in, reset``.
With ``batch=1`` and ``squeeze=True`` arrays lose the batch axis, i.e. the object is a literal
drop-in for the reference's solver handle.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import numpy as np

from . import _lib
from .config import OcpConfig

_SET = {"x": _lib.F_X, "u": _lib.F_U, "yref": _lib.F_YREF, "p": _lib.F_P}
_GET = {"x": _lib.F_X, "u": _lib.F_U, "sl": _lib.F_SL, "su": _lib.F_SU}
_CON = {"lbx": _lib.F_LBX, "ubx": _lib.F_UBX, "lbu": _lib.F_LBU, "ubu": _lib.F_UBU, "lh": _lib.F_LH, "uh": _lib.F_UH}
_COST = {"W": _lib.F_W, "zl": _lib.F_ZL_LIN, "zu": _lib.F_ZU_LIN, "Zl": _lib.F_ZL_QUAD, "Zu": _lib.F_ZU_QUAD}
_QPIN = {"A": (_lib.F_QP_A, (8, 8)), "B": (_lib.F_QP_B, (8, 2)), "b": (_lib.F_QP_b, (8,))}
_STATS = {"time_tot": _lib.S_TIME_TOT, "sqp_iter": _lib.S_SQP_ITER, "qp_iter": _lib.S_QP_ITER,
          "kkt": _lib.S_KKT, "time_batch": _lib.S_TIME_BATCH, "time_linearize": _lib.S_TIME_LINEARIZE,
          "time_qp": _lib.S_TIME_QP}
_SCALAR_STATS = ("time_batch", "time_linearize", "time_qp")


class SolverError(RuntimeError):
    pass


def _check(rc: int):
    if rc != 0:
        msg = _lib.last_error()
        # API misuse mirrors acados_template, which raises on bad field names / sizes
        if rc in (-1, -2, -3):
            raise ValueError(f"b200nmpc: {msg} (code {rc})")
        raise SolverError(f"b200nmpc: {msg} (code {rc})")


class BatchedOcpSolver:
    def __init__(self, ocp: OcpConfig, batch: int = 1, device: int = 0, squeeze: Optional[bool] = None):
        self.ocp = ocp
        self.N = int(ocp.N)
        self.batch = int(batch)
        self.squeeze = (batch == 1) if squeeze is None else bool(squeeze)
        self._L = _lib.lib()
        self._cfg = _lib.make_config(ocp)
        h = C.c_void_p()
        _check(self._L.b200nmpc_create(C.byref(self._cfg), self.batch, int(device), C.byref(h)))
        self._h = h
        self.status = np.zeros(self.batch, dtype=np.int32)
        self._children = weakref.WeakSet()
        self.snmpc = ocp.variant == "snmpc"
        self.n_samples = int(ocp.n_samples) if self.snmpc else 0
        self.nx_aug = 8 * (1 + self.n_samples)      # SNMPC: augmented state [x; x^1..x^ns]
        self.uph_alloc = max(int(ocp.uph), 1) if self.snmpc else 0
        self._np = (self._cfg.sn_L * self.n_samples + 2) if self.snmpc else 0

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None):
            for child in list(getattr(self, "_children", ())):      # closed-loop objects hold device memory tied to this handle
                child.close()
            self._L.b200nmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers
    def _stage_len(self, field: str, stage: int) -> int:
        N = self.N
        if field == "x": return self.nx_aug
        if field == "u": return 2
        if field == "yref": return 4 if stage == N else 6
        if field == "p": return self._np
        if field in ("lbx", "ubx"): return self.nx_aug if stage == 0 else 1
        if field in ("lbu", "ubu", "lh", "uh"): return 1
        if field == "W": return 4 if stage == N else 6
        if field in ("zl", "zu", "Zl", "Zu", "sl", "su"): return 1 if stage == 0 else (2 if stage == N else 3)
        raise ValueError(f"unknown field {field!r}")

    def _value(self, field, stage, value):
        """-> (contiguous float64 array, ld_batch).  Accepts [n] (broadcast) or [B, n]; a W matrix is
        reduced to its diagonal (the reference only ever sets diagonal W, NMPC_class.py:276-296)."""
        n = self._stage_len(field, stage)
        a = np.asarray(value, dtype=np.float64)
        # A weight MATRIX is (n, n) or (B, n, n); a 2-D (B, n) array holds per-instance diagonals.  With batch == n the
        # shape (n, n) is both: it is read as ONE matrix only if it is diagonal (what the reference passes,
        # NMPC_class.py:276-296), otherwise as B rows of diagonals (weights are positive, so such rows are never diagonal).
        is_matrix = False
        if field == "W" and a.ndim == 3 and a.shape[-2:] == (n, n):
            is_matrix = True
        elif field == "W" and a.ndim == 2 and a.shape == (n, n):
            is_matrix = self.batch != n or not np.any(a - np.diag(np.diag(a)))
        if is_matrix:
            off = a - np.einsum("...ii->...i", a)[..., None] * np.eye(a.shape[-1])
            if np.any(off != 0.0):
                raise ValueError("cost_set('W'): only diagonal weight matrices are supported")
            a = np.einsum("...ii->...i", a)
        if a.ndim == 0:
            a = a.reshape(1)
        if a.ndim == 1:
            if a.shape[0] != n:
                raise ValueError(f"{field!r} at stage {stage}: expected {n} values, got {a.shape[0]}")
            return np.ascontiguousarray(a), 0
        if a.ndim == 2 and a.shape == (self.batch, n):
            return np.ascontiguousarray(a), n
        raise ValueError(f"{field!r} at stage {stage}: expected shape ({n},) or ({self.batch}, {n}), got {a.shape}")

    def _out(self, a):
        return a[0] if self.squeeze else a

    # ---- AcadosOcpSolver surface -------------------------------------------------------------------
    def set(self, stage: int, field: str, value):
        if field not in _SET:
            raise ValueError(f"set: unknown field {field!r}")
        a, ld = self._value(field, stage, value)
        _check(self._L.b200nmpc_set(self._h, int(stage), _SET[field], a.ctypes.data, ld))

    def constraints_set(self, stage: int, field: str, value):
        if field not in _CON:
            raise ValueError(f"constraints_set: unknown field {field!r}")
        a, ld = self._value(field, stage, value)
        _check(self._L.b200nmpc_constraints_set(self._h, int(stage), _CON[field], a.ctypes.data, ld))

    def cost_set(self, stage: int, field: str, value):
        if field not in _COST:
            raise ValueError(f"cost_set: unknown field {field!r}")
        a, ld = self._value(field, stage, value)
        _check(self._L.b200nmpc_cost_set(self._h, int(stage), _COST[field], a.ctypes.data, ld))

    def options_set(self, field: str, value):
        """``AcadosOcpSolver.options_set``; fields of this implementation: ``"launch_order_hint"`` (see b200nmpc.h)."""
        _check(self._L.b200nmpc_options_set(self._h, str(field).encode(), float(value)))

    def constraints_get(self, stage: int, field: str):
        """The bound in force per instance, [B, 1] (R2NMPC: the tightened lbx / ubx / uh written on the device)."""
        if field not in _CON or (stage == 0 and field in ("lbx", "ubx")):
            raise ValueError(f"constraints_get: unknown field {field!r} at stage {stage}")
        out = np.empty((self.batch, 1))
        _check(self._L.b200nmpc_get(self._h, int(stage), _CON[field], out.ctypes.data))
        return self._out(out)

    def solve(self):
        """One RTI step for every instance.  Returns the acados status (int for a squeezed batch of 1,
        else int32[B]); 0 = success, 4 = QP failure (iterate left untouched)."""
        _check(self._L.b200nmpc_solve(self._h, self.status.ctypes.data))
        return int(self.status[0]) if self.squeeze else self.status.copy()

    def solve_async(self):
        """Enqueue the RTI step on the handle's stream without synchronising."""
        _check(self._L.b200nmpc_solve(self._h, None))

    def get(self, stage: int, field: str):
        if field not in _GET:
            raise ValueError(f"get: unknown field {field!r}")
        out = np.empty((self.batch, self._stage_len(field, stage)))
        _check(self._L.b200nmpc_get(self._h, int(stage), _GET[field], out.ctypes.data))
        return self._out(out)

    def get_from_qp_in(self, stage: int, field: str):
        if field not in _QPIN:
            raise ValueError(f"get_from_qp_in: unknown field {field!r}")
        fid, shp = _QPIN[field]
        out = np.empty((self.batch,) + shp)
        _check(self._L.b200nmpc_get(self._h, int(stage), fid, out.ctypes.data))
        return self._out(out)

    def get_cost(self):
        out = np.empty(self.batch)
        _check(self._L.b200nmpc_get_cost(self._h, out.ctypes.data))
        return float(out[0]) if self.squeeze else out

    def get_stats(self, name: str):
        if name not in _STATS:
            raise ValueError(f"get_stats: unknown field {name!r}")
        out = np.empty(1 if name in _SCALAR_STATS else self.batch)
        _check(self._L.b200nmpc_get_stats(self._h, _STATS[name], out.ctypes.data))
        if name in _SCALAR_STATS:
            return float(out[0])
        if name == "qp_iter" and self.squeeze:
            return np.array([out[0]])          # acados returns an array over SQP iterations
        return float(out[0]) if self.squeeze else out

    def reset(self):
        _check(self._L.b200nmpc_reset(self._h))

    # ---- bulk variants (O(1) calls per solve) -----------------------------------------------------------
    def set_x0(self, x0):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(x0, np.float64), (self.batch, 8)))
        _check(self._L.b200nmpc_set_x0(self._h, a.ctypes.data))

    def set_yref_all(self, yref):
        a = np.asarray(yref, np.float64)
        if a.shape[-2:] != (self.N + 1, 6):
            raise ValueError(f"set_yref_all: expected [..., {self.N + 1}, 6], got {a.shape}")
        a = np.ascontiguousarray(np.broadcast_to(a, (self.batch, self.N + 1, 6)))
        _check(self._L.b200nmpc_set_yref_all(self._h, a.ctypes.data))

    def get_yref_all(self):
        """The references currently on the device, ``[B, N+1, 6]`` (written by ``set_yref_all`` or the device planner)."""
        out = np.empty((self.batch, self.N + 1, 6))
        _check(self._L.b200nmpc_get_yref_all(self._h, out.ctypes.data))
        return self._out(out)

    def set_iterate(self, X=None, U=None):
        Xa = None if X is None else np.ascontiguousarray(np.broadcast_to(np.asarray(X, np.float64), (self.batch, self.N + 1, 8)))
        Ua = None if U is None else np.ascontiguousarray(np.broadcast_to(np.asarray(U, np.float64), (self.batch, self.N, 2)))
        _check(self._L.b200nmpc_set_iterate(self._h, None if Xa is None else Xa.ctypes.data,
                                            None if Ua is None else Ua.ctypes.data))

    def init_iterate_from_x0(self):
        _check(self._L.b200nmpc_init_iterate_from_x0(self._h))

    def get_iterate(self, out_X=None, out_U=None):
        X = np.empty((self.batch, self.N + 1, 8)) if out_X is None else out_X
        U = np.empty((self.batch, self.N, 2)) if out_U is None else out_U
        _check(self._L.b200nmpc_get_iterate(self._h, X.ctypes.data, U.ctypes.data))
        return (X[0], U[0]) if self.squeeze else (X, U)

    def get_solver_debug(self):
        """[B, 5] rows ``[cost, time_tot, sqp_iter, qp_iter, status]`` (``NMPC_class.py:202-206``)."""
        out = np.empty((self.batch, 5))
        _check(self._L.b200nmpc_get_solver_debug(self._h, out.ctypes.data))
        return self._out(out)

    def update_cost_function_weights(self, params):
        """``update_cost_function_weights(params[7])`` of ``NMPC_class.py:269-317`` for all stages;
        ``params`` is ``[7]`` (all instances) or ``[B, 7]`` (one weight set per instance)."""
        a = np.asarray(params, np.float64)
        if a.shape == (7,):
            _check(self._L.b200nmpc_update_cost_function_weights(self._h, np.ascontiguousarray(a).ctypes.data, 0))
        elif a.shape == (self.batch, 7):
            a = np.ascontiguousarray(a)
            _check(self._L.b200nmpc_update_cost_function_weights(self._h, a.ctypes.data, 7))
        else:
            raise ValueError(f"update_cost_function_weights: expected (7,) or ({self.batch}, 7), got {a.shape}")

    def r2_tighten(self):
        _check(self._L.b200nmpc_r2_tighten(self._h))

    # ---- SNMPC bulk variants -----------------------------------------------------------------------------
    def set_x0_samples(self, x0_samples):
        """x0 sample matrix ``[B, 1+ns, 8]`` (``compute_x0dist``): bulk ``constraints_set(0, 'lbx'/'ubx', ...)``."""
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(x0_samples, np.float64).reshape(-1, 1 + self.n_samples, 8),
                                                 (self.batch, 1 + self.n_samples, 8)))
        _check(self._L.b200nmpc_set_x0_samples(self._h, a.ctypes.data))

    def get_samples(self):
        """Sample iterates ``[B, UPH+1, ns, 8]`` (stage UPH is the frozen state of all later stages)."""
        out = np.empty((self.batch, self.uph_alloc + 1, self.n_samples, 8))
        _check(self._L.b200nmpc_get_samples(self._h, out.ctypes.data))
        return self._out(out)

    def set_samples(self, XS):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(XS, np.float64), (self.batch, self.uph_alloc + 1, self.n_samples, 8)))
        _check(self._L.b200nmpc_set_samples(self._h, a.ctypes.data))

    @property
    def uph(self) -> int:
        return int(self._L.b200nmpc_snmpc_uph(self._h))

    def snmpc_head(self):
        """Diagnostics: the condensed head of the last solve as a dict of arrays (layout ``nmpc_kernels.cuh``)."""
        stride = C.c_int(0)
        _check(self._L.b200nmpc_snmpc_get_head(self._h, None, C.byref(stride)))
        raw = np.empty((self.batch, stride.value))
        _check(self._L.b200nmpc_snmpc_get_head(self._h, raw.ctypes.data, C.byref(stride)))
        H = self.uph_alloc; nh = 2 * H
        o = 0
        out = {}
        for name, shape in (("G", (H, 8, nh)), ("c", (H, 8)), ("rows", (H, nh)), ("v0", (H,)), ("HH", (nh, nh)), ("gH", (nh,))):
            n = int(np.prod(shape))
            out[name] = raw[:, o:o + n].reshape((self.batch,) + shape)
            o += n
        return out

    def synchronize(self):
        _check(self._L.b200nmpc_synchronize(self._h))

    def timer_start(self):
        """CUDA-event stopwatch on the stream the kernels are launched on."""
        _check(self._L.b200nmpc_timer_start(self._h))

    def timer_stop(self) -> float:
        sec = C.c_double(0.0)
        _check(self._L.b200nmpc_timer_stop(self._h, C.byref(sec)))
        return float(sec.value)

    @property
    def launch_count(self) -> int:
        return int(self._L.b200nmpc_launch_count(self._h))

    # ---- raw pointer entry points (pinned host buffers / device tensors) ----------------------------------
    def set_x0_ptr(self, ptr: int, device: bool = False):
        _check((self._L.b200nmpc_set_x0_dev if device else self._L.b200nmpc_set_x0)(self._h, ptr))

    def set_yref_all_ptr(self, ptr: int, device: bool = False):
        _check((self._L.b200nmpc_set_yref_all_dev if device else self._L.b200nmpc_set_yref_all)(self._h, ptr))

    def get_results_async_ptr(self, X_ptr, U_ptr, debug_ptr):
        """Asynchronous read-back into pinned host buffers; complete after :meth:`synchronize`."""
        _check(self._L.b200nmpc_get_results_async(self._h, X_ptr, U_ptr, debug_ptr))

    def get_iterate_ptr(self, X_ptr, U_ptr, device: bool = False):
        _check((self._L.b200nmpc_get_iterate_dev if device else self._L.b200nmpc_get_iterate)(self._h, X_ptr, U_ptr))
