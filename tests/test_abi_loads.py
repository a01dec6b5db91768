"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/b200nmpc.h declares,
agrees with the Python mirror of the config struct, and REFUSES to run without a CUDA device (no fallback)."""
import os
import re

import numpy as np
import pytest

from tum_control_b200 import _lib
from tum_control_b200.config import OcpConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "b200nmpc.h")).read()
    declared = set(re.findall(r"\b(b200nmpc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200nmpc_handle", "b200nmpc_config"}
    L = _lib.lib()
    missing = sorted(s for s in declared if not hasattr(L, s))
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared | {"b200nmpc_config"}, sorted(set(_lib.EXPORTS) - declared)


def test_config_struct_layout_and_defaults():
    L = _lib.lib()
    import ctypes as C
    assert L.b200nmpc_sizeof_config() == C.sizeof(_lib.B200Config)
    c = _lib.make_config(OcpConfig(N=30, Tp=2.4))
    assert c.N == 30 and abs(c.dt - 0.08) < 1e-15 and c.rk_steps == 3 and c.n_ggv == 10
    np.testing.assert_allclose(list(c.W), 0.01 * np.array([2.8, 2.8, 0.4, 0.2, 38.1, 101.4]))
    assert (c.lbu, c.ubu, c.lbx, c.ubx, c.lh, c.uh) == (-0.322, 0.322, -0.610865, 0.610865, 0.0, 1.0)
    assert c.zl == 106.7 and c.Zu == 9.9


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from tum_control_b200.solver import BatchedOcpSolver, SolverError
    with pytest.raises(SolverError, match="no CPU fallback"):
        BatchedOcpSolver(OcpConfig(), batch=2)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may touch oracle/."""
    pkg = os.path.join(ROOT, "tum_control_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
                assert 'include "../../oracle' not in src and "nmpc_oracle.h" not in src, f
