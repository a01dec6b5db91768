"""tum_control_b200 -- Blackwell-native batched NMPC hot path for bzarr/TUM-CONTROL.

One data-parallel path is accelerated: the ``acados_ocp_solver.solve()`` real-time-iteration step behind
the reference's three NMPC controllers.  ``BatchedOcpSolver`` keeps the AcadosOcpSolver surface with a
leading batch axis and calls hand-written sm_100a CUDA through the C-ABI of ``include/b200nmpc.h``.
"""
from .config import OcpConfig, VehicleModel, load_reftraj  # noqa: F401

__all__ = ["OcpConfig", "VehicleModel", "load_reftraj", "BatchedOcpSolver", "BatchPipeline"]


def __getattr__(name):
    if name == "BatchedOcpSolver":
        from .solver import BatchedOcpSolver
        return BatchedOcpSolver
    if name == "BatchPipeline":
        from .pipeline import BatchPipeline
        return BatchPipeline
    raise AttributeError(name)
