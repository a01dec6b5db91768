"""Small nominal + SNMPC + closed-loop run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tum_control_b200.closed_loop import ClosedLoopSim
from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch
from tum_control_b200.solver import BatchedOcpSolver

N, Tp, B = 12, 0.96, 10
x0, yref = synthetic_batch("monteblanco", B, N, Tp, seed=0)
for variant in ("nominal", "r2", "snmpc"):
    kw = dict(rk_steps=1) if variant == "snmpc" else {}
    S = BatchedOcpSolver(OcpConfig(N=N, Tp=Tp, variant=variant, **kw), batch=B, device=0, squeeze=False)
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    for _ in range(2):
        st = S.solve()
    print(variant, "status", st.tolist(), "iters", S.get_stats("qp_iter").tolist())
    sim = ClosedLoopSim(S, "monteblanco", log_capacity=4)
    sim.configure_disturbances("uniform", np.full(7, 0.1), "gaussian", np.full(7, 0.01), seed=1)
    sim.reset(sim.initial_states(np.arange(B) * 50))
    sim.run(3)
    print(" loop ok", np.isfinite(sim.logs()["simU"]).all())
    # device-side stop conditions: one instance stops at once (end index = start index), the others run on
    sim.set_stop_conditions(end_idx=np.arange(B) * 50 + np.where(np.arange(B) == 0, 0, 400), max_lat_dev=5.0)
    sim.reset(sim.initial_states(np.arange(B) * 50))
    sim.run(3)
    print(" stop state", sim.stop_state()[0].tolist(), sim.logs_range(1, 2, keys=("CiLX", "ref_idx"))["ref_idx"].shape)
