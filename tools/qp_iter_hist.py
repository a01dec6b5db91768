"""Distribution of interior-point iterations over a batched closed loop for several step-length rules (diagnostic)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tum_control_b200.closed_loop import ClosedLoopSim
from tum_control_b200.config import OcpConfig, load_reftraj
from tum_control_b200.planner import synthetic_batch
from tum_control_b200.solver import BatchedOcpSolver

B, n = 4096, 600
nref = len(load_reftraj("monteblanco")["pos_x"])
x0c, yrefc = synthetic_batch("monteblanco", B, 30, 2.4, seed=0)
SETTINGS = [tuple(float(v) for v in a.split(',')) for a in sys.argv[1:]] or [(0.99, 0, 1.0), (0.99, 0, 0.3), (0.99, 0, 0.1), (0.99, 0, 0.03)]
for eta, corr, s0 in SETTINGS:
    ocp = OcpConfig(N=30, Tp=2.4, ipm_eta_max=eta, ipm_corr_mode=int(corr), ipm_s0=s0)
    S = BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)
    # cold batch
    ts = []
    for _ in range(4):
        S.set_x0(x0c); S.set_yref_all(yrefc); S.init_iterate_from_x0(); S.synchronize()
        S.timer_start(); S.solve_async(); ts.append(S.timer_stop())
    itc = S.get_stats("qp_iter")
    sim = ClosedLoopSim(S, "monteblanco", log_capacity=n)
    sim.reset(sim.initial_states((np.arange(B) * nref) // B))
    sec = sim.run(n)
    it = sim.logs()["simSolverDebug"][:, :, 3]
    print(json.dumps({"s0": s0, "eta_max": eta, "corr_mode": corr, "cold_ms": 1e3 * float(np.mean(ts[1:])), "cold_iter_mean": float(itc.mean()),
                      "cold_iter_max": int(itc.max()), "loop_ms_per_step": 1e3 * sec / n, "loop_iter_mean": float(it.mean()),
                      "loop_iter_ge25": int((it >= 25).sum()), "loop_iter_eq50": int((it >= 50).sum()), "loop_iter_max": int(it.max())}))
    del sim, S
