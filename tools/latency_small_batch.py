"""Latency and throughput of one RTI solve over the batch size, from the literal drop-in case B = 1 to 64 Ki instances
(cold start and the following warm solve; device time from CUDA events) -- diagnostic."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch
from tum_control_b200.solver import BatchedOcpSolver

N, Tp = 30, 2.4
ocp = OcpConfig(N=N, Tp=Tp)
x0a, yrefa = synthetic_batch("monteblanco", 4096, N, Tp, seed=0)
BATCHES = tuple(int(a) for a in sys.argv[1:]) or (1, 8, 148, 592, 2368, 4096, 16384, 65536)
for B in BATCHES:
    reps = (B + 4095) // 4096
    x0, yref = np.tile(x0a, (reps, 1))[:B], np.tile(yrefa, (reps, 1, 1))[:B]
    S = BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)
    cold, warm = [], []
    for r in range(6):
        S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0(); S.synchronize()
        S.timer_start(); S.solve_async(); cold.append(S.timer_stop())
        S.timer_start(); S.solve_async(); warm.append(S.timer_stop())
    it = S.get_stats("qp_iter")
    print(json.dumps({"B": B, "cold_ms": 1e3 * float(np.median(cold[1:])), "warm_ms": 1e3 * float(np.median(warm[1:])),
                      "cold_solves_per_s": B / float(np.median(cold[1:])), "warm_solves_per_s": B / float(np.median(warm[1:])),
                      "qp_iter_max_warm": int(it.max()), "qp_iter_mean_warm": float(it.mean())}))
    S.close()
