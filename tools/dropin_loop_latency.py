"""Wall-clock cost of the reference's OWN per-step call sequence through the drop-in shim (NMPC_class.py:111-112,163-206):
constraints_set(0, 'lbx'/'ubx', x0); N x set(j, 'yref', ...); set(N, 'yref', ...); solve(); get(0, 'u'); N x get(j, 'x');
get_stats x 3; get_cost -- at the batch sizes the reference's callers use (1: main.py, 8: BO pool, 16: RL envs).
Closed loop in MPC-sim mode (next x0 = predicted x_1, planner window from the host mirror) -- diagnostic, wall clock."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tum_control_b200.config import OcpConfig, load_reftraj
from tum_control_b200.planner import PlannerEmulator, initial_state
from tum_control_b200.solver import BatchedOcpSolver

N, Tp = 30, 2.4
ocp = OcpConfig(N=N, Tp=Tp)
ref = load_reftraj("monteblanco")
planner = PlannerEmulator(ref)

for B in (1, 8, 16):
    S = BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)
    x = np.stack([initial_state(ref, 40 * b) for b in range(B)])
    S.set_x0(x); S.init_iterate_from_x0()
    t_api, t_solve, t_plan = [], [], []
    for step in range(60):
        t0 = time.perf_counter()
        yref = np.stack([planner.yref([x[b, 0], x[b, 1], x[b, 2], np.hypot(x[b, 3], x[b, 4])], N, Tp) for b in range(B)])
        t1 = time.perf_counter()
        S.constraints_set(0, "lbx", x); S.constraints_set(0, "ubx", x)
        for j in range(N):
            S.set(j, "yref", yref[:, j, :])
        S.set(N, "yref", yref[:, N, :4])
        t2 = time.perf_counter()
        status = S.solve()
        t3 = time.perf_counter()
        u0 = S.get(0, "u")
        pred = np.stack([S.get(j, "x") for j in range(N)], axis=1)
        tt = S.get_stats("time_tot"); S.get_stats("sqp_iter"); S.get_stats("qp_iter"); S.get_cost()
        t4 = time.perf_counter()
        assert np.all(status == 0)
        x = pred[:, 1, :].copy()
        if step >= 10:
            t_plan.append(t1 - t0); t_api.append((t2 - t1) + (t4 - t3)); t_solve.append(t3 - t2)
    print(json.dumps({"B": B, "set_get_calls_ms": 1e3 * float(np.median(t_api)), "solve_call_ms": 1e3 * float(np.median(t_solve)),
                      "step_ms_without_planner": 1e3 * float(np.median(np.add(t_api, t_solve))),
                      "host_planner_ms": 1e3 * float(np.median(t_plan)), "device_time_tot_ms": 1e3 * float(tt.sum())}))
    S.close()
