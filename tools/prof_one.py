"""One cold-start solve of the headline batch (and one SNMPC solve) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch
from tum_control_b200.solver import BatchedOcpSolver

which = sys.argv[1] if len(sys.argv) > 1 else "nominal"
if which == "nominal":
    B, N, Tp = 4096, 30, 2.4
    ocp = OcpConfig(N=N, Tp=Tp)
else:
    B, N, Tp = 1024, 30, 2.4
    ocp = OcpConfig(N=N, Tp=Tp, variant="snmpc", rk_steps=1, n_samples=13, uph=5)
x0, yref = synthetic_batch("monteblanco", B, N, Tp, seed=0)
S = BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)
for _ in range(3):
    S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0()
    S.solve()
print("done", S.get_stats("qp_iter").mean())
