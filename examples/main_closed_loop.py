#!/usr/bin/env python
"""The reference's main.py (closed-loop driver, main.py:20-81) on the batched device loop.

    python examples/main_closed_loop.py --controller nominal|snmpc|r2 --track monteblanco --batch 1 --steps 5000 --out Logs/

One vehicle (batch 1) reproduces the reference run; a larger batch simulates that many vehicles started at reference
indices spread over the lap.  Writes one full_logs.npz per vehicle in the reference's schema."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tum_control_b200.closed_loop import ClosedLoopSim  # noqa: E402
from tum_control_b200.config import SIM_MAIN_PARAMS, load_reftraj  # noqa: E402
from tum_control_b200.controllers import ocp_from_reference_params  # noqa: E402
from tum_control_b200.solver import BatchedOcpSolver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--controller", default="nominal", choices=["nominal", "snmpc", "r2"])
    ap.add_argument("--track", default="monteblanco")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=int(SIM_MAIN_PARAMS["T"] / SIM_MAIN_PARAMS["Ts"]))      # Nsim, simMode 0
    ap.add_argument("--start", type=int, default=0, help="idx_ref_start of vehicle 0")
    ap.add_argument("--out", default=None, help="directory for full_logs_<b>.npz")
    args = ap.parse_args()

    kw = dict(rk_steps=1) if args.controller == "snmpc" else {}
    ocp = ocp_from_reference_params(variant=args.controller, **kw)          # MPC_params.yaml / sim_main_params.yaml values
    solver = BatchedOcpSolver(ocp, batch=args.batch, device=0, squeeze=False)
    sim = ClosedLoopSim(solver, args.track, log_capacity=args.steps)
    n_ref = len(load_reftraj(args.track)["pos_x"])
    starts = (args.start + (np.arange(args.batch) * n_ref) // args.batch) % n_ref
    sim.reset(sim.initial_states(starts))
    sec = sim.run(args.steps)
    lg = sim.logs()
    dbg = lg["simSolverDebug"]
    print(f"{args.controller} on {args.track}: {args.batch} vehicle(s) x {args.steps} steps in {sec:.2f} s device time "
          f"({args.batch * args.steps / sec:.0f} closed-loop steps/s); status != 0: {int((dbg[:, :, 4] != 0).sum())}; "
          f"qp_iter mean {dbg[:, :, 3].mean():.2f} max {int(dbg[:, :, 3].max())}")
    if args.out:
        os.makedirs(args.out, exist_ok=True)
        for b in range(args.batch):
            sim.save_logs(os.path.join(args.out, f"full_logs_{b}.npz"), b, logs=lg)
        print("logs written to", args.out)


if __name__ == "__main__":
    main()
