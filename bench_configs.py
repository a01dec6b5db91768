#!/usr/bin/env python
"""bench_configs.py -- the other BASELINE.json configurations (parity-test cases, not the headline bench line):

    configs[2]  Stochastic NMPC, 13 sigma points, Monteblanco, batch = 1024 closed-loop simulations (on-device loop)
    configs[3]  Nominal NMPC horizon sweep N in {20, 40, 80}, LVMS, batch = 8192
    configs[4]  R2NMPC, Modena, N = 30, 2048 instances per GPU (16384 over 8 GPUs)
    + the nominal on-device closed loop at batch = 4096 (planner -> solve -> plant -> filter -> logs per step)
    + the batched BO objective: 2048 (weight set, segment) pairs in one evaluate() call, wall clock

One JSON line per configuration; device time from CUDA events on the solver's stream; 1 GPU.
    python bench_configs.py [--steps K] [--only snmpc,sweep,r2,loop,eval]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tum_control_b200.closed_loop import ClosedLoopSim  # noqa: E402
from tum_control_b200.config import OcpConfig, load_reftraj  # noqa: E402
from tum_control_b200.planner import synthetic_batch  # noqa: E402
from tum_control_b200.solver import BatchedOcpSolver  # noqa: E402


def algorithmic_bytes(N, nx=8, nu=2, extra=0):
    return 8 * (8 + 6 * N + 4 + 2 * ((N + 1) * nx + nu * N) + 4 + extra)


def time_cold_solves(S, x0, yref, steps, warm=3):
    ts, k1, k2 = [], [], []
    for i in range(warm + steps):
        S.set_x0(x0); S.set_yref_all(yref); S.init_iterate_from_x0(); S.synchronize()
        S.timer_start(); S.solve_async(); t = S.timer_stop()
        if i >= warm:
            ts.append(t); k1.append(S.get_stats("time_linearize")); k2.append(S.get_stats("time_qp"))
    return float(np.mean(ts)), float(np.mean(k1)), float(np.mean(k2))


def stats_of(S):
    it = S.get_stats("qp_iter")
    return {"qp_iter_mean": float(it.mean()), "qp_iter_max": int(it.max()), "frac_interior_point": float((it > 0).mean()),
            "kkt_max": float(S.get_stats("kkt").max())}


def closed_loop(ocp, track, B, nsteps, label, weights=None):
    S = BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)
    sim = ClosedLoopSim(S, track, log_capacity=nsteps + 8)
    n_ref = len(load_reftraj(track)["pos_x"])
    starts = (np.arange(B) * n_ref) // B                     # start indices spread uniformly over the lap
    sim.reset(sim.initial_states(starts))
    sim.run(8)                                               # warm-up: first cold solves
    sec = sim.run(nsteps)
    lg = sim.logs()
    dbg = lg["simSolverDebug"][:, 8:]
    out = {"config": label, "metric": "closed-loop steps/s (planner + RTI solve + plant + filter, all on device)",
           "value": B * nsteps / sec, "unit": "solves/s", "batch": B, "steps": nsteps, "ms_per_step": 1e3 * sec / nsteps,
           "status_nonzero": int((dbg[:, :, 4] != 0).sum()), "qp_iter_mean": float(dbg[:, :, 3].mean()),
           "qp_iter_max": int(dbg[:, :, 3].max()),
           "lat_dev_max_m": float(np.abs(_lat_dev(lg)).max()), "launches_per_step": None}
    return out


def _lat_dev(lg):
    from tum_control_b200.closed_loop import deviations
    n = lg["simU"].shape[1]
    return deviations(lg["CiLX"][:, :n], lg["simREF"])[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--loop-steps", type=int, default=100)
    ap.add_argument("--only", default="snmpc,sweep,r2,loop,eval")
    args = ap.parse_args()
    only = set(args.only.split(","))
    peak = 6538.3
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])

    if "snmpc" in only:
        # configs[2]: SNMPC, 13 samples, batch 1024: cold RTI solves and the closed loop
        N, Tp, B, ns = 30, 2.4, 1024, 13
        ocp = OcpConfig(N=N, Tp=Tp, variant="snmpc", rk_steps=1, n_samples=ns, uph=5)
        x0, yref = synthetic_batch("monteblanco", B, N, Tp, seed=0)
        S = BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)
        t, k1, k2 = time_cold_solves(S, x0, yref, args.steps)
        nb = algorithmic_bytes(N, nx=8 * (1 + ns))
        print(json.dumps({"config": "configs[2] SNMPC n_s=13 UPH=5 monteblanco N=30 B=1024, cold-start RTI step",
                          "value": B / t, "unit": "solves/s", "ms_per_step": 1e3 * t, "linearize_condense_ms": 1e3 * k1,
                          "qp_ms": 1e3 * k2, "algorithmic_bytes_per_solve": nb,
                          "roofline": {"bound": "hbm", "achieved": nb * B / t / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": nb * B / t / 1e9 / peak}, **stats_of(S)}), flush=True)
        del S
        print(json.dumps(closed_loop(ocp, "monteblanco", B, args.loop_steps,
                                     "configs[2] SNMPC n_s=13 UPH=5 monteblanco N=30, 1024 closed-loop simulations")), flush=True)
    if "sweep" in only:
        for N in (20, 40, 80):
            Tp, B = 0.08 * N, 8192
            ocp = OcpConfig(N=N, Tp=Tp)
            x0, yref = synthetic_batch("lvms", B, N, Tp, seed=0)
            S = BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)
            t, k1, k2 = time_cold_solves(S, x0, yref, args.steps)
            nb = algorithmic_bytes(N)
            print(json.dumps({"config": f"configs[3] nominal NMPC lvms N={N} B=8192, cold-start RTI step",
                              "value": B / t, "unit": "solves/s", "ms_per_step": 1e3 * t, "linearize_ms": 1e3 * k1,
                              "qp_ms": 1e3 * k2, "algorithmic_bytes_per_solve": nb,
                              "roofline": {"bound": "hbm", "achieved": nb * B / t / 1e9, "peak": peak, "unit": "GB/s",
                                           "frac": nb * B / t / 1e9 / peak}, **stats_of(S)}), flush=True)
            del S
    if "r2" in only:
        N, Tp, B = 30, 2.4, 2048
        ocp = OcpConfig(N=N, Tp=Tp, variant="r2")
        x0, yref = synthetic_batch("modena", B, N, Tp, seed=0)
        S = BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)
        t, k1, k2 = time_cold_solves(S, x0, yref, args.steps)
        nb = algorithmic_bytes(N, extra=2 * 3 * N)
        print(json.dumps({"config": "configs[4] R2NMPC modena N=30, 2048 instances per GPU, cold-start RTI step incl. tightening",
                          "value": B / t, "unit": "solves/s", "ms_per_step": 1e3 * t, "linearize_ms": 1e3 * k1,
                          "qp_and_tighten_ms": 1e3 * k2, "algorithmic_bytes_per_solve": nb,
                          "roofline": {"bound": "hbm", "achieved": nb * B / t / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": nb * B / t / 1e9 / peak}, **stats_of(S)}), flush=True)
        del S
    if "loop" in only:
        print(json.dumps(closed_loop(OcpConfig(N=30, Tp=2.4), "monteblanco", 4096, args.loop_steps,
                                     "nominal NMPC monteblanco N=30, 4096 closed-loop simulations (warm RTI steady state)")),
              flush=True)
    if "eval" in only:
        # BO_WMPC objective_function semantics: each pair = one weight set on one 60-index segment of the lap (N=38, Tp=3.04)
        import time
        from tum_control_b200.evaluation import BatchedSegmentEvaluator
        B, seg = 2048, 60                                                    # 60 indices = 120 m: ~250 closed-loop steps of Ts = 0.02 s
        n_ref = len(load_reftraj("monteblanco")["pos_x"])
        rng = np.random.default_rng(0)
        base = np.array([2.8, 0.4, 0.2, 38.1, 101.4, 106.7, 9.9])           # MPC_params.yaml:17-24 [q_xy q_yaw q_vel r_jerk r_steer L1 L2]
        params = base * np.exp(rng.uniform(-0.7, 0.7, size=(B, 7)))
        starts = rng.integers(0, n_ref - seg - 1, size=B)
        ev = BatchedSegmentEvaluator("monteblanco", B, max_steps=800, chunk=100)
        ev.evaluate(params, starts, starts + seg)                           # warm-up (allocations, first launches)
        t0 = time.perf_counter()
        out = ev.evaluate(params, starts, starts + seg)
        wall = time.perf_counter() - t0
        steps_run = ev.sim.steps_done
        print(json.dumps({"config": "BO objective: 2048 (weight set, 60-index monteblanco segment) pairs, N=38, one evaluate() call",
                          "value": B / wall, "unit": "pairs/s (wall clock, host post-processing included)", "wall_s": wall,
                          "closed_loop_steps_per_pair_mean": float(out["steps"].mean()), "batch_steps_run": int(steps_run),
                          "closed_loop_steps_per_s": B * steps_run / wall, "feasible_frac": float(out["feasible"].mean()),
                          "done_frac": float(out["done"].mean())}), flush=True)


if __name__ == "__main__":
    main()
