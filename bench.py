#!/usr/bin/env python
"""bench.py -- headline benchmark: NMPC solves/sec (N=30, 8-state Pacejka STM, one RTI-SQP step per solve).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4-20|cfg4-40|cfg4-80|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch.  Default workload = BASELINE.json configs[1] (`cfg2`): nominal
NMPC, Monteblanco, N=30, batch 4096 perturbed x0 per GPU, cold start (x_k = x0, u = 0: the state of a freshly
constructed / reset() reference solver, which needs the interior-point QP path on most instances).  The other
BASELINE configurations run with --workload (same JSON contract, `config.workload` names them):
   cfg3     configs[2]  Stochastic NMPC, 13 samples, UPH 5, Monteblanco, 1024 closed-loop simulations per GPU; a step is
                        one closed-loop step of every simulation (planner -> RTI solve -> plant -> filter, on the device)
   cfg4-N   configs[3]  nominal NMPC, LVMS, horizon N in {20, 40, 80}, batch 8192, cold-start RTI step
   cfg5     configs[4]  R2NMPC, Modena, N=30, 2048 instances per GPU (16384 over 8), cold-start RTI step + tightening
   value : whole-job solves/s with (x0, yref) already resident in HBM (device arrays in the caller's
           instance-major layout); timed with CUDA events on the stream the kernels run on, L2 flushed
           between timed steps, max over ranks.
   e2e   : the same step through the public host API (BatchedOcpSolver): pinned host x0/yref -> H2D ->
           solve -> D2H of the full (x, u) prediction + [cost, time, sqp_iter, qp_iter, status] per instance
           (cfg3: ClosedLoopSim, per step H2D of that step's disturbance rows and D2H of the step's log rows).
The batch is embarrassingly parallel: at N GPUs every rank owns its own batch (weak scaling), no
data-path collective; one NCCL all-gather of the solved trajectories is timed separately (logging only).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tum_control_b200.config import OcpConfig  # noqa: E402
from tum_control_b200.planner import synthetic_batch  # noqa: E402

METRIC = "NMPC solves/sec (N=30, 8-state Pacejka STM, one RTI-SQP step per solve)"
NX, NU = 8, 2

# BASELINE.json configs[1..4]; batch = instances per GPU
WORKLOADS = {
    "cfg2": dict(variant="nominal", track="monteblanco", N=30, batch=4096, closed_loop=False,
                 label="nominal NMPC, monteblanco reftraj, N=30, batch=4096 perturbed x0 per GPU, cold-start RTI step "
                       "(BASELINE.json configs[1])"),
    "cfg3": dict(variant="snmpc", track="monteblanco", N=30, batch=1024, closed_loop=True, n_samples=13, uph=5,
                 label="Stochastic NMPC, 13 samples, UPH=5, monteblanco reftraj, N=30, 1024 closed-loop simulations per GPU, "
                       "one closed-loop step (planner, RTI solve, plant, filter) per simulation and step (BASELINE.json configs[2])"),
    "cfg5": dict(variant="r2", track="modena", N=30, batch=2048, closed_loop=False,
                 label="R2NMPC, modena reftraj, N=30, 2048 instances per GPU, cold-start RTI step incl. tightening "
                       "(BASELINE.json configs[4])"),
}
for _n in (20, 40, 80):
    WORKLOADS[f"cfg4-{_n}"] = dict(variant="nominal", track="lvms", N=_n, batch=8192, closed_loop=False,
                                   label=f"nominal NMPC, lvms reftraj, N={_n}, batch=8192 perturbed x0 per GPU, cold-start RTI "
                                         "step (BASELINE.json configs[3])")


def make_ocp(w):
    N = w["N"]
    kw = dict(N=N, Tp=0.08 * N, variant=w["variant"])
    if w["variant"] == "snmpc":
        kw.update(rk_steps=1, n_samples=w["n_samples"], uph=w["uph"])
    return OcpConfig(**kw)


def workload_config(name, world):
    """The `config` object of the JSON line: identical in both arms (`--impl ours` / `--impl reference`)."""
    w = WORKLOADS[name]
    return {"workload": w["label"], "name": name, "variant": w["variant"], "track": w["track"], "N": w["N"], "Tp": 0.08 * w["N"],
            "batch_per_gpu": w["batch"], "global_batch": w["batch"] * world,
            "parallelism": f"dp{world} (independent instances, no collective)",
            "l2": "flushed between timed steps (256 MiB write)", "seed": "rank",
            "qp_tolerance": "KKT <= 1e-9 (OcpConfig.ipm_tol default; the reference's HPIPM default is 1e-8)"}


def algorithmic_bytes(w):
    """compulsory fp64 traffic of one solve = read x0, yref, warm start; write primal + 4 stats
    (SNMPC: the augmented state of 8 (1 + n_s); R2NMPC: + the tightened bound words)."""
    N = w["N"]
    nx = NX * (1 + w["n_samples"]) if w["variant"] == "snmpc" else NX
    extra = 2 * 3 * N if w["variant"] == "r2" else 0
    return 8 * (8 + 6 * N + 4 + 2 * ((N + 1) * nx + NU * N) + 4 + extra)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(name):
    """DRAM bytes per launch of the dominant kernel: NOT measured by this run (ncu cannot run inside a timed bench) --
    the figure of the committed `ncu --set full` capture of the same kernel and workload, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            if name == "cfg2":
                return d.get("k_qp_solve_dram_bytes_per_launch"), d.get("source", "profiles/traffic.json")
        except Exception:
            pass
    return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag, self.ok = index, [], set(), None, False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:        # CUDA index -> NVML handle through the PCI bus id (CUDA_VISIBLE_DEVICES may remap indices)
                import torch
                pr = torch.cuda.get_device_properties(index)
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(
                    f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------------
# CPU side: the oracle (kind "port": a restatement pinned against the reference's logs / model dump, NOT acados)
# ---------------------------------------------------------------------------------------------------------
class CpuStepper:
    """One bounded step of the workload on the host cores: the CPU oracle on the first `n` instances of the batch."""

    def __init__(self, name, x0, yref, n, cores):
        import oracle
        self.w, self.name, self.n, self.cores = WORKLOADS[name], name, int(n), int(cores)
        self.ocp = make_ocp(self.w)
        self.x0, self.yref = x0[:n], yref[:n]
        N = self.w["N"]
        v = self.w["variant"]
        if v == "nominal":
            self.O = oracle.Oracle(self.ocp)
        elif v == "snmpc":
            self.O = oracle.SnmpcOracle(self.ocp)
        else:
            self.O = [oracle.Oracle(self.ocp) for _ in range(self.cores)]
            self.qp = [oracle.QpView() for _ in range(self.cores)]
        self.N = N
        self.desc = {"nominal": "oracle/nmpc_oracle.c (Riccati interior point), orc_rti_batch over pthreads",
                     "snmpc": "oracle/snmpc_oracle.c (dense full-condensing interior point), one RTI step per instance, thread pool",
                     "r2": "oracle/nmpc_oracle.c RTI step + orc_r2_tighten per instance, thread pool"}[v]

    def step(self):
        """cold-start RTI step of the n instances; returns wall seconds"""
        from concurrent.futures import ThreadPoolExecutor
        N, n, v = self.N, self.n, self.w["variant"]
        if v == "nominal":
            X = np.repeat(self.x0[:, None, :], N + 1, axis=1).copy()
            U = np.zeros((n, N, 2))
            t = time.perf_counter()
            self.O.rti_batch(self.x0, self.yref, X, U, nthreads=self.cores)
            return time.perf_counter() - t
        if v == "snmpc":
            its = [self.O.init_guess(self.x0[b]) for b in range(n)]

            def one(b):
                Xo, XSo, Uo = its[b]
                return self.O.rti_step(self.x0[b], self.yref[b], Xo, XSo, Uo)[0]
        else:
            its = [self.O[0].init_guess(self.x0[b]) for b in range(n)]

            def one(b):
                w = b % self.cores          # one oracle object (bounds are per instance and mutated by the tightening) per worker
                O, qp = self.O[w], self.qp[w]
                lb, ub = O.lb.copy(), O.ub.copy()
                Xo, Uo = its[b]
                st = O.rti_step(self.x0[b], self.yref[b], Xo, Uo, qp=qp)[0]
                O.r2_tighten(qp.get("A", N).reshape(N, 8, 8).copy(), Xo)
                O.lb[...] = lb; O.ub[...] = ub
                return st
        t = time.perf_counter()
        with ThreadPoolExecutor(self.cores) as ex:
            if v == "r2":       # a worker's oracle object must not be shared by two threads: stride the instances by worker
                list(ex.map(lambda wk: [one(b) for b in range(wk, n, self.cores)], range(self.cores)))
            else:
                list(ex.map(one, range(n), chunksize=max(1, n // (4 * self.cores))))
        return time.perf_counter() - t


def cpu_sample_size(name):
    """instances per CPU step: the whole single-GPU batch where that is ~0.1 s of wall, a bounded sample otherwise"""
    return {"cfg2": 4096, "cfg3": 256, "cfg5": 1024}.get(name, 2048)


def cpu_baseline(name, x0, yref, budget_s=15.0):
    cores = os.cpu_count() or 1
    n = min(cpu_sample_size(name), len(x0))
    st = CpuStepper(name, x0, yref, n, cores)
    CpuStepper(name, x0, yref, min(64, n), cores).step()                  # page in
    t1 = st.step()
    reps = max(1, int(round(budget_s / max(t1 * cores, 1e-9))))            # ~budget_s core-seconds of CPU work
    ts = sum(st.step() for _ in range(reps))
    kind = "RTI solves of the closed-loop initial states (planner / plant / filter excluded)" if WORKLOADS[name]["closed_loop"] \
        else "cold-start instances of the same batch"
    return {"value": reps * n / ts, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {n} {kind}, {st.desc}, {cores} threads, {ts:.2f} s wall ({ts * cores:.0f} core-seconds)"}


def run_reference(args):
    """--impl reference: the reference's CPU path.  acados/HPIPM/CasADi are absent from this image and from
    /root/reference, so this arm times the CPU oracle (a restatement pinned against the
    reference's committed logs and model dump) with all host threads, on a bounded sample of the same workload per step."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    name = args.workload
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    n = cpu_sample_size(name)
    x0, yref = synthetic_batch(w["track"], n, w["N"], 0.08 * w["N"], seed=0)
    st = CpuStepper(name, x0, yref, n, cores)
    for _ in range(args.warmup):
        st.step()
    times = [st.step() for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(times))
    value = n / (ms * 1e-3)
    kind = "RTI solves of the closed-loop initial states (planner / plant / filter excluded)" if w["closed_loop"] \
        else "cold-start instances"
    sample = f"{n} {kind} per step (seed-0 batch of one GPU{'' if n == w['batch'] else ', first ' + str(n)}), {st.desc}, {cores} threads"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(name, max(1, args.gpus)),
           "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "acados/HPIPM are not installable here; this is the CPU oracle (oracle/), pinned against the reference's "
                   "committed closed-loop logs and acados_ocp_SNMPC.json; the host cores process a bounded sample per step"}
    print(json.dumps(out), flush=True)


def run_ours(args):
    import torch
    from tum_control_b200 import _lib
    from tum_control_b200.solver import BatchedOcpSolver

    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)

    name = args.workload
    w = WORKLOADS[name]
    B, N = w["batch"], w["N"]
    Tp = 0.08 * N
    ocp = make_ocp(w)
    # every rank draws its own shard (different seed per rank): independent instances, no exchange
    x0, yref = synthetic_batch(w["track"], B, N, Tp, seed=rank)
    S = BatchedOcpSolver(ocp, batch=B, device=dev, squeeze=False)
    handles = [S]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for Sh in handles:
            Sh.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # L2-flush buffer (2x the 126 MB L2)
    extras = {}
    if not w["closed_loop"]:
        # device-resident inputs (instance-major, as a caller would hold them)
        x0_d = torch.from_numpy(x0).cuda()
        yref_d = torch.from_numpy(yref).cuda()
        torch.cuda.synchronize()

        def device_step():
            S.set_x0_ptr(x0_d.data_ptr(), device=True)
            S.set_yref_all_ptr(yref_d.data_ptr(), device=True)
            S.init_iterate_from_x0()
            S.solve_async()

        # end-to-end arm: tum_control_b200.pipeline.BatchPipeline, the package's API for a stream of independent batches.
        # Every batch in flight has its own handles (own streams) and pinned host buffers; a batch is split over `nsplit`
        # handles so that the H2D / D2H copies of one part overlap the solves of the others.  submit() enqueues pinned H2D
        # of (x0, yref), the RTI step and the async D2H of (x, u, stats); result() waits for one batch.  With depth = 2 step i
        # is submitted before the host waits for the results of step i - 1, so the draining SMs of one batch are filled by
        # the next.  The synchronous sequence (submit, wait, read: one batch at a time) is timed beside it.
        from tum_control_b200.pipeline import BatchPipeline
        nsplit = int(os.environ.get("BENCH_E2E_SPLIT", "2"))
        # batches in flight: two cover the tail of a batch of >= 4096 instances; smaller batches (cfg5: 2048 = 0.86 waves
        # of resident warps) need more of them to keep the SMs full
        depth = int(os.environ.get("BENCH_E2E_DEPTH", str(max(2, -(-8192 // B)))))
        pipe = BatchPipeline(ocp, batch=B, device=dev, depth=depth, split=nsplit)
        handles += pipe.handles
        for d in range(depth):                      # the synthetic inputs, written once into each slot's pinned buffers
            px0, pyr = pipe.inputs(d)
            px0[...] = x0; pyr[...] = yref
        hx0, hy = pipe.inputs(0)
        hX, hU, hD = (pipe._slots[0][1][k] for k in ("x", "u", "stats"))

        def e2e_step(i=0, pipelined=False):
            if not pipelined:
                pipe.drain()
                return pipe.result(pipe.submit())[2]
            pipe.submit()
            return pipe.result(pipe.oldest)[2] if pipe.in_flight >= depth else hD.array

        h2d = int(hx0.nbytes + hy.nbytes)
        d2h = int(hX.array.nbytes + hU.array.nbytes + hD.array.nbytes)
        api = (f"tum_control_b200.pipeline.BatchPipeline(depth={depth}, split={nsplit}): {nsplit} BatchedOcpSolver handles of "
               f"batch/{nsplit} per batch in flight; submit() = pinned H2D of (x0, yref), solve, async D2H of (x, u, stats) per "
               "handle, result() waits for one batch; the copies of one part overlap the solves of the others"
               + (f"; {depth} batches in flight (the host waits for the results of step i - {depth - 1} after submitting step i, "
                  "each batch in flight has its own handles and pinned buffers)" if depth > 1 else ""))
    else:
        # closed loop on the device: planner -> solve -> plant -> filter -> logs per step (ClosedLoopSim)
        from tum_control_b200.closed_loop import ClosedLoopSim
        from tum_control_b200.config import load_reftraj
        n_ref = len(load_reftraj(w["track"])["pos_x"])
        starts = ((np.arange(B) + rank * 7) * n_ref) // B % n_ref          # start indices spread uniformly over the lap
        cap = 2 * (max(args.warmup, 3) + args.steps) + 64
        sim = ClosedLoopSim(S, w["track"], log_capacity=cap)
        sim.reset(sim.initial_states(starts))

        def device_step():
            sim.run(1, sync=False)

        # end to end: a second simulation driven step by step from the host -- every step the host uploads that step's
        # state-estimation noise rows [B, 7] (seeded; sim_main_params.yaml:54-60 stds) and reads back the step's log rows
        Se = BatchedOcpSolver(ocp, batch=B, device=dev, squeeze=False)
        handles.append(Se)
        sime = ClosedLoopSim(Se, w["track"], log_capacity=cap)
        sime.reset(sime.initial_states(starts))
        from tum_control_b200.config import SIM_MAIN_PARAMS as sp
        sig = 0.1 * np.array([sp["w_posx"], sp["w_posy"], sp["w_yaw"], sp["w_vlong"], sp["w_vlat"], sp["w_yawrate"], sp["w_delta_f"]])
        noise = np.random.default_rng(rank).standard_normal((cap, B, 7)) * sig
        keys = ("CiLX", "simU", "simSolverDebug")

        depth = 1

        def e2e_step(i=0, pipelined=False):
            n0 = sime.steps_done
            sime.set_disturbances(None, noise[n0])
            sime.run(1, sync=False)
            return sime.logs_range(n0, 1, keys=keys)["simSolverDebug"][:, 0]

        h2d = int(noise[0].nbytes)
        d2h = int(B * (2 * 7 + 2 + 5) * 8)
        api = ("ClosedLoopSim driven step by step from the host: per step H2D of the step's state-estimation noise rows, "
               "run(1) (planner, solve, plant, filter on the device), D2H of the step's CiLX / simU / simSolverDebug rows")

    W = max(args.warmup, 3)
    for _ in range(W):
        device_step()
        e2e_step()
    barrier()

    sampler = ClockSampler(dev)
    sampler.start()
    # ---- device-resident timing: CUDA events on the solver's stream, L2 flushed between steps
    l0 = S.launch_count
    t_steps, t_k1, t_k2 = [], [], []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        S.timer_start()
        device_step()
        t_steps.append(S.timer_stop())
        t_k1.append(S.get_stats("time_linearize")); t_k2.append(S.get_stats("time_qp"))
    barrier()
    launches = S.launch_count - l0
    dev_s = allmax(float(np.sum(t_steps)))
    qp_iter = S.get_stats("qp_iter")
    status = S.status.copy(); S._L.b200nmpc_get_status(S._h, status.ctypes.data)
    kkt = S.get_stats("kkt")

    # ---- the same step without scheduling history: the launch-order hint (instances that iterated longest in the handle's
    # previous solve start first) forgotten before every step = what the very first solve of a handle costs
    if not w["closed_loop"]:
        t_first = []
        for _ in range(min(args.steps, 10)):
            S.options_set("launch_order_hint", 1)
            flush.zero_()
            torch.cuda.synchronize()
            S.timer_start()
            device_step()
            t_first.append(S.timer_stop())
        first_s = allmax(float(np.mean(t_first)))
        extras["first_solve_of_a_handle"] = {"ms_per_step": 1e3 * first_s, "solves_per_s": B * world / first_s,
                                             "note": "launch-order hint cleared before every step (index order)"}
        extras["launch_order"] = ("instances that needed the most interior-point iterations in the handle's previous solve "
                                  "start first (options_set('launch_order_hint')); results do not depend on the order")

    # ---- the same step at the QP tolerance of round 1 (KKT <= 1e-11 instead of 1e-9: ~1.1 more interior-point iterations)
    if not w["closed_loop"]:
        import dataclasses
        St = BatchedOcpSolver(dataclasses.replace(ocp, ipm_tol=1e-11), batch=B, device=dev, squeeze=False)
        handles.append(St)
        t_tight = []
        for i in range(3 + min(args.steps, 10)):
            St.set_x0_ptr(x0_d.data_ptr(), device=True); St.set_yref_all_ptr(yref_d.data_ptr(), device=True)
            St.init_iterate_from_x0()
            flush.zero_()
            torch.cuda.synchronize(); St.synchronize()
            St.timer_start(); St.solve_async(); t = St.timer_stop()
            if i >= 3:
                t_tight.append(t + float(np.mean(t_steps)) - float(np.mean(t_k1)) - float(np.mean(t_k2)))   # + set / init share of a step
        tight_s = allmax(float(np.mean(t_tight)))
        extras["at_qp_tolerance_1e-11"] = {"ms_per_step": 1e3 * tight_s, "solves_per_s": B * world / tight_s,
                                           "qp_iter_mean": float(St.get_stats("qp_iter").mean()),
                                           "note": "ipm_tol = 1e-11, the round-1 default; everything else identical"}

    # ---- device-resident with two batches in flight: step i runs on handle i % 2 (own stream), nothing waits between steps,
    # so the draining SMs of one batch (B = 4096 is 1.73 waves of resident warps) are filled by the next one.  The L2 flush
    # of a step is ordered before the step's kernels by an event; device time between the first flush and the last kernel.
    if not w["closed_loop"]:
        Sb = BatchedOcpSolver(ocp, batch=B, device=dev, squeeze=False)
        handles.append(Sb)
        pair = [S, Sb]
        ext = [torch.cuda.ExternalStream(int(Sh._L.b200nmpc_stream(Sh._h)), device=dev) for Sh in pair]

        def pair_step(i):
            Sh = pair[i % 2]
            flush.zero_()
            ev = torch.cuda.Event(); ev.record()
            ext[i % 2].wait_event(ev)
            Sh.set_x0_ptr(x0_d.data_ptr(), device=True); Sh.set_yref_all_ptr(yref_d.data_ptr(), device=True)
            Sh.init_iterate_from_x0()
            Sh.solve_async()

        for i in range(4):
            pair_step(i)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        for i in range(args.steps):
            pair_step(i)
        ends = []
        for st in ext:
            e1 = torch.cuda.Event(enable_timing=True); e1.record(st); ends.append(e1)
        barrier()
        pair_s = allmax(max(e0.elapsed_time(e1) for e1 in ends) * 1e-3)
        extras["two_batches_in_flight"] = {
            "ms_per_step": 1e3 * pair_s / args.steps, "solves_per_s": B * world * args.steps / pair_s,
            "status_nonzero": int((Sb.solve() != 0).sum()),
            "note": "device-resident inputs, steps alternate between two handles (own streams) with no wait in between; "
                    "L2 flushed before every step's kernels (event-ordered 256 MiB write)"}

    # ---- end-to-end timing through the host API (wall clock between synchronisations)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dbg = e2e_step()
    barrier()
    e2e_sync_s = e2e_s = allmax(time.perf_counter() - t0)
    if depth > 1:       # the same K steps with `depth` batches in flight; the timed region ends when the last results are in
        for i in range(depth):
            e2e_step(i, True)
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            dbg = e2e_step(i, True)
        last = pipe.drain()
        dbg = last[2] if last is not None else dbg
        barrier()
        e2e_s = allmax(time.perf_counter() - t0)
        # ... and once more with the launch-order history of every handle cleared before every submit: what a stream of
        # DIFFERENT batches gets (the bench sends the same synthetic batch every step, so the hint of the previous solve is
        # exact; with batches in flight the order matters little because the tail of a batch is hidden anyway)
        if not w["closed_loop"]:
            t0 = time.perf_counter()
            for i in range(args.steps):
                for Sh in pipe._slots[pipe.next_ticket % depth][0]:
                    Sh.options_set("launch_order_hint", 1)
                e2e_step(i, True)
            pipe.drain()
            barrier()
            extras["e2e_without_launch_order_history"] = {"solves_per_s": B * world * args.steps / allmax(time.perf_counter() - t0),
                                                          "note": "launch-order hint of every handle cleared before every submit"}
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    assert dbg.shape[-1] == 5

    # ---- SURVEY 8(d) warm protocol: solves #2..#11 of a sequence, each warm-started from the previous solution with x0
    # advanced to the predicted x_1 (RTI steady state); yref stays.  Reported beside the cold-start headline.
    if not w["closed_loop"]:
        Xd = torch.empty((B, N + 1, 8), dtype=torch.float64, device="cuda")
        Ud = torch.empty((B, N, 2), dtype=torch.float64, device="cuda")
        device_step()                                                     # solve #1 (cold)
        warm_t, warm_it = [], []
        for _ in range(10):
            S.get_iterate_ptr(Xd.data_ptr(), Ud.data_ptr(), device=True)
            S.synchronize()
            x1 = Xd[:, 1, :].contiguous()
            if w["variant"] != "snmpc":
                flush.zero_()
            torch.cuda.synchronize()
            S.timer_start()
            S.set_x0_ptr(x1.data_ptr(), device=True)
            S.solve_async()
            warm_t.append(S.timer_stop())
            warm_it.append(float(S.get_stats("qp_iter").mean()))
        warm_ms = 1e3 * float(np.mean(warm_t))
        extras["warm_rti"] = {"protocol": "SURVEY 8(d): solves #2..#11, warm start = previous solution, x0 <- predicted x_1, L2 flushed",
                              "ms_per_step": warm_ms, "solves_per_s": B * world / (warm_ms * 1e-3),
                              "qp_iter_mean": float(np.mean(warm_it)), "status_nonzero": int((S.solve() != 0).sum())}

    # ---- logging all-gather (not part of the solve): [B, (N+1)*8 + N*2 + 5] per rank
    if world > 1:
        rowlen = (N + 1) * 8 + N * 2 + 5
        Xg = torch.empty((B, (N + 1) * 8), dtype=torch.float64, device="cuda")
        Ug = torch.empty((B, N * 2), dtype=torch.float64, device="cuda")
        Dg = torch.empty((B, 5), dtype=torch.float64, device="cuda")
        S.get_iterate_ptr(Xg.data_ptr(), Ug.data_ptr(), device=True)
        S._L.b200nmpc_get_solver_debug_dev(S._h, Dg.data_ptr())
        S.synchronize()
        local_rows = torch.cat([Xg, Ug, Dg], dim=1).contiguous()
        gathered = torch.empty((world * B, rowlen), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(gathered, local_rows)         # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_gather_into_tensor(gathered, local_rows); e1.record()
        torch.cuda.synchronize()
        extras["allgather_ms"] = allmax(e0.elapsed_time(e1))
        extras["allgather_bytes_per_rank"] = int(local_rows.numel() * 8)

    if rank == 0:
        total = B * world
        ms_per_step = 1e3 * dev_s / args.steps
        value = total / (dev_s / args.steps)
        peak, peak_src = measured_peak()
        bytes_per_solve = algorithmic_bytes(w)
        k2_s = float(np.mean(t_k2))
        achieved = bytes_per_solve * B / k2_s / 1e9
        traffic, traffic_src = measured_traffic(name)
        out = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, world),
            "roofline": {"bound": "hbm", "kernel": "k_qp_solve_warp", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src if traffic is not None else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_solve": bytes_per_solve,
                         "kernel_ms": 1e3 * k2_s, "linearize_ms": 1e3 * float(np.mean(t_k1)),
                         "note": "path is FP64-issue/latency bound, not HBM bound (SURVEY section 8d); see DESIGN.md",
                         "solver": {"qp_iter_mean": float(qp_iter.mean()), "qp_iter_max": int(qp_iter.max()),
                                    "frac_interior_point": float((qp_iter > 0).mean()),
                                    "status_nonzero": int((status != 0).sum()), "kkt_max": float(kkt.max())},
                         **extras},
            "e2e": {"value": total / (e2e_s / args.steps), "unit": "solves/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "api": api, "batches_in_flight": depth,
                    "synchronous": {"value": total / (e2e_sync_s / args.steps), "ms_per_step": 1e3 * e2e_sync_s / args.steps,
                                    "note": "one batch at a time: issue, wait, read"},
                    "without_launch_order_history": extras.pop("e2e_without_launch_order_history", None)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(name, x0, yref)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
