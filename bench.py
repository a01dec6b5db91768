#!/usr/bin/env python
"""bench.py -- headline benchmark: NMPC solves/sec (N=30, 8-state Pacejka STM, one RTI-SQP step per solve).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: BASELINE.json configs[1] -- nominal NMPC,
Monteblanco, N=30, batch 4096 perturbed x0 per GPU, cold start (x_k = x0, u = 0: the state of a freshly
constructed / reset() reference solver, which needs the interior-point QP path on most instances).
   value : whole-job solves/s with (x0, yref) already resident in HBM (device arrays in the caller's
           instance-major layout); timed with CUDA events on the stream the kernels run on, L2 flushed
           between timed steps, max over ranks.
   e2e   : the same step through the public host API (BatchedOcpSolver): pinned host x0/yref -> H2D ->
           solve -> D2H of the full (x, u) prediction + [cost, time, sqp_iter, qp_iter, status] per instance.
The batch is embarrassingly parallel: at N GPUs every rank owns its own 4096 instances (weak scaling), no
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
TRACK, N_H, TP, B_PER_GPU = "monteblanco", 30, 2.4, 4096
NX, NU = 8, 2


def algorithmic_bytes(N, nx=NX, nu=NU):
    """compulsory fp64 traffic of one solve = read x0, yref, warm start; write primal + 4 stats."""
    return 8 * (8 + 6 * N + 4 + 2 * ((N + 1) * nx + nu * N) + 4)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (or None)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f).get("k_qp_solve_dram_bytes_per_launch")
        except Exception:
            pass
    return None


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


def cpu_baseline(ocp, x0, yref, budget_s=15.0):
    """The CPU oracle (kind 'port': a restatement, not acados) on the host cores of this box, bounded sample."""
    import oracle
    O = oracle.Oracle(ocp)
    cores = os.cpu_count() or 1
    N = ocp.N

    def run(n):
        X = np.repeat(x0[:n, None, :], N + 1, axis=1).copy()
        U = np.zeros((n, N, 2))
        t = time.perf_counter()
        O.rti_batch(x0[:n], yref[:n], X, U, nthreads=cores)
        return time.perf_counter() - t

    run(min(256, len(x0)))                      # page in
    n = len(x0)
    t1 = run(n)
    reps = max(1, int(round(budget_s / max(t1 * cores, 1e-9))))       # ~budget_s core-seconds of CPU work
    t0 = time.perf_counter()
    for _ in range(reps):
        run(n)
    ts = time.perf_counter() - t0
    return {"value": reps * n / ts, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {n} cold-start instances of the same batch, oracle/nmpc_oracle.c, {cores} pthreads, "
                      f"{ts:.2f} s wall ({ts * cores:.0f} core-seconds)"}


def run_reference(args):
    """--impl reference: the reference's CPU path.  acados/HPIPM/CasADi are absent from this image and from
    /root/reference, so this arm times the CPU oracle (a restatement pinned against the
    reference's committed logs) with all host threads, on a bounded sample of the same workload per step."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    import oracle
    ocp = OcpConfig(N=N_H, Tp=TP)
    cores = os.cpu_count() or 1
    n = B_PER_GPU                      # the whole single-GPU batch of the headline workload: ~0.1 s of wall per step
    x0, yref = synthetic_batch(TRACK, n, N_H, TP, seed=0)
    O = oracle.Oracle(ocp)

    def step():
        X = np.repeat(x0[:, None, :], N_H + 1, axis=1).copy()
        U = np.zeros((n, N_H, 2))
        t = time.perf_counter()
        O.rti_batch(x0, yref, X, U, nthreads=cores)
        return time.perf_counter() - t

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(times))
    value = n / (ms * 1e-3)
    sample = f"{n} cold-start instances per step (the seed-0 batch of one GPU), {cores} pthreads"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"nominal NMPC, {TRACK} reftraj, N={N_H}, batch={n} perturbed x0, cold-start RTI step "
                                  f"(BASELINE.json configs[1]; the host cores process one GPU's batch per step)",
                      "track": TRACK, "N": N_H, "Tp": TP, "batch_per_step": n},
           "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "acados/HPIPM are not installable here; this is the CPU oracle (oracle/nmpc_oracle.c), "
                   "pinned against the reference's committed closed-loop logs"}
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

    B, N = B_PER_GPU, N_H
    ocp = OcpConfig(N=N, Tp=TP)
    # every rank draws its own shard (different seed per rank): independent instances, no exchange
    x0, yref = synthetic_batch(TRACK, B, N, TP, seed=rank)
    S = BatchedOcpSolver(ocp, batch=B, device=dev, squeeze=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        S.synchronize()
        for Sh in S2:
            Sh.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # device-resident inputs (instance-major, as a caller would hold them) and an L2-flush buffer
    x0_d = torch.from_numpy(x0).cuda()
    yref_d = torch.from_numpy(yref).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def device_step():
        S.set_x0_ptr(x0_d.data_ptr(), device=True)
        S.set_yref_all_ptr(yref_d.data_ptr(), device=True)
        S.init_iterate_from_x0()
        S.solve_async()

    # pinned host buffers for the end-to-end arm
    hx0 = _lib.PinnedArray((B, 8)); hy = _lib.PinnedArray((B, N + 1, 6))
    hX = _lib.PinnedArray((B, N + 1, 8)); hU = _lib.PinnedArray((B, N, 2)); hD = _lib.PinnedArray((B, 5))
    hx0.array[...] = x0; hy.array[...] = yref
    # two handles of B/2 instances each: the H2D / D2H copies of one half overlap the solve of the other half
    halves = [(0, B // 2), (B // 2, B)]
    S2 = [BatchedOcpSolver(ocp, batch=hi - lo, device=dev, squeeze=False) for lo, hi in halves]

    def e2e_step():
        for Sh, (lo, hi) in zip(S2, halves):
            Sh.set_x0_ptr(hx0.array[lo:hi].ctypes.data)
            Sh.set_yref_all_ptr(hy.array[lo:hi].ctypes.data)
            Sh.init_iterate_from_x0()
            Sh.solve_async()
            Sh.get_results_async_ptr(hX.array[lo:hi].ctypes.data, hU.array[lo:hi].ctypes.data, hD.array[lo:hi].ctypes.data)
        for Sh in S2:
            Sh.synchronize()
        return hD.array

    for _ in range(max(args.warmup, 3)):
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

    # ---- end-to-end timing through the host API (wall clock between synchronisations)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dbg = e2e_step()
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    # ---- warm (steady-state RTI) step, reported beside the cold-start headline
    warm_times = []
    for _ in range(3):
        S.timer_start(); S.solve_async(); warm_times.append(S.timer_stop())
    warm_ms = 1e3 * float(np.mean(warm_times[1:]))

    # ---- logging all-gather (not part of the solve): [B, (N+1)*8 + N*2 + 5] per rank
    allgather_ms = None
    if world > 1:
        rowlen = (N + 1) * 8 + N * 2 + 5
        Xd = torch.empty((B, (N + 1) * 8), dtype=torch.float64, device="cuda")
        Ud = torch.empty((B, N * 2), dtype=torch.float64, device="cuda")
        Dd = torch.empty((B, 5), dtype=torch.float64, device="cuda")
        S.get_iterate_ptr(Xd.data_ptr(), Ud.data_ptr(), device=True)
        S._L.b200nmpc_get_solver_debug_dev(S._h, Dd.data_ptr())
        S.synchronize()
        local_rows = torch.cat([Xd, Ud, Dd], dim=1).contiguous()
        gathered = torch.empty((world * B, rowlen), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(gathered, local_rows)         # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_gather_into_tensor(gathered, local_rows); e1.record()
        torch.cuda.synchronize()
        allgather_ms = allmax(e0.elapsed_time(e1))

    if rank == 0:
        total = B * world
        ms_per_step = 1e3 * dev_s / args.steps
        value = total / (dev_s / args.steps)
        peak, peak_src = measured_peak()
        bytes_per_solve = algorithmic_bytes(N)
        k2_s = float(np.mean(t_k2))
        achieved = bytes_per_solve * B / k2_s / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"nominal NMPC, {TRACK} reftraj, N={N}, batch={B} perturbed x0 per GPU, "
                                   "cold-start RTI step (BASELINE.json configs[1])",
                       "track": TRACK, "N": N, "Tp": TP, "batch_per_gpu": B, "global_batch": total,
                       "parallelism": f"dp{world} (independent instances, no collective)",
                       "l2": "flushed between timed steps (256 MiB write)", "seed": "rank"},
            "roofline": {"bound": "hbm", "kernel": "k_qp_solve", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(),
                         "peak_source": peak_src, "algorithmic_bytes_per_solve": bytes_per_solve,
                         "kernel_ms": 1e3 * k2_s, "linearize_ms": 1e3 * float(np.mean(t_k1)),
                         "note": "path is FP64-issue/latency bound, not HBM bound (SURVEY section 8d); see DESIGN.md"},
            "e2e": {"value": total / (e2e_s / args.steps), "unit": "solves/s",
                    "h2d_bytes_per_step": int(hx0.array.nbytes + hy.array.nbytes),
                    "d2h_bytes_per_step": int(hX.array.nbytes + hU.array.nbytes + dbg.nbytes),
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "two BatchedOcpSolver handles of batch/2 on one device: pinned H2D of (x0, yref), solve, async D2H "
                           "of (x, u, stats) per half; the copies of one half overlap the solve of the other"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "solver": {"qp_iter_mean": float(qp_iter.mean()), "qp_iter_max": int(qp_iter.max()),
                       "frac_interior_point": float((qp_iter > 0).mean()), "status_nonzero": int((status != 0).sum()),
                       "kkt_max": float(kkt.max()), "warm_step_ms": warm_ms,
                       "warm_solves_per_s": total / (warm_ms * 1e-3)},
            "allgather_ms": allgather_ms,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(ocp, x0, yref)
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
