"""GPU parity at the FULL sizes of BASELINE.json configs 3, 4 and 5 (config 2 lives in test_gpu_parity.py).

Each case runs the exact kernel instantiation `bench.py --workload` / `bench_configs.py` time at that size
(`k_qp_solve_warp<0,0>` at N != 30 and B = 8192, the SNMPC head kernel at n_s = 13 / B = 1024 / 128 blocks, the R2
tightening at B = 2048) and checks
  * >= 64 instances spread over the batch against the CPU oracle (1e-6 * max(1, |.|)),
  * size-independent properties over the WHOLE batch: every status 0, QP KKT residual <= 1e-9, x_0 equals the given
    state, and `sharded == whole` bit for bit (a batch solved in two halves equals the batch solved at once: the
    multi-GPU sharding property).
"""
import numpy as np
import pytest

import oracle
from tum_control_b200.config import OcpConfig
from tum_control_b200.planner import synthetic_batch

pytestmark = pytest.mark.gpu

TOL = 1e-6


def _solver(ocp, B):
    from tum_control_b200.solver import BatchedOcpSolver
    return BatchedOcpSolver(ocp, batch=B, device=0, squeeze=False)


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def _run(ocp, x0, yref, sl, steps):
    S = _solver(ocp, sl.stop - sl.start)
    S.set_x0(x0[sl]); S.set_yref_all(yref[sl]); S.init_iterate_from_x0()
    st = [S.solve().copy() for _ in range(steps)]
    X, U = S.get_iterate()
    return S, st, X, U


@pytest.mark.parametrize("N", [20, 40, 80])
def test_config4_horizon_sweep_full_size(N):
    """configs[3]: nominal NMPC, LVMS, B = 8192, N in {20, 40, 80}: the throughput instantiation of the QP kernel at
    horizons other than 30 (smaller / larger per-warp shared memory, 4 / 4 / 2 instances per block)."""
    Tp, B = 0.08 * N, 8192
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("lvms", B, N, Tp, seed=0)
    S, st, X, U = _run(ocp, x0, yref, slice(0, B), 2)
    assert all(np.all(s == 0) for s in st)
    kkt, its = S.get_stats("kkt"), S.get_stats("qp_iter")
    assert np.all(np.isfinite(X)) and np.all(np.isfinite(U)) and np.all(np.isfinite(S.get_cost()))
    assert kkt.max() <= 1e-9 and its.max() < 50
    assert np.abs(X[:, 0, :] - x0).max() <= 1e-9
    _, _, Xa, Ua = _run(ocp, x0, yref, slice(0, B // 2), 2)
    _, _, Xb, Ub = _run(ocp, x0, yref, slice(B // 2, B), 2)
    assert np.array_equal(np.concatenate([Xa, Xb]), X) and np.array_equal(np.concatenate([Ua, Ub]), U)
    # 128 instances spread over the batch against the oracle
    O = oracle.Oracle(ocp)
    idx = np.arange(0, B, 64)
    Xo = np.repeat(x0[idx, None, :], N + 1, axis=1).copy(); Uo = np.zeros((len(idx), N, 2))
    for _ in range(2):
        O.rti_batch(x0[idx], yref[idx], Xo, Uo, nthreads=8)
    assert _rel(X[idx], Xo) <= TOL and _rel(U[idx], Uo) <= TOL


def test_config3_snmpc_full_size():
    """configs[2]: Stochastic NMPC, 13 collocation samples, UPH = 5, Monteblanco, B = 1024: head condensing + head
    kernel at the benchmarked size against the dense full-condensing oracle."""
    N, Tp, B, ns, uph = 30, 2.4, 1024, 13, 5
    ocp = OcpConfig(N=N, Tp=Tp, variant="snmpc", rk_steps=1, n_samples=ns, uph=uph, ipm_tol=1e-13)
    x0, yref = synthetic_batch("monteblanco", B, N, Tp, seed=0)
    steps = 2
    S, st, X, U = _run(ocp, x0, yref, slice(0, B), steps)
    assert all(np.all(s == 0) for s in st)
    assert S.get_stats("kkt").max() <= 1e-9 and S.get_stats("qp_iter").max() < 50
    assert np.all(np.isfinite(X)) and np.all(np.isfinite(U))
    assert np.abs(X[:, 0, :] - x0).max() <= 1e-9
    XS = S.get_samples()
    Sa, _, Xa, Ua = _run(ocp, x0, yref, slice(0, B // 2), steps)
    Sb, _, Xb, Ub = _run(ocp, x0, yref, slice(B // 2, B), steps)
    assert np.array_equal(np.concatenate([Xa, Xb]), X) and np.array_equal(np.concatenate([Ua, Ub]), U)
    assert np.array_equal(np.concatenate([Sa.get_samples(), Sb.get_samples()]), XS)
    # 64 instances against the dense oracle (nominal iterate, inputs, sample iterates)
    O = oracle.SnmpcOracle(ocp)
    assert O.ns == ns and O.H == uph
    n_ipm = 0
    for b in range(0, B, 16):
        Xo, XSo, Uo = O.init_guess(x0[b])
        for _ in range(steps):
            code, stats = O.rti_step(x0[b], yref[b], Xo, XSo, Uo)
            assert code == 0
            n_ipm += stats[1] > 0
        assert _rel(X[b], Xo) <= TOL and _rel(U[b], Uo) <= TOL, b
        assert _rel(XS[b, :uph], XSo[:uph, :ns]) <= TOL, b
    assert n_ipm > 0


def test_config5_r2_full_size():
    """configs[4]: R2NMPC, Modena, N = 30, 2048 instances per GPU: three solves with the fused tightening in between;
    iterates AND the tightened lbx / ubx / uh of 64 instances against the oracle's restatement of
    Reduced_Robustified_NMPC_class.py:287-365."""
    N, Tp, B = 30, 2.4, 2048
    ocp = OcpConfig(N=N, Tp=Tp, variant="r2")
    x0, yref = synthetic_batch("modena", B, N, Tp, seed=0)
    steps = 3
    S, st, X, U = _run(ocp, x0, yref, slice(0, B), steps)
    assert all(np.all(s == 0) for s in st)
    assert S.get_stats("kkt").max() <= 1e-9 and np.all(np.isfinite(X)) and np.all(np.isfinite(U))
    lbx = np.stack([S.constraints_get(k, "lbx")[:, 0] for k in range(1, N)], axis=1)
    ubx = np.stack([S.constraints_get(k, "ubx")[:, 0] for k in range(1, N)], axis=1)
    uh = np.stack([S.constraints_get(k, "uh")[:, 0] for k in range(1, N)], axis=1)
    assert np.all(ubx < ocp.vehicle.delta_f_max) and np.all(lbx > -ocp.vehicle.delta_f_max) and np.all(uh < 1.0)
    assert np.all(S.constraints_get(N, "uh") == 1.0) and np.all(S.constraints_get(N, "ubx") == ocp.vehicle.delta_f_max)
    Sa, _, Xa, Ua = _run(ocp, x0, yref, slice(0, B // 2), steps)
    Sb, _, Xb, Ub = _run(ocp, x0, yref, slice(B // 2, B), steps)
    assert np.array_equal(np.concatenate([Xa, Xb]), X) and np.array_equal(np.concatenate([Ua, Ub]), U)
    uh2 = np.concatenate([np.stack([s.constraints_get(k, "uh")[:, 0] for k in range(1, N)], axis=1) for s in (Sa, Sb)])
    assert np.array_equal(uh2, uh)
    qp = oracle.QpView()
    for b in range(0, B, 32):
        O = oracle.Oracle(ocp)
        Xo, Uo = O.init_guess(x0[b])
        for _ in range(steps):
            code, _ = O.rti_step(x0[b], yref[b], Xo, Uo, qp=qp)
            assert code == 0
            O.r2_tighten(qp.get("A", N).reshape(N, 8, 8).copy(), Xo)
        assert _rel(X[b], Xo) <= TOL and _rel(U[b], Uo) <= TOL, b
        assert np.abs(lbx[b] - O.lb[1:N, 1]).max() <= 1e-9 and np.abs(ubx[b] - O.ub[1:N, 1]).max() <= 1e-9, b
        assert np.abs(uh[b] - O.ub[1:N, 2]).max() <= 1e-8, b


def test_r2_with_zero_uncertainty_is_the_nominal_solver_bit_for_bit():
    """Exact reduction (the reference commits no R2NMPC outputs, so this is the strongest pin there is): with
    Sigma_0 = 0 and W_d = 0 every back-off is sqrt(0) = 0, the tightened bounds equal the nominal bounds exactly, and
    the R2 handle (per-instance bound arrays + tightening kernel) must reproduce the nominal handle BIT FOR BIT."""
    N, Tp, B = 30, 2.4, 256
    x0, yref = synthetic_batch("modena", B, N, Tp, seed=2)
    nom = OcpConfig(N=N, Tp=Tp)
    r2 = OcpConfig(N=N, Tp=Tp, variant="r2", r2_sigma0=np.zeros(8), r2_wd=np.zeros(4))
    Sn, stn, Xn, Un = _run(nom, x0, yref, slice(0, B), 4)
    Sr, str_, Xr, Ur = _run(r2, x0, yref, slice(0, B), 4)
    assert all(np.array_equal(a, b) for a, b in zip(stn, str_))
    assert np.array_equal(Xn, Xr) and np.array_equal(Un, Ur)
    assert np.array_equal(Sn.get_cost(), Sr.get_cost()) and np.array_equal(Sn.get_stats("qp_iter"), Sr.get_stats("qp_iter"))
    assert (Sn.get_stats("qp_iter") > 0).any()
    for k in (1, N // 2, N - 1):
        assert np.all(Sr.constraints_get(k, "ubx") == nom.vehicle.delta_f_max) and np.all(Sr.constraints_get(k, "uh") == 1.0)


@pytest.mark.parametrize("N", [30, 80])
def test_every_instantiation_of_the_qp_kernel_gives_the_same_bits(N):
    """The same instances solved in a small handle and inside a large one: the small batch runs the latency instantiation
    (8 resident warps per SM, short layout of the per-stage shared block, every octet runs the vector sweeps), the large
    one the throughput instantiation (16 warps; at N = 80 the long-horizon layout with kff / gus / ddu in the stage record;
    corrector sweep and forward roll-out on octet 0 only).  All of them are the same arithmetic: bit-identical iterates,
    costs and iteration counts."""
    Tp, Bs, Bl = 0.08 * N, 48, 1536          # 1536 > 8 warps x 148 SMs: the throughput instantiation
    ocp = OcpConfig(N=N, Tp=Tp)
    x0, yref = synthetic_batch("lvms", Bl, N, Tp, seed=3)
    Sl, stl, Xl, Ul = _run(ocp, x0, yref, slice(0, Bl), 3)
    Ss, sts, Xs, Us = _run(ocp, x0, yref, slice(0, Bs), 3)
    assert all(np.array_equal(a, b[:Bs]) for a, b in zip(sts, stl)) and all(np.all(a == 0) for a in sts)
    assert np.array_equal(Xs, Xl[:Bs]) and np.array_equal(Us, Ul[:Bs])
    assert np.array_equal(Ss.get_cost(), Sl.get_cost()[:Bs])
    assert np.array_equal(Ss.get_stats("qp_iter"), Sl.get_stats("qp_iter")[:Bs])
