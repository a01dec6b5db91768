// k_qp_solve.cu -- K2 + K3: per-instance structured QP solve and full step, one thread per instance.
//
// Replaces acados' feedback phase for this OCP: full condensing + HPIPM dense
// interior point + `x += dx, u += du` + get_cost().  The QP optimum is unique (R > 0, Z > 0), so the
// method is free: here a Riccati recursion over the N stages inside a Mehrotra predictor-corrector
// primal-dual interior point iteration, started from the equality-constrained optimum.
//   * every soft row (omega_f bound, delta_f bound, h) contributes a scalar-weighted rank-1 term to
//     its stage Hessian; the slack pair is eliminated in closed form, so stage blocks stay 8+2 wide;
//   * the trivial rows/columns of A_k, B_k are never multiplied;
//   * a dual proximal term caps the barrier weights at 1/ipm_reg and the u-stationarity residual is
//     fed back every iteration, which keeps the recursion accurate to KKT <= 1e-9 in fp64;
//   * if the equality-constrained optimum already satisfies every bound it IS the optimum and the
//     interior-point loop is skipped.
// The algorithm is the one restated in oracle/nmpc_oracle.c (orc_qp_solve); tests compare the two.
//
// Layout: batch-contiguous SoA in global memory (see DevMem); lane i of a warp owns instance 32w+i,
// so every load/store of a warp is one coalesced 256-byte line.
#include "nmpc_kernels.cuh"

namespace b200 {

#define AT(p, idx) (p)[(size_t)(idx) * B + b]

__device__ __forceinline__ double stage_val2(const double *shared, const double *inst, int idx, int b, int B)
{
    return inst ? inst[(size_t)idx * B + b] : shared[idx];
}

// ---- structured products with A_k = [e0 e1 | A36 ; rows 6,7 = e6', e7'], B_k = [B12 ; [0 dt] ; [dt 0]]
__device__ __forceinline__ void A_mul(const double *A, const double *x, double *y)
{
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 6; c++) s += A[r * 6 + c] * x[2 + c];
        y[r] = s;
    }
    y[0] += x[0]; y[1] += x[1]; y[6] = x[6]; y[7] = x[7];
}
__device__ __forceinline__ void B_mul_add(const double *Bm, double dt, const double *u, double *y)
{
#pragma unroll
    for (int r = 0; r < 6; r++) y[r] += Bm[r * 2] * u[0] + Bm[r * 2 + 1] * u[1];
    y[6] += dt * u[1]; y[7] += dt * u[0];
}
__device__ __forceinline__ void At_mul(const double *A, const double *v, double *y)
{
    y[0] = v[0]; y[1] = v[1];
#pragma unroll
    for (int c = 0; c < 6; c++) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < 6; r++) s += A[r * 6 + c] * v[r];
        y[2 + c] = s;
    }
    y[6] += v[6]; y[7] += v[7];
}
__device__ __forceinline__ void Bt_mul(const double *Bm, double dt, const double *v, double *y)
{
    double s0 = dt * v[7], s1 = dt * v[6];
#pragma unroll
    for (int r = 0; r < 6; r++) { s0 += Bm[r * 2] * v[r]; s1 += Bm[r * 2 + 1] * v[r]; }
    y[0] = s0; y[1] = s1;
}

// soft rows of one stage
struct StageCon {
    double v0[3];          // row value at the linearisation point: [u1, x6, h]
    double lo[3], up[3], zl[3], zu[3], Zl[3], Zu[3];
    double g4[4];          // grad h w.r.t. (vl, vt, r, a) -> dx indices (3, 4, 5, 7)
    bool act[3];
};

__device__ __forceinline__ void load_con(const DevCfg &cfg, const DevMem &m, int k, int b, StageCon &c)
{
    const int B = cfg.B, N = cfg.N;
    const double sig = (k < N) ? cfg.dt : 1.0;
    c.act[0] = (k < N); c.act[1] = (k >= 1); c.act[2] = (k >= 1);
    c.v0[0] = (k < N) ? AT(m.U, k * 2 + 1) : 0.0;
    c.v0[1] = AT(m.X, k * 8 + 6);
    c.v0[2] = AT(m.hval, k);
#pragma unroll
    for (int i = 0; i < 4; i++) c.g4[i] = AT(m.gh, k * 4 + i);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        c.lo[j] = stage_val2(m.lb_s, m.lb_i, k * 3 + j, b, B);
        c.up[j] = stage_val2(m.ub_s, m.ub_i, k * 3 + j, b, B);
        c.zl[j] = sig * stage_val2(m.zl_s, m.zl_i, k * 3 + j, b, B);
        c.zu[j] = sig * stage_val2(m.zu_s, m.zu_i, k * 3 + j, b, B);
        c.Zl[j] = sig * stage_val2(m.Zl_s, m.Zl_i, k * 3 + j, b, B);
        c.Zu[j] = sig * stage_val2(m.Zu_s, m.Zu_i, k * 3 + j, b, B);
        // a soft row without any penalty is vacuous
        if (!((c.zl[j] + c.Zl[j] > 0.0) && (c.zu[j] + c.Zu[j] > 0.0))) c.act[j] = false;
    }
}

__device__ __forceinline__ double row_lin(const StageCon &c, int j, const double *dx, const double *du)
{
    if (j == 0) return du[1];
    if (j == 1) return dx[6];
    return c.g4[0] * dx[3] + c.g4[1] * dx[4] + c.g4[2] * dx[5] + c.g4[3] * dx[7];
}

// gradient accumulation: qx += coef * g_x, ru += coef * g_u
__device__ __forceinline__ void row_add_grad(const StageCon &c, int j, double coef, double *qx, double *ru)
{
    if (j == 0) ru[1] += coef;
    else if (j == 1) qx[6] += coef;
    else { qx[3] += coef * c.g4[0]; qx[4] += coef * c.g4[1]; qx[5] += coef * c.g4[2]; qx[7] += coef * c.g4[3]; }
}

// Elimination of one soft side (see the derivation in oracle/nmpc_oracle.c above orc_qp_solve)
struct SideK { double th, Dp, rho, cc, om; };
__device__ __forceinline__ SideK side_coeffs(double reg, double Z, double s, double t, double lam, double mu,
                                            double rt, double rs, double r1, double r2)
{
    SideK o;
    const double tt = t + reg * lam;
    o.th = lam / tt;
    const double ms = mu / s;
    o.Dp = Z + o.th + ms;
    o.rho = -rs + r1 / tt + r2 / s + o.th * rt;
    o.cc = r1 / tt + o.th * rt - o.th * o.rho / o.Dp;
    o.om = o.th * (Z + ms) / o.Dp;
    return o;
}

// One backward Riccati step.  In: P,p of stage k+1 (overwritten with stage k), stage matrices, modified
// Hessian (diagonal qd[8] + om_h g4 g4' on (3,4,5,7) + optional (vl,vt) block Qv), gradient (qt, rt2),
// R diagonal rd[2].  bvec != nullptr adds the affine term P b.  Stores K, Hi, kff.
template <bool AFFINE, bool NEED_P>
__device__ __forceinline__ void riccati_step(const DevCfg &cfg, const DevMem &m, int k, int b, double P[8][8],
                                             double *p, const double *A, const double *Bm, const double *qd,
                                             double om_h, const double *g4, const double *Qv, const double *qt,
                                             const double *rd, const double *rt2)
{
    const int B = cfg.B;
    const double dt = cfg.dt;
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = p[i];
    if (AFFINE) {
        double bv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) bv[i] = AT(m.bd, k * 8 + i);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < 8; l++) s += P[i][l] * bv[l];
            v[i] += s;
        }
    }
    // PB, Huu
    double PB[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double s0 = dt * P[i][7], s1 = dt * P[i][6];
#pragma unroll
        for (int r = 0; r < 6; r++) { s0 += P[i][r] * Bm[r * 2]; s1 += P[i][r] * Bm[r * 2 + 1]; }
        PB[i][0] = s0; PB[i][1] = s1;
    }
    double c0[8], c1[8], t2[2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c0[i] = PB[i][0]; c1[i] = PB[i][1]; }
    Bt_mul(Bm, dt, c0, t2);
    double h00 = rd[0] + t2[0], h01a = t2[1];
    Bt_mul(Bm, dt, c1, t2);
    double h01 = 0.5 * (h01a + t2[0]), h11 = rd[1] + t2[1];
    // 2x2 Cholesky, Huu^-1 = L^-T L^-1
    const double l11 = sqrt(h00), l21 = h01 / l11;
    const double l22 = sqrt(h11 - l21 * l21);
    const double i11 = 1.0 / l11, i22 = 1.0 / l22;
    const double m21 = -l21 * i11 * i22;
    const double Hi0 = i11 * i11 + m21 * m21, Hi1 = m21 * i22, Hi3 = i22 * i22;
    AT(m.Hi, k * 3 + 0) = Hi0; AT(m.Hi, k * 3 + 1) = Hi1; AT(m.Hi, k * 3 + 2) = Hi3;
    // gradient part
    double hu[2], hx[8];
    Bt_mul(Bm, dt, v, hu);
    hu[0] += rt2[0]; hu[1] += rt2[1];
    const double kf0 = -(Hi0 * hu[0] + Hi1 * hu[1]), kf1 = -(Hi1 * hu[0] + Hi3 * hu[1]);
    AT(m.kff, k * 2 + 0) = kf0; AT(m.kff, k * 2 + 1) = kf1;
    if (!NEED_P) return;
    At_mul(A, v, hx);
#pragma unroll
    for (int i = 0; i < 8; i++) hx[i] += qt[i];
    // column by column: PA[:,j], Hux[:,j] = B' PA[:,j], Hxx[:,j] = A' PA[:,j]
    double Pn[8][8], Hux[2][8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        double col[8];
        if (j < 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) col[i] = P[i][j];
        } else {
            const int c = j - 2;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                double s = (c == 4) ? P[i][6] : ((c == 5) ? P[i][7] : 0.0);
#pragma unroll
                for (int r = 0; r < 6; r++) s += P[i][r] * A[r * 6 + c];
                col[i] = s;
            }
        }
        double hxj[8], huj[2];
        Bt_mul(Bm, dt, col, huj);
        At_mul(A, col, hxj);
        Hux[0][j] = huj[0]; Hux[1][j] = huj[1];
#pragma unroll
        for (int i = 0; i < 8; i++) Pn[i][j] = hxj[i];
    }
    // Y = L^-1 Hux, K = -L^-T Y, P_k = Q~ + Hxx - Y'Y
    double Y0[8], Y1[8], K0[8], K1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        Y0[j] = Hux[0][j] * i11;
        Y1[j] = (Hux[1][j] - l21 * Y0[j]) * i22;
        K1[j] = -Y1[j] * i22;
        K0[j] = (-Y0[j] - l21 * K1[j]) * i11;
        AT(m.K, k * 16 + j) = K0[j];
        AT(m.K, k * 16 + 8 + j) = K1[j];
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = i; j < 8; j++) {
            double s = 0.5 * (Pn[i][j] + Pn[j][i]) - (Y0[i] * Y0[j] + Y1[i] * Y1[j]);
            P[i][j] = s; P[j][i] = s;
        }
#pragma unroll
    for (int i = 0; i < 8; i++) P[i][i] += qd[i];
    {
        const int ix[4] = {3, 4, 5, 7};
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) P[ix[a]][ix[c]] += om_h * g4[a] * g4[c];
    }
    if (Qv) { P[3][3] += Qv[0]; P[3][4] += Qv[1]; P[4][3] += Qv[1]; P[4][4] += Qv[2]; }
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = hx[i] + K0[i] * hu[0] + K1[i] * hu[1];
}

// terminal value function P_N = Q~_N, p_N = q~_N
__device__ __forceinline__ void riccati_terminal(double P[8][8], double *p, const double *qd, double om_h,
                                                 const double *g4, const double *Qv, const double *qt)
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) P[i][j] = 0.0;
        P[i][i] = qd[i];
        p[i] = qt[i];
    }
    const int ix[4] = {3, 4, 5, 7};
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) P[ix[a]][ix[c]] += om_h * g4[a] * g4[c];
    if (Qv) { P[3][3] += Qv[0]; P[3][4] += Qv[1]; P[4][3] += Qv[1]; P[4][4] += Qv[2]; }
}

__device__ __forceinline__ void load_AB(const DevMem &m, int k, int b, int B, double *A, double *Bm)
{
#pragma unroll
    for (int i = 0; i < 36; i++) A[i] = AT(m.A36, k * 36 + i);
#pragma unroll
    for (int i = 0; i < 12; i++) Bm[i] = AT(m.B12, k * 12 + i);
}

// stage cost Hessian diagonal (dt-scaled) of the Gauss-Newton NLS cost
__device__ __forceinline__ void load_cost_diag(const DevCfg &cfg, const DevMem &m, int k, int b, double *qd, double *rd)
{
    const int B = cfg.B;
    const double sig = (k < cfg.N) ? cfg.dt : 1.0;
#pragma unroll
    for (int i = 0; i < 8; i++) qd[i] = 0.0;
    qd[0] = sig * stage_val2(m.W_s, m.W_i, k * 6 + 0, b, B);
    qd[1] = sig * stage_val2(m.W_s, m.W_i, k * 6 + 1, b, B);
    qd[2] = sig * stage_val2(m.W_s, m.W_i, k * 6 + 2, b, B);
    if (!cfg.cost_vel_abs) qd[3] = sig * stage_val2(m.W_s, m.W_i, k * 6 + 3, b, B);
    rd[0] = (k < cfg.N) ? sig * stage_val2(m.W_s, m.W_i, k * 6 + 4, b, B) : 0.0;
    rd[1] = (k < cfg.N) ? sig * stage_val2(m.W_s, m.W_i, k * 6 + 5, b, B) : 0.0;
}

#define IPX(k, c, sd) (((k) * 3 + (c)) * 2 + (sd))

// Backward sweep of one interior-point iteration: costates / u-stationarity residual and, if FULL, the
// barrier Hessian, the Riccati factorisation and the predictor right-hand side.  Returns max |r_u|.
template <bool FULL>
__device__ double sweep_backward(const DevCfg &cfg, const DevMem &m, int b)
{
    const int B = cfg.B, N = cfg.N;
    const double reg = cfg.ipm_reg;
    double P[8][8], p[8], pi[8];
    double rstat = 0.0;
    for (int k = N; k >= 0; k--) {
        StageCon sc;
        load_con(cfg, m, k, b, sc);
        double dx[8], du[2] = {0.0, 0.0}, qd[8], rd[2];
#pragma unroll
        for (int i = 0; i < 8; i++) dx[i] = AT(m.dx, k * 8 + i);
        if (k < N) { du[0] = AT(m.du, k * 2); du[1] = AT(m.du, k * 2 + 1); }
        load_cost_diag(cfg, m, k, b, qd, rd);
        double Qv[3] = {0.0, 0.0, 0.0};
        if (cfg.cost_vel_abs) { Qv[0] = AT(m.Qv, k * 3); Qv[1] = AT(m.Qv, k * 3 + 1); Qv[2] = AT(m.Qv, k * 3 + 2); }
        // stationarity rows at the current iterate
        double gx[8], gu[2] = {0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 8; i++) gx[i] = qd[i] * dx[i];
#pragma unroll
        for (int i = 0; i < 5; i++) gx[i] += AT(m.qx, k * 5 + i);
        if (cfg.cost_vel_abs) { gx[3] += Qv[0] * dx[3] + Qv[1] * dx[4]; gx[4] += Qv[1] * dx[3] + Qv[2] * dx[4]; }
        if (k < N) {
            gu[0] = AT(m.rg, k * 2) + rd[0] * du[0];
            gu[1] = AT(m.rg, k * 2 + 1) + rd[1] * du[1];
        }
        double qt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, rt2[2] = {0.0, 0.0};
        double om_h = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (!sc.act[j]) continue;
            const double v = sc.v0[j] + row_lin(sc, j, dx, du);
            double om = 0.0, coef = 0.0;
#pragma unroll
            for (int sd = 0; sd < 2; sd++) {
                const double sgn = sd ? -1.0 : 1.0;
                const double z = sd ? sc.zu[j] : sc.zl[j], Z = sd ? sc.Zu[j] : sc.Zl[j];
                const double s = AT(m.ips, IPX(k, j, sd)), t = AT(m.ipt, IPX(k, j, sd));
                const double lam = AT(m.iplam, IPX(k, j, sd)), mu = AT(m.ipmu, IPX(k, j, sd));
                row_add_grad(sc, j, -sgn * lam, gx, gu);
                if (FULL) {
                    const double rt = t - ((sd ? sc.up[j] - v : v - sc.lo[j]) + s);
                    const double rs = z + Z * s - lam - mu;
                    const SideK sk = side_coeffs(reg, Z, s, t, lam, mu, rt, rs, -lam * t, -mu * s);
                    om += sk.om;
                    coef -= sgn * sk.cc;
                }
            }
            if (FULL) {
                if (j == 0) rd[1] += om;
                else if (j == 1) qd[6] += om;
                else om_h = om;
                row_add_grad(sc, j, coef, qt, rt2);
            }
        }
        double A[36], Bm[12];
        if (k < N) {
            load_AB(m, k, b, B, A, Bm);
            double t8[8], t2[2];
            At_mul(A, pi, t8);
#pragma unroll
            for (int i = 0; i < 8; i++) gx[i] += t8[i];
            Bt_mul(Bm, cfg.dt, pi, t2);
            gu[0] += t2[0]; gu[1] += t2[1];
            AT(m.gus, k * 2) = gu[0]; AT(m.gus, k * 2 + 1) = gu[1];
            rstat = fmax(rstat, fmax(fabs(gu[0]), fabs(gu[1])));
            if (!(gu[0] == gu[0]) || !(gu[1] == gu[1])) rstat = nan("");
        }
#pragma unroll
        for (int i = 0; i < 8; i++) pi[i] = gx[i];
        if (FULL) {
            if (k == N) {
                riccati_terminal(P, p, qd, om_h, sc.g4, cfg.cost_vel_abs ? Qv : nullptr, qt);
            } else {
                rt2[0] += gu[0]; rt2[1] += gu[1];
                if (k > 0)
                    riccati_step<false, true>(cfg, m, k, b, P, p, A, Bm, qd, om_h, sc.g4,
                                              cfg.cost_vel_abs ? Qv : nullptr, qt, rd, rt2);
                else
                    riccati_step<false, false>(cfg, m, k, b, P, p, A, Bm, qd, om_h, sc.g4,
                                               cfg.cost_vel_abs ? Qv : nullptr, qt, rd, rt2);
            }
        }
    }
    return rstat;
}

// Backward vector sweep of the corrector: same factorisation, new right-hand side.
__device__ void sweep_backward_corrector(const DevCfg &cfg, const DevMem &m, int b, double sigma_mu)
{
    const int B = cfg.B, N = cfg.N;
    const double reg = cfg.ipm_reg;
    double p[8];
    for (int k = N; k >= 0; k--) {
        StageCon sc;
        load_con(cfg, m, k, b, sc);
        double dx[8], du[2] = {0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 8; i++) dx[i] = AT(m.dx, k * 8 + i);
        if (k < N) { du[0] = AT(m.du, k * 2); du[1] = AT(m.du, k * 2 + 1); }
        double qt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, rt2[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (!sc.act[j]) continue;
            const double v = sc.v0[j] + row_lin(sc, j, dx, du);
            double coef = 0.0;
#pragma unroll
            for (int sd = 0; sd < 2; sd++) {
                const double sgn = sd ? -1.0 : 1.0;
                const double z = sd ? sc.zu[j] : sc.zl[j], Z = sd ? sc.Zu[j] : sc.Zl[j];
                const double s = AT(m.ips, IPX(k, j, sd)), t = AT(m.ipt, IPX(k, j, sd));
                const double lam = AT(m.iplam, IPX(k, j, sd)), mu = AT(m.ipmu, IPX(k, j, sd));
                const double rt = t - ((sd ? sc.up[j] - v : v - sc.lo[j]) + s);
                const double rs = z + Z * s - lam - mu;
                const double r1 = -lam * t + sigma_mu - AT(m.pr1, IPX(k, j, sd));
                const double r2 = -mu * s + sigma_mu - AT(m.pr2, IPX(k, j, sd));
                const SideK sk = side_coeffs(reg, Z, s, t, lam, mu, rt, rs, r1, r2);
                coef -= sgn * sk.cc;
            }
            row_add_grad(sc, j, coef, qt, rt2);
        }
        if (k == N) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = qt[i];
        } else {
            double A[36], Bm[12];
            load_AB(m, k, b, B, A, Bm);
            rt2[0] += AT(m.gus, k * 2); rt2[1] += AT(m.gus, k * 2 + 1);
            double hu[2], hx[8];
            Bt_mul(Bm, cfg.dt, p, hu);
            hu[0] += rt2[0]; hu[1] += rt2[1];
            const double Hi0 = AT(m.Hi, k * 3), Hi1 = AT(m.Hi, k * 3 + 1), Hi3 = AT(m.Hi, k * 3 + 2);
            AT(m.kff, k * 2 + 0) = -(Hi0 * hu[0] + Hi1 * hu[1]);
            AT(m.kff, k * 2 + 1) = -(Hi1 * hu[0] + Hi3 * hu[1]);
            if (k > 0) {
                At_mul(A, p, hx);
#pragma unroll
                for (int i = 0; i < 8; i++)
                    p[i] = hx[i] + qt[i] + AT(m.K, k * 16 + i) * hu[0] + AT(m.K, k * 16 + 8 + i) * hu[1];
            }
        }
    }
}

// Forward sweep: Newton direction in (x,u), recovery of the slack / multiplier directions, step to the
// boundary.  PASS 0 (predictor) accumulates the complementarity polynomial and stores the second-order
// products; PASS 1 (corrector) stores the directions.  Returns alpha_max.
template <int PASS>
__device__ double sweep_forward(const DevCfg &cfg, const DevMem &m, int b, double sigma_mu, double *poly)
{
    const int B = cfg.B, N = cfg.N;
    const double reg = cfg.ipm_reg;
    double ddx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double amax = 1e300;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int k = 0; k <= N; k++) {
        StageCon sc;
        load_con(cfg, m, k, b, sc);
        double dx[8], du[2] = {0.0, 0.0}, ddu[2] = {0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 8; i++) dx[i] = AT(m.dx, k * 8 + i);
        if (k < N) {
            du[0] = AT(m.du, k * 2); du[1] = AT(m.du, k * 2 + 1);
            ddu[0] = AT(m.kff, k * 2); ddu[1] = AT(m.kff, k * 2 + 1);
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    ddu[0] += AT(m.K, k * 16 + i) * ddx[i];
                    ddu[1] += AT(m.K, k * 16 + 8 + i) * ddx[i];
                }
            }
            if (PASS == 1) { AT(m.ddu, k * 2) = ddu[0]; AT(m.ddu, k * 2 + 1) = ddu[1]; }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (!sc.act[j]) continue;
            const double v = sc.v0[j] + row_lin(sc, j, dx, du);
            const double dv = row_lin(sc, j, ddx, ddu);
#pragma unroll
            for (int sd = 0; sd < 2; sd++) {
                const double sgn = sd ? -1.0 : 1.0;
                const double z = sd ? sc.zu[j] : sc.zl[j], Z = sd ? sc.Zu[j] : sc.Zl[j];
                const double s = AT(m.ips, IPX(k, j, sd)), t = AT(m.ipt, IPX(k, j, sd));
                const double lam = AT(m.iplam, IPX(k, j, sd)), mu = AT(m.ipmu, IPX(k, j, sd));
                const double rt = t - ((sd ? sc.up[j] - v : v - sc.lo[j]) + s);
                const double rs = z + Z * s - lam - mu;
                double r1 = -lam * t, r2 = -mu * s;
                if (PASS == 1) {
                    r1 += sigma_mu - AT(m.pr1, IPX(k, j, sd));
                    r2 += sigma_mu - AT(m.pr2, IPX(k, j, sd));
                }
                const SideK sk = side_coeffs(reg, Z, s, t, lam, mu, rt, rs, r1, r2);
                const double tt = t + reg * lam;
                const double ds = (sk.rho - sk.th * sgn * dv) / sk.Dp;
                const double dl = (r1 - lam * (sgn * dv + ds - rt)) / tt;
                const double dtt = sgn * dv + ds - rt + reg * dl;
                const double dm = (r2 - mu * ds) / s;
                if (ds < 0.0) amax = fmin(amax, -s / ds);
                if (dl < 0.0) amax = fmin(amax, -lam / dl);
                if (dtt < 0.0) amax = fmin(amax, -t / dtt);
                if (dm < 0.0) amax = fmin(amax, -mu / dm);
                if (PASS == 0) {
                    AT(m.pr1, IPX(k, j, sd)) = dl * dtt;
                    AT(m.pr2, IPX(k, j, sd)) = dm * ds;
                    c0 += lam * t + mu * s;
                    c1 += lam * dtt + t * dl + mu * ds + s * dm;
                    c2 += dl * dtt + dm * ds;
                } else {
                    AT(m.ids, IPX(k, j, sd)) = ds; AT(m.idlam, IPX(k, j, sd)) = dl;
                    AT(m.idt, IPX(k, j, sd)) = dtt; AT(m.idmu, IPX(k, j, sd)) = dm;
                }
            }
        }
        if (k < N) {
            double A[36], Bm[12], nx[8];
            load_AB(m, k, b, B, A, Bm);
            A_mul(A, ddx, nx);
            B_mul_add(Bm, cfg.dt, ddu, nx);
#pragma unroll
            for (int i = 0; i < 8; i++) ddx[i] = nx[i];
        }
    }
    if (PASS == 0) { poly[0] = c0; poly[1] = c1; poly[2] = c2; }
    return amax;
}

// Step of length alpha, re-roll of the states through the linearised dynamics, new residual measures.
template <bool MEASURE_ONLY>
__device__ void sweep_update(const DevCfg &cfg, const DevMem &m, int b, double alpha, int nact, double *gap,
                             double *cmax, double *rmax)
{
    const int B = cfg.B, N = cfg.N;
    double dx[8];
#pragma unroll
    for (int i = 0; i < 8; i++) dx[i] = AT(m.dx, i);
    double g = 0.0, cm = 0.0, rm = 0.0;
    for (int k = 0; k <= N; k++) {
        StageCon sc;
        load_con(cfg, m, k, b, sc);
        double du[2] = {0.0, 0.0};
        if (k < N) {
            du[0] = AT(m.du, k * 2); du[1] = AT(m.du, k * 2 + 1);
            if (!MEASURE_ONLY) {
                du[0] += alpha * AT(m.ddu, k * 2); du[1] += alpha * AT(m.ddu, k * 2 + 1);
                AT(m.du, k * 2) = du[0]; AT(m.du, k * 2 + 1) = du[1];
            }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (!sc.act[j]) continue;
            const double v = sc.v0[j] + row_lin(sc, j, dx, du);
#pragma unroll
            for (int sd = 0; sd < 2; sd++) {
                const double z = sd ? sc.zu[j] : sc.zl[j], Z = sd ? sc.Zu[j] : sc.Zl[j];
                double s = AT(m.ips, IPX(k, j, sd)), t = AT(m.ipt, IPX(k, j, sd));
                double lam = AT(m.iplam, IPX(k, j, sd)), mu = AT(m.ipmu, IPX(k, j, sd));
                if (!MEASURE_ONLY) {
                    s += alpha * AT(m.ids, IPX(k, j, sd)); t += alpha * AT(m.idt, IPX(k, j, sd));
                    lam += alpha * AT(m.idlam, IPX(k, j, sd)); mu += alpha * AT(m.idmu, IPX(k, j, sd));
                    AT(m.ips, IPX(k, j, sd)) = s; AT(m.ipt, IPX(k, j, sd)) = t;
                    AT(m.iplam, IPX(k, j, sd)) = lam; AT(m.ipmu, IPX(k, j, sd)) = mu;
                }
                const double p1 = lam * t, p2 = mu * s;
                g += p1 + p2;
                cm = fmax(cm, fmax(p1, p2));
                if (!(p1 == p1) || !(p2 == p2)) cm = nan("");
                const double rt = t - ((sd ? sc.up[j] - v : v - sc.lo[j]) + s);
                const double rs = z + Z * s - lam - mu;
                rm = fmax(rm, fmax(fabs(rt), fabs(rs)));
            }
        }
        if (k < N) {
            double A[36], Bm[12], nx[8];
            load_AB(m, k, b, B, A, Bm);
            A_mul(A, dx, nx);
            B_mul_add(Bm, cfg.dt, du, nx);
#pragma unroll
            for (int i = 0; i < 8; i++) { dx[i] = nx[i] + AT(m.bd, k * 8 + i); AT(m.dx, (k + 1) * 8 + i) = dx[i]; }
        }
    }
    *gap = g / (4.0 * nact); *cmax = cm; *rmax = rm;
}

__global__ void __launch_bounds__(32) k_qp_solve(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int B = cfg.B, N = cfg.N;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double tol = cfg.ipm_tol;

    // ---------------- phase 0: equality-constrained optimum (affine Riccati solve)
    {
        double P[8][8], p[8];
        for (int k = N; k >= 0; k--) {
            double qd[8], rd[2], qt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, g4z[4] = {0, 0, 0, 0};
            load_cost_diag(cfg, m, k, b, qd, rd);
#pragma unroll
            for (int i = 0; i < 5; i++) qt[i] = AT(m.qx, k * 5 + i);
            double Qv[3] = {0.0, 0.0, 0.0};
            if (cfg.cost_vel_abs) { Qv[0] = AT(m.Qv, k * 3); Qv[1] = AT(m.Qv, k * 3 + 1); Qv[2] = AT(m.Qv, k * 3 + 2); }
            if (k == N) {
                riccati_terminal(P, p, qd, 0.0, g4z, cfg.cost_vel_abs ? Qv : nullptr, qt);
            } else {
                double A[36], Bm[12], rt2[2] = {AT(m.rg, k * 2), AT(m.rg, k * 2 + 1)};
                load_AB(m, k, b, B, A, Bm);
                riccati_step<true, true>(cfg, m, k, b, P, p, A, Bm, qd, 0.0, g4z, cfg.cost_vel_abs ? Qv : nullptr, qt,
                                         rd, rt2);
            }
        }
    }
    int nact = 0;
    bool feasible = true;
    {
        double dx[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { dx[i] = AT(m.x0, i) - AT(m.X, i); AT(m.dx, i) = dx[i]; }
        for (int k = 0; k <= N; k++) {
            StageCon sc;
            load_con(cfg, m, k, b, sc);
            double du[2] = {0.0, 0.0};
            if (k < N) {
                du[0] = AT(m.kff, k * 2); du[1] = AT(m.kff, k * 2 + 1);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    du[0] += AT(m.K, k * 16 + i) * dx[i];
                    du[1] += AT(m.K, k * 16 + 8 + i) * dx[i];
                }
                AT(m.du, k * 2) = du[0]; AT(m.du, k * 2 + 1) = du[1];
            }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                AT(m.sl, k * 3 + j) = 0.0; AT(m.su, k * 3 + j) = 0.0;
                if (!sc.act[j]) continue;
                nact++;
                const double v = sc.v0[j] + row_lin(sc, j, dx, du);
                if (v < sc.lo[j] || v > sc.up[j]) feasible = false;
                // interior start with zero linear residuals: s_l = s_u = S, lam_l = lam_u
                const double viol = fmax(fmax(sc.lo[j] - v, v - sc.up[j]), 0.0);
                const double S = 1.0 + viol;
                const double gl = sc.zl[j] + sc.Zl[j] * S, gu = sc.zu[j] + sc.Zu[j] * S;
                const double l0 = 0.5 * fmin(gl, gu);
                AT(m.ips, IPX(k, j, 0)) = S; AT(m.ips, IPX(k, j, 1)) = S;
                AT(m.iplam, IPX(k, j, 0)) = l0; AT(m.iplam, IPX(k, j, 1)) = l0;
                AT(m.ipt, IPX(k, j, 0)) = v - sc.lo[j] + S; AT(m.ipt, IPX(k, j, 1)) = sc.up[j] - v + S;
                AT(m.ipmu, IPX(k, j, 0)) = gl - l0; AT(m.ipmu, IPX(k, j, 1)) = gu - l0;
            }
            if (k < N) {
                double A[36], Bm[12], nx[8];
                load_AB(m, k, b, B, A, Bm);
                A_mul(A, dx, nx);
                B_mul_add(Bm, cfg.dt, du, nx);
#pragma unroll
                for (int i = 0; i < 8; i++) { dx[i] = nx[i] + AT(m.bd, k * 8 + i); AT(m.dx, (k + 1) * 8 + i) = dx[i]; }
            }
        }
    }

    int iters = 0, qp_status = 0;
    double kkt = 0.0;
    if (!((feasible && !cfg.qp_no_shortcut) || nact == 0)) {
        // ---------------- interior-point iterations
        double gap = 0.0, cmax = 0.0, rmax = 0.0;
        sweep_update<true>(cfg, m, b, 0.0, nact, &gap, &cmax, &rmax);     // measures only
        bool converged = false;
        int it;
        for (it = 0; it <= cfg.ipm_max_iter; it++) {
            if (!(cmax == cmax) || !(rmax == rmax)) { qp_status = 3; break; }
            double rstat = -1.0;
            if (cmax <= tol && rmax <= tol) {
                rstat = sweep_backward<false>(cfg, m, b);
                if (!(rstat == rstat)) { qp_status = 3; break; }
                if (rstat <= tol) { converged = true; kkt = fmax(fmax(cmax, rmax), rstat); break; }
            }
            if (it == cfg.ipm_max_iter) { kkt = fmax(cmax, rmax); break; }
            rstat = sweep_backward<true>(cfg, m, b);
            if (!(rstat == rstat)) { qp_status = 3; break; }
            double poly[3];
            const double amax_aff = sweep_forward<0>(cfg, m, b, 0.0, poly);
            const double a = fmin(1.0, amax_aff);
            const double gap_aff = (poly[0] + a * poly[1] + a * a * poly[2]) / (4.0 * nact);
            const double sg = gap_aff / gap;
            const double sigma_mu = fmax(sg * sg * sg * gap, 0.1 * tol);
            sweep_backward_corrector(cfg, m, b, sigma_mu);
            const double amax = sweep_forward<1>(cfg, m, b, sigma_mu, poly);
            const double eta = fmax(0.995, 1.0 - gap);
            const double alpha = fmin(1.0, eta * amax);
            sweep_update<false>(cfg, m, b, alpha, nact, &gap, &cmax, &rmax);
        }
        iters = it;
        if (!converged && qp_status == 0) qp_status = 1;
        for (int k = 0; k <= N; k++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const bool act = (j == 0) ? (k < N) : (k >= 1);
                if (!act) continue;
                AT(m.sl, k * 3 + j) = AT(m.ips, IPX(k, j, 0));
                AT(m.su, k * 3 + j) = AT(m.ips, IPX(k, j, 1));
            }
    }

    // ---------------- K3: full step (nlp_solver_step_length 1.0) and cost at the new iterate
    bool bad = (qp_status == 3);
    if (!bad) {
        for (int k = 0; k <= N && !bad; k++) {
#pragma unroll
            for (int i = 0; i < 8; i++) { const double d = AT(m.dx, k * 8 + i); if (!(d == d)) bad = true; }
            if (k < N) { const double d0 = AT(m.du, k * 2), d1 = AT(m.du, k * 2 + 1); if (!(d0 == d0) || !(d1 == d1)) bad = true; }
        }
    }
    double J = 0.0;
    for (int k = 0; k <= N; k++) {
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[i] = AT(m.X, k * 8 + i);
            if (!bad) { x[i] += AT(m.dx, k * 8 + i); AT(m.X, k * 8 + i) = x[i]; }
        }
        const double sig = (k < N) ? cfg.dt : 1.0;
        const double vv = cfg.cost_vel_abs ? sqrt(x[3] * x[3] + x[4] * x[4]) : x[3];
        const double e0 = x[0] - AT(m.yref, k * 6 + 0), e1 = x[1] - AT(m.yref, k * 6 + 1);
        const double e2 = wrap_yaw(x[2]) - AT(m.yref, k * 6 + 2), e3 = vv - AT(m.yref, k * 6 + 3);
        double c = stage_val2(m.W_s, m.W_i, k * 6 + 0, b, B) * e0 * e0 + stage_val2(m.W_s, m.W_i, k * 6 + 1, b, B) * e1 * e1 +
                   stage_val2(m.W_s, m.W_i, k * 6 + 2, b, B) * e2 * e2 + stage_val2(m.W_s, m.W_i, k * 6 + 3, b, B) * e3 * e3;
        if (k < N) {
            double u0 = AT(m.U, k * 2), u1 = AT(m.U, k * 2 + 1);
            if (!bad) {
                u0 += AT(m.du, k * 2); u1 += AT(m.du, k * 2 + 1);
                AT(m.U, k * 2) = u0; AT(m.U, k * 2 + 1) = u1;
            }
            const double e4 = u0 - AT(m.yref, k * 6 + 4), e5 = u1 - AT(m.yref, k * 6 + 5);
            c += stage_val2(m.W_s, m.W_i, k * 6 + 4, b, B) * e4 * e4 + stage_val2(m.W_s, m.W_i, k * 6 + 5, b, B) * e5 * e5;
        }
        c *= 0.5;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const bool act = (j == 0) ? (k < N) : (k >= 1);
            if (!act) continue;
            const double sl = AT(m.sl, k * 3 + j), su = AT(m.su, k * 3 + j);
            c += stage_val2(m.zl_s, m.zl_i, k * 3 + j, b, B) * sl + 0.5 * stage_val2(m.Zl_s, m.Zl_i, k * 3 + j, b, B) * sl * sl +
                 stage_val2(m.zu_s, m.zu_i, k * 3 + j, b, B) * su + 0.5 * stage_val2(m.Zu_s, m.Zu_i, k * 3 + j, b, B) * su * su;
        }
        J += sig * c;
    }
    m.cost[b] = J;
    m.kkt[b] = kkt;
    m.qp_iter[b] = iters;
    m.status[b] = bad ? 4 : 0;      // ACADOS_QP_FAILURE; RTI keeps status 0 when the QP hits max-iter (SURVEY B-9)
}

void launch_qp_solve(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    const int threads = 32;      // one warp per block: B/32 blocks spread over all SMs
    k_qp_solve<<<(cfg.B + threads - 1) / threads, threads, 0, st>>>(cfg, m);
}

}  // namespace b200
