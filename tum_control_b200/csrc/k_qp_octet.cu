// k_qp_octet.cu -- K2 + K3, second generation: EIGHT lanes ("an octet") cooperate on one instance,
// four instances per warp.
//
// Why: the per-instance QP solve is a chain of ~12 interior-point iterations x ~5 sweeps x N stages of
// small dense algebra.  One thread per instance leaves a B200 with B/32 = 128 resident warps at the
// benchmark batch (< 1 warp per SM) and 255-register threads that spill: ncu showed IPC 0.09 and an
// FP64 pipe 0.6 % busy (profiles/r01_*).  Splitting every 8x8 / 8-vector object across 8 lanes gives
// 8x the resident warps, 8x shorter dependency chains per lane and no spills:
//   * lane j owns component j of every 8-vector (dx, p, pi, ...) and row j (= column j) of the
//     symmetric Riccati matrix P;
//   * lanes 0..5 own one soft-constraint SIDE each ((bu,bx,h) x (lower,upper));
//   * small reductions (B'PB, B'p, g'dx, step length, residual norms) are xor-shuffle butterflies
//     inside the octet; 8-vector gathers are width-8 shuffles; the one 8x8 transpose per stage goes
//     through a padded shared-memory tile.
// Data stays in the batch-contiguous SoA arrays of DevMem: lane j of the four octets of a warp reads
// element (idx_j, b..b+3), i.e. whole 32-byte sectors.
//
// The arithmetic is the algorithm of oracle/nmpc_oracle.c (orc_qp_solve) / k_qp_solve.cu, reordered.
#include "nmpc_kernels.cuh"

namespace b200 {

#define AT(p, idx) (p)[(size_t)(idx) * B + b]
#define FULL 0xffffffffu

namespace {

__device__ __forceinline__ double sval(const double *shared, const double *inst, int idx, int b, int B)
{
    return inst ? inst[(size_t)idx * B + b] : shared[idx];
}
__device__ __forceinline__ double oct_bcast(double v, int src) { return __shfl_sync(FULL, v, src, 8); }
__device__ __forceinline__ double oct_sum(double v)
{
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 4);
    return v;
}
__device__ __forceinline__ double oct_min(double v)
{
    v = fmin(v, __shfl_xor_sync(FULL, v, 1));
    v = fmin(v, __shfl_xor_sync(FULL, v, 2));
    v = fmin(v, __shfl_xor_sync(FULL, v, 4));
    return v;
}
// max that propagates NaN (fmax would drop it)
__device__ __forceinline__ double nmax(double a, double b) { return (a != a || b != b) ? nan("") : fmax(a, b); }
__device__ __forceinline__ double oct_nmax(double v)
{
    v = nmax(v, __shfl_xor_sync(FULL, v, 1));
    v = nmax(v, __shfl_xor_sync(FULL, v, 2));
    v = nmax(v, __shfl_xor_sync(FULL, v, 4));
    return v;
}
__device__ __forceinline__ void oct_gather(double mine, double *all)
{
#pragma unroll
    for (int i = 0; i < 8; i++) all[i] = oct_bcast(mine, i);
}

// per-lane view of the soft rows of one stage
struct OctCon {
    double gfull;      // lane j: coefficient of dx_j in the linearised h row  [0,0,0,g_vl,g_vt,g_r,0,g_a]
    double g4[4];      // replicated
    double v0;         // lanes 0..5: value of my row at the linearisation point
    double bound, z, Z;    // lanes 0..5: my side's bound and (sigma-scaled) penalties
    bool act;          // lanes 0..5: my row exists at this stage and is penalised
    bool rowact[3];    // replicated: row c active
};

__device__ __forceinline__ void oct_load_con(const DevCfg &cfg, const DevMem &m, int k, int b, int lane, OctCon &c)
{
    const int B = cfg.B, N = cfg.N;
    const double sig = (k < N) ? cfg.dt : 1.0;
#pragma unroll
    for (int i = 0; i < 4; i++) c.g4[i] = AT(m.gh, k * 4 + i);
    c.gfull = (lane == 3) ? c.g4[0] : (lane == 4) ? c.g4[1] : (lane == 5) ? c.g4[2] : (lane == 7) ? c.g4[3] : 0.0;
    const int row = lane >> 1, sd = lane & 1;
    const bool side = lane < 6;
    const int r = side ? row : 0;
    const bool exists = (r == 0) ? (k < N) : (k >= 1);
    // every lane evaluates the activity of all three rows (cheap, keeps the flags replicated)
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const bool ex = (j == 0) ? (k < N) : (k >= 1);
        const double zl = sval(m.zl_s, m.zl_i, k * 3 + j, b, B), Zl = sval(m.Zl_s, m.Zl_i, k * 3 + j, b, B);
        const double zu = sval(m.zu_s, m.zu_i, k * 3 + j, b, B), Zu = sval(m.Zu_s, m.Zu_i, k * 3 + j, b, B);
        c.rowact[j] = ex && (sig * zl + sig * Zl > 0.0) && (sig * zu + sig * Zu > 0.0);
    }
    c.act = side && exists && c.rowact[r];
    c.v0 = (r == 0) ? ((k < N) ? AT(m.U, k * 2 + 1) : 0.0) : (r == 1) ? AT(m.X, k * 8 + 6) : AT(m.hval, k);
    c.bound = sd ? sval(m.ub_s, m.ub_i, k * 3 + r, b, B) : sval(m.lb_s, m.lb_i, k * 3 + r, b, B);
    c.z = sig * (sd ? sval(m.zu_s, m.zu_i, k * 3 + r, b, B) : sval(m.zl_s, m.zl_i, k * 3 + r, b, B));
    c.Z = sig * (sd ? sval(m.Zu_s, m.Zu_i, k * 3 + r, b, B) : sval(m.Zl_s, m.Zl_i, k * 3 + r, b, B));
}

// linearised value of my row: lanes 0,1 -> du1; 2,3 -> dx6; 4,5 -> g'dx   (all lanes must call: shuffles)
__device__ __forceinline__ double oct_row_lin(const OctCon &c, int lane, double dxj, double du1)
{
    const double hl = oct_sum(c.gfull * dxj);
    const double x6 = oct_bcast(dxj, 6);
    const int row = lane >> 1;
    return (row == 0) ? du1 : (row == 1) ? x6 : hl;
}

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

// row sums of a per-side quantity: returns (bu, bx, h) sums replicated on every lane
__device__ __forceinline__ void oct_rowsum(double side_val, double &r0, double &r1, double &r2)
{
    const double pair = side_val + __shfl_xor_sync(FULL, side_val, 1);
    r0 = oct_bcast(pair, 0); r1 = oct_bcast(pair, 2); r2 = oct_bcast(pair, 4);
}

__device__ __forceinline__ void load_AB(const DevMem &m, int k, int b, int B, double *A, double *Bm)
{
#pragma unroll
    for (int i = 0; i < 36; i++) A[i] = AT(m.A36, k * 36 + i);
#pragma unroll
    for (int i = 0; i < 12; i++) Bm[i] = AT(m.B12, k * 12 + i);
}

// row j of the full 8x8 A_k and 8x2 B_k (SURVEY A.2 structure)
__device__ __forceinline__ void load_AB_row(const DevCfg &cfg, const DevMem &m, int k, int b, int lane, double *Ar, double *Br)
{
    const int B = cfg.B;
    const int r = lane < 6 ? lane : 0;
#pragma unroll
    for (int c = 0; c < 6; c++) Ar[c] = AT(m.A36, k * 36 + r * 6 + c);
    Br[0] = AT(m.B12, k * 12 + r * 2); Br[1] = AT(m.B12, k * 12 + r * 2 + 1);
    if (lane >= 6) {
#pragma unroll
        for (int c = 0; c < 6; c++) Ar[c] = 0.0;
        Ar[lane - 2] = 1.0;                                  // rows 6,7 = e6', e7'
        Br[0] = (lane == 7) ? cfg.dt : 0.0; Br[1] = (lane == 6) ? cfg.dt : 0.0;
    }
}

// (A x)_j + (B u)_j with x gathered across the octet
__device__ __forceinline__ double oct_dyn(const double *Ar, const double *Br, int lane, double xj, double u0, double u1)
{
    double x[8];
    oct_gather(xj, x);
    double s = (lane < 2) ? xj : 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) s += Ar[c] * x[2 + c];
    return s + Br[0] * u0 + Br[1] * u1;
}

// (A' v)_j with v gathered across the octet; column j of the full A
__device__ __forceinline__ double At_comp(const double *A, int lane, const double *v)
{
    if (lane < 2) return v[lane];
    const int c = lane - 2;
    double s = (c == 4) ? v[6] : (c == 5) ? v[7] : 0.0;
#pragma unroll
    for (int r = 0; r < 6; r++) {
        // A[r*6 + c] with a lane-dependent c: select without dynamic register indexing
        const double a = (c == 0) ? A[r * 6] : (c == 1) ? A[r * 6 + 1] : (c == 2) ? A[r * 6 + 2]
                       : (c == 3) ? A[r * 6 + 3] : (c == 4) ? A[r * 6 + 4] : A[r * 6 + 5];
        s += a * v[r];
    }
    return s;
}
// (B' v), replicated (v gathered)
__device__ __forceinline__ void Bt_full(const double *Bm, double dt, const double *v, double &y0, double &y1)
{
    double s0 = dt * v[7], s1 = dt * v[6];
#pragma unroll
    for (int r = 0; r < 6; r++) { s0 += Bm[r * 2] * v[r]; s1 += Bm[r * 2 + 1] * v[r]; }
    y0 = s0; y1 = s1;
}

// stage cost Hessian diagonal entry of my component (dt-scaled) and the replicated R diagonal
__device__ __forceinline__ void oct_cost_diag(const DevCfg &cfg, const DevMem &m, int k, int b, int lane, double &qdj,
                                              double &rd0, double &rd1)
{
    const int B = cfg.B;
    const double sig = (k < cfg.N) ? cfg.dt : 1.0;
    const int nw = cfg.cost_vel_abs ? 3 : 4;
    qdj = (lane < nw) ? sig * sval(m.W_s, m.W_i, k * 6 + (lane < 4 ? lane : 0), b, B) : 0.0;
    rd0 = (k < cfg.N) ? sig * sval(m.W_s, m.W_i, k * 6 + 4, b, B) : 0.0;
    rd1 = (k < cfg.N) ? sig * sval(m.W_s, m.W_i, k * 6 + 5, b, B) : 0.0;
}

// Riccati state carried backward through a sweep: lane j holds row j of P and p_j
struct RicState { double P[8]; double p; };

// One backward Riccati step on the octet.  Hessian = diag(qd) + om_bx e6e6' + om_h g g' (+ Qv block),
// gradient (qt_j ; rt0, rt1), R diagonal (rd0, rd1 incl. om_bu).  Stores K, Hi, kff when `store`.
template <bool AFFINE, bool NEED_P>
__device__ __forceinline__ void oct_riccati_step(const DevCfg &cfg, const DevMem &m, int k, int b, int lane, bool store,
                                                 RicState &R, const double *A, const double *Bm, double qdj, double om_bx,
                                                 double om_h, const OctCon &c, const double *Qv, double qtj, double rd0,
                                                 double rd1, double rt0, double rt1, double *tile)
{
    const int B = cfg.B;
    const double dt = cfg.dt;
    // v = p (+ P b): component j
    double vj = R.p;
    if (AFFINE) {
#pragma unroll
        for (int l = 0; l < 8; l++) vj += R.P[l] * AT(m.bd, k * 8 + l);
    }
    // row j of P B, then Huu = R~ + B'PB by an octet reduction
    double pb0 = dt * R.P[7], pb1 = dt * R.P[6];
#pragma unroll
    for (int r = 0; r < 6; r++) { pb0 += R.P[r] * Bm[r * 2]; pb1 += R.P[r] * Bm[r * 2 + 1]; }
    // B_full[j][a]
    const double bj0 = (lane < 6) ? ((lane == 0) ? Bm[0] : (lane == 1) ? Bm[2] : (lane == 2) ? Bm[4] : (lane == 3) ? Bm[6]
                                     : (lane == 4) ? Bm[8] : Bm[10]) : ((lane == 7) ? dt : 0.0);
    const double bj1 = (lane < 6) ? ((lane == 0) ? Bm[1] : (lane == 1) ? Bm[3] : (lane == 2) ? Bm[5] : (lane == 3) ? Bm[7]
                                     : (lane == 4) ? Bm[9] : Bm[11]) : ((lane == 6) ? dt : 0.0);
    const double h00 = rd0 + oct_sum(bj0 * pb0);
    const double h01 = 0.5 * (oct_sum(bj0 * pb1) + oct_sum(bj1 * pb0));
    const double h11 = rd1 + oct_sum(bj1 * pb1);
    const double l11 = sqrt(h00), l21 = h01 / l11;
    const double l22 = sqrt(h11 - l21 * l21);
    const double i11 = 1.0 / l11, i22 = 1.0 / l22;
    const double m21 = -l21 * i11 * i22;
    const double Hi0 = i11 * i11 + m21 * m21, Hi1 = m21 * i22, Hi3 = i22 * i22;
    // gradient part
    const double hu0 = rt0 + oct_sum(bj0 * vj), hu1 = rt1 + oct_sum(bj1 * vj);
    const double kf0 = -(Hi0 * hu0 + Hi1 * hu1), kf1 = -(Hi1 * hu0 + Hi3 * hu1);
    if (store && lane == 0) {
        AT(m.Hi, k * 3 + 0) = Hi0; AT(m.Hi, k * 3 + 1) = Hi1; AT(m.Hi, k * 3 + 2) = Hi3;
        AT(m.kff, k * 2 + 0) = kf0; AT(m.kff, k * 2 + 1) = kf1;
    }
    if (!NEED_P) return;
    double vall[8];
    oct_gather(vj, vall);
    const double hxj = qtj + At_comp(A, lane, vall);
    // row j of P A
    double PAj[8];
    PAj[0] = R.P[0]; PAj[1] = R.P[1];
#pragma unroll
    for (int cc = 0; cc < 6; cc++) {
        double s = (cc == 4) ? R.P[6] : (cc == 5) ? R.P[7] : 0.0;
#pragma unroll
        for (int r = 0; r < 6; r++) s += R.P[r] * A[r * 6 + cc];
        PAj[2 + cc] = s;
    }
    // transpose through the padded tile: write row j, read column j
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) tile[lane * 9 + i] = PAj[i];
    __syncwarp();
    double col[8];
#pragma unroll
    for (int l = 0; l < 8; l++) col[l] = tile[l * 9 + lane];
    // column j of Hxx = A' PAcol and of Hux = B' PAcol
    double Hj[8];
    Hj[0] = col[0]; Hj[1] = col[1];
#pragma unroll
    for (int cc = 0; cc < 6; cc++) {
        double s = (cc == 4) ? col[6] : (cc == 5) ? col[7] : 0.0;
#pragma unroll
        for (int r = 0; r < 6; r++) s += A[r * 6 + cc] * col[r];
        Hj[2 + cc] = s;
    }
    double hux0, hux1;
    Bt_full(Bm, dt, col, hux0, hux1);
    // Y = L^-1 Hux, K = -L^-T Y  (column j)
    const double Y0 = hux0 * i11;
    const double Y1 = (hux1 - l21 * Y0) * i22;
    const double K1 = -Y1 * i22;
    const double K0 = (-Y0 - l21 * K1) * i11;
    if (store) { AT(m.K, k * 16 + lane) = K0; AT(m.K, k * 16 + 8 + lane) = K1; }
    // row j of P_k = Q~ + Hxx - Y'Y
    double Y0all[8], Y1all[8];
    oct_gather(Y0, Y0all);
    oct_gather(Y1, Y1all);
#pragma unroll
    for (int i = 0; i < 8; i++) R.P[i] = Hj[i] - (Y0all[i] * Y0 + Y1all[i] * Y1);
    // Q~ row j: diag + om_bx e6 e6' + om_h g g' + Qv block
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double gi = (i == 3) ? c.g4[0] : (i == 4) ? c.g4[1] : (i == 5) ? c.g4[2] : (i == 7) ? c.g4[3] : 0.0;
        double add = om_h * c.gfull * gi;
        if (i == lane) add += qdj + ((lane == 6) ? om_bx : 0.0);
        R.P[i] += add;
    }
    if (Qv) {
        if (lane == 3) { R.P[3] += Qv[0]; R.P[4] += Qv[1]; }
        if (lane == 4) { R.P[3] += Qv[1]; R.P[4] += Qv[2]; }
    }
    R.p = hxj + K0 * hu0 + K1 * hu1;
}

__device__ __forceinline__ void oct_riccati_terminal(RicState &R, int lane, double qdj, double om_bx, double om_h,
                                                     const OctCon &c, const double *Qv, double qtj)
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double gi = (i == 3) ? c.g4[0] : (i == 4) ? c.g4[1] : (i == 5) ? c.g4[2] : (i == 7) ? c.g4[3] : 0.0;
        double v = om_h * c.gfull * gi;
        if (i == lane) v += qdj + ((lane == 6) ? om_bx : 0.0);
        R.P[i] = v;
    }
    if (Qv) {
        if (lane == 3) { R.P[3] += Qv[0]; R.P[4] += Qv[1]; }
        if (lane == 4) { R.P[3] += Qv[1]; R.P[4] += Qv[2]; }
    }
    R.p = qtj;
}

#define SIDE(k) ((k) * 6 + sidx)

// Backward sweep of one IPM iteration (FULL: costates + barrier Hessian + factorisation + predictor rhs;
// !FULL: costates / u-stationarity residual only).  Returns max |r_u| (NaN-propagating), replicated.
template <bool FULLSWEEP>
__device__ double oct_sweep_backward(const DevCfg &cfg, const DevMem &m, int b, int lane, bool store, double *tile)
{
    const int B = cfg.B, N = cfg.N;
    const double reg = cfg.ipm_reg;
    const int sidx = lane < 6 ? lane : 0;
    const double sgn = (lane & 1) ? -1.0 : 1.0;
    RicState R;
    double pij = 0.0;
    double rstat = 0.0;
    for (int k = N; k >= 0; k--) {
        OctCon c;
        oct_load_con(cfg, m, k, b, lane, c);
        const double dxj = AT(m.dx, k * 8 + lane);
        double du0 = 0.0, du1 = 0.0;
        if (k < N) { du0 = AT(m.du, k * 2); du1 = AT(m.du, k * 2 + 1); }
        double qdj, rd0, rd1;
        oct_cost_diag(cfg, m, k, b, lane, qdj, rd0, rd1);
        double Qv[3] = {0.0, 0.0, 0.0};
        if (cfg.cost_vel_abs) { Qv[0] = AT(m.Qv, k * 3); Qv[1] = AT(m.Qv, k * 3 + 1); Qv[2] = AT(m.Qv, k * 3 + 2); }
        // stationarity rows at the current iterate: my component of g_x, replicated g_u
        double gxj = qdj * dxj + ((lane < 5) ? AT(m.qx, k * 5 + (lane < 5 ? lane : 0)) : 0.0);
        if (cfg.cost_vel_abs) {
            const double d3 = oct_bcast(dxj, 3), d4 = oct_bcast(dxj, 4);
            if (lane == 3) gxj += Qv[0] * d3 + Qv[1] * d4;
            if (lane == 4) gxj += Qv[1] * d3 + Qv[2] * d4;
        }
        double gu0 = 0.0, gu1 = 0.0;
        if (k < N) { gu0 = AT(m.rg, k * 2) + rd0 * du0; gu1 = AT(m.rg, k * 2 + 1) + rd1 * du1; }
        // my side
        const double vlin = oct_row_lin(c, lane, dxj, du1);
        double s = 1.0, t = 1.0, lam = 0.0, mu = 1.0;
        if (c.act) { s = AT(m.ips, SIDE(k)); t = AT(m.ipt, SIDE(k)); lam = AT(m.iplam, SIDE(k)); mu = AT(m.ipmu, SIDE(k)); }
        double om_s = 0.0, cc_s = 0.0;
        if (FULLSWEEP && c.act) {
            const double v = c.v0 + vlin;
            const double rt = t - (sgn * (v - c.bound) + s);
            const double rs = c.z + c.Z * s - lam - mu;
            const SideK sk = side_coeffs(reg, c.Z, s, t, lam, mu, rt, rs, -lam * t, -mu * s);
            om_s = sk.om; cc_s = sgn * sk.cc;
        }
        double ld_bu, ld_bx, ld_h;                 // sum_sides sgn * lam per row
        oct_rowsum(c.act ? sgn * lam : 0.0, ld_bu, ld_bx, ld_h);
        gu1 -= ld_bu;
        gxj -= ((lane == 6) ? ld_bx : 0.0) + ld_h * c.gfull;
        double om_bu = 0.0, om_bx = 0.0, om_h = 0.0, cf_bu = 0.0, cf_bx = 0.0, cf_h = 0.0;
        if (FULLSWEEP) {
            oct_rowsum(om_s, om_bu, om_bx, om_h);
            oct_rowsum(cc_s, cf_bu, cf_bx, cf_h);
        }
        double A[36], Bm[12];
        if (k < N) {
            load_AB(m, k, b, B, A, Bm);
            double pall[8];
            oct_gather(pij, pall);
            gxj += At_comp(A, lane, pall);
            double t0, t1;
            Bt_full(Bm, cfg.dt, pall, t0, t1);
            gu0 += t0; gu1 += t1;
            if (store && lane == 0) { AT(m.gus, k * 2) = gu0; AT(m.gus, k * 2 + 1) = gu1; }
            rstat = nmax(rstat, nmax(fabs(gu0), fabs(gu1)));
        }
        pij = gxj;
        if (FULLSWEEP) {
            // predictor rhs: grad_x = -sum sgn cc g_x, grad_u = r_u - sum sgn cc g_u
            const double qtj = -(((lane == 6) ? cf_bx : 0.0) + cf_h * c.gfull);
            if (k == N) {
                oct_riccati_terminal(R, lane, qdj, om_bx, om_h, c, cfg.cost_vel_abs ? Qv : nullptr, qtj);
            } else {
                const double rt0 = gu0, rt1 = gu1 - cf_bu;
                if (k > 0)
                    oct_riccati_step<false, true>(cfg, m, k, b, lane, store, R, A, Bm, qdj, om_bx, om_h, c,
                                                  cfg.cost_vel_abs ? Qv : nullptr, qtj, rd0, rd1 + om_bu, rt0, rt1, tile);
                else
                    oct_riccati_step<false, false>(cfg, m, k, b, lane, store, R, A, Bm, qdj, om_bx, om_h, c,
                                                   cfg.cost_vel_abs ? Qv : nullptr, qtj, rd0, rd1 + om_bu, rt0, rt1, tile);
            }
        }
    }
    return rstat;
}

// Corrector: backward vector sweep with the stored factorisation.
__device__ void oct_sweep_corrector(const DevCfg &cfg, const DevMem &m, int b, int lane, bool store, double sigma_mu)
{
    const int B = cfg.B, N = cfg.N;
    const double reg = cfg.ipm_reg;
    const int sidx = lane < 6 ? lane : 0;
    const double sgn = (lane & 1) ? -1.0 : 1.0;
    double pj = 0.0;
    for (int k = N; k >= 0; k--) {
        OctCon c;
        oct_load_con(cfg, m, k, b, lane, c);
        const double dxj = AT(m.dx, k * 8 + lane);
        const double du1 = (k < N) ? AT(m.du, k * 2 + 1) : 0.0;
        const double vlin = oct_row_lin(c, lane, dxj, du1);
        double cc_s = 0.0;
        if (c.act) {
            const double s = AT(m.ips, SIDE(k)), t = AT(m.ipt, SIDE(k)), lam = AT(m.iplam, SIDE(k)), mu = AT(m.ipmu, SIDE(k));
            const double v = c.v0 + vlin;
            const double rt = t - (sgn * (v - c.bound) + s);
            const double rs = c.z + c.Z * s - lam - mu;
            const double r1 = -lam * t + sigma_mu - AT(m.pr1, SIDE(k));
            const double r2 = -mu * s + sigma_mu - AT(m.pr2, SIDE(k));
            cc_s = sgn * side_coeffs(reg, c.Z, s, t, lam, mu, rt, rs, r1, r2).cc;
        }
        double cf_bu, cf_bx, cf_h;
        oct_rowsum(cc_s, cf_bu, cf_bx, cf_h);
        const double qtj = -(((lane == 6) ? cf_bx : 0.0) + cf_h * c.gfull);
        if (k == N) {
            pj = qtj;
        } else {
            double A[36], Bm[12];
            load_AB(m, k, b, B, A, Bm);
            double pall[8];
            oct_gather(pj, pall);
            double hu0, hu1;
            Bt_full(Bm, cfg.dt, pall, hu0, hu1);
            hu0 += AT(m.gus, k * 2);
            hu1 += AT(m.gus, k * 2 + 1) - cf_bu;
            const double Hi0 = AT(m.Hi, k * 3), Hi1 = AT(m.Hi, k * 3 + 1), Hi3 = AT(m.Hi, k * 3 + 2);
            if (store && lane == 0) {
                AT(m.kff, k * 2 + 0) = -(Hi0 * hu0 + Hi1 * hu1);
                AT(m.kff, k * 2 + 1) = -(Hi1 * hu0 + Hi3 * hu1);
            }
            if (k > 0)
                pj = At_comp(A, lane, pall) + qtj + AT(m.K, k * 16 + lane) * hu0 + AT(m.K, k * 16 + 8 + lane) * hu1;
        }
    }
}

// Forward sweep: Newton direction, slack / multiplier directions of my side, step to the boundary.
template <int PASS>
__device__ double oct_sweep_forward(const DevCfg &cfg, const DevMem &m, int b, int lane, bool store, double sigma_mu,
                                    double *poly)
{
    const int B = cfg.B, N = cfg.N;
    const double reg = cfg.ipm_reg;
    const int sidx = lane < 6 ? lane : 0;
    const double sgn = (lane & 1) ? -1.0 : 1.0;
    double ddxj = 0.0;
    double amax = 1e300, c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int k = 0; k <= N; k++) {
        OctCon c;
        oct_load_con(cfg, m, k, b, lane, c);
        const double dxj = AT(m.dx, k * 8 + lane);
        double du1 = 0.0, ddu0 = 0.0, ddu1 = 0.0;
        if (k < N) {
            du1 = AT(m.du, k * 2 + 1);
            double k0 = 0.0, k1 = 0.0;
            if (k > 0) { k0 = AT(m.K, k * 16 + lane) * ddxj; k1 = AT(m.K, k * 16 + 8 + lane) * ddxj; }
            ddu0 = AT(m.kff, k * 2) + oct_sum(k0);
            ddu1 = AT(m.kff, k * 2 + 1) + oct_sum(k1);
            if (PASS == 1 && store && lane == 0) { AT(m.ddu, k * 2) = ddu0; AT(m.ddu, k * 2 + 1) = ddu1; }
        }
        const double vlin = oct_row_lin(c, lane, dxj, du1);
        const double dv = oct_row_lin(c, lane, ddxj, ddu1);
        if (c.act) {
            const double s = AT(m.ips, SIDE(k)), t = AT(m.ipt, SIDE(k)), lam = AT(m.iplam, SIDE(k)), mu = AT(m.ipmu, SIDE(k));
            const double v = c.v0 + vlin;
            const double rt = t - (sgn * (v - c.bound) + s);
            const double rs = c.z + c.Z * s - lam - mu;
            double r1 = -lam * t, r2 = -mu * s;
            if (PASS == 1) { r1 += sigma_mu - AT(m.pr1, SIDE(k)); r2 += sigma_mu - AT(m.pr2, SIDE(k)); }
            const SideK sk = side_coeffs(reg, c.Z, s, t, lam, mu, rt, rs, r1, r2);
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
                if (store) { AT(m.pr1, SIDE(k)) = dl * dtt; AT(m.pr2, SIDE(k)) = dm * ds; }
                c0 += lam * t + mu * s;
                c1 += lam * dtt + t * dl + mu * ds + s * dm;
                c2 += dl * dtt + dm * ds;
            } else if (store) {
                AT(m.ids, SIDE(k)) = ds; AT(m.idlam, SIDE(k)) = dl; AT(m.idt, SIDE(k)) = dtt; AT(m.idmu, SIDE(k)) = dm;
            }
        }
        if (k < N) {
            double Ar[6], Br[2];
            load_AB_row(cfg, m, k, b, lane, Ar, Br);
            ddxj = oct_dyn(Ar, Br, lane, ddxj, ddu0, ddu1);
        }
    }
    if (PASS == 0) { poly[0] = oct_sum(c0); poly[1] = oct_sum(c1); poly[2] = oct_sum(c2); }
    return oct_min(amax);
}

// Step of length alpha (MEASURE_ONLY: alpha = 0, no stores), re-roll of the states, new residual measures.
template <bool MEASURE_ONLY>
__device__ void oct_sweep_update(const DevCfg &cfg, const DevMem &m, int b, int lane, bool store, double alpha, int nact,
                                 double *gap, double *cmax, double *rmax)
{
    const int B = cfg.B, N = cfg.N;
    const int sidx = lane < 6 ? lane : 0;
    const double sgn = (lane & 1) ? -1.0 : 1.0;
    double dxj = AT(m.dx, lane);
    double g = 0.0, cm = 0.0, rm = 0.0;
    for (int k = 0; k <= N; k++) {
        OctCon c;
        oct_load_con(cfg, m, k, b, lane, c);
        double du0 = 0.0, du1 = 0.0;
        if (k < N) {
            du0 = AT(m.du, k * 2); du1 = AT(m.du, k * 2 + 1);
            if (!MEASURE_ONLY) {
                du0 += alpha * AT(m.ddu, k * 2); du1 += alpha * AT(m.ddu, k * 2 + 1);
                if (store && lane == 0) { AT(m.du, k * 2) = du0; AT(m.du, k * 2 + 1) = du1; }
            }
        }
        const double vlin = oct_row_lin(c, lane, dxj, du1);
        if (c.act) {
            double s = AT(m.ips, SIDE(k)), t = AT(m.ipt, SIDE(k)), lam = AT(m.iplam, SIDE(k)), mu = AT(m.ipmu, SIDE(k));
            if (!MEASURE_ONLY) {
                s += alpha * AT(m.ids, SIDE(k)); t += alpha * AT(m.idt, SIDE(k));
                lam += alpha * AT(m.idlam, SIDE(k)); mu += alpha * AT(m.idmu, SIDE(k));
                if (store) { AT(m.ips, SIDE(k)) = s; AT(m.ipt, SIDE(k)) = t; AT(m.iplam, SIDE(k)) = lam; AT(m.ipmu, SIDE(k)) = mu; }
            }
            const double p1 = lam * t, p2 = mu * s;
            g += p1 + p2;
            cm = nmax(cm, nmax(p1, p2));
            const double v = c.v0 + vlin;
            const double rt = t - (sgn * (v - c.bound) + s);
            const double rs = c.z + c.Z * s - lam - mu;
            rm = nmax(rm, nmax(fabs(rt), fabs(rs)));
        }
        if (k < N) {
            double Ar[6], Br[2];
            load_AB_row(cfg, m, k, b, lane, Ar, Br);
            dxj = oct_dyn(Ar, Br, lane, dxj, du0, du1) + AT(m.bd, k * 8 + lane);
            if (store) AT(m.dx, (k + 1) * 8 + lane) = dxj;
        }
    }
    *gap = oct_sum(g) / (4.0 * nact); *cmax = oct_nmax(cm); *rmax = oct_nmax(rm);
}

}  // namespace

// 64 threads = 2 warps = 8 instances per block; grid covers ceil(B/8) blocks.
__global__ void __launch_bounds__(64) k_qp_solve_octet(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    __shared__ double tiles[8][72];                      // one padded 8x9 transpose tile per octet
    const int B = cfg.B, N = cfg.N;
    const int lane = threadIdx.x & 7;
    const int oct = threadIdx.x >> 3;
    const int inst = blockIdx.x * 8 + oct;
    const bool valid = inst < B;
    const int b = valid ? inst : B - 1;                  // padding octets shadow the last instance, never store
    double *tile = tiles[oct];
    const double tol = cfg.ipm_tol;
    const int sidx = lane < 6 ? lane : 0;

    // ---------------- phase 0: equality-constrained optimum (affine Riccati solve)
    {
        RicState R;
        for (int k = N; k >= 0; k--) {
            OctCon c;
            oct_load_con(cfg, m, k, b, lane, c);
            double qdj, rd0, rd1;
            oct_cost_diag(cfg, m, k, b, lane, qdj, rd0, rd1);
            const double qtj = (lane < 5) ? AT(m.qx, k * 5 + (lane < 5 ? lane : 0)) : 0.0;
            double Qv[3] = {0.0, 0.0, 0.0};
            if (cfg.cost_vel_abs) { Qv[0] = AT(m.Qv, k * 3); Qv[1] = AT(m.Qv, k * 3 + 1); Qv[2] = AT(m.Qv, k * 3 + 2); }
            if (k == N) {
                oct_riccati_terminal(R, lane, qdj, 0.0, 0.0, c, cfg.cost_vel_abs ? Qv : nullptr, qtj);
            } else {
                double A[36], Bm[12];
                load_AB(m, k, b, B, A, Bm);
                oct_riccati_step<true, true>(cfg, m, k, b, lane, valid, R, A, Bm, qdj, 0.0, 0.0, c,
                                             cfg.cost_vel_abs ? Qv : nullptr, qtj, rd0, rd1, AT(m.rg, k * 2),
                                             AT(m.rg, k * 2 + 1), tile);
            }
        }
    }
    __syncwarp();
    int nact = 0;
    bool feasible = true;
    {
        double dxj = AT(m.x0, lane) - AT(m.X, lane);
        if (valid) AT(m.dx, lane) = dxj;
        for (int k = 0; k <= N; k++) {
            OctCon c;
            oct_load_con(cfg, m, k, b, lane, c);
            double du0 = 0.0, du1 = 0.0;
            if (k < N) {
                du0 = AT(m.kff, k * 2) + oct_sum(AT(m.K, k * 16 + lane) * dxj);
                du1 = AT(m.kff, k * 2 + 1) + oct_sum(AT(m.K, k * 16 + 8 + lane) * dxj);
                if (valid && lane == 0) { AT(m.du, k * 2) = du0; AT(m.du, k * 2 + 1) = du1; }
            }
            const double vlin = oct_row_lin(c, lane, dxj, du1);
            nact += (int)c.rowact[0] + (int)c.rowact[1] + (int)c.rowact[2];
            if (valid && lane < 3) { AT(m.sl, k * 3 + lane) = 0.0; AT(m.su, k * 3 + lane) = 0.0; }
            // interior start with zero linear residuals: s_l = s_u = S, lam_l = lam_u (lanes 2c, 2c+1 agree on S, l0)
            const double v = c.v0 + vlin;
            const double lo = (lane & 1) ? __shfl_xor_sync(FULL, c.bound, 1) : c.bound;
            const double up = (lane & 1) ? c.bound : __shfl_xor_sync(FULL, c.bound, 1);
            const double gother_z = __shfl_xor_sync(FULL, c.z, 1), gother_Z = __shfl_xor_sync(FULL, c.Z, 1);
            if (c.act) {
                if (v < lo || v > up) feasible = false;
                const double viol = fmax(fmax(lo - v, v - up), 0.0);
                const double S = 1.0 + viol;
                const double gme = c.z + c.Z * S, goth = gother_z + gother_Z * S;
                const double l0 = 0.5 * fmin(gme, goth);
                if (valid) {
                    AT(m.ips, SIDE(k)) = S; AT(m.iplam, SIDE(k)) = l0;
                    AT(m.ipt, SIDE(k)) = ((lane & 1) ? up - v : v - lo) + S;
                    AT(m.ipmu, SIDE(k)) = gme - l0;
                }
            }
            if (k < N) {
                double Ar[6], Br[2];
                load_AB_row(cfg, m, k, b, lane, Ar, Br);
                dxj = oct_dyn(Ar, Br, lane, dxj, du0, du1) + AT(m.bd, k * 8 + lane);
                if (valid) AT(m.dx, (k + 1) * 8 + lane) = dxj;
            }
        }
        // feasibility is an octet-wide AND
        feasible = oct_min(feasible ? 1.0 : 0.0) > 0.5;
    }
    __syncwarp();

    // ---------------- interior-point iterations (warp-uniform control flow; finished octets stop storing)
    int iters = 0, qp_status = 0;
    double kkt = 0.0;
    bool active = valid && !((feasible && !cfg.qp_no_shortcut) || nact == 0);
    const bool ran_ipm = active;
    double gap = 0.0, cmax = 0.0, rmax = 0.0;
    if (__any_sync(FULL, active)) {
        oct_sweep_update<true>(cfg, m, b, lane, false, 0.0, nact > 0 ? nact : 1, &gap, &cmax, &rmax);
        bool converged = false;
        for (int it = 0; it <= cfg.ipm_max_iter; it++) {
            if (active && (!(cmax == cmax) || !(rmax == rmax))) { qp_status = 3; iters = it; active = false; }
            // octets that pass the complementarity / residual test check stationarity with a light sweep
            const bool want_check = active && cmax <= tol && rmax <= tol;
            if (__any_sync(FULL, want_check)) {
                const double rs = oct_sweep_backward<false>(cfg, m, b, lane, false, tile);
                if (want_check) {
                    if (!(rs == rs)) { qp_status = 3; iters = it; active = false; }
                    else if (rs <= tol) { converged = true; kkt = fmax(fmax(cmax, rmax), rs); iters = it; active = false; }
                }
            }
            if (active && it == cfg.ipm_max_iter) { kkt = fmax(cmax, rmax); iters = it; active = false; }
            if (!__any_sync(FULL, active)) break;
            const double rstat = oct_sweep_backward<true>(cfg, m, b, lane, active, tile);
            if (active && !(rstat == rstat)) { qp_status = 3; iters = it; active = false; }
            __syncwarp();
            double poly[3];
            const double amax_aff = oct_sweep_forward<0>(cfg, m, b, lane, active, 0.0, poly);
            const double a = fmin(1.0, amax_aff);
            const double gap_aff = (poly[0] + a * poly[1] + a * a * poly[2]) / (4.0 * (nact > 0 ? nact : 1));
            const double sg = gap_aff / gap;
            const double sigma_mu = fmax(sg * sg * sg * gap, 0.1 * tol);
            __syncwarp();
            oct_sweep_corrector(cfg, m, b, lane, active, sigma_mu);
            __syncwarp();
            const double amax = oct_sweep_forward<1>(cfg, m, b, lane, active, sigma_mu, poly);
            const double eta = fmax(0.995, 1.0 - gap);
            const double alpha = fmin(1.0, eta * amax);
            __syncwarp();
            double g2, c2, r2;
            oct_sweep_update<false>(cfg, m, b, lane, active, alpha, nact > 0 ? nact : 1, &g2, &c2, &r2);
            if (active) { gap = g2; cmax = c2; rmax = r2; }
            __syncwarp();
        }
        if (ran_ipm && !converged && qp_status == 0) qp_status = 1;
    }
    __syncwarp();
    if (ran_ipm && lane < 6) {
        // slack values of the last QP (acados get(stage,"sl"/"su")); side lane = row*2 + (upper)
        for (int k = 0; k <= N; k++) {
            const int row = lane >> 1;
            const bool ex = (row == 0) ? (k < N) : (k >= 1);
            if (!ex) continue;
            const double sv = AT(m.ips, SIDE(k));
            if (lane & 1) AT(m.su, k * 3 + row) = sv; else AT(m.sl, k * 3 + row) = sv;
        }
    }
    __syncwarp();

    // ---------------- K3: full step and cost at the new iterate
    double badf = (qp_status == 3) ? 1.0 : 0.0;
    for (int k = 0; k <= N; k++) {
        const double d = AT(m.dx, k * 8 + lane);
        if (!(d == d)) badf = 1.0;
        if (k < N && lane < 2) { const double e = AT(m.du, k * 2 + lane); if (!(e == e)) badf = 1.0; }
    }
    const bool bad = oct_nmax(badf) != 0.0;
    double Jl = 0.0;
    for (int k = 0; k <= N; k++) {
        const double sig = (k < N) ? cfg.dt : 1.0;
        double xj = AT(m.X, k * 8 + lane);
        if (!bad) { xj += AT(m.dx, k * 8 + lane); if (valid) AT(m.X, k * 8 + lane) = xj; }
        double c = 0.0;
        // tracking error of my output component: lanes 0..3 -> y0..y3, lanes 4,5 -> jerk, omega
        const double x4 = oct_bcast(xj, 4);
        if (lane < 4) {
            double y = xj;
            if (lane == 2) y = wrap_yaw(xj);
            if (lane == 3 && cfg.cost_vel_abs) y = sqrt(xj * xj + x4 * x4);
            const double e = y - AT(m.yref, k * 6 + lane);
            c = 0.5 * sval(m.W_s, m.W_i, k * 6 + lane, b, B) * e * e;
        } else if (lane < 6 && k < N) {
            double u = AT(m.U, k * 2 + (lane - 4));
            if (!bad) { u += AT(m.du, k * 2 + (lane - 4)); if (valid) AT(m.U, k * 2 + (lane - 4)) = u; }
            const double e = u - AT(m.yref, k * 6 + lane);
            c = 0.5 * sval(m.W_s, m.W_i, k * 6 + lane, b, B) * e * e;
        }
        // slack cost of my side
        if (lane < 6) {
            const int row = lane >> 1;
            const bool ex = (row == 0) ? (k < N) : (k >= 1);
            if (ex) {
                const double sv = (lane & 1) ? AT(m.su, k * 3 + row) : AT(m.sl, k * 3 + row);
                const double z = (lane & 1) ? sval(m.zu_s, m.zu_i, k * 3 + row, b, B) : sval(m.zl_s, m.zl_i, k * 3 + row, b, B);
                const double Z = (lane & 1) ? sval(m.Zu_s, m.Zu_i, k * 3 + row, b, B) : sval(m.Zl_s, m.Zl_i, k * 3 + row, b, B);
                c += z * sv + 0.5 * Z * sv * sv;
            }
        }
        Jl += sig * c;
    }
    const double J = oct_sum(Jl);
    if (valid && lane == 0) {
        m.cost[b] = J;
        m.kkt[b] = kkt;
        m.qp_iter[b] = iters;
        m.status[b] = bad ? 4 : 0;       // ACADOS_QP_FAILURE; RTI keeps status 0 when the QP hits max-iter (SURVEY B-9)
    }
}

void launch_qp_solve_octet(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    k_qp_solve_octet<<<(cfg.B + 7) / 8, 64, 0, st>>>(cfg, m);
}

}  // namespace b200
