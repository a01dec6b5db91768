// snmpc_head.cuh -- the Stochastic NMPC specific parts of the QP kernel (included by k_qp_warp.cu only): head condensing as
// the PROLOGUE and the sample-iterate step as the EPILOGUE of k_qp_solve_warp<HEAD>, so that one SNMPC solve is two
// launches (k_linearize with the sample copies as extra work items, then the QP kernel).
//
// Reference being replaced (paths relative to /root/reference): the OCP of
// Model_Predictive_Controller/Stochastic_NMPC -- augmented discrete model pred_model_dynamic_disc.py:115-220
// (1 + n_s copies of the 8-state model, samples propagated while stop_flag == 0, nominal next state = PCE mean
// A[0,:] * samples_next), chance constraint SNMPC_acados_settings.py:86-133,184 (h~ = E[h] + kappa sqrt(Var[h]) from
// the PCE coefficients A * [h(sample_i)]), all 8(1+n_s) initial states fixed (SNMPC_class.py:259-264).
//
// acados treats this as a dense 88-state OCP (that is why the reference's SNMPC solve is 6x slower than its NMPC
// solve).  Here the structure is used instead: because every sample's initial state is fixed, the
// sample states of the propagation horizon (stages <= H = UPH) are affine functions of the first nh = 2H inputs
// only.  snmpc_condense_warp eliminates stages 0..H-1 onto those nh "head" inputs (per-sample 8-vector chains, the
// PCE mean for the nominal state, the chance-constraint rows), and the Riccati recursion of k_qp_warp.cu starts
// at stage H with dx_H = G_H u~ + c_H; the head itself is an nh x nh dense solve inside the same warp.
#pragma once
#include "nmpc_kernels.cuh"

namespace b200 {

__device__ __forceinline__ double stage_val_s(const double *shared, const double *inst, int b, int NS, int k, int ncomp, int comp)
{
    return inst ? inst[((size_t)b * NS + k) * ncomp + comp] : shared[k * ncomp + comp];
}

// scratch of the condensing prologue, doubles per warp (aliases the QP kernel's per-warp shared memory, which is
// initialised afterwards)
__host__ __device__ inline int condense_smem_doubles(int H, int ns)
{
    const int nh = 2 * H, ldc = nh + 1;
    // G accumulators [H][8][ldc] | row accumulators [H][ldc] | dd [H][ns] | hv [ns] | cj [PCE_L] | v0 base [H]
    return (H * 8 * ldc + H * ldc + H * ns + ns + B200NMPC_PCE_L + H + 1) & ~1;
}

// One warp per instance.  Lane a <= nh owns column a of every affine map (column nh is the constant part).
// `sm`: condense_smem_doubles(H, n_s) doubles of per-warp scratch.  Writes the condensed head block of instance b (HBM, read
// back by the same warp after a warp barrier; b200nmpc_snmpc_get_head hands it to tests).
__device__ void snmpc_condense_warp(const DevCfg &cfg, const DevMem &m, int b, int lane, double *sm)
{
    const int N = cfg.N, NS = N + 1, H = cfg.head_H, ns = cfg.n_samples, L = cfg.sn_L;
    const HeadLayout hl(H);
    const int nh = hl.nh, ldc = nh + 1;
    double *Gacc = sm;                          // [H][8][ldc]   index k -> map of stage k+1
    double *Racc = Gacc + H * 8 * ldc;          // [H][ldc]      chance-constraint rows of stage k (col nh: value part)
    double *dd = Racc + H * ldc;                // [H][ns]       d h~_k / d h(sample s)
    double *hv = dd + H * ns;                   // [ns]
    double *cj = hv + ns;                       // [PCE_L]
    double *v0b = cj + B200NMPC_PCE_L;          // [H]           h~ at the sample iterates
    const double *slin = m.slin + (size_t)b * m.HA * ns * SREC;
    const double *XS = m.XS + (size_t)b * (m.HA + 1) * ns * 8;
    const double *x0s = m.x0s + (size_t)b * ns * 8;
    const double kappa = cfg.sn_kappa;

    for (int i = lane; i < H * 8 * ldc + H * ldc; i += 32) sm[i] = 0.0;
    if (lane < H) v0b[lane] = 0.0;
    __syncwarp();

    // ---- 1. PCE coefficients of h at the sample iterates, h~ and its derivative w.r.t. every sample value
    for (int k = 1; k < H; k++) {
        if (lane < ns) hv[lane] = slin[((size_t)k * ns + lane) * SREC + SR_H];
        __syncwarp();
        if (lane < L) {
            double s2 = 0.0;
            for (int i = 0; i < ns; i++) s2 += m.sn_A[lane * ns + i] * hv[i];
            cj[lane] = s2;
        }
        __syncwarp();
        double var = 0.0;
        for (int j = 1; j < L; j++) var += cj[j] * cj[j];
        const double sd = sqrt(var);
        if (lane < ns) {
            double g = m.sn_A[lane];
            if (sd > 0.0)
                for (int j = 1; j < L; j++) g += kappa * cj[j] * m.sn_A[j * ns + lane] / sd;
            dd[k * ns + lane] = g;
        }
        if (lane == 0) v0b[k] = cj[0] + kappa * sd;
        __syncwarp();
    }

    // ---- 2. per-sample column chains g <- A_s,k g (+ B_s,k column, + defect for the constant column)
    if (lane <= nh) {
        const int a = lane;
        const bool affine = (a == nh);
        const int kstart = affine ? 0 : (a >> 1);          // input column a enters at stage a/2
        for (int s = 0; s < ns; s++) {
            const double a0 = m.sn_A[s];                   // PCE mean weights (row 0 of A)
            double g[8];
#pragma unroll
            for (int i = 0; i < 8; i++) g[i] = affine ? x0s[s * 8 + i] - XS[(size_t)s * 8 + i] : 0.0;
            for (int k = kstart; k < H; k++) {
                const double *rec = slin + ((size_t)k * ns + s) * SREC;
                if (k >= 1) {
                    const double dot = rec[SR_GH] * g[3] + rec[SR_GH + 1] * g[4] + rec[SR_GH + 2] * g[5] + rec[SR_GH + 3] * g[7];
                    Racc[k * ldc + a] += dd[k * ns + s] * dot;
                }
                double t[8];
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    double v = (r < 2) ? g[r] : 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) v += rec[SR_A + r * 6 + c] * g[2 + c];
                    t[r] = v;
                }
                t[6] = g[6]; t[7] = g[7];
                if (affine) {
#pragma unroll
                    for (int i = 0; i < 8; i++) t[i] += rec[SR_XN + i];
                } else if ((a >> 1) == k) {
                    const int c = a & 1;
#pragma unroll
                    for (int r = 0; r < 6; r++) t[r] += rec[SR_B + r * 2 + c];
                    if (c == 0) t[7] += cfg.dt; else t[6] += cfg.dt;      // rows delta, a of B: [0 dt], [dt 0]
                }
#pragma unroll
                for (int i = 0; i < 8; i++) Gacc[(k * 8 + i) * ldc + a] += a0 * t[i];
                if (affine && k + 1 < H) {
                    const double *xn = XS + ((size_t)(k + 1) * ns + s) * 8;
#pragma unroll
                    for (int i = 0; i < 8; i++) g[i] = t[i] - xn[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) g[i] = t[i];
                }
            }
        }
    }
    __syncwarp();

    // ---- 3. head block: maps, rows, Hessian and gradient of the eliminated stages
    double *hd = m.head + (size_t)b * m.head_stride;
    const double *Xb = m.X + (size_t)b * NS * 8;
    // constant part of the nominal maps: c_k = sum_s a_s (A c_s + xn_s) - x_k
    if (lane < 8)
        for (int k = 1; k <= H; k++) Gacc[((k - 1) * 8 + lane) * ldc + nh] -= Xb[k * 8 + lane];
    __syncwarp();
    for (int i = lane; i < H * 8 * nh; i += 32) {
        const int row = i / nh, a = i - row * nh;
        hd[hl.G + i] = Gacc[row * ldc + a];
    }
    for (int i = lane; i < H * 8; i += 32) hd[hl.C + i] = Gacc[i * ldc + nh];
    for (int i = lane; i < H * nh; i += 32) {
        const int k = i / nh, a = i - k * nh;
        hd[hl.ROW + i] = Racc[k * ldc + a];
    }
    if (lane < H) hd[hl.V0 + lane] = v0b[lane] + Racc[lane * ldc + nh];
    // Hessian / gradient: stages 1..H-1 state cost through G_k, stages 0..H-1 input cost
    const int nw = cfg.cost_vel_abs ? 3 : 4;
    for (int i = lane; i < nh * nh + nh; i += 32) {
        const bool grad = i >= nh * nh;
        const int a = grad ? i - nh * nh : i / nh, c = grad ? 0 : i - a * nh;
        double acc = 0.0;
        for (int k = 1; k < H; k++) {
            const double *Gk = Gacc + (k - 1) * 8 * ldc;
            const double *rec = m.stg + ((size_t)b * NS + k) * STG;
            double w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) w[j] = (j < nw) ? cfg.dt * stage_val_s(m.W_s, m.W_i, b, NS, k, 6, j) : 0.0;
            double Qv0 = 0.0, Qv1 = 0.0, Qv2 = 0.0;
            if (cfg.cost_vel_abs) { Qv0 = rec[L_QV]; Qv1 = rec[L_QV + 1]; Qv2 = rec[L_QV + 2]; }
            // y = Q_k v with v = G_k[:, c] (Hessian) or c_k with the gradient q_k added (gradient)
            double v[5];
#pragma unroll
            for (int j = 0; j < 5; j++) v[j] = Gk[j * ldc + (grad ? nh : c)];
            double y[5];
#pragma unroll
            for (int j = 0; j < 3; j++) y[j] = w[j] * v[j];
            y[3] = w[3] * v[3] + Qv0 * v[3] + Qv1 * v[4];
            y[4] = Qv1 * v[3] + Qv2 * v[4];
            if (grad) {
#pragma unroll
                for (int j = 0; j < 5; j++) y[j] += rec[L_QX + j];
            }
#pragma unroll
            for (int j = 0; j < 5; j++) acc += Gk[j * ldc + a] * y[j];
        }
        const int ka = a >> 1, ca = a & 1;
        if (grad) {
            acc += m.stg[((size_t)b * NS + ka) * STG + L_RG + ca];
            hd[hl.GH + a] = acc;
        } else {
            if (a == c) acc += cfg.dt * stage_val_s(m.W_s, m.W_i, b, NS, ka, 6, 4 + ca);
            hd[hl.HH + i] = acc;
        }
    }
}

// K3 for the samples: the full step of every sample iterate of stages 0..H.
// dx_s,0 = x0_s - xs_0; dx_s,k+1 = A_s,k dx_s,k + B_s,k du_k + Phi(xs_k, u_k) - xs_k+1 with the head inputs du_k
// of the QP just solved; stage H (the state every later stage freezes, pred_model_dynamic_disc.py:204-210)
// receives the propagated value itself.
// Epilogue of the QP kernel (the caller skips it for a failed QP: iterate left untouched); lane s < n_s owns sample s.
__device__ __forceinline__ void snmpc_step_warp(const DevCfg &cfg, const DevMem &m, int b, int s)
{
    const int H = cfg.head_H, ns = cfg.n_samples;
    if (s >= ns) return;
    double *XS = m.XS + (size_t)b * (m.HA + 1) * ns * 8;
    const double *slin = m.slin + (size_t)b * m.HA * ns * SREC;
    const double *wk = m.stg + (size_t)b * (cfg.N + 1) * STG;             // stage records (W_* offsets are record-relative)
    double dx[8];
#pragma unroll
    for (int i = 0; i < 8; i++) dx[i] = m.x0s[((size_t)b * ns + s) * 8 + i] - XS[(size_t)s * 8 + i];
    for (int k = 0; k < H; k++) {
        const double *rec = slin + ((size_t)k * ns + s) * SREC;
        double *xk = XS + ((size_t)k * ns + s) * 8;
        const double du0 = wk[(size_t)k * STG + W_DU], du1 = wk[(size_t)k * STG + W_DU + 1];
        double t[8];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double v = (r < 2) ? dx[r] : 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) v += rec[SR_A + r * 6 + c] * dx[2 + c];
            t[r] = v + rec[SR_B + r * 2] * du0 + rec[SR_B + r * 2 + 1] * du1 + rec[SR_XN + r];
        }
        t[6] = dx[6] + cfg.dt * du1 + rec[SR_XN + 6];
        t[7] = dx[7] + cfg.dt * du0 + rec[SR_XN + 7];
#pragma unroll
        for (int i = 0; i < 8; i++) xk[i] += dx[i];
        double *xn = XS + ((size_t)(k + 1) * ns + s) * 8;
        if (k + 1 < H) {
#pragma unroll
            for (int i = 0; i < 8; i++) dx[i] = t[i] - xn[i];
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) xn[i] = t[i];
        }
    }
}

}  // namespace b200
