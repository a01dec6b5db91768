// nmpc_model.cuh -- device-side model of the 8-state dynamic single-track Pacejka vehicle and the
// quantities the RTI step needs from it.  Hand-written for sm_100a (fp64 throughout: the reference
// computes in double and parity is stated at 1e-6).
//
// Reference being replaced (paths relative to /root/reference):
//   dynamics      Prediction_Models/pred_model_dynamic_stm_pacejka.py:118-177
//   integrator    Model_Predictive_Controller/Nominal_NMPC/NMPC_STM_acados_settings.py:238-240 (ERK 4x3)
//   constraint h  NMPC_STM_acados_settings.py:70-74,107-109 + NMPC_class.py:322-335 (ggv interpolant)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/b200nmpc.h"

namespace b200 {

// Device copy of the configuration with derived constants folded once on the host.
struct DevCfg {
    double lf, lr, inv_m, inv_Iz, m;
    double Bf, Cf, Df, Ef, Br, Cr, Dr, Er;
    double fr0, fr1, fr4;
    double Fz_f, Fz_r, inv_Fmax_f, inv_Fmax_r, kaero;
    double ax_brake;
    double ggv_v[B200NMPC_MAX_GGV], ggv_ax[B200NMPC_MAX_GGV], ggv_ay[B200NMPC_MAX_GGV];
    double dt;
    double ipm_tol, ipm_reg, ipm_eta_max, ipm_s0;
    double r2_sigma0_diag[8], r2_wd[4], delta_max;
    int n_ggv, N, rk_steps, variant, ipm_max_iter, qp_no_shortcut, uph, n_samples;
    int cost_vel_abs, ipm_corr_mode;
    int head_H;        // SNMPC: stages 0..head_H-1 propagate the samples (stop_flag == 0); 0 = no head
    int sn_L;          // SNMPC: PCE terms
    double sn_kappa;   // SNMPC: sqrt((1-gamma)/gamma)
    int B;             // batch (leading dimension of every SoA array)
};

struct Tire { double Fy, dFy; };

// Reciprocal / reciprocal square root of the model evaluation.  K1 (k_linearize.cu defines NMPC_MODEL_FAST_RECIP) uses the
// branch-free form: hardware seed (2^-23) + two Newton steps, ~1 ulp -- `1.0 / x` and `sqrt(x)` compile to the same seed and
// refinement PLUS a divergent slow-path call (BSSY / BRA / BSYNC) for denormal / huge operands, 13 of them per evaluation of f.
// The plant of the closed loop (k_sim.cu) keeps the IEEE operations.
#ifdef NMPC_MODEL_FAST_RECIP
__device__ __forceinline__ double mdl_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double mdl_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx, y * y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx, y * y, 0.5);
    return fma(y, e, y);
}
#else
__device__ __forceinline__ double mdl_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ double mdl_rsqrt(double x) { return 1.0 / sqrt(x); }
#endif

// Pacejka magic formula D sin(C atan(B a - E (B a - atan(B a)))) and its slope (:152-153)
__device__ __forceinline__ Tire pacejka(double B, double C, double D, double E, double alpha)
{
    double ba = B * alpha;
    double at = atan(ba);
    double phi = ba - E * (ba - at);
    double dphi = B - E * (B - B * mdl_rcp(1.0 + ba * ba));
    double s, c;
    sincos(C * atan(phi), &s, &c);
    Tire t;
    t.Fy = D * s;
    t.dFy = D * c * C * mdl_rcp(1.0 + phi * phi) * dphi;
    return t;
}

// Jacobian entries of one evaluation of f, packed for the structured sensitivity propagation:
//   rows 0,1 (px, py) depend on (yaw, vl, vt):  j02 = -f1, j12 = f0, [j03 j04; j13 j14] = [c -s; s c]
//   row 2 (yaw) = e_r
//   rows 3..5 x cols (vl, vt, r, delta, a): J[3][5]
struct StageJac {
    double sy, cy, f0, f1;
    double J[3][5];
};

// xdot[0..5] and the packed Jacobian at (x, u).  xdot[6] = u[1], xdot[7] = u[0] are handled by callers.
template <bool WITH_JAC>
__device__ __forceinline__ void stm_eval(const DevCfg &M, const double *x, double *f, StageJac *sj)
{
    const double vl = x[3], vt = x[4], r = x[5], de = x[6], a = x[7];
    double sy, cy;
    sincos(x[2], &sy, &cy);
    f[0] = vl * cy - vt * sy;                       // :165
    f[1] = vl * sy + vt * cy;                       // :166
    f[2] = r;                                       // :167
    // rolling resistance (:118-124)
#ifdef NMPC_MODEL_FAST_RECIP
    const double v2 = vl * vl + vt * vt;
    const double iV = v2 > 0.0 ? mdl_rsqrt(v2) : 0.0;
    const double V = v2 * iV;
#else
    const double V = sqrt(vl * vl + vt * vt);
#endif
    const double v100 = V * 3.6 / 100.0;
    const double v100_3 = v100 * v100 * v100;
    const double fr = M.fr0 + M.fr1 * v100 + M.fr4 * (v100_3 * v100);
    const double Fx_f = -fr * M.Fz_f;               // :144
    const double Fx_r = M.m * a - fr * M.Fz_r;      // :145
    const double Faero = M.kaero * vl * vl;         // :134
    // slip angles (:148-149)
    double af = 0.0, ar = 0.0, qf = 0.0, qr = 0.0, cf = 0.0, cr = 0.0, ivl = 0.0;
    const bool moving = vl > 0.001;
    if (moving) {
        ivl = mdl_rcp(vl);
        qf = (vt + M.lf * r) * ivl;
        qr = (M.lr * r - vt) * ivl;
        af = de - atan(qf);
        ar = atan(qr);
    }
    const Tire tf = pacejka(M.Bf, M.Cf, M.Df, M.Ef, af);
    const Tire tr = pacejka(M.Br, M.Cr, M.Dr, M.Er, ar);
    // combined slip (:156-161): cos(asin(G)) = sqrt(1 - G^2), G clipped to +-0.98
    double Gf = Fx_f * M.inv_Fmax_f, Gr = Fx_r * M.inv_Fmax_r;
    const bool in_f = (Gf <= 0.98) && (Gf >= -0.98), in_r = (Gr <= 0.98) && (Gr >= -0.98);
    Gf = fmin(fmax(Gf, -0.98), 0.98);
    Gr = fmin(fmax(Gr, -0.98), 0.98);
#ifdef NMPC_MODEL_FAST_RECIP
    const double icsf = mdl_rsqrt(1.0 - Gf * Gf), icsr = mdl_rsqrt(1.0 - Gr * Gr);
    const double csf = (1.0 - Gf * Gf) * icsf, csr = (1.0 - Gr * Gr) * icsr;
#else
    const double csf = sqrt(1.0 - Gf * Gf), csr = sqrt(1.0 - Gr * Gr);
    const double icsf = 1.0 / csf, icsr = 1.0 / csr;      // (used by the Jacobian only)
#endif
    const double Fy_f = tf.Fy * csf, Fy_r = tr.Fy * csr;
    double sd, cd;
    sincos(de, &sd, &cd);
    const double lat = Fy_f * cd + Fx_f * sd;
    f[3] = (Fx_r - Faero - Fy_f * sd + Fx_f * cd) * M.inv_m + vt * r;      // :169
    f[4] = (Fy_r + lat) * M.inv_m - vl * r;                                // :170
    f[5] = (M.lf * lat - M.lr * Fy_r) * M.inv_Iz;                          // :171
    if (WITH_JAC) {
        sj->sy = sy; sj->cy = cy; sj->f0 = f[0]; sj->f1 = f[1];
        const double dfr_dV = (3.6 / 100.0) * (M.fr1 + 4.0 * M.fr4 * v100_3);
#ifndef NMPC_MODEL_FAST_RECIP
        const double iV = V > 0.0 ? 1.0 / V : 0.0;
#endif
        const double dV0 = vl * iV, dV1 = vt * iV;
        // d Fx_f, d Fx_r w.r.t. (vl, vt) through fr; d Fx_r / d a = m
        const double dFxf0 = -M.Fz_f * dfr_dV * dV0, dFxf1 = -M.Fz_f * dfr_dV * dV1;
        const double dFxr0 = -M.Fz_r * dfr_dV * dV0, dFxr1 = -M.Fz_r * dfr_dV * dV1;
        double daf0 = 0, daf1 = 0, daf2 = 0, dar0 = 0, dar1 = 0, dar2 = 0;
        if (moving) {
            cf = mdl_rcp(1.0 + qf * qf);
            cr = mdl_rcp(1.0 + qr * qr);
            daf0 = cf * qf * ivl;  daf1 = -cf * ivl;  daf2 = -cf * M.lf * ivl;
            dar0 = -cr * qr * ivl; dar1 = -cr * ivl;  dar2 = cr * M.lr * ivl;
        }
        const double dcsf = in_f ? -Gf * icsf * M.inv_Fmax_f : 0.0;       // d csf / d Fx_f
        const double dcsr = in_r ? -Gr * icsr * M.inv_Fmax_r : 0.0;
        const double a_f = tf.dFy * csf, b_f = tf.Fy * dcsf;
        const double a_r = tr.dFy * csr, b_r = tr.Fy * dcsr;
        // d Fy_f, d Fy_r w.r.t. (vl, vt, r, delta, a)
        const double dFyf[5] = {a_f * daf0 + b_f * dFxf0, a_f * daf1 + b_f * dFxf1, a_f * daf2, moving ? a_f : 0.0, 0.0};
        const double dFyr[5] = {a_r * dar0 + b_r * dFxr0, a_r * dar1 + b_r * dFxr1, a_r * dar2, 0.0, b_r * M.m};
        const double dFxf[5] = {dFxf0, dFxf1, 0.0, 0.0, 0.0};
        const double dFxr[5] = {dFxr0, dFxr1, 0.0, 0.0, M.m};
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const double dlat = dFyf[i] * cd + dFxf[i] * sd;
            sj->J[0][i] = (dFxr[i] - dFyf[i] * sd + dFxf[i] * cd) * M.inv_m;
            sj->J[1][i] = (dFyr[i] + dlat) * M.inv_m;
            sj->J[2][i] = (M.lf * dlat - M.lr * dFyr[i]) * M.inv_Iz;
        }
        sj->J[0][0] -= 2.0 * M.kaero * vl * M.inv_m;
        sj->J[0][1] += r;  sj->J[0][2] += vt;
        sj->J[1][0] -= r;  sj->J[1][2] -= vl;
        const double dlat_d = -Fy_f * sd + Fx_f * cd;
        sj->J[0][3] += (-Fy_f * cd - Fx_f * sd) * M.inv_m;
        sj->J[1][3] += dlat_d * M.inv_m;
        sj->J[2][3] += M.lf * dlat_d * M.inv_Iz;
    }
}

// casadi.interpolant('linear'): piecewise linear, end intervals extrapolated (NMPC_class.py:332-333)
__device__ __forceinline__ void interp_lin(int n, const double *xs, const double *ys, double x, double &y, double &dy)
{
    int i = 0;
    while (i < n - 2 && x >= xs[i + 1]) i++;
    const double sl = (ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]);
    y = ys[i] + sl * (x - xs[i]);
    dy = sl;
}

// h = (a/ax_max)^2 + (vl r/ay_max)^2 and its gradient w.r.t. (vl, vt, r, a)
__device__ __forceinline__ double h_eval(const DevCfg &M, const double *x, double *g4)
{
    const double vl = x[3], vt = x[4], r = x[5], a = x[7];
    double v = vl, dv0 = 1.0, dv1 = 0.0;
    if (M.cost_vel_abs) {                       // SNMPC_acados_settings.py:62-65,92-94
        v = sqrt(vl * vl + vt * vt);
        const double iv = v > 0.0 ? 1.0 / v : 0.0;
        dv0 = vl * iv; dv1 = vt * iv;
    }
    double ax, dax, ay, day;
    interp_lin(M.n_ggv, M.ggv_v, M.ggv_ax, v, ax, dax);
    interp_lin(M.n_ggv, M.ggv_v, M.ggv_ay, v, ay, day);
    if (a < 0.0) { ax = M.ax_brake; dax = 0.0; }          // NMPC_STM_acados_settings.py:73
    const double alat = vl * r;
    const double ra = a / ax, rl = alat / ay;
    const double h = ra * ra + rl * rl;
    const double dh_dv = -2.0 * ra * ra / ax * dax - 2.0 * rl * rl / ay * day;
    const double dh_dal = 2.0 * rl / ay;
    g4[0] = dh_dal * r + dh_dv * dv0;
    g4[1] = dh_dv * dv1;
    g4[2] = dh_dal * vl;
    g4[3] = 2.0 * ra / ax;
    return h;
}

// wrap = fmod(yaw, 2pi) shifted to [0, 2pi)  (NMPC_STM_acados_settings.py:41-42)
__device__ __forceinline__ double wrap_yaw(double yaw)
{
    const double two_pi = 6.283185307179586476925286766559;
    double y = fmod(yaw, two_pi);
    if (y < 0.0) y += two_pi;
    return y;
}

}  // namespace b200
