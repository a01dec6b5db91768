/*
 * nmpc_oracle.c -- CPU ORACLE (test infrastructure, NOT the product; see nmpc_oracle.h).
 *
 * Plain-C restatement of the RTI-SQP step acados performs for TUM-CONTROL's NMPC.
 * Every function cites the reference file:line it follows.  Paths are relative to
 * /root/reference (bzarr/TUM-CONTROL @ f59dd020).
 *
 * Parity status: PINNED for the nominal controller by the reference's committed logs
 * Learning_To_Adapt/SafeRL_WMPC/_baseline/F/{monteblanco,lvms}/k.npz
 * (tests/test_oracle_golden.py replays them teacher-forced).  The SNMPC model is pinned
 * numerically against the expressions committed in acados_ocp_SNMPC.json (snmpc_oracle.c).
 * The R2NMPC tightening is "parity unpinned" by data (the reference commits no outputs of
 * it); it is pinned through its exact reduction to the nominal OCP.
 */
#include "nmpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdio.h>
int orc_verbose = 0;

#define NX ORC_NX
#define NU ORC_NU
#define NC ORC_NC

/* ------------------------------------------------------------------------------------------
 * Dynamics  (Prediction_Models/pred_model_dynamic_stm_pacejka.py:118-177)
 * x = [px, py, yaw, vlong, vlat, yawrate, delta_f, a],  u = [jerk, steering_rate]
 * ---------------------------------------------------------------------------------------- */

/* Pacejka lateral force D sin(C atan(B a - E (B a - atan(B a)))) and d/d alpha (:152-153) */
static void pacejka(double B, double C, double D, double E, double alpha, double *Fy, double *dFy)
{
    double ba = B * alpha;
    double at = atan(ba);
    double phi = ba - E * (ba - at);
    double dphi = B - E * (B - B / (1.0 + ba * ba));
    double ap = C * atan(phi);
    *Fy = D * sin(ap);
    *dFy = D * cos(ap) * C / (1.0 + phi * phi) * dphi;
}

/* core: xdot and the 3x5 block d(xdot[3..5])/d(vl, vt, r, delta, a) */
static void stm_core(const orc_model *M, double vl, double vt, double r, double de, double a,
                     double *f3, double *f4, double *f5, double J[3][5])
{
    const double lf = M->lf, lr = M->lr, m = M->m, Iz = M->Iz;
    /* rolling resistance (:118-124) */
    double V = sqrt(vl * vl + vt * vt);
    double v100 = V * 3.6 / 100.0;
    double fr = M->fr0 + M->fr1 * v100 + M->fr4 * (v100 * v100 * v100 * v100);
    double dfr_dV = (3.6 / 100.0) * (M->fr1 + 4.0 * M->fr4 * v100 * v100 * v100);
    double dV[5] = {0, 0, 0, 0, 0};
    if (V > 0.0) { dV[0] = vl / V; dV[1] = vt / V; }
    double Fz_f = m * lr * M->g / (lf + lr);
    double Fz_r = m * lf * M->g / (lf + lr);
    /* longitudinal forces (:135-145); aero (:134) */
    double Fx_f = -fr * Fz_f;
    double Fx_r = m * a - fr * Fz_r;
    double kaero = 0.5 * M->ro * M->S * M->Cd;
    double Faero = kaero * vl * vl;
    double dFxf[5], dFxr[5];
    for (int i = 0; i < 5; i++) {
        dFxf[i] = -Fz_f * dfr_dV * dV[i];
        dFxr[i] = -Fz_r * dfr_dV * dV[i];
    }
    dFxr[4] += m;
    /* slip angles (:148-149); if_else has zero derivative w.r.t. the condition */
    double af = 0.0, ar = 0.0, daf[5] = {0, 0, 0, 0, 0}, dar[5] = {0, 0, 0, 0, 0};
    if (vl > 0.001) {
        double qf = (vt + lf * r) / vl;
        double qr = (lr * r - vt) / vl;
        double cf = 1.0 / (1.0 + qf * qf), cr = 1.0 / (1.0 + qr * qr);
        af = de - atan(qf);
        ar = atan(qr);
        daf[0] = cf * qf / vl;  daf[1] = -cf / vl;  daf[2] = -cf * lf / vl;  daf[3] = 1.0;
        dar[0] = -cr * qr / vl; dar[1] = -cr / vl;  dar[2] = cr * lr / vl;
    }
    /* lateral forces (:152-153) */
    double Fyl_f, dFyl_f, Fyl_r, dFyl_r;
    pacejka(M->Bf, M->Cf, M->Df, M->Ef, af, &Fyl_f, &dFyl_f);
    pacejka(M->Br, M->Cr, M->Dr, M->Er, ar, &Fyl_r, &dFyl_r);
    /* combined slip (:156-161): Fy = Fy_lat * cos(asin(clip(Fx/Fmax, +-0.98))) */
    double Fmax_f = sqrt(Fz_f * Fz_f + (M->Cf * Fz_f) * (M->Cf * Fz_f));
    double Fmax_r = sqrt(Fz_r * Fz_r + (M->Cr * Fz_r) * (M->Cr * Fz_r));
    double Gf = Fx_f / Fmax_f, Gr = Fx_r / Fmax_r;
    int inf_ = 1, inr = 1;
    if (Gf > 0.98) { Gf = 0.98; inf_ = 0; } else if (Gf < -0.98) { Gf = -0.98; inf_ = 0; }
    if (Gr > 0.98) { Gr = 0.98; inr = 0; } else if (Gr < -0.98) { Gr = -0.98; inr = 0; }
    double csf = sqrt(1.0 - Gf * Gf), csr = sqrt(1.0 - Gr * Gr);
    double dcsf = inf_ ? -Gf / csf / Fmax_f : 0.0;      /* d cs / d Fx */
    double dcsr = inr ? -Gr / csr / Fmax_r : 0.0;
    double Fy_f = Fyl_f * csf, Fy_r = Fyl_r * csr;
    double dFyf[5], dFyr[5];
    for (int i = 0; i < 5; i++) {
        dFyf[i] = dFyl_f * daf[i] * csf + Fyl_f * dcsf * dFxf[i];
        dFyr[i] = dFyl_r * dar[i] * csr + Fyl_r * dcsr * dFxr[i];
    }
    double sd = sin(de), cd = cos(de);
    /* state derivatives (:169-171); banking = 0 */
    *f3 = (Fx_r - Faero - Fy_f * sd + Fx_f * cd) / m + vt * r;
    *f4 = (Fy_r + Fy_f * cd + Fx_f * sd) / m - vl * r;
    *f5 = (lf * (Fy_f * cd + Fx_f * sd) - lr * Fy_r) / Iz;
    if (J) {
        for (int i = 0; i < 5; i++) {
            double dlat = dFyf[i] * cd + dFxf[i] * sd;
            J[0][i] = (dFxr[i] - dFyf[i] * sd + dFxf[i] * cd) / m;
            J[1][i] = (dFyr[i] + dlat) / m;
            J[2][i] = (lf * dlat - lr * dFyr[i]) / Iz;
        }
        J[0][0] -= 2.0 * kaero * vl / m;
        J[0][1] += r;  J[0][2] += vt;
        J[1][0] -= r;  J[1][2] -= vl;
        J[0][3] += (-Fy_f * cd - Fx_f * sd) / m;
        J[1][3] += (-Fy_f * sd + Fx_f * cd) / m;
        J[2][3] += lf * (-Fy_f * sd + Fx_f * cd) / Iz;
    }
}

void orc_f(const orc_model *M, const double *x, const double *u, double *xdot)
{
    double sy = sin(x[2]), cy = cos(x[2]);
    xdot[0] = x[3] * cy - x[4] * sy;      /* :165 */
    xdot[1] = x[3] * sy + x[4] * cy;      /* :166 */
    xdot[2] = x[5];                       /* :167 */
    stm_core(M, x[3], x[4], x[5], x[6], x[7], &xdot[3], &xdot[4], &xdot[5], NULL);
    xdot[6] = u[1];                       /* :172 delta_f_dot = steering_rate */
    xdot[7] = u[0];                       /* :173 a_dot = jerk */
}

void orc_f_jac(const orc_model *M, const double *x, const double *u, double *xdot, double *Jx,
               double *Ju)
{
    double J[3][5];
    double sy = sin(x[2]), cy = cos(x[2]);
    xdot[0] = x[3] * cy - x[4] * sy;
    xdot[1] = x[3] * sy + x[4] * cy;
    xdot[2] = x[5];
    stm_core(M, x[3], x[4], x[5], x[6], x[7], &xdot[3], &xdot[4], &xdot[5], J);
    xdot[6] = u[1];
    xdot[7] = u[0];
    memset(Jx, 0, 64 * sizeof(double));
    memset(Ju, 0, 16 * sizeof(double));
    Jx[0 * 8 + 2] = -x[3] * sy - x[4] * cy;  Jx[0 * 8 + 3] = cy;  Jx[0 * 8 + 4] = -sy;
    Jx[1 * 8 + 2] = x[3] * cy - x[4] * sy;   Jx[1 * 8 + 3] = sy;  Jx[1 * 8 + 4] = cy;
    Jx[2 * 8 + 5] = 1.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 5; j++) Jx[(3 + i) * 8 + 3 + j] = J[i][j];
    Ju[6 * 2 + 1] = 1.0;
    Ju[7 * 2 + 0] = 1.0;
}

/* ------------------------------------------------------------------------------------------
 * Integrator: ERK, 4 stages, num_steps sub-steps, forward sensitivities
 * (NMPC_STM_acados_settings.py:238-240; acados sim_erk with forward VDE)
 * ---------------------------------------------------------------------------------------- */
static void vde(const orc_model *M, const double *x, const double *u, const double *S /*8x10*/,
                double *k, double *kS /*8x10*/)
{
    double Jx[64], Ju[16];
    orc_f_jac(M, x, u, k, Jx, Ju);
    for (int i = 0; i < 8; i++)
        for (int c = 0; c < 10; c++) {
            double acc = (c >= 8) ? Ju[i * 2 + (c - 8)] : 0.0;
            for (int j = 0; j < 8; j++) acc += Jx[i * 8 + j] * S[j * 10 + c];
            kS[i * 10 + c] = acc;
        }
}

void orc_integrate(const orc_model *M, const double *x, const double *u, double *xn, double *A,
                   double *B)
{
    double X[8], S[80];
    memcpy(X, x, sizeof X);
    memset(S, 0, sizeof S);
    for (int i = 0; i < 8; i++) S[i * 10 + i] = 1.0;
    int ns = M->rk_steps;
    double h = M->dt / ns;
    for (int s = 0; s < ns; s++) {
        double k1[8], k2[8], k3[8], k4[8], S1[80], S2[80], S3[80], S4[80], Xt[8], St[80];
        vde(M, X, u, S, k1, S1);
        for (int i = 0; i < 8; i++) Xt[i] = X[i] + 0.5 * h * k1[i];
        for (int i = 0; i < 80; i++) St[i] = S[i] + 0.5 * h * S1[i];
        vde(M, Xt, u, St, k2, S2);
        for (int i = 0; i < 8; i++) Xt[i] = X[i] + 0.5 * h * k2[i];
        for (int i = 0; i < 80; i++) St[i] = S[i] + 0.5 * h * S2[i];
        vde(M, Xt, u, St, k3, S3);
        for (int i = 0; i < 8; i++) Xt[i] = X[i] + h * k3[i];
        for (int i = 0; i < 80; i++) St[i] = S[i] + h * S3[i];
        vde(M, Xt, u, St, k4, S4);
        for (int i = 0; i < 8; i++) X[i] += h / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
        for (int i = 0; i < 80; i++) S[i] += h / 6.0 * (S1[i] + 2.0 * S2[i] + 2.0 * S3[i] + S4[i]);
    }
    memcpy(xn, X, sizeof X);
    if (A)
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) A[i * 8 + j] = S[i * 10 + j];
    if (B)
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 2; j++) B[i * 2 + j] = S[i * 10 + 8 + j];
}

void orc_flow(const orc_model *M, const double *x, const double *u, double T, int n_steps,
              double *xn)
{
    double X[8];
    memcpy(X, x, sizeof X);
    double h = T / n_steps;
    for (int s = 0; s < n_steps; s++) {
        double k1[8], k2[8], k3[8], k4[8], Xt[8];
        orc_f(M, X, u, k1);
        for (int i = 0; i < 8; i++) Xt[i] = X[i] + 0.5 * h * k1[i];
        orc_f(M, Xt, u, k2);
        for (int i = 0; i < 8; i++) Xt[i] = X[i] + 0.5 * h * k2[i];
        orc_f(M, Xt, u, k3);
        for (int i = 0; i < 8; i++) Xt[i] = X[i] + h * k3[i];
        orc_f(M, Xt, u, k4);
        for (int i = 0; i < 8; i++) X[i] += h / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    }
    memcpy(xn, X, sizeof X);
}

/* 7-state plant: same STM with inputs (a, steering_rate), RK4 with nsub finite elements
 * (Vehicle_Simulator/VehicleSimulator.py:60-94, sim_model_dynamic_stm_pacejka.py:137-196) */
void orc_plant_step(const orc_model *Ms, const double *x7, const double *u2, const double *w7,
                    double Ts, int nsub, double *x7n)
{
    double X[7];
    memcpy(X, x7, sizeof X);
    double h = Ts / nsub;
    for (int s = 0; s < nsub; s++) {
        double k[4][7], Xt[7];
        for (int st = 0; st < 4; st++) {
            const double *xe = X;
            if (st > 0) {
                double c = (st == 3) ? h : 0.5 * h;
                for (int i = 0; i < 7; i++) Xt[i] = X[i] + c * k[st - 1][i];
                xe = Xt;
            }
            double sy = sin(xe[2]), cy = cos(xe[2]);
            k[st][0] = xe[3] * cy - xe[4] * sy;
            k[st][1] = xe[3] * sy + xe[4] * cy;
            k[st][2] = xe[5];
            stm_core(Ms, xe[3], xe[4], xe[5], xe[6], u2[0], &k[st][3], &k[st][4], &k[st][5], NULL);
            k[st][6] = u2[1];
            if (w7)
                for (int i = 0; i < 7; i++) k[st][i] += w7[i];
        }
        for (int i = 0; i < 7; i++)
            X[i] += h / 6.0 * (k[0][i] + 2.0 * k[1][i] + 2.0 * k[2][i] + k[3][i]);
    }
    memcpy(x7n, X, sizeof X);
}

/* ------------------------------------------------------------------------------------------
 * g-g-v lookup: casadi.interpolant('linear') (NMPC_class.py:322-335) -- piecewise linear,
 * linearly extrapolated from the end intervals.
 * ---------------------------------------------------------------------------------------- */
static void interp_lin(int n, const double *xs, const double *ys, double x, double *y, double *dy)
{
    int i = 0;
    while (i < n - 2 && x >= xs[i + 1]) i++;
    double sl = (ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]);
    *y = ys[i] + sl * (x - xs[i]);
    *dy = sl;
}

/* h = (a/ax_max)^2 + (vlong*yawrate/ay_max)^2   (NMPC_STM_acados_settings.py:70-74,107-109) */
double orc_h(const orc_model *M, const double *x, double *grad)
{
    double vl = x[3], vt = x[4], r = x[5], a = x[7];
    double v, dv_dvl, dv_dvt;
    if (M->cost_vel_abs) {   /* SNMPC_acados_settings.py:62-65,92-94 */
        v = sqrt(vl * vl + vt * vt);
        dv_dvl = v > 0 ? vl / v : 0.0;
        dv_dvt = v > 0 ? vt / v : 0.0;
    } else {
        v = vl; dv_dvl = 1.0; dv_dvt = 0.0;
    }
    double ax, dax, ay, day;
    interp_lin(M->n_ggv, M->ggv_v, M->ggv_ax, v, &ax, &dax);
    interp_lin(M->n_ggv, M->ggv_v, M->ggv_ay, v, &ay, &day);
    if (a < 0.0) { ax = M->ax_brake; dax = 0.0; }
    double alat = vl * r;
    double h = (a / ax) * (a / ax) + (alat / ay) * (alat / ay);
    if (grad) {
        double dh_dv = -2.0 * a * a / (ax * ax * ax) * dax - 2.0 * alat * alat / (ay * ay * ay) * day;
        double dh_dal = 2.0 * alat / (ay * ay);
        for (int i = 0; i < 8; i++) grad[i] = 0.0;
        grad[3] = dh_dal * r + dh_dv * dv_dvl;
        grad[4] = dh_dv * dv_dvt;
        grad[5] = dh_dal * vl;
        grad[7] = 2.0 * a / (ax * ax);
    }
    return h;
}

/* ------------------------------------------------------------------------------------------
 * NLS output  y = [px, py, wrap(yaw), v, jerk, steering_rate]
 * (NMPC_STM_acados_settings.py:41-52).  wrap = fmod(yaw, 2pi) shifted to [0, 2pi).
 * ---------------------------------------------------------------------------------------- */
static double wrap_yaw(double yaw)
{
    double y = fmod(yaw, 2.0 * M_PI);
    if (y < 0.0) y += 2.0 * M_PI;
    return y;
}

static void set_active(int N, int k, int *act)
{
    /* stage 0: bu only (ns_0=1); 1..N-1: bu,bx,h (ns=3); N: bx,h (ns_e=2)
     * (NMPC_STM_acados_settings.py:126-139,164-224; acados_ocp_SNMPC.json dims nh_0=0) */
    act[0] = (k < N);
    act[1] = (k >= 1);
    act[2] = (k >= 1);
}

void orc_linearize(const orc_model *M, const double *x0, const double *X, const double *U,
                   const double *yref, const double *W, const double *lb, const double *ub,
                   const double *zl, const double *zu, const double *Zl, const double *Zu,
                   orc_qp *qp)
{
    int N = M->N;
    qp->N = N;
    for (int i = 0; i < 8; i++) qp->dx0[i] = x0[i] - X[i];
    for (int k = 0; k <= N; k++) {
        const double *x = X + 8 * k;
        double sig = (k < N) ? M->dt : 1.0;       /* cost_discretization EULER: dt-scaled stages */
        if (k < N) {
            double xn[8];
            orc_integrate(M, x, U + 2 * k, xn, qp->A[k], qp->B[k]);
            for (int i = 0; i < 8; i++) qp->b[k][i] = xn[i] - X[8 * (k + 1) + i];
        }
        /* Gauss-Newton cost: H = sig J'WJ, g = sig J'W (y - yref) */
        const double *w = W + 6 * k;
        const double *yr = yref + 6 * k;
        memset(qp->Q[k], 0, sizeof qp->Q[k]);
        memset(qp->q[k], 0, sizeof qp->q[k]);
        qp->Q[k][0 * 8 + 0] = sig * w[0];  qp->q[k][0] = sig * w[0] * (x[0] - yr[0]);
        qp->Q[k][1 * 8 + 1] = sig * w[1];  qp->q[k][1] = sig * w[1] * (x[1] - yr[1]);
        qp->Q[k][2 * 8 + 2] = sig * w[2];  qp->q[k][2] = sig * w[2] * (wrap_yaw(x[2]) - yr[2]);
        if (M->cost_vel_abs) {
            double V = sqrt(x[3] * x[3] + x[4] * x[4]);
            double g3 = V > 0 ? x[3] / V : 0.0, g4 = V > 0 ? x[4] / V : 0.0;
            qp->Q[k][3 * 8 + 3] = sig * w[3] * g3 * g3;
            qp->Q[k][3 * 8 + 4] = qp->Q[k][4 * 8 + 3] = sig * w[3] * g3 * g4;
            qp->Q[k][4 * 8 + 4] = sig * w[3] * g4 * g4;
            qp->q[k][3] = sig * w[3] * (V - yr[3]) * g3;
            qp->q[k][4] = sig * w[3] * (V - yr[3]) * g4;
        } else {
            qp->Q[k][3 * 8 + 3] = sig * w[3];
            qp->q[k][3] = sig * w[3] * (x[3] - yr[3]);
        }
        if (k < N) {
            const double *u = U + 2 * k;
            qp->R[k][0] = sig * w[4];  qp->R[k][1] = 0.0;  qp->R[k][2] = 0.0;  qp->R[k][3] = sig * w[5];
            qp->r[k][0] = sig * w[4] * (u[0] - yr[4]);
            qp->r[k][1] = sig * w[5] * (u[1] - yr[5]);
        }
        /* soft rows [bu, bx, h] */
        set_active(N, k, qp->act[k]);
        qp->cval[k][0] = (k < N) ? U[2 * k + 1] : 0.0;
        qp->cval[k][1] = x[6];
        qp->cval[k][2] = orc_h(M, x, qp->gh[k]);
        for (int c = 0; c < NC; c++) {
            qp->lo[k][c] = lb[3 * k + c];  qp->up[k][c] = ub[3 * k + c];
            qp->zl[k][c] = sig * zl[3 * k + c];  qp->zu[k][c] = sig * zu[3 * k + c];
            qp->Zl[k][c] = sig * Zl[3 * k + c];  qp->Zu[k][c] = sig * Zu[3 * k + c];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Structured QP.  acados condenses it and calls HPIPM's dense IPM; the optimum is
 * unique (R > 0, Z > 0), so any exact method reproduces it.  Here: Riccati recursion inside a
 * Mehrotra predictor-corrector primal-dual IPM that starts from the equality-constrained
 * optimum (all linear KKT residuals are zero from then on; only complementarity is driven to 0).
 * ---------------------------------------------------------------------------------------- */
#ifndef ORC_REAL
#define ORC_REAL double      /* -DORC_REAL="long double" builds an extended-precision Riccati for diagnostics */
#endif
typedef ORC_REAL real;
typedef struct ric_ws {
    real K[ORC_MAX_N][16];     /* feedback du = K dx + kff */
    real Hi[ORC_MAX_N][4];     /* (R + B'PB)^-1 */
    real P[ORC_MAX_N + 1][64];
    real p[ORC_MAX_N + 1][8];
    real Qt[ORC_MAX_N + 1][64], Rt[ORC_MAX_N][4];
    real qt[ORC_MAX_N + 1][8], rt[ORC_MAX_N][2];
    real dx[ORC_MAX_N + 1][8], du[ORC_MAX_N][2];
} ric_ws;

/* backward matrix sweep */
static void ric_factor(const orc_qp *qp, ric_ws *w)
{
    int N = qp->N;
    for (int i = 0; i < 64; i++) w->P[N][i] = w->Qt[N][i];
    for (int k = N - 1; k >= 0; k--) {
        const double *A = qp->A[k], *B = qp->B[k]; const real *P = w->P[k + 1];
        real PA[64], PB[16], Huu[4], Hux[16];
        for (int i = 0; i < 8; i++) {
            for (int j = 0; j < 8; j++) {
                real s = 0; for (int l = 0; l < 8; l++) s += P[i * 8 + l] * A[l * 8 + j];
                PA[i * 8 + j] = s;
            }
            for (int j = 0; j < 2; j++) {
                real s = 0; for (int l = 0; l < 8; l++) s += P[i * 8 + l] * B[l * 2 + j];
                PB[i * 2 + j] = s;
            }
        }
        for (int i = 0; i < 2; i++) {
            for (int j = 0; j < 2; j++) {
                real s = w->Rt[k][i * 2 + j];
                for (int l = 0; l < 8; l++) s += B[l * 2 + i] * PB[l * 2 + j];
                Huu[i * 2 + j] = s;
            }
            for (int j = 0; j < 8; j++) {
                real s = 0; for (int l = 0; l < 8; l++) s += B[l * 2 + i] * PA[l * 8 + j];
                Hux[i * 8 + j] = s;
            }
        }
        /* 2x2 Cholesky of Huu, Y = L^-1 Hux, P_k = Hxx - Y'Y, K = -L^-T Y */
        real h01 = 0.5 * (Huu[1] + Huu[2]);
        real l11 = sqrt(Huu[0]), l21 = h01 / l11;
        real l22 = sqrt(Huu[3] - l21 * l21);
        real i11 = 1.0 / l11, i22 = 1.0 / l22;
        real *Hi = w->Hi[k];
        /* Huu^-1 = L^-T L^-1 */
        real m21 = -l21 * i11 * i22;          /* (L^-1)[1][0] */
        Hi[0] = i11 * i11 + m21 * m21;  Hi[1] = Hi[2] = m21 * i22;  Hi[3] = i22 * i22;
        real *K = w->K[k];
        real Y[16];
        for (int j = 0; j < 8; j++) {
            Y[0 * 8 + j] = Hux[0 * 8 + j] * i11;
            Y[1 * 8 + j] = (Hux[1 * 8 + j] - l21 * Y[0 * 8 + j]) * i22;
            K[1 * 8 + j] = -Y[1 * 8 + j] * i22;
            K[0 * 8 + j] = (-Y[0 * 8 + j] - l21 * K[1 * 8 + j]) * i11;
        }
        real *Pk = w->P[k];
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) {
                real s = w->Qt[k][i * 8 + j];
                for (int l = 0; l < 8; l++) s += A[l * 8 + i] * PA[l * 8 + j];
                s -= Y[0 * 8 + i] * Y[0 * 8 + j] + Y[1 * 8 + i] * Y[1 * 8 + j];
                Pk[i * 8 + j] = s;
            }
        for (int i = 0; i < 8; i++)
            for (int j = i + 1; j < 8; j++) {
                real s = 0.5 * (Pk[i * 8 + j] + Pk[j * 8 + i]);
                Pk[i * 8 + j] = Pk[j * 8 + i] = s;
            }
    }
}

/* backward vector sweep + forward rollout.  affine != 0: uses qp->b and dx0. */
static void ric_solve(const orc_qp *qp, ric_ws *w, int affine)
{
    int N = qp->N;
    real kff[ORC_MAX_N][2];
    for (int i = 0; i < 8; i++) w->p[N][i] = w->qt[N][i];
    for (int k = N - 1; k >= 0; k--) {
        const double *A = qp->A[k], *B = qp->B[k]; const real *P = w->P[k + 1];
        real v[8], hu[2], hx[8];
        for (int i = 0; i < 8; i++) {
            real s = w->p[k + 1][i];
            if (affine) for (int l = 0; l < 8; l++) s += P[i * 8 + l] * qp->b[k][l];
            v[i] = s;
        }
        for (int i = 0; i < 2; i++) {
            real s = w->rt[k][i];
            for (int l = 0; l < 8; l++) s += B[l * 2 + i] * v[l];
            hu[i] = s;
        }
        for (int i = 0; i < 8; i++) {
            real s = w->qt[k][i];
            for (int l = 0; l < 8; l++) s += A[l * 8 + i] * v[l];
            hx[i] = s;
        }
        const real *Hi = w->Hi[k], *K = w->K[k];
        kff[k][0] = -(Hi[0] * hu[0] + Hi[1] * hu[1]);
        kff[k][1] = -(Hi[2] * hu[0] + Hi[3] * hu[1]);
        for (int i = 0; i < 8; i++) w->p[k][i] = hx[i] + K[0 * 8 + i] * hu[0] + K[1 * 8 + i] * hu[1];
    }
    for (int i = 0; i < 8; i++) w->dx[0][i] = affine ? qp->dx0[i] : 0.0;
    for (int k = 0; k < N; k++) {
        const double *A = qp->A[k], *B = qp->B[k]; const real *K = w->K[k];
        for (int i = 0; i < 2; i++) {
            real s = kff[k][i];
            for (int l = 0; l < 8; l++) s += K[i * 8 + l] * w->dx[k][l];
            w->du[k][i] = s;
        }
        for (int i = 0; i < 8; i++) {
            real s = affine ? qp->b[k][i] : 0.0;
            for (int l = 0; l < 8; l++) s += A[i * 8 + l] * w->dx[k][l];
            s += B[i * 2 + 0] * w->du[k][0] + B[i * 2 + 1] * w->du[k][1];
            w->dx[k + 1][i] = s;
        }
    }
    if (orc_verbose > 2) {
        double ru = 0, rx = 0, pmax = 0, pimax = 0;
        for (int k = 0; k < N; k++) {
            real pin[8], pik[8];
            for (int i = 0; i < 8; i++) {
                real s = w->p[k + 1][i], s2 = w->p[k][i];
                for (int l = 0; l < 8; l++) { s += w->P[k + 1][i * 8 + l] * w->dx[k + 1][l]; s2 += w->P[k][i * 8 + l] * w->dx[k][l]; }
                pin[i] = s; pik[i] = s2;
                if (fabs((double)s) > pimax) pimax = fabs((double)s);
            }
            for (int i = 0; i < 64; i++) if (fabs((double)w->P[k][i]) > pmax) pmax = fabs((double)w->P[k][i]);
            for (int i = 0; i < 2; i++) {
                real s = w->rt[k][i] + w->Rt[k][i * 2] * w->du[k][0] + w->Rt[k][i * 2 + 1] * w->du[k][1];
                for (int l = 0; l < 8; l++) s += qp->B[k][l * 2 + i] * pin[l];
                if (fabs((double)s) > ru) ru = fabs((double)s);
            }
            if (k > 0) for (int i = 0; i < 8; i++) {
                real s = w->qt[k][i] - pik[i];
                for (int l = 0; l < 8; l++) s += w->Qt[k][i * 8 + l] * w->dx[k][l] + qp->A[k][l * 8 + i] * pin[l];
                if (fabs((double)s) > rx) rx = fabs((double)s);
            }
        }
        printf("      [ric] affine %d  res_u %.3e res_x %.3e  max|P| %.3e max|pi| %.3e\n", affine, ru, rx, pmax, pimax);
    }
}

/* g'w for soft row c of stage k */
static inline double row_dot(const orc_qp *qp, int k, int c, const double *dx, const double *du)
{
    if (c == 0) return du[1];
    if (c == 1) return dx[6];
    double s = 0;
    for (int i = 0; i < 8; i++) s += qp->gh[k][i] * dx[i];
    return s;
}

/* add coef * g (x) to a stage gradient / coef * g g' to a stage Hessian */
static inline void row_add_grad(const orc_qp *qp, int k, int c, double coef, real *qx, real *ru)
{
    if (c == 0) ru[1] += coef;
    else if (c == 1) qx[6] += coef;
    else for (int i = 0; i < 8; i++) qx[i] += coef * qp->gh[k][i];
}
static inline void row_add_hess(const orc_qp *qp, int k, int c, double coef, real *Qx, real *Ru)
{
    if (c == 0) Ru[3] += coef;
    else if (c == 1) Qx[6 * 8 + 6] += coef;
    else
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) Qx[i * 8 + j] += coef * qp->gh[k][i] * qp->gh[k][j];
}

typedef struct ipm_ws {
    /* index [k][c][side], side 0 = lower (sgn +1), 1 = upper (sgn -1) */
    double s[ORC_MAX_N + 1][NC][2], lam[ORC_MAX_N + 1][NC][2];
    double t[ORC_MAX_N + 1][NC][2], mu[ORC_MAX_N + 1][NC][2];
    double ds[ORC_MAX_N + 1][NC][2], dlam[ORC_MAX_N + 1][NC][2];
    double dt[ORC_MAX_N + 1][NC][2], dmu[ORC_MAX_N + 1][NC][2];
    double rc1[ORC_MAX_N + 1][NC][2], rc2[ORC_MAX_N + 1][NC][2];
    double rt[ORC_MAX_N + 1][NC][2], rs[ORC_MAX_N + 1][NC][2];
    double ru[ORC_MAX_N][2];          /* u-stationarity residual (x rows define the costates) */
} ipm_ws;

static double qp_kkt(const orc_qp *qp);

/* Costates from the x-stationarity rows, residual of the u-stationarity rows.  In exact arithmetic the
 * residual stays zero along the iteration (see below); feeding it back into the right-hand side is
 * iterative refinement of the Riccati solves, whose barrier weights reach 1/ipm_reg. */
static double stationarity(const orc_qp *qp, const ipm_ws *ip, int act[][NC], double ru[][2])
{
    int N = qp->N;
    real pin[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double rmax = 0;
    for (int k = N; k >= 0; k--) {
        real gx[8], gu[2] = {0, 0};
        for (int i = 0; i < 8; i++) {
            real s = qp->q[k][i];
            for (int l = 0; l < 8; l++) s += qp->Q[k][i * 8 + l] * qp->dx[k][l];
            gx[i] = s;
        }
        if (k < N) {
            gu[0] = qp->r[k][0] + qp->R[k][0] * qp->du[k][0] + qp->R[k][1] * qp->du[k][1];
            gu[1] = qp->r[k][1] + qp->R[k][2] * qp->du[k][0] + qp->R[k][3] * qp->du[k][1];
        }
        for (int c = 0; c < NC; c++)
            if (act[k][c]) row_add_grad(qp, k, c, -(ip->lam[k][c][0] - ip->lam[k][c][1]), gx, gu);
        if (k < N) {
            for (int i = 0; i < 8; i++)
                for (int l = 0; l < 8; l++) gx[i] += qp->A[k][l * 8 + i] * pin[l];
            for (int i = 0; i < 2; i++) {
                for (int l = 0; l < 8; l++) gu[i] += qp->B[k][l * 2 + i] * pin[l];
                ru[k][i] = (double)gu[i];
                rmax = fmax(rmax, fabs(ru[k][i]));
            }
        }
        for (int i = 0; i < 8; i++) pin[i] = gx[i];
    }
    return rmax;
}

/*
 * Newton system of the perturbed KKT conditions, one soft side (sgn = +1 lower, -1 upper),
 * v = cval + g'w:
 *     t = sgn (v - bound) + s >= 0   (multiplier lam),   s >= 0   (multiplier mu)
 *     r_s = z + Z s - lam - mu,   r_t = t - (sgn (v - bound) + s),   lam t = tau,   mu s = tau
 * with the dual proximal term reg (tt = t + reg*lam replaces t as the pivot of the lam-row, which
 * caps the barrier weight lam/t at 1/reg and vanishes at the fixed point).  Eliminating (ds, dt, dmu):
 *     dlam = cc - om * sgn * dv,
 *     th = lam/tt,  Dp = Z + th + mu/s,  om = th (Z + mu/s) / Dp,
 *     rho = -r_s + rc1/tt + rc2/s + th r_t,   cc = rc1/tt + th r_t - th rho / Dp,
 *     ds = (rho - th sgn dv) / Dp,  dt = sgn dv + ds - r_t + reg dlam,  dmu = (rc2 - mu ds)/s
 * so each side adds  om g g'  to the stage Hessian and the Riccati solve stays 10x10 per stage:
 *     (H + sum om g g') dw + A' dpi = -r_u + sum sgn cc g,   dynamics homogeneous.
 */
void orc_qp_solve(const orc_model *M, orc_qp *qp)
{
    int N = qp->N;
    const double reg = M->ipm_reg;
    ric_ws *w = (ric_ws *)malloc(sizeof(ric_ws));
    ipm_ws *ip = (ipm_ws *)calloc(1, sizeof(ipm_ws));
    int act[ORC_MAX_N + 1][NC];
    int nact = 0;
    for (int k = 0; k <= N; k++)
        for (int c = 0; c < NC; c++) {
            /* a soft row with no penalty at all is vacuous */
            int pen = (qp->zl[k][c] + qp->Zl[k][c] > 0.0) && (qp->zu[k][c] + qp->Zu[k][c] > 0.0);
            act[k][c] = qp->act[k][c] && pen;
            nact += act[k][c];
        }

    /* phase 0: equality-constrained optimum (affine Riccati solve with b and dx0) */
    for (int k = 0; k <= N; k++) {
        for (int i = 0; i < 64; i++) w->Qt[k][i] = qp->Q[k][i];
        for (int i = 0; i < 8; i++) w->qt[k][i] = qp->q[k][i];
        if (k < N) {
            for (int i = 0; i < 4; i++) w->Rt[k][i] = qp->R[k][i];
            for (int i = 0; i < 2; i++) w->rt[k][i] = qp->r[k][i];
        }
    }
    ric_factor(qp, w);
    ric_solve(qp, w, 1);
    for (int k = 0; k <= N; k++) {
        for (int i = 0; i < 8; i++) qp->dx[k][i] = (double)w->dx[k][i];
        if (k < N) for (int i = 0; i < 2; i++) qp->du[k][i] = (double)w->du[k][i];
    }
    memset(qp->sl, 0, sizeof qp->sl);  memset(qp->su, 0, sizeof qp->su);
    memset(qp->lam_l, 0, sizeof qp->lam_l);  memset(qp->lam_u, 0, sizeof qp->lam_u);
    qp->iters = 0;
    qp->status = 0;

    double zero2[2] = {0, 0};
    int feasible = 1;
    for (int k = 0; k <= N; k++)
        for (int c = 0; c < NC; c++) {
            if (!act[k][c]) continue;
            double v = qp->cval[k][c] + row_dot(qp, k, c, qp->dx[k], k < N ? qp->du[k] : zero2);
            if (v < qp->lo[k][c] || v > qp->up[k][c]) feasible = 0;
        }
    if ((feasible && !M->qp_no_shortcut) || nact == 0) {
        /* the equality-constrained optimum satisfies every bound: it is the QP optimum (s = lam = 0) */
        qp->kkt = qp_kkt(qp);
        free(w); free(ip);
        return;
    }

    /* interior start with zero linear residuals: s_l = s_u = S, lam_l = lam_u (they cancel in the
     * stationarity rows, so the phase-0 point stays stationary) */
    for (int k = 0; k <= N; k++)
        for (int c = 0; c < NC; c++) {
            if (!act[k][c]) continue;
            double v = qp->cval[k][c] + row_dot(qp, k, c, qp->dx[k], k < N ? qp->du[k] : zero2);
            double viol = fmax(fmax(qp->lo[k][c] - v, v - qp->up[k][c]), 0.0);
            double S = M->ipm_s0 + viol;
            double gl = qp->zl[k][c] + qp->Zl[k][c] * S, gu = qp->zu[k][c] + qp->Zu[k][c] * S;
            double l0 = 0.5 * fmin(gl, gu);
            ip->s[k][c][0] = ip->s[k][c][1] = S;
            ip->lam[k][c][0] = ip->lam[k][c][1] = l0;
            ip->t[k][c][0] = v - qp->lo[k][c] + S;
            ip->t[k][c][1] = qp->up[k][c] - v + S;
            ip->mu[k][c][0] = gl - l0;
            ip->mu[k][c][1] = gu - l0;
        }

    int it;
    int converged = 0, polish = 0;
    for (it = 0; it <= M->ipm_max_iter; it++) {
        /* residuals and complementarity */
        double gap = 0, cmax = 0, rmax = 0;
        for (int k = 0; k <= N; k++)
            for (int c = 0; c < NC; c++) {
                if (!act[k][c]) continue;
                double v = qp->cval[k][c] + row_dot(qp, k, c, qp->dx[k], k < N ? qp->du[k] : zero2);
                for (int sd = 0; sd < 2; sd++) {
                    double z = sd ? qp->zu[k][c] : qp->zl[k][c], Z = sd ? qp->Zu[k][c] : qp->Zl[k][c];
                    double s = ip->s[k][c][sd], t = ip->t[k][c][sd], lam = ip->lam[k][c][sd], mu = ip->mu[k][c][sd];
                    double c1 = lam * t, c2 = mu * s;
                    gap += c1 + c2;
                    cmax = fmax(cmax, fmax(c1, c2));
                    ip->rt[k][c][sd] = t - ((sd ? qp->up[k][c] - v : v - qp->lo[k][c]) + s);
                    ip->rs[k][c][sd] = z + Z * s - lam - mu;
                    rmax = fmax(rmax, fmax(fabs(ip->rt[k][c][sd]), fabs(ip->rs[k][c][sd])));
                }
            }
        gap /= (4.0 * nact);
        double rstat = stationarity(qp, ip, act, ip->ru);
        if (orc_verbose) printf("it %d gap %.3e cmax %.3e rmax %.3e rstat %.3e\n", it, gap, cmax, rmax, rstat);
        if (!(cmax == cmax) || !(rmax == rmax) || !(rstat == rstat)) { qp->status = 3; break; }
        if (cmax <= M->ipm_tol && rmax <= M->ipm_tol && rstat <= M->ipm_tol) { converged = 1; break; }
        if (it == M->ipm_max_iter) break;
        /* complementarity converged but the linear residuals sit on their round-off floor (badly scaled instances:
         * diverged iterates with large slacks): a few polishing iterations, then stop */
        if (cmax <= M->ipm_tol && ++polish > ORC_POLISH_ITER) break;

        /* barrier Hessian */
        for (int k = 0; k <= N; k++) {
            for (int i = 0; i < 64; i++) w->Qt[k][i] = qp->Q[k][i];
            if (k < N) for (int i = 0; i < 4; i++) w->Rt[k][i] = qp->R[k][i];
            for (int c = 0; c < NC; c++) {
                if (!act[k][c]) continue;
                double om = 0;
                for (int sd = 0; sd < 2; sd++) {
                    double Z = sd ? qp->Zu[k][c] : qp->Zl[k][c];
                    double th = ip->lam[k][c][sd] / (ip->t[k][c][sd] + reg * ip->lam[k][c][sd]);
                    double ms = ip->mu[k][c][sd] / ip->s[k][c][sd];
                    om += th * (Z + ms) / (Z + th + ms);
                }
                row_add_hess(qp, k, c, om, w->Qt[k], k < N ? w->Rt[k] : NULL);
            }
        }
        ric_factor(qp, w);

        double sigma_mu = 0.0, alpha = 1.0, corr = 1.0;
        for (int pass = 0; pass < 2; pass++) {
            /* right-hand side: r_u - sum sgn cc g */
            for (int k = 0; k <= N; k++) {
                for (int i = 0; i < 8; i++) w->qt[k][i] = 0.0;
                if (k < N) for (int i = 0; i < 2; i++) w->rt[k][i] = ip->ru[k][i];
                for (int c = 0; c < NC; c++) {
                    if (!act[k][c]) continue;
                    double coef = 0.0;
                    for (int sd = 0; sd < 2; sd++) {
                        double sgn = sd ? -1.0 : 1.0;
                        double Z = sd ? qp->Zu[k][c] : qp->Zl[k][c];
                        double s = ip->s[k][c][sd], t = ip->t[k][c][sd], lam = ip->lam[k][c][sd], mu = ip->mu[k][c][sd];
                        double tt = t + reg * lam;
                        double th = lam / tt, ms = mu / s, Dp = Z + th + ms;
                        double r1 = -lam * t, r2 = -mu * s;
                        if (pass) {
                            r1 += sigma_mu - corr * ip->dlam[k][c][sd] * ip->dt[k][c][sd];
                            r2 += sigma_mu - corr * ip->dmu[k][c][sd] * ip->ds[k][c][sd];
                        }
                        ip->rc1[k][c][sd] = r1;  ip->rc2[k][c][sd] = r2;
                        double rho = -ip->rs[k][c][sd] + r1 / tt + r2 / s + th * ip->rt[k][c][sd];
                        double cc = r1 / tt + th * ip->rt[k][c][sd] - th * rho / Dp;
                        coef -= sgn * cc;
                    }
                    row_add_grad(qp, k, c, coef, w->qt[k], k < N ? w->rt[k] : NULL);
                }
            }
            ric_solve(qp, w, 0);
            /* recover slack / multiplier directions, step to the boundary */
            double amax = 1e300;
            for (int k = 0; k <= N; k++)
                for (int c = 0; c < NC; c++) {
                    if (!act[k][c]) continue;
                    double ddx[8], ddu[2] = {0, 0};
                    for (int i = 0; i < 8; i++) ddx[i] = (double)w->dx[k][i];
                    if (k < N) { ddu[0] = (double)w->du[k][0]; ddu[1] = (double)w->du[k][1]; }
                    double dv = row_dot(qp, k, c, ddx, ddu);
                    for (int sd = 0; sd < 2; sd++) {
                        double sgn = sd ? -1.0 : 1.0;
                        double Z = sd ? qp->Zu[k][c] : qp->Zl[k][c];
                        double s = ip->s[k][c][sd], t = ip->t[k][c][sd], lam = ip->lam[k][c][sd], mu = ip->mu[k][c][sd];
                        double tt = t + reg * lam;
                        double th = lam / tt, ms = mu / s, Dp = Z + th + ms;
                        double r1 = ip->rc1[k][c][sd], r2 = ip->rc2[k][c][sd];
                        double rho = -ip->rs[k][c][sd] + r1 / tt + r2 / s + th * ip->rt[k][c][sd];
                        double ds = (rho - th * sgn * dv) / Dp;
                        double dl = (r1 - lam * (sgn * dv + ds - ip->rt[k][c][sd])) / tt;
                        double dtt = sgn * dv + ds - ip->rt[k][c][sd] + reg * dl;
                        double dm = (r2 - mu * ds) / s;
                        ip->ds[k][c][sd] = ds;  ip->dlam[k][c][sd] = dl;
                        ip->dt[k][c][sd] = dtt; ip->dmu[k][c][sd] = dm;
                        if (ds < 0) amax = fmin(amax, -s / ds);
                        if (dl < 0) amax = fmin(amax, -lam / dl);
                        if (dtt < 0) amax = fmin(amax, -t / dtt);
                        if (dm < 0) amax = fmin(amax, -mu / dm);
                    }
                }
            if (pass == 0) {
                double a = fmin(1.0, amax), gap_aff = 0;
                for (int k = 0; k <= N; k++)
                    for (int c = 0; c < NC; c++) {
                        if (!act[k][c]) continue;
                        for (int sd = 0; sd < 2; sd++)
                            gap_aff += (ip->lam[k][c][sd] + a * ip->dlam[k][c][sd]) * (ip->t[k][c][sd] + a * ip->dt[k][c][sd])
                                     + (ip->mu[k][c][sd] + a * ip->dmu[k][c][sd]) * (ip->s[k][c][sd] + a * ip->ds[k][c][sd]);
                    }
                gap_aff /= (4.0 * nact);
                double sg = gap_aff / gap;
                /* Mehrotra centring, floored so that no product is driven far below the tolerance */
                sigma_mu = fmax(sg * sg * sg * gap, 0.1 * M->ipm_tol);
                /* second-order (Mehrotra) term scaled by how far the predictor could actually go:
                 * ipm_corr_mode 0: full term, 1: a, 2: a^2 */
                corr = M->ipm_corr_mode == 1 ? a : M->ipm_corr_mode == 2 ? a * a : 1.0;
                if (it >= ORC_SAFE_ITER) corr = a * a;      /* safe mode: damped corrector, short steps */
                if (orc_verbose) printf("    aff: amax %.4f gap_aff %.3e\n", amax, gap_aff);
            } else {
                /* fraction to the boundary: 1 - gap (superlinear tail) capped at ipm_eta_max.  Without the cap a blocking
                 * variable shrinks by the factor `gap` in one step; on the degenerate tail rows (linearised lh = 0 of
                 * neighbouring stages, SURVEY B-3) that de-centres the iterate into a limit cycle. */
                double eta = fmin(fmax(0.995, 1.0 - gap), M->ipm_eta_max);
                if (it >= ORC_SAFE_ITER) eta = fmin(eta, 0.9);
                alpha = fmin(1.0, eta * amax);
                if (orc_verbose) printf("    sigma_mu %.3e amax %.4f alpha %.4f\n", sigma_mu, amax, alpha);
            }
        }
        /* step; the states are re-rolled from the inputs so the dynamics rows hold to round-off */
        for (int k = 0; k < N; k++)
            for (int i = 0; i < 2; i++) qp->du[k][i] += alpha * (double)w->du[k][i];
        for (int k = 0; k < N; k++)
            for (int i = 0; i < 8; i++) {
                double s = qp->b[k][i];
                for (int l = 0; l < 8; l++) s += qp->A[k][i * 8 + l] * qp->dx[k][l];
                s += qp->B[k][i * 2] * qp->du[k][0] + qp->B[k][i * 2 + 1] * qp->du[k][1];
                qp->dx[k + 1][i] = s;
            }
        for (int k = 0; k <= N; k++)
            for (int c = 0; c < NC; c++) {
                if (!act[k][c]) continue;
                for (int sd = 0; sd < 2; sd++) {
                    ip->s[k][c][sd] += alpha * ip->ds[k][c][sd];
                    ip->lam[k][c][sd] += alpha * ip->dlam[k][c][sd];
                    ip->t[k][c][sd] += alpha * ip->dt[k][c][sd];
                    ip->mu[k][c][sd] += alpha * ip->dmu[k][c][sd];
                }
            }
    }
    qp->iters = it;
    if (!converged && qp->status == 0) qp->status = 1;
    for (int k = 0; k <= N; k++)
        for (int c = 0; c < NC; c++) {
            if (!act[k][c]) continue;
            qp->sl[k][c] = ip->s[k][c][0];  qp->su[k][c] = ip->s[k][c][1];
            qp->lam_l[k][c] = ip->lam[k][c][0];  qp->lam_u[k][c] = ip->lam[k][c][1];
        }
    qp->kkt = qp_kkt(qp);
    free(w);
    free(ip);
}

/* KKT residual of the returned point.  Costates are defined by the x-stationarity rows, so the
 * reported number is max(|u-stationarity|, |dynamics defect|, complementarity, bound violation,
 * slack stationarity violation). */
static double qp_kkt(const orc_qp *qp)
{
    int N = qp->N;
    double res = 0;
    double zero2[2] = {0, 0};
    /* dynamics */
    for (int i = 0; i < 8; i++) res = fmax(res, fabs(qp->dx[0][i] - qp->dx0[i]));
    for (int k = 0; k < N; k++)
        for (int i = 0; i < 8; i++) {
            double s = qp->b[k][i];
            for (int l = 0; l < 8; l++) s += qp->A[k][i * 8 + l] * qp->dx[k][l];
            s += qp->B[k][i * 2] * qp->du[k][0] + qp->B[k][i * 2 + 1] * qp->du[k][1];
            res = fmax(res, fabs(s - qp->dx[k + 1][i]));
        }
    /* costates from x-stationarity, residual from u-stationarity */
    double pi[8], pin[8];
    for (int k = N; k >= 0; k--) {
        real gx[8], gu[2] = {0, 0};
        for (int i = 0; i < 8; i++) {
            double s = qp->q[k][i];
            for (int l = 0; l < 8; l++) s += qp->Q[k][i * 8 + l] * qp->dx[k][l];
            gx[i] = s;
        }
        if (k < N) {
            gu[0] = qp->r[k][0] + qp->R[k][0] * qp->du[k][0] + qp->R[k][1] * qp->du[k][1];
            gu[1] = qp->r[k][1] + qp->R[k][2] * qp->du[k][0] + qp->R[k][3] * qp->du[k][1];
        }
        for (int c = 0; c < NC; c++) {
            if (!qp->act[k][c]) continue;
            double coef = -(qp->lam_l[k][c] - qp->lam_u[k][c]);
            row_add_grad(qp, k, c, coef, gx, gu);
            double v = qp->cval[k][c] + row_dot(qp, k, c, qp->dx[k], k < N ? qp->du[k] : zero2);
            double tl = v - qp->lo[k][c] + qp->sl[k][c], tu = qp->up[k][c] - v + qp->su[k][c];
            double ml = qp->zl[k][c] + qp->Zl[k][c] * qp->sl[k][c] - qp->lam_l[k][c];
            double mu_ = qp->zu[k][c] + qp->Zu[k][c] * qp->su[k][c] - qp->lam_u[k][c];
            res = fmax(res, fmax(-tl, -tu));
            res = fmax(res, fmax(-ml, -mu_));
            res = fmax(res, fmax(fabs(tl * qp->lam_l[k][c]), fabs(tu * qp->lam_u[k][c])));
            res = fmax(res, fmax(fabs(ml * qp->sl[k][c]), fabs(mu_ * qp->su[k][c])));
        }
        if (k < N) {
            for (int i = 0; i < 8; i++)
                for (int l = 0; l < 8; l++) gx[i] += qp->A[k][l * 8 + i] * pin[l];
            for (int i = 0; i < 2; i++)
                for (int l = 0; l < 8; l++) gu[i] += qp->B[k][l * 2 + i] * pin[l];
            res = fmax(res, fmax(fabs((double)gu[0]), fabs((double)gu[1])));
        }
        for (int i = 0; i < 8; i++) pi[i] = (double)gx[i];
        memcpy(pin, pi, sizeof pin);
        if (orc_verbose > 1) printf("  kkt k=%d res so far %.3e gu %.3e %.3e\n", k, res, (double)gu[0], (double)gu[1]);
    }
    return res;
}

/* NLS output y of one stage (NMPC_STM_acados_settings.py:41-52; with cost_vel_abs the SNMPC variant
 * SNMPC_acados_settings.py:62-65 = acados_ocp_SNMPC.json model.cost_y_expr): u == NULL -> the 4 terminal outputs */
void orc_cost_output(const orc_model *M, const double *x, const double *u, double *y)
{
    y[0] = x[0]; y[1] = x[1]; y[2] = wrap_yaw(x[2]);
    y[3] = M->cost_vel_abs ? sqrt(x[3] * x[3] + x[4] * x[4]) : x[3];
    if (u) { y[4] = u[0]; y[5] = u[1]; }
}

/* cost (acados get_cost at the current iterate incl. slack terms, NMPC_class.py:202) */
double orc_cost(const orc_model *M, const double *X, const double *U, const double *yref,
                const double *W, const double *zl, const double *zu, const double *Zl,
                const double *Zu, const double *sl, const double *su)
{
    int N = M->N;
    double J = 0;
    for (int k = 0; k <= N; k++) {
        const double *x = X + 8 * k, *w = W + 6 * k, *yr = yref + 6 * k;
        double sig = (k < N) ? M->dt : 1.0;
        double y[6];
        orc_cost_output(M, x, k < N ? U + 2 * k : NULL, y);
        double e0 = y[0] - yr[0], e1 = y[1] - yr[1], e2 = y[2] - yr[2], e3 = y[3] - yr[3];
        double c = w[0] * e0 * e0 + w[1] * e1 * e1 + w[2] * e2 * e2 + w[3] * e3 * e3;
        if (k < N) {
            double e4 = y[4] - yr[4], e5 = y[5] - yr[5];
            c += w[4] * e4 * e4 + w[5] * e5 * e5;
        }
        c *= 0.5;
        int act[NC];
        set_active(N, k, act);
        for (int j = 0; j < NC; j++) {
            if (!act[j]) continue;
            double a = sl[3 * k + j], b = su[3 * k + j];
            c += zl[3 * k + j] * a + 0.5 * Zl[3 * k + j] * a * a + zu[3 * k + j] * b + 0.5 * Zu[3 * k + j] * b * b;
        }
        J += sig * c;
    }
    return J;
}

int orc_rti_step(const orc_model *M, const double *x0, const double *yref, const double *W,
                 const double *lb, const double *ub, const double *zl, const double *zu,
                 const double *Zl, const double *Zu, double *X, double *U, double *stats,
                 double *sl_out, double *su_out, orc_qp *qp_ws)
{
    int N = M->N;
    orc_qp *qp = qp_ws ? qp_ws : (orc_qp *)malloc(sizeof(orc_qp));
    orc_linearize(M, x0, X, U, yref, W, lb, ub, zl, zu, Zl, Zu, qp);
    orc_qp_solve(M, qp);
    int status = 0;
    int nan = 0;
    for (int k = 0; k <= N && !nan; k++) {
        for (int i = 0; i < 8; i++) if (!(qp->dx[k][i] == qp->dx[k][i])) nan = 1;
        if (k < N) for (int i = 0; i < 2; i++) if (!(qp->du[k][i] == qp->du[k][i])) nan = 1;
    }
    if (nan || qp->status == 3) {
        status = 4;     /* ACADOS_QP_FAILURE; iterate left untouched */
    } else {
        /* full step (nlp_solver_step_length 1.0, FIXED_STEP) */
        for (int k = 0; k <= N; k++) {
            for (int i = 0; i < 8; i++) X[8 * k + i] += qp->dx[k][i];
            if (k < N) for (int i = 0; i < 2; i++) U[2 * k + i] += qp->du[k][i];
        }
    }
    double sl[(ORC_MAX_N + 1) * NC], su[(ORC_MAX_N + 1) * NC];
    for (int k = 0; k <= N; k++)
        for (int c = 0; c < NC; c++) { sl[3 * k + c] = qp->sl[k][c]; su[3 * k + c] = qp->su[k][c]; }
    if (stats) {
        stats[0] = orc_cost(M, X, U, yref, W, zl, zu, Zl, Zu, sl, su);
        stats[1] = qp->iters;
        stats[2] = status;
        stats[3] = qp->kkt;
    }
    if (sl_out) memcpy(sl_out, sl, sizeof(double) * 3 * (N + 1));
    if (su_out) memcpy(su_out, su, sizeof(double) * 3 * (N + 1));
    if (!qp_ws) free(qp);
    return status;
}

/* batch driver: a plain pthread pool pulling chunks of instances off an atomic counter */
typedef struct batch_job {
    const orc_model *M; int B; const double *x0, *yref, *W, *lb, *ub, *zl, *zu, *Zl, *Zu;
    double *X, *U, *stats; int next; int bad; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    int N = j->M->N;
    orc_qp *qp = (orc_qp *)malloc(sizeof(orc_qp));
    int bad = 0;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int lo = j->next;
        j->next += 8;
        pthread_mutex_unlock(&j->mu);
        if (lo >= j->B) break;
        int hi = lo + 8 < j->B ? lo + 8 : j->B;
        for (int i = lo; i < hi; i++) {
            int st = orc_rti_step(j->M, j->x0 + 8 * i, j->yref + (size_t)6 * (N + 1) * i, j->W, j->lb, j->ub,
                                  j->zl, j->zu, j->Zl, j->Zu, j->X + (size_t)8 * (N + 1) * i,
                                  j->U + (size_t)2 * N * i, j->stats + 4 * i, NULL, NULL, qp);
            bad += (st != 0);
        }
    }
    free(qp);
    pthread_mutex_lock(&j->mu);
    j->bad += bad;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}

int orc_rti_batch(const orc_model *M, int B, int nthreads, const double *x0, const double *yref,
                  const double *W, const double *lb, const double *ub, const double *zl,
                  const double *zu, const double *Zl, const double *Zu, double *X, double *U,
                  double *stats)
{
    batch_job j = {M, B, x0, yref, W, lb, ub, zl, zu, Zl, Zu, X, U, stats, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &j);
    batch_worker(&j);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return j.bad;
}

/* ------------------------------------------------------------------------------------------
 * R2NMPC tightening (Reduced_Robustified_NMPC_class.py:287-365,
 * Robust_NMPC_pred_model_utils.py:221-223): Sigma_{s+1} = A_s Sigma_s A_s' + B Wd B',
 * back-offs sqrt(Sigma[6,6]) on delta_f and sqrt(grad_h' Sigma grad_h) on h for 1 <= s < UPH,
 * stages UPH..N-1 reuse the last back-off.  B selects states 2..5.
 * ---------------------------------------------------------------------------------------- */
void orc_r2_tighten(const orc_model *M, const double *A, const double *X, const double *Sigma0,
                    const double *Wd, int uph, double delta_max, double *lbx, double *ubx, double *uh)
{
    int N = M->N;
    double Sg[64], T[64];
    memcpy(Sg, Sigma0, sizeof Sg);
    double cd = 0.0, ch = 0.0;
    for (int s = 0; s < N; s++) {
        if (s < uph) {
            if (s >= 1) {
                double gh[8];
                orc_h(M, X + 8 * s, gh);
                cd = sqrt(Sg[6 * 8 + 6]);
                double q = 0;
                for (int i = 0; i < 8; i++)
                    for (int j = 0; j < 8; j++) q += gh[i] * Sg[i * 8 + j] * gh[j];
                ch = sqrt(q);
            }
            /* propagate */
            const double *As = A + 64 * s;
            for (int i = 0; i < 8; i++)
                for (int j = 0; j < 8; j++) {
                    double v = 0; for (int l = 0; l < 8; l++) v += As[i * 8 + l] * Sg[l * 8 + j];
                    T[i * 8 + j] = v;
                }
            for (int i = 0; i < 8; i++)
                for (int j = 0; j < 8; j++) {
                    double v = 0; for (int l = 0; l < 8; l++) v += T[i * 8 + l] * As[j * 8 + l];
                    Sg[i * 8 + j] = v;
                }
            for (int i = 0; i < 4; i++) Sg[(2 + i) * 8 + 2 + i] += Wd[i];
        }
        if (s >= 1) {
            lbx[s] = -delta_max + cd;
            ubx[s] = delta_max - cd;
            uh[s] = 1.0 - ch;
        }
    }
}

int orc_sizeof_qp(void) { return (int)sizeof(orc_qp); }
int orc_sizeof_model(void) { return (int)sizeof(orc_model); }
