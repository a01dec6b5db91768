/*
 * snmpc_oracle.c -- CPU ORACLE for the Stochastic NMPC variant (test infrastructure, NOT the product).
 *
 * Restates the OCP of Model_Predictive_Controller/Stochastic_NMPC (paths relative to /root/reference):
 *   augmented discrete model   pred_model_dynamic_disc.py:115-220   (1 + n_s copies of the 8-state STM, one RK4
 *                              step of dt each; samples frozen and nominal = RK4 when stop_flag == 1, nominal =
 *                              PCE mean A[0,:] * samples_next otherwise)
 *   cost / chance constraint   SNMPC_acados_settings.py:57-142,184-194 (cost on the nominal block with
 *                              v = sqrt(vl^2+vt^2); h~ = E[h] + kappa sqrt(Var[h]) from the PCE coefficients
 *                              A * [h(sample_i)], kappa = sqrt((1-gamma)/gamma); nominal h when stop_flag == 1)
 *   stage-0 fixes all 8(1+n_s) states (SNMPC_class.py:120-121,259-264); stop_k = 1 for k >= UPH (:97-98)
 *
 * Method: deliberately NOT the structured Riccati path of the product.  Like acados (FULL_CONDENSING_HPIPM) it
 * condenses the multiple-shooting QP onto the 2N inputs and solves the dense soft-constrained QP with a dense
 * primal-dual interior point method (same side elimination / Mehrotra rules as orc_qp_solve).  With n_s = 0
 * and UPH = 0 it is the nominal OCP with a 1-step integrator, which cross-checks the Riccati oracle.
 *
 * Parity status: the MODEL is PINNED numerically against the reference's own folded expressions in the committed
 * acados_ocp_SNMPC.json (model.disc_dyn_expr :775, con_h_expr :745, cost_y_expr :772), evaluated by
 * tests/golden/make_snmpc_model_golden.py: orc_snmpc_disc / orc_snmpc_h / orc_cost_output agree to 1e-12 with the
 * full-precision constants and to 1e-5 with the literals as printed (tests/test_oracle_snmpc.py).  Structure (dims,
 * slack layout, options, x0 sample matrix) is pinned by the same JSON.  The reference commits no SNMPC *solver outputs*
 * made with the committed parameters: the QP solve on top of the pinned model is checked against the
 * golden-pinned Riccati oracle (n_s = 0) and through reductions.
 */
#include "nmpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NC ORC_NC

static double wrap_yaw2(double yaw)
{
    double y = fmod(yaw, 2.0 * M_PI);
    if (y < 0.0) y += 2.0 * M_PI;
    return y;
}

/* dense Cholesky solve of the n x n SPD system M x = r (M overwritten by its factor); returns 0 on success */
static int chol_factor(int n, double *M)
{
    for (int j = 0; j < n; j++) {
        double d = M[j * n + j];
        for (int l = 0; l < j; l++) d -= M[j * n + l] * M[j * n + l];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        M[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = M[i * n + j];
            for (int l = 0; l < j; l++) s -= M[i * n + l] * M[j * n + l];
            M[i * n + j] = s / d;
        }
    }
    return 0;
}
static void chol_solve(int n, const double *L, double *x)
{
    for (int i = 0; i < n; i++) {
        double s = x[i];
        for (int l = 0; l < i; l++) s -= L[i * n + l] * x[l];
        x[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int l = i + 1; l < n; l++) s -= L[l * n + i] * x[l];
        x[i] = s / L[i * n + i];
    }
}

/* chance-constraint value from the sample values of h and its derivative w.r.t. each sample value */
static double pce_h(const orc_snmpc *S, const double *hv, double *dh_dhv)
{
    double c[ORC_MAX_L];
    for (int j = 0; j < S->L; j++) {
        double s = 0;
        for (int i = 0; i < S->ns; i++) s += S->A[j * S->ns + i] * hv[i];
        c[j] = s;
    }
    double var = 0;
    for (int j = 1; j < S->L; j++) var += c[j] * c[j];
    const double sd = sqrt(var);
    if (dh_dhv)
        for (int i = 0; i < S->ns; i++) {
            double g = S->A[i];
            if (sd > 0.0)
                for (int j = 1; j < S->L; j++) g += S->kappa * c[j] * S->A[j * S->ns + i] / sd;
            dh_dhv[i] = g;
        }
    return c[0] + S->kappa * sd;
}

/*
 * One RTI step of the SNMPC OCP.  X: nominal iterate (N+1)x8; XS: sample iterates [UPH][ns][8] (stages
 * 0..UPH-1; later stages are frozen copies that touch nothing); U: Nx2.  x0: measured state; the sample
 * initial states are x0 + w_off[i] (compute_x0dist).  stats = [cost, qp_iter, status, kkt].
 */
int orc_snmpc_rti_step(const orc_model *M0, const orc_snmpc *S, const double *x0, const double *yref,
                       const double *W, const double *lb, const double *ub, const double *zl, const double *zu,
                       const double *Zl, const double *Zu, double *X, double *XS, double *U, double *stats,
                       double *sl_out, double *su_out, double *dbg_G, double *dbg_rows, double *dbg_v0)
{
    orc_model Mv = *M0;             /* DISCRETE dynamics: one RK4 step of dt; cost and h use |v| */
    Mv.rk_steps = 1;
    Mv.cost_vel_abs = 1;
    const orc_model *M = &Mv;
    const int N = M->N, H = S->H, ns = S->ns, nv = 2 * N;
    const int nrow = 3 * (N + 1);
    /* condensed affine maps: state = G u~ + c */
    double *G0 = (double *)calloc((size_t)8 * nv, sizeof(double)), c0[8];
    double *Gn = (double *)calloc((size_t)8 * nv, sizeof(double)), cn[8];
    double *GS = (double *)calloc((size_t)(ns > 0 ? ns : 1) * 8 * nv, sizeof(double));
    double *GSn = (double *)calloc((size_t)(ns > 0 ? ns : 1) * 8 * nv, sizeof(double));
    double *cS = (double *)calloc((size_t)(ns > 0 ? ns : 1) * 8, sizeof(double));
    double *cSn = (double *)calloc((size_t)(ns > 0 ? ns : 1) * 8, sizeof(double));
    double *Hm = (double *)calloc((size_t)nv * nv, sizeof(double));
    double *g = (double *)calloc(nv, sizeof(double));
    double *rows = (double *)calloc((size_t)nrow * nv, sizeof(double));
    double *v0 = (double *)calloc(nrow, sizeof(double));
    int *act = (int *)calloc(nrow, sizeof(int));
    double *Gall = (double *)calloc((size_t)(N + 1) * 8 * nv, sizeof(double));   /* nominal maps kept for the step */
    double *call = (double *)calloc((size_t)(N + 1) * 8, sizeof(double));
    double *GSall = (double *)calloc((size_t)(H > 0 ? H : 1) * (ns > 0 ? ns : 1) * 8 * nv, sizeof(double));
    double *cSall = (double *)calloc((size_t)(H > 0 ? H : 1) * (ns > 0 ? ns : 1) * 8, sizeof(double));

    for (int i = 0; i < 8; i++) c0[i] = x0[i] - X[i];
    for (int s = 0; s < ns; s++)
        for (int i = 0; i < 8; i++) cS[s * 8 + i] = (H > 0) ? (x0[i] + S->w_off[s * 8 + i]) - XS[(0 * ns + s) * 8 + i] : 0.0;

    for (int k = 0; k <= N; k++) {
        const double *x = X + 8 * k, *w = W + 6 * k, *yr = yref + 6 * k;
        const double sig = (k < N) ? M->dt : 1.0;
        memcpy(Gall + (size_t)k * 8 * nv, G0, sizeof(double) * 8 * nv);
        memcpy(call + 8 * k, c0, sizeof c0);
        if (k < H) {
            memcpy(GSall + (size_t)k * ns * 8 * nv, GS, sizeof(double) * ns * 8 * nv);
            memcpy(cSall + (size_t)k * ns * 8, cS, sizeof(double) * ns * 8);
        }
        /* ---- cost on the nominal block (SNMPC_acados_settings.py:184-194), Gauss-Newton */
        double Q[64], q[8];
        memset(Q, 0, sizeof Q); memset(q, 0, sizeof q);
        Q[0] = sig * w[0]; q[0] = sig * w[0] * (x[0] - yr[0]);
        Q[9] = sig * w[1]; q[1] = sig * w[1] * (x[1] - yr[1]);
        Q[18] = sig * w[2]; q[2] = sig * w[2] * (wrap_yaw2(x[2]) - yr[2]);
        {
            double V = sqrt(x[3] * x[3] + x[4] * x[4]);
            double g3 = V > 0 ? x[3] / V : 0.0, g4 = V > 0 ? x[4] / V : 0.0;
            Q[3 * 8 + 3] = sig * w[3] * g3 * g3; Q[3 * 8 + 4] = Q[4 * 8 + 3] = sig * w[3] * g3 * g4; Q[4 * 8 + 4] = sig * w[3] * g4 * g4;
            q[3] = sig * w[3] * (V - yr[3]) * g3; q[4] = sig * w[3] * (V - yr[3]) * g4;
        }
        /* H += G0' Q G0, g += G0' (Q c0 + q) */
        for (int a = 0; a < nv; a++) {
            double t[8];
            for (int i = 0; i < 8; i++) {
                double s2 = 0; for (int l = 0; l < 8; l++) s2 += Q[i * 8 + l] * G0[l * nv + a];
                t[i] = s2;
            }
            for (int b2 = 0; b2 < nv; b2++) {
                double s2 = 0; for (int i = 0; i < 8; i++) s2 += G0[i * nv + b2] * t[i];
                Hm[b2 * nv + a] += s2;
            }
            double s3 = 0;
            for (int i = 0; i < 8; i++) {
                double qc = q[i]; for (int l = 0; l < 8; l++) qc += Q[i * 8 + l] * c0[l];
                s3 += G0[i * nv + a] * qc;
            }
            g[a] += s3;
        }
        if (k < N) {
            Hm[(2 * k) * nv + 2 * k] += sig * w[4];          g[2 * k] += sig * w[4] * (U[2 * k] - yr[4]);
            Hm[(2 * k + 1) * nv + 2 * k + 1] += sig * w[5];  g[2 * k + 1] += sig * w[5] * (U[2 * k + 1] - yr[5]);
        }
        /* ---- soft rows [bu, bx, h] (same slack layout as the nominal OCP; JSON dims ns_0=1, ns=3, ns_e=2) */
        act[3 * k + 0] = (k < N); act[3 * k + 1] = (k >= 1); act[3 * k + 2] = (k >= 1);
        if (k < N) { rows[(size_t)(3 * k) * nv + 2 * k + 1] = 1.0; v0[3 * k] = U[2 * k + 1]; }
        for (int a = 0; a < nv; a++) rows[(size_t)(3 * k + 1) * nv + a] = G0[6 * nv + a];
        v0[3 * k + 1] = x[6] + c0[6];
        if (k >= 1) {
            if (k < H) {                    /* stop_flag == 0: chance constraint from the samples */
                double hv[ORC_MAX_NS], gh[ORC_MAX_NS][8], dd[ORC_MAX_NS];
                for (int s = 0; s < ns; s++) hv[s] = orc_h(M, XS + ((size_t)k * ns + s) * 8, gh[s]);
                double ht = pce_h(S, hv, dd);
                for (int s = 0; s < ns; s++) {
                    for (int a = 0; a < nv; a++) {
                        double s2 = 0; for (int i = 0; i < 8; i++) s2 += gh[s][i] * GS[((size_t)s * 8 + i) * nv + a];
                        rows[(size_t)(3 * k + 2) * nv + a] += dd[s] * s2;
                    }
                    double s3 = 0; for (int i = 0; i < 8; i++) s3 += gh[s][i] * cS[s * 8 + i];
                    ht += dd[s] * s3;
                }
                v0[3 * k + 2] = ht;
            } else {
                double gh[8];
                double hv = orc_h(M, x, gh);
                for (int a = 0; a < nv; a++) {
                    double s2 = 0; for (int i = 0; i < 8; i++) s2 += gh[i] * G0[i * nv + a];
                    rows[(size_t)(3 * k + 2) * nv + a] = s2;
                }
                double s3 = 0; for (int i = 0; i < 8; i++) s3 += gh[i] * c0[i];
                v0[3 * k + 2] = hv + s3;
            }
        }
        if (k == N) break;
        /* ---- dynamics k -> k+1 */
        const double *u = U + 2 * k;
        memset(Gn, 0, sizeof(double) * 8 * nv); memset(cn, 0, sizeof cn);
        if (k < H) {
            for (int s = 0; s < ns; s++) {
                double xn[8], A[64], B[16];
                orc_integrate(M, XS + ((size_t)k * ns + s) * 8, u, xn, A, B);
                const double *Gs = GS + (size_t)s * 8 * nv; double *Gsn = GSn + (size_t)s * 8 * nv;
                for (int i = 0; i < 8; i++) {
                    for (int a = 0; a < nv; a++) {
                        double s2 = 0; for (int l = 0; l < 8; l++) s2 += A[i * 8 + l] * Gs[l * nv + a];
                        Gsn[i * nv + a] = s2;
                    }
                    Gsn[i * nv + 2 * k] += B[i * 2]; Gsn[i * nv + 2 * k + 1] += B[i * 2 + 1];
                    double s3 = 0; for (int l = 0; l < 8; l++) s3 += A[i * 8 + l] * cS[s * 8 + l];
                    /* dx_next = A dx + B du + (Phi(xs,u) - xs_next): the defect needs the next sample iterate */
                    cSn[s * 8 + i] = s3 + xn[i] - ((k + 1 < H) ? XS[((size_t)(k + 1) * ns + s) * 8 + i] : xn[i]);
                    /* nominal next = PCE mean of the samples' next states (pred_model_dynamic_disc.py:215-218) */
                    const double a0 = S->A[s];
                    for (int a = 0; a < nv; a++) Gn[i * nv + a] += a0 * Gsn[i * nv + a];
                    cn[i] += a0 * (s3 + xn[i]);
                }
            }
            for (int i = 0; i < 8; i++) cn[i] -= X[8 * (k + 1) + i];
            double *t1 = GS; GS = GSn; GSn = t1;
            double *t2 = cS; cS = cSn; cSn = t2;
        } else {
            double xn[8], A[64], B[16];
            orc_integrate(M, x, u, xn, A, B);
            for (int i = 0; i < 8; i++) {
                for (int a = 0; a < nv; a++) {
                    double s2 = 0; for (int l = 0; l < 8; l++) s2 += A[i * 8 + l] * G0[l * nv + a];
                    Gn[i * nv + a] = s2;
                }
                Gn[i * nv + 2 * k] += B[i * 2]; Gn[i * nv + 2 * k + 1] += B[i * 2 + 1];
                double s3 = 0; for (int l = 0; l < 8; l++) s3 += A[i * 8 + l] * c0[l];
                cn[i] = s3 + xn[i] - X[8 * (k + 1) + i];
            }
        }
        memcpy(G0, Gn, sizeof(double) * 8 * nv); memcpy(c0, cn, sizeof c0);
    }

    /* optional diagnostics: condensed nominal maps, soft rows and their values at the linearisation point */
    if (dbg_G) memcpy(dbg_G, Gall, sizeof(double) * (size_t)(N + 1) * 8 * nv);
    if (dbg_rows) memcpy(dbg_rows, rows, sizeof(double) * (size_t)nrow * nv);
    if (dbg_v0) memcpy(dbg_v0, v0, sizeof(double) * nrow);

    /* ------------------------------------------------------------------ dense soft-constrained QP */
    double *ut = (double *)calloc(nv, sizeof(double)), *dut = (double *)calloc(nv, sizeof(double));
    double *Mm = (double *)malloc(sizeof(double) * nv * nv), *rhs = (double *)malloc(sizeof(double) * nv);
    double *s_ = (double *)calloc((size_t)nrow * 2, sizeof(double)), *t_ = (double *)calloc((size_t)nrow * 2, sizeof(double));
    double *lam = (double *)calloc((size_t)nrow * 2, sizeof(double)), *mu = (double *)calloc((size_t)nrow * 2, sizeof(double));
    double *ds = (double *)calloc((size_t)nrow * 2, sizeof(double)), *dt_ = (double *)calloc((size_t)nrow * 2, sizeof(double));
    double *dl = (double *)calloc((size_t)nrow * 2, sizeof(double)), *dm = (double *)calloc((size_t)nrow * 2, sizeof(double));
    double *pz = (double *)calloc((size_t)nrow * 2, sizeof(double)), *pZ = (double *)calloc((size_t)nrow * 2, sizeof(double));
    double *bnd = (double *)calloc((size_t)nrow * 2, sizeof(double));
    double *vv = (double *)calloc(nrow, sizeof(double)), *dvv = (double *)calloc(nrow, sizeof(double));
    const double reg = M->ipm_reg, tol = M->ipm_tol;
    int nact = 0;
    for (int k = 0; k <= N; k++)
        for (int c = 0; c < NC; c++) {
            const int r = 3 * k + c;
            const double sig = (k < N) ? M->dt : 1.0;
            pz[2 * r] = sig * zl[r]; pz[2 * r + 1] = sig * zu[r]; pZ[2 * r] = sig * Zl[r]; pZ[2 * r + 1] = sig * Zu[r];
            bnd[2 * r] = lb[r]; bnd[2 * r + 1] = ub[r];
            if (!((pz[2 * r] + pZ[2 * r] > 0.0) && (pz[2 * r + 1] + pZ[2 * r + 1] > 0.0))) act[r] = 0;
            nact += act[r];
        }
    /* phase 0: unconstrained optimum */
    memcpy(Mm, Hm, sizeof(double) * nv * nv);
    int status = 0, iters = 0, ran_ipm = 0;
    double kkt = 0.0;
    if (chol_factor(nv, Mm)) status = 3;
    for (int a = 0; a < nv; a++) ut[a] = -g[a];
    if (!status) chol_solve(nv, Mm, ut);
    int feasible = 1;
    for (int r = 0; r < nrow; r++) {
        double s2 = v0[r]; for (int a = 0; a < nv; a++) s2 += rows[(size_t)r * nv + a] * ut[a];
        vv[r] = s2;
        if (act[r] && (s2 < bnd[2 * r] || s2 > bnd[2 * r + 1])) feasible = 0;
    }
    if (!status && !((feasible && !M->qp_no_shortcut) || nact == 0)) {
        for (int r = 0; r < nrow; r++) {
            if (!act[r]) continue;
            const double viol = fmax(fmax(bnd[2 * r] - vv[r], vv[r] - bnd[2 * r + 1]), 0.0), Sv = M->ipm_s0 + viol;
            const double gl = pz[2 * r] + pZ[2 * r] * Sv, gu = pz[2 * r + 1] + pZ[2 * r + 1] * Sv, l0 = 0.5 * fmin(gl, gu);
            s_[2 * r] = s_[2 * r + 1] = Sv; lam[2 * r] = lam[2 * r + 1] = l0;
            t_[2 * r] = vv[r] - bnd[2 * r] + Sv; t_[2 * r + 1] = bnd[2 * r + 1] - vv[r] + Sv;
            mu[2 * r] = gl - l0; mu[2 * r + 1] = gu - l0;
        }
        int it, polish = 0;
        for (it = 0; it <= M->ipm_max_iter; it++) {
            double gap = 0, cmax = 0, rmax = 0, rstat = 0;
            for (int r = 0; r < nrow; r++) {
                if (!act[r]) continue;
                for (int sd = 0; sd < 2; sd++) {
                    const int j = 2 * r + sd;
                    gap += lam[j] * t_[j] + mu[j] * s_[j];
                    cmax = fmax(cmax, fmax(lam[j] * t_[j], mu[j] * s_[j]));
                    const double rt = t_[j] - ((sd ? bnd[j] - vv[r] : vv[r] - bnd[j]) + s_[j]);
                    const double rs = pz[j] + pZ[j] * s_[j] - lam[j] - mu[j];
                    rmax = fmax(rmax, fmax(fabs(rt), fabs(rs)));
                }
            }
            gap /= (4.0 * nact);
            /* stationarity residual H u + g - sum (lam_l - lam_u) row */
            for (int a = 0; a < nv; a++) {
                double s2 = g[a]; for (int b2 = 0; b2 < nv; b2++) s2 += Hm[a * nv + b2] * ut[b2];
                rhs[a] = s2;
            }
            for (int r = 0; r < nrow; r++) {
                if (!act[r]) continue;
                const double ld = lam[2 * r] - lam[2 * r + 1];
                for (int a = 0; a < nv; a++) rhs[a] -= ld * rows[(size_t)r * nv + a];
            }
            for (int a = 0; a < nv; a++) rstat = fmax(rstat, fabs(rhs[a]));
            if (!(cmax == cmax) || !(rmax == rmax) || !(rstat == rstat)) { status = 3; break; }
            if (cmax <= tol && rmax <= tol && rstat <= tol) { kkt = fmax(fmax(cmax, rmax), rstat); break; }
            if (it == M->ipm_max_iter) { kkt = fmax(fmax(cmax, rmax), rstat); break; }
            if (cmax <= tol && ++polish > ORC_POLISH_ITER) { kkt = fmax(fmax(cmax, rmax), rstat); break; }
            memcpy(Mm, Hm, sizeof(double) * nv * nv);
            for (int r = 0; r < nrow; r++) {
                if (!act[r]) continue;
                double om = 0;
                for (int sd = 0; sd < 2; sd++) {
                    const int j = 2 * r + sd;
                    const double th = lam[j] / (t_[j] + reg * lam[j]), ms = mu[j] / s_[j];
                    om += th * (pZ[j] + ms) / (pZ[j] + th + ms);
                }
                const double *gr = rows + (size_t)r * nv;
                for (int a = 0; a < nv; a++) {
                    if (gr[a] == 0.0) continue;
                    for (int b2 = 0; b2 < nv; b2++) Mm[a * nv + b2] += om * gr[a] * gr[b2];
                }
            }
            if (chol_factor(nv, Mm)) { status = 3; break; }
            double sigma_mu = 0.0, alpha = 1.0, corr = 1.0;
            double *stat = (double *)malloc(sizeof(double) * nv);
            memcpy(stat, rhs, sizeof(double) * nv);
            for (int pass = 0; pass < 2; pass++) {
                for (int a = 0; a < nv; a++) dut[a] = -stat[a];
                for (int r = 0; r < nrow; r++) {
                    if (!act[r]) continue;
                    double coef = 0;
                    for (int sd = 0; sd < 2; sd++) {
                        const int j = 2 * r + sd;
                        const double sgn = sd ? -1.0 : 1.0;
                        const double tt = t_[j] + reg * lam[j], th = lam[j] / tt, ms = mu[j] / s_[j], Dp = pZ[j] + th + ms;
                        const double rt = t_[j] - ((sd ? bnd[j] - vv[r] : vv[r] - bnd[j]) + s_[j]);
                        const double rs = pz[j] + pZ[j] * s_[j] - lam[j] - mu[j];
                        double r1 = -lam[j] * t_[j], r2 = -mu[j] * s_[j];
                        if (pass) { r1 += sigma_mu - corr * dl[j] * dt_[j]; r2 += sigma_mu - corr * dm[j] * ds[j]; }
                        const double rho = -rs + r1 / tt + r2 / s_[j] + th * rt;
                        coef += sgn * (r1 / tt + th * rt - th * rho / Dp);
                    }
                    const double *gr = rows + (size_t)r * nv;
                    for (int a = 0; a < nv; a++) dut[a] += coef * gr[a];
                }
                chol_solve(nv, Mm, dut);
                double amax = 1e300;
                for (int r = 0; r < nrow; r++) {
                    if (!act[r]) continue;
                    double dv = 0; for (int a = 0; a < nv; a++) dv += rows[(size_t)r * nv + a] * dut[a];
                    dvv[r] = dv;
                    for (int sd = 0; sd < 2; sd++) {
                        const int j = 2 * r + sd;
                        const double sgn = sd ? -1.0 : 1.0;
                        const double tt = t_[j] + reg * lam[j], th = lam[j] / tt, ms = mu[j] / s_[j], Dp = pZ[j] + th + ms;
                        const double rt = t_[j] - ((sd ? bnd[j] - vv[r] : vv[r] - bnd[j]) + s_[j]);
                        const double rs = pz[j] + pZ[j] * s_[j] - lam[j] - mu[j];
                        double r1 = -lam[j] * t_[j], r2 = -mu[j] * s_[j];
                        if (pass) { r1 += sigma_mu - corr * dl[j] * dt_[j]; r2 += sigma_mu - corr * dm[j] * ds[j]; }
                        const double rho = -rs + r1 / tt + r2 / s_[j] + th * rt;
                        const double dsj = (rho - th * sgn * dv) / Dp;
                        const double dlj = (r1 - lam[j] * (sgn * dv + dsj - rt)) / tt;
                        const double dtj = sgn * dv + dsj - rt + reg * dlj;
                        const double dmj = (r2 - mu[j] * dsj) / s_[j];
                        /* the products of the predictor are only consumed after both passes read them */
                        if (dsj < 0) amax = fmin(amax, -s_[j] / dsj);
                        if (dlj < 0) amax = fmin(amax, -lam[j] / dlj);
                        if (dtj < 0) amax = fmin(amax, -t_[j] / dtj);
                        if (dmj < 0) amax = fmin(amax, -mu[j] / dmj);
                        /* pass 0 keeps the predictor directions for the second-order terms (read above, before
                         * this overwrite, when pass == 1); pass 1 keeps the final directions */
                        ds[j] = dsj; dl[j] = dlj; dt_[j] = dtj; dm[j] = dmj;
                    }
                }
                if (pass == 0) {
                    const double a = fmin(1.0, amax);
                    double gap_aff = 0;
                    for (int r = 0; r < nrow; r++) {
                        if (!act[r]) continue;
                        for (int sd = 0; sd < 2; sd++) {
                            const int j = 2 * r + sd;
                            gap_aff += (lam[j] + a * dl[j]) * (t_[j] + a * dt_[j]) + (mu[j] + a * dm[j]) * (s_[j] + a * ds[j]);
                        }
                    }
                    gap_aff /= (4.0 * nact);
                    const double sg = gap_aff / gap;
                    sigma_mu = fmax(sg * sg * sg * gap, 0.1 * tol);
                    corr = M->ipm_corr_mode == 1 ? a : M->ipm_corr_mode == 2 ? a * a : 1.0;
                    if (it >= ORC_SAFE_ITER) corr = a * a;
                } else {
                    double eta = fmin(fmax(0.995, 1.0 - gap), M->ipm_eta_max);
                    if (it >= ORC_SAFE_ITER) eta = fmin(eta, 0.9);
                    alpha = fmin(1.0, eta * amax);
                }
            }
            free(stat);
            for (int a = 0; a < nv; a++) ut[a] += alpha * dut[a];
            for (int r = 0; r < nrow; r++) {
                vv[r] += alpha * dvv[r];
                if (!act[r]) continue;
                for (int sd = 0; sd < 2; sd++) {
                    const int j = 2 * r + sd;
                    s_[j] += alpha * ds[j]; lam[j] += alpha * dl[j]; t_[j] += alpha * dt_[j]; mu[j] += alpha * dm[j];
                }
            }
        }
        iters = it;
        ran_ipm = 1;
    }
    /* ------------------------------------------------------------------ full step and cost */
    int nan = (status == 3);
    for (int a = 0; a < nv && !nan; a++) if (!(ut[a] == ut[a])) nan = 1;
    int ret = 0;
    if (nan) {
        ret = 4;
    } else {
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < 8; i++) {
                double s2 = call[8 * k + i];
                for (int a = 0; a < nv; a++) s2 += Gall[((size_t)k * 8 + i) * nv + a] * ut[a];
                X[8 * k + i] += s2;
            }
        for (int k = 0; k < H; k++)
            for (int s = 0; s < ns; s++)
                for (int i = 0; i < 8; i++) {
                    double s2 = cSall[((size_t)k * ns + s) * 8 + i];
                    for (int a = 0; a < nv; a++) s2 += GSall[(((size_t)k * ns + s) * 8 + i) * nv + a] * ut[a];
                    XS[((size_t)k * ns + s) * 8 + i] += s2;
                }
        for (int a = 0; a < nv; a++) U[a] += ut[a];
    }
    double *sl = (double *)calloc(nrow, sizeof(double)), *su = (double *)calloc(nrow, sizeof(double));
    if (ran_ipm)
        for (int r = 0; r < nrow; r++) if (act[r]) { sl[r] = s_[2 * r]; su[r] = s_[2 * r + 1]; }
    if (stats) {
        stats[0] = orc_cost(M, X, U, yref, W, zl, zu, Zl, Zu, sl, su);
        stats[1] = iters; stats[2] = ret; stats[3] = kkt;
    }
    if (sl_out) memcpy(sl_out, sl, sizeof(double) * nrow);
    if (su_out) memcpy(su_out, su, sizeof(double) * nrow);
    free(sl); free(su);
    free(G0); free(Gn); free(GS); free(GSn); free(cS); free(cSn); free(Hm); free(g); free(rows); free(v0); free(act);
    free(Gall); free(call); free(GSall); free(cSall); free(ut); free(dut); free(Mm); free(rhs);
    free(s_); free(t_); free(lam); free(mu); free(ds); free(dt_); free(dl); free(dm); free(pz); free(pZ); free(bnd);
    free(vv); free(dvv);
    return ret;
}

/* the augmented discrete map itself (pred_model_dynamic_disc.py:187-220), for finite-difference tests:
 * xa = [x0; x1..x_ns] (8(1+ns)), returns xa_next.  stop != 0: samples frozen, nominal = RK4. */
void orc_snmpc_disc(const orc_model *M0, const orc_snmpc *S, const double *xa, const double *u, int stop, double *xa_next)
{
    orc_model Mv = *M0;
    Mv.rk_steps = 1;
    const int ns = S->ns;
    double mean[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < ns; s++) {
        double xn[8];
        if (stop) memcpy(xn, xa + 8 * (s + 1), sizeof xn);
        else orc_flow(&Mv, xa + 8 * (s + 1), u, Mv.dt, 1, xn);
        memcpy(xa_next + 8 * (s + 1), xn, sizeof xn);
        for (int i = 0; i < 8; i++) mean[i] += S->A[s] * xn[i];
    }
    if (stop) orc_flow(&Mv, xa, u, Mv.dt, 1, xa_next);
    else memcpy(xa_next, mean, sizeof mean);
}

/* chance-constraint value of an augmented state (SNMPC_acados_settings.py:131-133,184) */
double orc_snmpc_h(const orc_model *M0, const orc_snmpc *S, const double *xa, int stop)
{
    orc_model Mv = *M0;
    Mv.cost_vel_abs = 1;
    if (stop) return orc_h(&Mv, xa, NULL);
    double hv[ORC_MAX_NS];
    for (int s = 0; s < S->ns; s++) hv[s] = orc_h(&Mv, xa + 8 * (s + 1), NULL);
    return pce_h(S, hv, NULL);
}

int orc_sizeof_snmpc(void) { return (int)sizeof(orc_snmpc); }
