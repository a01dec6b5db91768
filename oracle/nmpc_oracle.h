/*
 * nmpc_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C, double-precision restatement of the one hot path this repository
 * accelerates: the real-time-iteration SQP step that acados executes for
 * bzarr/TUM-CONTROL's NMPC controllers (reference call site
 * Model_Predictive_Controller/Nominal_NMPC/NMPC_class.py:183, OCP declared in
 * Model_Predictive_Controller/Nominal_NMPC/NMPC_STM_acados_settings.py:16-245,
 * dynamics in Prediction_Models/pred_model_dynamic_stm_pacejka.py:118-177).
 *
 * The arithmetic of that path lives in third-party libraries that are absent
 * from /root/reference and from this image (acados + HPIPM + BLASFEO, unpinned;
 * CasADi 3.5.5), so this file restates their published algorithm
 * (multiple-shooting ERK4 with forward sensitivities, Gauss-Newton NLS cost,
 * soft BGH constraints, structured QP, full step) and is PINNED against the
 * reference's own committed closed-loop logs
 * (Learning_To_Adapt/SafeRL_WMPC/_baseline/F/...) by tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this code.  The product (libb200nmpc.so)
 * never links, imports or falls back to it.
 */
#ifndef NMPC_ORACLE_H
#define NMPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NX 8
#define ORC_NU 2
#define ORC_NY 6
#define ORC_NC 3        /* soft rows per stage, acados slack order [bu, bx, h] */
#define ORC_MAX_GGV 16
#define ORC_MAX_N 128
#define ORC_MAX_NS 16   /* SNMPC: max collocation samples */
#define ORC_MAX_L 10    /* SNMPC: PCE basis size (degree 2, 3 variables) */

/* Vehicle / tyre constants and horizon.  Values come from
 * Config/EDGAR/{veh_params_pred,pacejka_params}.yaml and the literals in
 * pred_model_dynamic_stm_pacejka.py:38-46. */
/* interior-point safe mode: from this iteration on the corrector term is damped by a_aff^2 and the fraction to the
 * boundary is 0.9 (breaks limit cycles of a de-centred iterate; never reached on the pinned golden laps) */
#define ORC_SAFE_ITER 24
/* iterations granted to the linear residuals once every complementarity product is below the tolerance */
#define ORC_POLISH_ITER 4

typedef struct orc_model {
    double lf, lr, m, Iz, ro, S, Cd;
    double Bf, Cf, Df, Ef, Br, Cr, Dr, Er;
    double g, fr0, fr1, fr4;
    double ax_brake;                 /* -acc_min (NMPC_STM_acados_settings.py:73) */
    double ggv_v[ORC_MAX_GGV], ggv_ax[ORC_MAX_GGV], ggv_ay[ORC_MAX_GGV];
    double dt;                       /* Tp / N */
    double ipm_tol;                  /* complementarity tolerance of the QP solve */
    double ipm_reg;                  /* dual proximal regularisation (caps barrier weights at 1/ipm_reg) */
    double ipm_eta_max;              /* cap of the fraction-to-the-boundary parameter max(0.995, 1 - gap) */
    double ipm_s0;                   /* initial slack of the interior start: s = ipm_s0 + violation */
    int n_ggv;
    int N;
    int rk_steps;                    /* sim_method_num_steps (3); RK4 stages fixed */
    int cost_vel_abs;                /* 0: y[3]=v_lon (nominal, R2); 1: sqrt(vl^2+vt^2) (SNMPC) */
    int ipm_max_iter;
    int qp_no_shortcut;              /* 1: always run the IPM even if the equality-constrained optimum is feasible */
    int ipm_corr_mode;               /* Mehrotra second-order term scaled by 1 (0), a_aff (1), a_aff^2 (2) */
} orc_model;

/* dynamics and derivatives -------------------------------------------------- */
void orc_f(const orc_model *M, const double *x, const double *u, double *xdot);
void orc_f_jac(const orc_model *M, const double *x, const double *u, double *xdot,
               double *Jx /*8x8 row-major*/, double *Ju /*8x2*/);
/* rk_steps x RK4 over dt with forward sensitivities: A = dPhi/dx, B = dPhi/du */
void orc_integrate(const orc_model *M, const double *x, const double *u, double *xn,
                   double *A /*8x8*/, double *B /*8x2*/);
/* plain RK4 flow without sensitivities (n_steps sub-steps over T) */
void orc_flow(const orc_model *M, const double *x, const double *u, double T, int n_steps,
              double *xn);
/* combined-acceleration circle constraint and its gradient (dense, 8 entries) */
double orc_h(const orc_model *M, const double *x, double *grad /*8 or NULL*/);

/* linearisation of one instance about (X,U) --------------------------------- */
typedef struct orc_qp {
    int N;
    double A[ORC_MAX_N][64], B[ORC_MAX_N][16], b[ORC_MAX_N][8];
    double Q[ORC_MAX_N + 1][64];      /* Gauss-Newton state Hessian (dt-scaled for k<N) */
    double R[ORC_MAX_N][4];
    double q[ORC_MAX_N + 1][8], r[ORC_MAX_N][2];
    double cval[ORC_MAX_N + 1][ORC_NC];   /* [u1, x6, h] at the linearisation point */
    double gh[ORC_MAX_N + 1][8];          /* grad h (dense) */
    double lo[ORC_MAX_N + 1][ORC_NC], up[ORC_MAX_N + 1][ORC_NC];
    double zl[ORC_MAX_N + 1][ORC_NC], zu[ORC_MAX_N + 1][ORC_NC];   /* already sigma_k-scaled */
    double Zl[ORC_MAX_N + 1][ORC_NC], Zu[ORC_MAX_N + 1][ORC_NC];
    int act[ORC_MAX_N + 1][ORC_NC];
    double dx0[8];
    /* solution */
    double dx[ORC_MAX_N + 1][8], du[ORC_MAX_N][2];
    double sl[ORC_MAX_N + 1][ORC_NC], su[ORC_MAX_N + 1][ORC_NC];
    double lam_l[ORC_MAX_N + 1][ORC_NC], lam_u[ORC_MAX_N + 1][ORC_NC];
    double pi[ORC_MAX_N + 1][8];      /* costates of the final Newton system (diagnostics) */
    int iters;
    int status;                       /* 0 converged, 1 max iter, 3 NaN */
    double kkt;                       /* max KKT residual of the returned point */
} orc_qp;

/* per-instance, per-stage runtime data (all arrays are [(N+1)*...] row-major):
 *   W   [(N+1)*6]  diagonal of the NLS weight (terminal stage uses the first 4)
 *   lb,ub [(N+1)*3] bounds in slack order [bu(omega_f), bx(delta_f), h]
 *   zl,zu,Zl,Zu [(N+1)*3] slack penalties (unscaled; dt-scaling applied here) */
void orc_linearize(const orc_model *M, const double *x0, const double *X, const double *U,
                   const double *yref, const double *W, const double *lb, const double *ub,
                   const double *zl, const double *zu, const double *Zl, const double *Zu,
                   orc_qp *qp);

/* soft-constrained structured QP: Riccati-based Mehrotra primal-dual IPM */
void orc_qp_solve(const orc_model *M, orc_qp *qp);

/* cost at (X,U) with slack values (acados get_cost semantics) */
void orc_cost_output(const orc_model *M, const double *x, const double *u /*2 or NULL*/, double *y /*6 (4 if u NULL)*/);
double orc_cost(const orc_model *M, const double *X, const double *U, const double *yref,
                const double *W, const double *zl, const double *zu, const double *Zl,
                const double *Zu, const double *sl, const double *su);

/* one RTI step, in place on (X,U).  stats = [cost, qp_iter, status, kkt].
 * sl_out/su_out may be NULL.  qp_ws may be NULL (internal workspace). */
int orc_rti_step(const orc_model *M, const double *x0, const double *yref, const double *W,
                 const double *lb, const double *ub, const double *zl, const double *zu,
                 const double *Zl, const double *Zu, double *X, double *U, double *stats,
                 double *sl_out, double *su_out, orc_qp *qp_ws);

/* batch driver used for the CPU baseline: instance-major arrays, OpenMP over B.
 * Shared (broadcast) W / bounds / penalties.  Returns number of non-zero statuses. */
int orc_rti_batch(const orc_model *M, int B, int nthreads, const double *x0 /*B*8*/,
                  const double *yref /*B*(N+1)*6*/, const double *W, const double *lb,
                  const double *ub, const double *zl, const double *zu, const double *Zl,
                  const double *Zu, double *X /*B*(N+1)*8*/, double *U /*B*N*2*/,
                  double *stats /*B*4*/);

/* R2NMPC post-solve tightening (Reduced_Robustified_NMPC_class.py:287-365) */
void orc_r2_tighten(const orc_model *M, const double *A /*N*64, pre-step linearisation*/,
                    const double *X /*(N+1)*8, post-step iterate*/, const double *Sigma0 /*64*/,
                    const double *Wd /*4x4 diag: 4 entries*/, int uph, double delta_max,
                    double *lbx /*N+1*/, double *ubx /*N+1*/, double *uh /*N+1*/);

/* 7-state plant (Vehicle_Simulator/sim_model_dynamic_stm_pacejka.py) RK4, n sub-steps */
void orc_plant_step(const orc_model *Msim, const double *x7, const double *u2 /*[a, omega]*/,
                    const double *w7 /*additive xdot disturbance or NULL*/, double Ts, int nsub,
                    double *x7n);

/* Stochastic NMPC (snmpc_oracle.c) ------------------------------------------------------------- */
typedef struct orc_snmpc {
    double A[ORC_MAX_L * ORC_MAX_NS];   /* PCE matrix, L x ns row-major (stochastic_mpc_utils.py:73) */
    double w_off[ORC_MAX_NS * 8];       /* sample offsets: x0_i = x0 + w_off[i] (compute_x0dist, :78-91) */
    double kappa;                       /* sqrt((1 - gamma) / gamma) */
    int ns, L, H, pad_;                 /* samples, basis size, uncertainty propagation horizon */
} orc_snmpc;
int orc_snmpc_rti_step(const orc_model *M, const orc_snmpc *S, const double *x0, const double *yref,
                       const double *W, const double *lb, const double *ub, const double *zl, const double *zu,
                       const double *Zl, const double *Zu, double *X /*(N+1)*8*/, double *XS /*H*ns*8*/,
                       double *U /*N*2*/, double *stats /*4*/, double *sl_out, double *su_out,
                       double *dbg_G /*(N+1)*8*2N or NULL*/, double *dbg_rows /*3(N+1)*2N or NULL*/,
                       double *dbg_v0 /*3(N+1) or NULL*/);
void orc_snmpc_disc(const orc_model *M, const orc_snmpc *S, const double *xa, const double *u, int stop, double *xa_next);
double orc_snmpc_h(const orc_model *M, const orc_snmpc *S, const double *xa, int stop);
int orc_sizeof_snmpc(void);

int orc_sizeof_qp(void);
int orc_sizeof_model(void);

#ifdef __cplusplus
}
#endif
#endif
