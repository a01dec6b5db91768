/*
 * b200nmpc.h -- C-ABI of the batched, Blackwell-native RTI-SQP NMPC solver (libb200nmpc.so).
 *
 * This is the drop-in boundary for ONE hot path of bzarr/TUM-CONTROL: the
 * `self.acados_solver.solve()` call and the `set / get / constraints_set / cost_set / get_cost /
 * get_stats / get_from_qp_in / reset` calls around it inside the three controller classes
 * (reference, paths relative to /root/reference):
 *     Model_Predictive_Controller/Nominal_NMPC/NMPC_class.py:111-112,169-206,243-254,290-317
 *     Model_Predictive_Controller/Stochastic_NMPC/SNMPC_class.py:120-130,184-222,263-272
 *     Model_Predictive_Controller/Reduced_Robustified_NMPC/Reduced_Robustified_NMPC_class.py:90-91,254-370
 * In the reference those calls go through acados_template's ctypes binding into a generated
 * acados/HPIPM solver; here they go through this header into hand-written sm_100a CUDA.
 *
 * Conventions
 *  - plain C, no torch types; every array is C-contiguous float64 unless stated.
 *  - every per-instance array gains a LEADING batch axis: acados `double[n]` becomes `double[B][n]`.
 *    Functions taking `ld_batch` accept ld_batch == 0 to broadcast one value set to all instances.
 *  - host pointers by default; *_dev variants take device pointers on the handle's device and are
 *    stream-ordered on the handle's stream (zero-copy path for on-device closed loops).
 *  - return value: 0 = ok, < 0 = API misuse (B200NMPC_E_*).  Nothing throws across the ABI.
 *    Solver outcomes are per-instance status codes with acados' values, written by b200nmpc_solve / b200nmpc_get_status.
 *    This implementation produces 0 (ACADOS_SUCCESS) and 4 (ACADOS_QP_FAILURE: the QP broke down, iterate left untouched)
 *    only: like the reference's SQP_RTI configuration the solve keeps status 0 when the QP stops at its iteration limit
 *    (B200NMPC_S_QP_ITER / B200NMPC_S_KKT tell), and 1 (NaN) / 2 (max-iter) / 3 (min-step) are never produced.
 *  - one handle = one CUDA device + one stream; a handle is not thread-safe, distinct handles are
 *    independent.  One process per GPU for multi-GPU (the batch is sharded by the caller).
 *  - the warm start (x_k, u_k) persists across solves un-shifted, per-stage weights / penalties /
 *    bounds are runtime-mutable and persist; b200nmpc_create leaves x_k = 0, u_k = 0 (the reference
 *    then fixes x0 via constraints_set(0,"lbx"/"ubx") -- acados' initial guess is x_k = x0).
 */
#ifndef B200NMPC_H
#define B200NMPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NMPC_NX 8
#define B200NMPC_NU 2
#define B200NMPC_NY 6
#define B200NMPC_NYE 4
#define B200NMPC_NC 3          /* soft rows per stage in acados slack order [bu, bx, h] */
#define B200NMPC_MAX_GGV 16
#define B200NMPC_MAX_N 128
#define B200NMPC_MAX_NS 16     /* SNMPC collocation samples */
#define B200NMPC_PCE_L 10      /* SNMPC PCE basis size (degree 2, 3 uncertain states) */
#define B200NMPC_MAX_UPH 8     /* SNMPC uncertainty propagation horizon handled by the condensed head */

/* error codes */
enum {
    B200NMPC_OK = 0,
    B200NMPC_E_ARG = -1,        /* null pointer / bad size */
    B200NMPC_E_STAGE = -2,      /* stage out of range for this field */
    B200NMPC_E_FIELD = -3,      /* unknown field id */
    B200NMPC_E_CUDA = -4,       /* CUDA runtime error (see b200nmpc_last_error) */
    B200NMPC_E_CONFIG = -5,     /* invalid configuration */
    B200NMPC_E_NODEVICE = -6    /* no usable CUDA device: there is NO CPU fallback */
};

/* OCP variants (which reference controller the handle replaces) */
enum {
    B200NMPC_NOMINAL = 0,       /* Nominal_NMPC/NMPC_STM_acados_settings.py */
    B200NMPC_R2 = 1,            /* Reduced_Robustified_NMPC: nominal OCP + fused post-solve tightening */
    B200NMPC_SNMPC = 2          /* Stochastic_NMPC/SNMPC_acados_settings.py */
};

/* field ids: the strings of AcadosOcpSolver.set/get/constraints_set/cost_set */
enum {
    B200NMPC_F_X = 0,           /* set/get(stage,"x")          [B][8]   stage 0..N  (SNMPC: [B][8(1+ns)], the augmented
                                   state [x; x^1..x^ns] of pred_model_dynamic_disc.py:115-186)        */
    B200NMPC_F_U = 1,           /* set/get(stage,"u")          [B][2]   stage 0..N-1 */
    B200NMPC_F_YREF = 2,        /* set(stage,"yref")           [B][6] (stage N: [B][4]) */
    B200NMPC_F_P = 3,           /* set(stage,"p")   SNMPC: [A(L x ns) row-major, gamma, stop_flag] (SNMPC_class.py:120-125);
                                   one value set for all instances (ld_batch = 0); the stop flags must be 0 on a
                                   prefix of at most B200NMPC_MAX_UPH stages and 1 afterwards (:97-98) */
    B200NMPC_F_LBX = 10,        /* constraints_set: stage 0: x0 [B][8] (SNMPC: the x0 sample matrix [B][8(1+ns)],
                                   SNMPC_class.py:120-121,259-264); stage>=1: delta_f bound [B][1] */
    B200NMPC_F_UBX = 11,
    B200NMPC_F_LBU = 12,        /* steering-rate bound [B][1], stage 0..N-1 */
    B200NMPC_F_UBU = 13,
    B200NMPC_F_LH = 14,         /* combined-acceleration bound [B][1], stage 1..N */
    B200NMPC_F_UH = 15,
    B200NMPC_F_W = 20,          /* cost_set(stage,"W"): diagonal [B][6] (stage N: [B][4]) */
    B200NMPC_F_ZL_LIN = 21,     /* cost_set(stage,"zl"): [B][ns_stage]  ns = 1 / 3 / 2 for stage 0 / mid / N */
    B200NMPC_F_ZU_LIN = 22,     /* "zu" */
    B200NMPC_F_ZL_QUAD = 23,    /* "Zl" */
    B200NMPC_F_ZU_QUAD = 24,    /* "Zu" */
    B200NMPC_F_SL = 30,         /* get(stage,"sl") / "su": slack values of the last QP, [B][ns_stage] */
    B200NMPC_F_SU = 31,
    B200NMPC_F_QP_A = 40,       /* get_from_qp_in(stage,"A") [B][8][8], stage 0..N-1 */
    B200NMPC_F_QP_B = 41,       /* get_from_qp_in(stage,"B") [B][8][2] */
    B200NMPC_F_QP_b = 42        /* get_from_qp_in(stage,"b") [B][8]    */
};

/* get_stats ids */
enum {
    B200NMPC_S_TIME_TOT = 0,    /* seconds of device time of the last solve, divided by B (acados: time_tot) */
    B200NMPC_S_SQP_ITER = 1,    /* always 1 (real-time iteration) */
    B200NMPC_S_QP_ITER = 2,     /* interior-point iterations of the last solve, per instance */
    B200NMPC_S_KKT = 3,         /* max complementarity / residual of the returned QP point, per instance */
    B200NMPC_S_TIME_BATCH = 4,  /* seconds of device time of the last solve for the whole batch */
    B200NMPC_S_TIME_LINEARIZE = 5, /* double[1]: device seconds of the linearisation kernel of the last solve */
    B200NMPC_S_TIME_QP = 6      /* double[1]: device seconds of the QP/step kernel(s) of the last solve */
};

/* Everything acados_settings(...) bakes into the generated solver, as a POD.
 * Values: Config/EDGAR/{veh_params_pred,pacejka_params}.yaml, ggv.csv, MPC_params.yaml and the
 * literals of Prediction_Models/pred_model_dynamic_stm_pacejka.py:38-46. */
typedef struct b200nmpc_config {
    /* vehicle + tyres */
    double lf, lr, m, Iz, ro, S, Cd;
    double Bf, Cf, Df, Ef, Br, Cr, Dr, Er;
    double g, fr0, fr1, fr4;
    double ax_brake;                               /* -acc_min */
    double ggv_v[B200NMPC_MAX_GGV], ggv_ax[B200NMPC_MAX_GGV], ggv_ay[B200NMPC_MAX_GGV];
    /* horizon / integrator: Tf = N*dt, ERK4 with rk_steps sub-steps (SNMPC: 1 step, DISCRETE) */
    double dt;
    /* default stage data (runtime-mutable through cost_set / constraints_set) */
    double W[6], W_e[4];
    double zl, zu, Zl, Zu;                         /* L1 / L2 slack penalties, all soft rows */
    double lbu, ubu, lbx, ubx, lh, uh;
    /* QP solve */
    double ipm_tol, ipm_reg;
    double ipm_eta_max;                            /* cap of the fraction-to-the-boundary parameter max(0.995, 1 - gap) */
    double ipm_s0;                                 /* initial slack of the interior start: s = ipm_s0 + violation */
    /* R2NMPC (Reduced_Robustified_NMPC_class.py:96-199) */
    double r2_sigma0_diag[8];                      /* diag of Sigma_0 */
    double r2_wd[4];                               /* diag of B W_d B' on states 2..5 */
    /* SNMPC (SNMPC_class.py:76-99, stochastic_mpc_utils.py:56-91) */
    double sn_x0_std[8];                           /* stds of compute_x0dist */
    double sn_gamma;                               /* risk parameter: kappa = sqrt((1-gamma)/gamma) */
    double sn_A[B200NMPC_PCE_L][B200NMPC_MAX_NS];  /* PCE least-squares matrix A = (Phi'Phi)^-1 Phi', rows [0,sn_L) x cols [0,n_samples) */
    double sn_w_off[B200NMPC_MAX_NS][8];           /* x0 sample offsets std (.) w_i: x0_i = x0 + sn_w_off[i] */
    int n_ggv;
    int N;
    int rk_steps;
    int variant;
    int ipm_max_iter;
    int qp_no_shortcut;
    int uph;                                       /* uncertainty propagation horizon (R2, SNMPC) */
    int n_samples;                                 /* SNMPC collocation samples */
    int sn_L;                                      /* SNMPC number of PCE terms (<= B200NMPC_PCE_L) */
    int ipm_corr_mode;                             /* Mehrotra second-order term scaled by 1 (0), a_aff (1), a_aff^2 (2) */
} b200nmpc_config;

typedef struct b200nmpc_handle b200nmpc_handle;

/* lifetime ------------------------------------------------------------------------------------ */
int b200nmpc_create(const b200nmpc_config *cfg, int batch, int device, b200nmpc_handle **out);
int b200nmpc_destroy(b200nmpc_handle *h);
const char *b200nmpc_last_error(void);
int b200nmpc_version(void);
int b200nmpc_sizeof_config(void);
int b200nmpc_batch(const b200nmpc_handle *h);
int b200nmpc_horizon(const b200nmpc_handle *h);
/* CUDA stream the handle launches on (cudaStream_t as void*), for callers that order their own work */
void *b200nmpc_stream(const b200nmpc_handle *h);

/* AcadosOcpSolver.set(stage, field, value): field in {X, U, YREF, P}.
 * value: [B][n] with ld_batch = n, or [n] with ld_batch = 0 (broadcast).  n is implied by field/stage. */
int b200nmpc_set(b200nmpc_handle *h, int stage, int field, const double *value, int ld_batch);
/* AcadosOcpSolver.constraints_set(stage, field, value): field in {LBX, UBX, LBU, UBU, LH, UH} */
int b200nmpc_constraints_set(b200nmpc_handle *h, int stage, int field, const double *value, int ld_batch);
/* AcadosOcpSolver.cost_set(stage, field, value): field in {W, ZL_LIN, ZU_LIN, ZL_QUAD, ZU_QUAD}.
 * W is the DIAGONAL of the weight matrix (the reference only ever sets diagonal W). */
int b200nmpc_cost_set(b200nmpc_handle *h, int stage, int field, const double *value, int ld_batch);

/* AcadosOcpSolver.solve(): one RTI step for every instance.  status: int[B] (host) or NULL.
 * Asynchronous on the handle's stream when status == NULL; synchronises otherwise. */
int b200nmpc_solve(b200nmpc_handle *h, int *status);
int b200nmpc_get_status(b200nmpc_handle *h, int *status);
/* AcadosOcpSolver.options_set(field, value) for the options of this implementation:
 *   "launch_order_hint" (default 1): the QP kernel starts the instances that needed the most interior-point iterations
 *       in the handle's PREVIOUS solve first (one warp per instance, a batch is a few waves of warps: the instances that
 *       start last should be the short ones).  Results do not depend on it.  Setting it (to 0 or 1) forgets the history,
 *       so the next solve runs in index order.  B200NMPC_NO_LPT in the environment sets the default to 0. */
int b200nmpc_options_set(b200nmpc_handle *h, const char *field, double value);

/* AcadosOcpSolver.get(stage, field): field in {X, U, SL, SU}; get_from_qp_in: {QP_A, QP_B, QP_b}.
 * out: [B][n].  Synchronises the handle's stream. */
int b200nmpc_get(b200nmpc_handle *h, int stage, int field, double *out);
/* AcadosOcpSolver.get_cost(): double[B] */
int b200nmpc_get_cost(b200nmpc_handle *h, double *out);
/* AcadosOcpSolver.get_stats(name): double[B] (TIME_BATCH: double[1]) */
int b200nmpc_get_stats(b200nmpc_handle *h, int stat, double *out);
/* AcadosOcpSolver.reset(): zero the iterate (the reference then calls set(i,'x',x0) for all i,
 * NMPC_class.py:250-254) */
int b200nmpc_reset(b200nmpc_handle *h);

/* bulk variants: O(1) calls per solve instead of the reference's O(N) per-stage loop -------------- */
int b200nmpc_set_x0(b200nmpc_handle *h, const double *x0 /*[B][8]*/);                   /* = lbx_0 = ubx_0 */
int b200nmpc_set_yref_all(b200nmpc_handle *h, const double *yref /*[B][N+1][6]*/);
/* read the references back (the on-device planner of b200nmpc_sim_* writes them) */
int b200nmpc_get_yref_all(b200nmpc_handle *h, double *yref /*[B][N+1][6]*/);
int b200nmpc_set_iterate(b200nmpc_handle *h, const double *X /*[B][N+1][8] or NULL*/,
                         const double *U /*[B][N][2] or NULL*/);
/* x_k = x0 for all k, u = 0: the state right after the reference constructs a solver / calls reset(x0) */
int b200nmpc_init_iterate_from_x0(b200nmpc_handle *h);
int b200nmpc_get_iterate(b200nmpc_handle *h, double *X /*[B][N+1][8] or NULL*/, double *U /*[B][N][2] or NULL*/);
/* [B][5] rows = [cost, time_tot, sqp_iter, qp_iter, status]: the stats tuple of NMPC_class.py:202-206 */
int b200nmpc_get_solver_debug(b200nmpc_handle *h, double *out);
/* reference update_cost_function_weights(params[7]) (NMPC_class.py:269-317) for every stage at once;
 * params: [B][7] (ld_batch = 7) or [7] (ld_batch = 0) = [q_xy, q_yaw, q_vel, r_jerk, r_steer, L1, L2] */
int b200nmpc_update_cost_function_weights(b200nmpc_handle *h, const double *params, int ld_batch);

/* device-pointer variants (stream-ordered, no synchronisation) */
int b200nmpc_set_x0_dev(b200nmpc_handle *h, const double *x0_dev);
int b200nmpc_set_yref_all_dev(b200nmpc_handle *h, const double *yref_dev);
int b200nmpc_get_iterate_dev(b200nmpc_handle *h, double *X_dev, double *U_dev);
int b200nmpc_get_solver_debug_dev(b200nmpc_handle *h, double *out_dev);
int b200nmpc_synchronize(b200nmpc_handle *h);
/* asynchronous read-back of the iterate and of the [B][5] stats rows (time_tot = 0) into PINNED host buffers
 * (b200nmpc_alloc_host), ordered on the handle's stream: complete after b200nmpc_synchronize.  With two handles on
 * one device the copies of one overlap the solve of the other.  Any pointer may be NULL. */
int b200nmpc_get_results_async(b200nmpc_handle *h, double *X, double *U, double *solver_debug);

/* pinned (page-locked) host buffers so that the H2D/D2H copies of the bulk calls are true async DMA */
void *b200nmpc_alloc_host(size_t bytes);
void b200nmpc_free_host(void *p);

/* CUDA-event stopwatch on the handle's stream (the stream its kernels are launched on) */
int b200nmpc_timer_start(b200nmpc_handle *h);
int b200nmpc_timer_stop(b200nmpc_handle *h, double *seconds);   /* synchronises */

/* launch accounting: number of kernels launched by this handle since creation */
long long b200nmpc_launch_count(const b200nmpc_handle *h);

/* SNMPC (variant B200NMPC_SNMPC) -------------------------------------------------------------------------
 * The augmented state of the reference ([x; x^1..x^ns], SNMPC_acados_settings.py / pred_model_dynamic_disc.py)
 * is kept as the nominal iterate X plus the sample iterates XS of stages 0..UPH (later stages are frozen
 * copies of stage UPH, pred_model_dynamic_disc.py:204-210). */
/* x0 sample matrix of every instance, [B][1+ns][8] = compute_x0dist (bulk form of
 * constraints_set(0,"lbx"/"ubx", x0_samples.flatten()), SNMPC_class.py:259-264).  b200nmpc_set_x0 on an
 * SNMPC handle derives it as x0 + sn_w_off. */
int b200nmpc_set_x0_samples(b200nmpc_handle *h, const double *x0s);
/* sample iterates [B][UPH+1][ns][8] */
int b200nmpc_set_samples(b200nmpc_handle *h, const double *XS);
int b200nmpc_get_samples(b200nmpc_handle *h, double *XS);
/* current uncertainty propagation horizon (number of leading stages with stop_flag == 0) */
int b200nmpc_snmpc_uph(const b200nmpc_handle *h);
/* diagnostics for the parity tests: the condensed head of the last solve, [B][stride] doubles
 * (layout: tum_control_b200/csrc/nmpc_kernels.cuh, HD_*); out may be NULL to query the stride only */
int b200nmpc_snmpc_get_head(b200nmpc_handle *h, double *out, int *stride);

/* Closed loop on the device (the callers either side of the solve: SURVEY section 8f rows 1-2) ------------------
 * B independent closed-loop simulations advance without touching the host between solves:
 *   planner window      Utils/MPC_sim_utils.py:137-194 (PlannerEmulator)           -> yref of the solver
 *   solve               b200nmpc_solve
 *   plant step          Vehicle_Simulator/VehicleSimulator.py:60-94 (RK4, plant_substeps finite elements over Ts)
 *   state estimation    Utils/SimulationMode_main_class.py:10-19,152-156 (truncated moving averages)  -> next x0
 *   logging             Utils/Logging_Plotting.py:124-145 (MPC_SimX, CiLX, simU, simREF, simSolverDebug)
 * exactly in the order of main.py:48-78 (simMode 0, "controller in the loop"). */
typedef struct b200nmpc_sim b200nmpc_sim;
typedef struct b200nmpc_sim_config {
    double Ts;                  /* simulation step, sim_main_params.yaml:40 */
    double Tp;                  /* prediction horizon in seconds (planner window) */
    int plant_substeps;         /* VehicleSimulator.py:73 (4) */
    int log_capacity;           /* steps the log arrays hold */
    int n_ref;                  /* reference trajectory points */
    int windows[8];             /* SimulationMode_main_class.py:86 */
} b200nmpc_sim_config;
/* ref_*: host arrays [n_ref] of the reference trajectory JSON (pos_x, pos_y, ref_yaw, ref_v) */
int b200nmpc_sim_create(b200nmpc_handle *solver, const b200nmpc_sim_config *cfg, const double *ref_px, const double *ref_py,
                        const double *ref_yaw, const double *ref_v, b200nmpc_sim **out);
int b200nmpc_sim_destroy(b200nmpc_sim *s);
/* X0_MPC [B][8] (SimulationMode_main_class.py:66): plant state, pose, empty filter, solver x0 and initial guess */
int b200nmpc_sim_reset(b200nmpc_sim *s, const double *X0_MPC);
/* Reset of single instances while the batch keeps its step counter (an environment of a vectorised RL run ends its
 * episode: RL_WMPC/environment.py reset() -> MPC.reset(X0_MPC), NMPC_class.py:243-254): for every b with mask[b] != 0
 * the plant, pose, filter, stop state, x0 and the initial guess (x_k = x0, u = 0) are set from X0_MPC[b]; the
 * instance's log rows continue at the current step.  mask int [B], X0_MPC [B][8] (rows of unmasked instances ignored). */
int b200nmpc_sim_reset_instances(b200nmpc_sim *s, const int *mask, const double *X0_MPC);
/* nsteps closed-loop steps, asynchronous on the solver's stream unless seconds != NULL (device time of the run).
 * B200NMPC_E_ARG if steps_done + nsteps would exceed the log capacity (the logs are linear arrays and the device-side
 * stop predicates read the log row of the step). */
int b200nmpc_sim_run(b200nmpc_sim *s, int nsteps, double *seconds);
int b200nmpc_sim_steps_done(const b200nmpc_sim *s);
/* Restarts the log at the current state: row 0 of MPC_SimX / CiLX / DisturbedX becomes the state after the last step,
 * steps_done returns to 0; vehicles, controller and filter state are untouched.  For callers that only read the newest
 * rows (RL_WMPC/environment.py:112-240: one 20-step chunk per step()) and run longer than the log capacity.  Not allowed
 * while stop conditions are enabled (they count steps from the last reset).  Disturbance playback rows are indexed by the
 * step counter too and restart at row 0. */
int b200nmpc_sim_rebase_logs(b200nmpc_sim *s);
/* logs of the steps done so far: MPC_SimX [B][n+1][8], CiLX [B][n+1][7], simU [B][n][2], simREF [B][n][4],
 * simSolverDebug [B][n][5] (any pointer may be NULL) */
int b200nmpc_sim_get_logs(b200nmpc_sim *s, double *MPC_SimX, double *CiLX, double *simU, double *simREF, double *simSolverDebug);
/* the same for the steps [first, first + count) only -- what a chunked caller (BO objective, RL environment) needs
 * after every sim_run: MPC_SimX [B][count+1][8] and CiLX [B][count+1][7] start at row `first`, the others hold
 * `count` rows; ref_idx int [B][count]; any pointer may be NULL */
int b200nmpc_sim_get_logs_range(b200nmpc_sim *s, int first, int count, double *MPC_SimX, double *CiLX, double *simU, double *simREF,
                                double *simSolverDebug, int *ref_idx);
/* Stop conditions of the weight-adaptation callers, evaluated on the device after every closed-loop step
 * (BO_WMPC/objective_function.py:153 `done = current_ref_idx == segment['end']`, :186-200 crash predicates
 * `lat_devs[step-1] > max_lat_dev or acomb > max_a_comb`; RL_WMPC/environment.py:239-240 uses the first only:
 * pass max_a_comb = +inf).  A stopped instance keeps its state and logs and is skipped by every kernel of the
 * following steps (its warp exits at once), so a batch of segments of different lengths costs the sum of the
 * lengths, not B x the longest.  end_idx int [B] or NULL (no end condition); enable = 0 turns the predicates off.
 * sim_reset clears the stop state. */
int b200nmpc_sim_set_stop(b200nmpc_sim *s, int enable, const int *end_idx, double max_lat_dev, double max_a_comb);
/* stop_step int [B]: -1 while running, else the number of steps simulated including the one that stopped the
 * instance; flags int [B]: bit 0 done (end index reached), bit 1 crash.  Either pointer may be NULL. */
int b200nmpc_sim_get_stop(b200nmpc_sim *s, int *stop_step, int *flags);
/* closest reference index of every step done so far (current_ref_idx of PlannerEmulator), int [B][n] */
int b200nmpc_sim_get_ref_idx(b200nmpc_sim *s, int *out);
/* additive disturbances applied in every following step until changed (Utils/MPC_sim_utils.py:54-99): on the
 * state derivatives of the plant and on the state estimate, [B][7] each, NULL = none */
int b200nmpc_sim_set_disturbances(b200nmpc_sim *s, const double *deriv, const double *est);
/* disturbances generated on the device (Utils/MPC_sim_utils.py:54-84), one channel for the state derivatives of the
 * plant and one for the state estimate: type 0 off, 1 'uniform' (= sampleFromEllipsoid(0, diag(bounds)), :56-57,68-84),
 * 2 'gaussian' (N(0, bound_i^2), :60-62), 3 'absolute' (:63-64); bounds = the w_* of sim_main_params.yaml:54-79.
 * Counter-based generator (Philox4x32-10 keyed by seed; counter = instance, step, channel): reproducible per vehicle,
 * independent of batch size and sharding.  As in the reference (SimulationMode_main_class.py:119-143) the disturbed
 * plant state only feeds the controller's state estimate; the undisturbed state continues as the plant state. */
int b200nmpc_sim_config_disturbances(b200nmpc_sim *s, int type_deriv, const double *bounds_deriv /*[7]*/, int type_est,
                                     const double *bounds_est /*[7]*/, unsigned long long seed);
/* disturbance_playback (sim_main_params.yaml:45-48): rows [B][n][7] of a previous run's sim_disturbance_* logs; NULL = off */
int b200nmpc_sim_set_playback(b200nmpc_sim *s, const double *deriv, const double *est, int n);
/* DisturbedX [B][n+1][7], sim_disturbance_derivatives [B][n][7], sim_disturbance_state_estimation [B][n][7] (any may be NULL) */
int b200nmpc_sim_get_disturbance_logs(b200nmpc_sim *s, double *DisturbedX, double *deriv, double *est);
/* planner only: fill the solver's yref from poses [B][4] (x, y, yaw, |v|); closest: int[B] or NULL */
int b200nmpc_sim_plan(b200nmpc_sim *s, const double *pose, int *closest);

/* R2NMPC: run the fused tightening kernel on the last solve (done automatically inside solve for
 * variant R2; exposed for tests).  Reduced_Robustified_NMPC_class.py:287-365 */
int b200nmpc_r2_tighten(b200nmpc_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* B200NMPC_H */
