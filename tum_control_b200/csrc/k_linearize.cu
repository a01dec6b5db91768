// k_linearize.cu -- K1: multiple-shooting linearisation, one thread per (instance, stage).
//
// Replaces, for every shooting node of every instance, what acados' preparation phase computes
// through CasADi-generated C:
//   * ERK4 x rk_steps integration of f with forward sensitivities  -> A_k, B_k, b_k
//     (NMPC_STM_acados_settings.py:238-240, model Prediction_Models/pred_model_dynamic_stm_pacejka.py:118-177)
//   * Gauss-Newton gradient of the dt-scaled NLS tracking cost      -> q_k, r_k   (:41-60)
//   * combined-acceleration constraint h and its gradient           -> h_k, grad h_k (:70-74,107-109)
//
// B x (N+1) independent work items, stage index fastest: a warp reads 32 consecutive stage states of the
// instance-major iterate and writes 32 consecutive LIN records.  The sensitivity matrix is
// propagated column by column in its structural 6x8 form (rows px..r; columns yaw,vl,vt,r,delta,a,
// jerk,omega): columns px,py of A are unit vectors and the delta/a rows are pure integrators whose
// sensitivities are known in closed form.
#ifndef LIN_IEEE_DIV
#define NMPC_MODEL_FAST_RECIP      // branch-free reciprocals / reciprocal square roots in the model evaluation (nmpc_model.cuh)
#endif
#include "nmpc_kernels.cuh"

namespace b200 {

__device__ __forceinline__ double stage_val(const double *shared, const double *inst, int b, int NS, int k, int ncomp, int comp)
{
    return inst ? inst[((size_t)b * NS + k) * ncomp + comp] : shared[k * ncomp + comp];
}

// KIND: structure of the (delta, a) rows of this sensitivity column
//   0: both zero (columns yaw, vl, vt, r)   1: delta row = 1 (column delta)   2: a row = 1 (column a)
//   3: a row = tau (column jerk)            4: delta row = tau (column omega)
template <int KIND>
__device__ __forceinline__ void sens_deriv(const StageJac &J, const double *s, double tau, double *k)
{
    k[0] = -J.f1 * s[2] + J.cy * s[3] - J.sy * s[4];
    k[1] = J.f0 * s[2] + J.sy * s[3] + J.cy * s[4];
    k[2] = s[5];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double v = J.J[i][0] * s[3] + J.J[i][1] * s[4] + J.J[i][2] * s[5];
        if (KIND == 1) v += J.J[i][3];
        if (KIND == 2) v += J.J[i][4];
        if (KIND == 3) v += J.J[i][4] * tau;
        if (KIND == 4) v += J.J[i][3] * tau;
        k[3 + i] = v;
    }
}

template <int KIND>
__device__ __forceinline__ void sens_col_step(const StageJac *J, double h, double t0, double *s)
{
    double k[6], acc[6], t[6];
    sens_deriv<KIND>(J[0], s, t0, k);
#pragma unroll
    for (int i = 0; i < 6; i++) { acc[i] = k[i]; t[i] = s[i] + 0.5 * h * k[i]; }
    sens_deriv<KIND>(J[1], t, t0 + 0.5 * h, k);
#pragma unroll
    for (int i = 0; i < 6; i++) { acc[i] += 2.0 * k[i]; t[i] = s[i] + 0.5 * h * k[i]; }
    sens_deriv<KIND>(J[2], t, t0 + 0.5 * h, k);
#pragma unroll
    for (int i = 0; i < 6; i++) { acc[i] += 2.0 * k[i]; t[i] = s[i] + h * k[i]; }
    sens_deriv<KIND>(J[3], t, t0 + h, k);
#pragma unroll
    for (int i = 0; i < 6; i++) s[i] += (h / 6.0) * (acc[i] + k[i]);
}

// ERK4 x rk_steps of the 8-state model from x (overwritten by the end state) with the structured forward
// sensitivities S[c][row]: columns yaw, vl, vt, r, delta, a, jerk, omega of the rows px..r.
__device__ __forceinline__ void erk4_sens(const DevCfg &cfg, double *x, const double *u, double (*S)[6])
{
#pragma unroll
    for (int c = 0; c < 8; c++)
#pragma unroll
        for (int r = 0; r < 6; r++) S[c][r] = 0.0;
    S[0][2] = 1.0; S[1][3] = 1.0; S[2][4] = 1.0; S[3][5] = 1.0;

    const double h = cfg.dt / cfg.rk_steps;
    for (int step = 0; step < cfg.rk_steps; step++) {
        StageJac J[4];
        double k1[6], k2[6], k3[6], k4[6], xt[8];
        const double t0 = step * h;
        stm_eval<true>(cfg, x, k1, &J[0]);
#pragma unroll
        for (int i = 0; i < 6; i++) xt[i] = x[i] + 0.5 * h * k1[i];
        xt[6] = x[6] + 0.5 * h * u[1]; xt[7] = x[7] + 0.5 * h * u[0];
        stm_eval<true>(cfg, xt, k2, &J[1]);
#pragma unroll
        for (int i = 0; i < 6; i++) xt[i] = x[i] + 0.5 * h * k2[i];
        stm_eval<true>(cfg, xt, k3, &J[2]);
#pragma unroll
        for (int i = 0; i < 6; i++) xt[i] = x[i] + h * k3[i];
        xt[6] = x[6] + h * u[1]; xt[7] = x[7] + h * u[0];
        stm_eval<true>(cfg, xt, k4, &J[3]);
#pragma unroll
        for (int i = 0; i < 6; i++) x[i] += (h / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
        x[6] += h * u[1]; x[7] += h * u[0];
        sens_col_step<0>(J, h, t0, S[0]);
        sens_col_step<0>(J, h, t0, S[1]);
        sens_col_step<0>(J, h, t0, S[2]);
        sens_col_step<0>(J, h, t0, S[3]);
        sens_col_step<1>(J, h, t0, S[4]);
        sens_col_step<2>(J, h, t0, S[5]);
        sens_col_step<3>(J, h, t0, S[6]);
        sens_col_step<4>(J, h, t0, S[7]);
    }
}

// K4a: SNMPC per-sample linearisation of one (instance, stage < UPH, sample) item.  Replaces what CasADi's AD of the
// 8(1+n_s)-state discrete map computes for the sample blocks (pred_model_dynamic_disc.py:187-210: one RK4 step of dt per
// sample with the shared input) and the sample values of the acceleration constraint (SNMPC_acados_settings.py:86-130).
// Writes one SREC record: A36 | B12 | RK4 image | h | grad h.  The reference evaluates the nominal copy and the n_s
// sample copies in ONE map; here they are work items of ONE launch (k_linearize): items [0, B (N+1)) are the nominal
// (instance, stage) pairs, items [B (N+1), B (N+1) + B H n_s) the sample copies of the propagation horizon.
__device__ __noinline__ void linearize_sample(const DevCfg &cfg, const DevMem &m, long long g)
{
    const int N = cfg.N, H = cfg.head_H, ns = cfg.n_samples;
    const int b = (int)(g / (H * ns));
    if (m.halt && m.halt[b] >= 0) return;
    const int r = (int)(g - (long long)b * H * ns);
    const int k = r / ns, s = r - k * ns;
    const double *xs = m.XS + (((size_t)b * (m.HA + 1) + k) * ns + s) * 8;
    double *rec = m.slin + (((size_t)b * m.HA + k) * ns + s) * SREC;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = xs[i];
    {
        double g4[4];
        rec[SR_H] = h_eval(cfg, x, g4);
#pragma unroll
        for (int i = 0; i < 4; i++) rec[SR_GH + i] = g4[i];
    }
    const double *uk = m.U + ((size_t)b * N + k) * 2;
    const double u[2] = {uk[0], uk[1]};
    double S[8][6];
    erk4_sens(cfg, x, u, S);
#pragma unroll
    for (int rr = 0; rr < 6; rr++) {
#pragma unroll
        for (int c = 0; c < 6; c++) rec[SR_A + rr * 6 + c] = S[c][rr];
        rec[SR_B + rr * 2 + 0] = S[6][rr];
        rec[SR_B + rr * 2 + 1] = S[7][rr];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) rec[SR_XN + i] = x[i];
}

#ifndef LIN_MINB
#define LIN_MINB 2      // resident blocks per SM the kernel is compiled for (2: up to 255 registers)
#endif
// The LIN part of an item's stage record (LREC doubles from LIN_OFF) is built in a shared-memory row of the thread and
// written out by the whole warp: the LIN parts of 32 consecutive items are 32 x 640 contiguous bytes 992 bytes apart, so the
// warp stores them with 128-bit accesses, 512 contiguous bytes per instruction.  (Each thread storing its own record scalar by scalar put the 32 lanes of every
// store instruction 960 bytes apart: 32 sectors per instruction, 47 MB of partial-sector writes per launch.)
constexpr int LROW = LREC + 1;      // odd row stride: the per-thread rows are bank-conflict free
#ifdef LIN_DIRECT_STORE
#define LIN_OUT(rec, row) (rec)
#else
#define LIN_OUT(rec, row) (row)
#endif

// `rec`: start of the stage record (rec[L_*] are the fields; only the LIN part [LIN_OFF, LIN_OFF + LREC) is touched).
// returns 0 (LIN part written; a stage without dynamics -- terminal stage, SNMPC head stage -- gets zeros in A, B, b) or
// LREC (nothing: stopped instance)
__device__ __forceinline__ int linearize_nominal(const DevCfg &cfg, const DevMem &m, int b, int k, double *rec)
{
    const int N = cfg.N, NS = N + 1;
    if (m.halt && m.halt[b] >= 0) return LREC;                           // stopped closed-loop instance
    const double *xk = m.X + ((size_t)b * NS + k) * 8;
    const double *yr = m.yref + ((size_t)b * NS + k) * 6;

    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = xk[i];
    const double sig = (k < N) ? cfg.dt : 1.0;      // EULER cost discretisation: stages scaled by dt

    // ---- Gauss-Newton gradient of 1/2 |y - yref|^2_W, y = [px, py, wrap(yaw), v, jerk, omega]
    {
        const double w0 = stage_val(m.W_s, m.W_i, b, NS, k, 6, 0), w1 = stage_val(m.W_s, m.W_i, b, NS, k, 6, 1);
        const double w2 = stage_val(m.W_s, m.W_i, b, NS, k, 6, 2), w3 = stage_val(m.W_s, m.W_i, b, NS, k, 6, 3);
        rec[L_QX + 0] = sig * w0 * (x[0] - yr[0]);
        rec[L_QX + 1] = sig * w1 * (x[1] - yr[1]);
        rec[L_QX + 2] = sig * w2 * (wrap_yaw(x[2]) - yr[2]);
        const double vref = yr[3];
        if (cfg.cost_vel_abs) {                       // SNMPC_acados_settings.py:62-65
            const double V = sqrt(x[3] * x[3] + x[4] * x[4]);
            const double iV = V > 0.0 ? 1.0 / V : 0.0;
            const double g3 = x[3] * iV, g4 = x[4] * iV;
            rec[L_QX + 3] = sig * w3 * (V - vref) * g3;
            rec[L_QX + 4] = sig * w3 * (V - vref) * g4;
            rec[L_QV + 0] = sig * w3 * g3 * g3;
            rec[L_QV + 1] = sig * w3 * g3 * g4;
            rec[L_QV + 2] = sig * w3 * g4 * g4;
        } else {
            rec[L_QX + 3] = sig * w3 * (x[3] - vref);
            rec[L_QX + 4] = 0.0;
            rec[L_QV + 0] = 0.0; rec[L_QV + 1] = 0.0; rec[L_QV + 2] = 0.0;
        }
        rec[L_H + 1] = 0.0;                           // (pad)
    }
    // ---- soft nonlinear row h (stages 1..N; evaluated everywhere, stage 0 is masked in the QP)
    {
        double g4[4];
        rec[L_H] = h_eval(cfg, x, g4);
#pragma unroll
        for (int i = 0; i < 4; i++) rec[L_GH + i] = g4[i];
        rec[L_GF + 0] = 0.0; rec[L_GF + 1] = 0.0; rec[L_GF + 2] = 0.0; rec[L_GF + 3] = g4[0];
        rec[L_GF + 4] = g4[1]; rec[L_GF + 5] = g4[2]; rec[L_GF + 6] = 0.0; rec[L_GF + 7] = g4[3];
    }
    rec[L_RG + 0] = 0.0; rec[L_RG + 1] = 0.0;
    // SNMPC: stages with stop_flag == 0 are linearised per sample; the terminal stage has no dynamics
    const bool no_dyn = (k == N) || (cfg.head_H > 0 && k < cfg.head_H);
    if (no_dyn) {
        for (int i = 0; i < 48; i++) rec[L_A + i] = 0.0;
        for (int i = 0; i < 8; i++) rec[L_b + i] = 0.0;
    }
    if (k == N) return 0;

    const double *uk = m.U + ((size_t)b * N + k) * 2;
    double u[2] = {uk[0], uk[1]};
    {
        const double w4 = stage_val(m.W_s, m.W_i, b, NS, k, 6, 4), w5 = stage_val(m.W_s, m.W_i, b, NS, k, 6, 5);
        rec[L_RG + 0] = sig * w4 * (u[0] - yr[4]);
        rec[L_RG + 1] = sig * w5 * (u[1] - yr[5]);
    }

    // ---- ERK4 x rk_steps with structured forward sensitivities
    if (no_dyn) return 0;
#ifdef LIN_DIRECT_STORE
    double S[8][6];                 // S[c][row]: column c of the 6x8 block
    erk4_sens(cfg, x, u, S);
    // ---- outputs
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 6; c++) rec[L_A + c * 6 + r] = S[c][r];      // transposed: column c of A contiguous
        rec[L_B + r] = S[6][r];
        rec[L_B + 6 + r] = S[7][r];
    }
#else
    // the record stores A and B transposed, column c of [A | B] contiguous at L_A + 6 c: exactly S[c][row].  The
    // sensitivities are propagated IN PLACE in the thread's shared-memory row (odd row stride: conflict-free) instead of
    // a 48-double local array that the 255-register kernel kept in local memory (L1 / L2 round trips per column step)
    static_assert(L_B == L_A + 36, "S[c] = rec + L_A + 6 c");
    erk4_sens(cfg, x, u, reinterpret_cast<double (*)[6]>(rec + L_A));
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) rec[L_b + i] = x[i] - xk[8 + i];
    return 0;
}

__global__ void __launch_bounds__(128, LIN_MINB) k_linearize(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int B = cfg.B, NS = cfg.N + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_nom = (long long)B * NS;
#ifdef LIN_DIRECT_STORE
    if (g >= n_nom) {                                                    // SNMPC: the sample copies of the propagation horizon
        if (g - n_nom < (long long)B * cfg.head_H * cfg.n_samples) linearize_sample(cfg, m, g - n_nom);
        return;
    }
    const int b = (int)(g / NS), k = (int)(g - (long long)b * NS);      // stage index fastest: records are contiguous
    linearize_nominal(cfg, m, b, k, m.stg + (size_t)g * STG);
#else
    extern __shared__ __align__(16) double rows[];                       // [128][LROW]
    const int lane = threadIdx.x & 31;
    double *row = rows + (size_t)threadIdx.x * LROW;
    int first = LREC;
    if (g < n_nom) {
        const int b = (int)(g / NS), k = (int)(g - (long long)b * NS);  // stage index fastest: record g is item g
        first = linearize_nominal(cfg, m, b, k, row - LIN_OFF);       // (the row holds the LIN part only)
    } else if (g - n_nom < (long long)B * cfg.head_H * cfg.n_samples) {  // SNMPC: the sample copies of the propagation horizon
        linearize_sample(cfg, m, g - n_nom);
    }
    __syncwarp();
    // cooperative write-out of the warp's 32 rows: double2 units, 40 per record
    const long long g0 = g - lane;
    const double *wrows = rows + (size_t)(threadIdx.x - lane) * LROW;
    for (int r = 0; r < 32; r++) {
        const int fr = __shfl_sync(0xffffffffu, first, r);
        if (fr >= LREC) continue;                                        // (warp-uniform)
        double2 *dst = reinterpret_cast<double2 *>(m.stg + (size_t)(g0 + r) * STG + LIN_OFF);
        const double *src = wrows + r * LROW;
        for (int c = fr / 2 + lane; c < LREC / 2; c += 32) dst[c] = make_double2(src[2 * c], src[2 * c + 1]);
    }
#endif
}

void launch_linearize(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    // nominal (instance, stage) items, then (SNMPC) the sample copies of the propagation horizon: the item index selects the path
    const long long items = (long long)cfg.B * (cfg.N + 1) + (long long)cfg.B * cfg.head_H * cfg.n_samples;
#ifdef LIN_DIRECT_STORE
    k_linearize<<<(unsigned)((items + 127) / 128), 128, 0, st>>>(cfg, m);
#else
    const int smem = 128 * LROW * (int)sizeof(double);                   // 83 KB: two blocks per SM, as the registers allow
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(k_linearize, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set[dev] = true;
    }
    k_linearize<<<(unsigned)((items + 127) / 128), 128, smem, st>>>(cfg, m);
#endif
}

}  // namespace b200
