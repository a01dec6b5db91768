// k_linearize.cu -- K1: multiple-shooting linearisation, one thread per (instance, stage).
//
// Replaces, for every shooting node of every instance, what acados' preparation phase computes
// through CasADi-generated C:
//   * ERK4 x rk_steps integration of f with forward sensitivities  -> A_k, B_k, b_k
//     (NMPC_STM_acados_settings.py:238-240, model Prediction_Models/pred_model_dynamic_stm_pacejka.py:118-177)
//   * Gauss-Newton gradient of the dt-scaled NLS tracking cost      -> q_k, r_k   (:41-60)
//   * combined-acceleration constraint h and its gradient           -> h_k, grad h_k (:70-74,107-109)
//
// B x (N+1) independent work items; the batch index is the fastest thread index so that every global
// access of a warp is one contiguous 256-byte line of the SoA arrays.  The sensitivity matrix is
// propagated column by column in its structural 6x8 form (rows px..r; columns yaw,vl,vt,r,delta,a,
// jerk,omega): columns px,py of A are unit vectors and the delta/a rows are pure integrators whose
// sensitivities are known in closed form.
#include "nmpc_kernels.cuh"

namespace b200 {

#define AT(p, idx) (p)[(size_t)(idx) * B + b]

__device__ __forceinline__ double stage_val(const double *shared, const double *inst, int idx, int b, int B)
{
    return inst ? inst[(size_t)idx * B + b] : shared[idx];
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

__global__ void __launch_bounds__(128) k_linearize(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int B = cfg.B, N = cfg.N;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (b >= B) return;

    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = AT(m.X, k * 8 + i);
    const double sig = (k < N) ? cfg.dt : 1.0;      // EULER cost discretisation: stages scaled by dt

    // ---- Gauss-Newton gradient of 1/2 |y - yref|^2_W, y = [px, py, wrap(yaw), v, jerk, omega]
    {
        const double w0 = stage_val(m.W_s, m.W_i, k * 6 + 0, b, B), w1 = stage_val(m.W_s, m.W_i, k * 6 + 1, b, B);
        const double w2 = stage_val(m.W_s, m.W_i, k * 6 + 2, b, B), w3 = stage_val(m.W_s, m.W_i, k * 6 + 3, b, B);
        AT(m.qx, k * 5 + 0) = sig * w0 * (x[0] - AT(m.yref, k * 6 + 0));
        AT(m.qx, k * 5 + 1) = sig * w1 * (x[1] - AT(m.yref, k * 6 + 1));
        AT(m.qx, k * 5 + 2) = sig * w2 * (wrap_yaw(x[2]) - AT(m.yref, k * 6 + 2));
        const double vref = AT(m.yref, k * 6 + 3);
        if (cfg.cost_vel_abs) {                       // SNMPC_acados_settings.py:62-65
            const double V = sqrt(x[3] * x[3] + x[4] * x[4]);
            const double iV = V > 0.0 ? 1.0 / V : 0.0;
            const double g3 = x[3] * iV, g4 = x[4] * iV;
            AT(m.qx, k * 5 + 3) = sig * w3 * (V - vref) * g3;
            AT(m.qx, k * 5 + 4) = sig * w3 * (V - vref) * g4;
            AT(m.Qv, k * 3 + 0) = sig * w3 * g3 * g3;
            AT(m.Qv, k * 3 + 1) = sig * w3 * g3 * g4;
            AT(m.Qv, k * 3 + 2) = sig * w3 * g4 * g4;
        } else {
            AT(m.qx, k * 5 + 3) = sig * w3 * (x[3] - vref);
            AT(m.qx, k * 5 + 4) = 0.0;
        }
    }
    // ---- soft nonlinear row h (stages 1..N; evaluated everywhere, stage 0 is masked in the QP)
    {
        double g4[4];
        AT(m.hval, k) = h_eval(cfg, x, g4);
#pragma unroll
        for (int i = 0; i < 4; i++) AT(m.gh, k * 4 + i) = g4[i];
    }
    if (k == N) return;

    double u[2] = {AT(m.U, k * 2 + 0), AT(m.U, k * 2 + 1)};
    {
        const double w4 = stage_val(m.W_s, m.W_i, k * 6 + 4, b, B), w5 = stage_val(m.W_s, m.W_i, k * 6 + 5, b, B);
        AT(m.rg, k * 2 + 0) = sig * w4 * (u[0] - AT(m.yref, k * 6 + 4));
        AT(m.rg, k * 2 + 1) = sig * w5 * (u[1] - AT(m.yref, k * 6 + 5));
    }

    // ---- ERK4 x rk_steps with structured forward sensitivities
    double S[8][6];                 // S[c][row]: column c of the 6x8 block
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
    // ---- outputs
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 6; c++) AT(m.A36, k * 36 + r * 6 + c) = S[c][r];
        AT(m.B12, k * 12 + r * 2 + 0) = S[6][r];
        AT(m.B12, k * 12 + r * 2 + 1) = S[7][r];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) AT(m.bd, k * 8 + i) = x[i] - AT(m.X, (k + 1) * 8 + i);
}

void launch_linearize(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    dim3 block(128), grid((cfg.B + 127) / 128, cfg.N + 1);
    k_linearize<<<grid, block, 0, st>>>(cfg, m);
}

}  // namespace b200
