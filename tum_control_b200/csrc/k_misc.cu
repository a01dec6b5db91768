// k_misc.cu -- bookkeeping kernels of libb200nmpc (instance-major layout, see nmpc_kernels.cuh).
#include "nmpc_kernels.cuh"

namespace b200 {

__global__ void k_fill(double *dst, double v, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}
void launch_fill(double *dst, double v, size_t n, cudaStream_t st)
{
    if (n == 0) return;
    k_fill<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, v, n);
}

// dst[b*ld + off + i] = values[i]: broadcast of one row into a strided field of every instance
__global__ void k_fill_rows(double *dst, long long ld, int off, const double *values, int B, int n)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)B * n) return;
    const int b = (int)(g / n), i = (int)(g - (long long)b * n);
    dst[(size_t)b * ld + off + i] = values[i];
}
void launch_fill_rows(double *dst, long long ld, int off, const double *values_dev, int B, int n, cudaStream_t st)
{
    const long long items = (long long)B * n;
    if (items == 0) return;
    k_fill_rows<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(dst, ld, off, values_dev, B, n);
}

// x_k = x0 for all k, u = 0: acados' initial guess after construction / the reference's reset(x0)
// (NMPC_class.py:250-254)
__global__ void k_init_iterate(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int B = cfg.B, N = cfg.N, NS = N + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)B * NS * 8) return;
    const int b = (int)(g / (NS * 8));
    const int r = (int)(g - (long long)b * NS * 8);
    m.X[g] = m.x0[(size_t)b * 8 + (r & 7)];
    const int k = r >> 3, i = r & 7;
    if (k < N && i < 2) m.U[((size_t)b * N + k) * 2 + i] = 0.0;
}
void launch_init_iterate(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    const long long items = (long long)cfg.B * (cfg.N + 1) * 8;
    k_init_iterate<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(cfg, m);
}

// [B][5] rows = [cost, time_tot, sqp_iter, qp_iter, status]  (NMPC_class.py:202-206)
__global__ void k_solver_debug(const __grid_constant__ DevCfg cfg, const DevMem m, double time_tot, double *out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= cfg.B) return;
    out[(size_t)b * 5 + 0] = m.cost[b];
    out[(size_t)b * 5 + 1] = time_tot;
    out[(size_t)b * 5 + 2] = 1.0;
    out[(size_t)b * 5 + 3] = (double)m.qp_iter[b];
    out[(size_t)b * 5 + 4] = (double)m.status[b];
}
void launch_solver_debug(const DevCfg &cfg, const DevMem &m, double time_tot, double *out_b5, cudaStream_t st)
{
    k_solver_debug<<<(cfg.B + 127) / 128, 128, 0, st>>>(cfg, m, time_tot, out_b5);
}

// update_cost_function_weights(params[7]) for every stage (NMPC_class.py:269-317):
// W = diag(p0,p0,p1,p2,p3,p4), W_e = diag(p0,p0,p1,p2) -- WITHOUT the 0.01 of the constructor --
// and zl = zu = p5, Zl = Zu = p6 on every soft row.  params: [B][7] (per_instance) or [7].
__global__ void k_update_weights(const __grid_constant__ DevCfg cfg, const DevMem m, const double *params, int per_instance)
{
    const int B = cfg.B, N = cfg.N, NS = N + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = per_instance ? B : 1;
    if (g >= (long long)nb * NS) return;
    const int b = (int)(g / NS), k = (int)(g - (long long)b * NS);
    const double *p = params + (per_instance ? (size_t)b * 7 : 0);
    double *W = per_instance ? m.W_i + ((size_t)b * NS + k) * 6 : m.W_s + k * 6;
    W[0] = p[0]; W[1] = p[0]; W[2] = p[1]; W[3] = p[2];
    W[4] = (k < N) ? p[3] : 0.0; W[5] = (k < N) ? p[4] : 0.0;
    const size_t o = per_instance ? ((size_t)b * NS + k) * 3 : (size_t)k * 3;
    double *zl = per_instance ? m.zl_i : m.zl_s, *zu = per_instance ? m.zu_i : m.zu_s;
    double *Zl = per_instance ? m.Zl_i : m.Zl_s, *Zu = per_instance ? m.Zu_i : m.Zu_s;
    for (int j = 0; j < 3; j++) { zl[o + j] = p[5]; zu[o + j] = p[5]; Zl[o + j] = p[6]; Zu[o + j] = p[6]; }
}
void launch_update_weights(const DevCfg &cfg, const DevMem &m, const double *params_b7, int per_instance, cudaStream_t st)
{
    const long long items = (long long)(per_instance ? cfg.B : 1) * (cfg.N + 1);
    k_update_weights<<<(unsigned)((items + 127) / 128), 128, 0, st>>>(cfg, m, params_b7, per_instance);
}

// K5: R2NMPC back-off computation (Reduced_Robustified_NMPC_class.py:287-365,
// Robust_NMPC_pred_model_utils.py:221-223), fused behind the solve, one thread per instance:
//   Sigma_{s+1} = A_s Sigma_s A_s' + B W_d B'      (A_s: pre-step linearisation of the solve just done)
//   1 <= s < UPH: c_delta = sqrt(Sigma_s[6,6]), c_h = sqrt(grad h(x_s)' Sigma_s grad h(x_s)) with x_s the
//   POST-step iterate; lbx/ubx[s] = -+delta_max +- c_delta, uh[s] = 1 - c_h; stages UPH..N-1 reuse the last c.
__global__ void __launch_bounds__(64) k_r2_tighten(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int B = cfg.B, N = cfg.N, NS = N + 1;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (m.status[b] != 0) return;                  // tightening only after a successful solve (:283)
    double Sg[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) Sg[i][j] = (i == j) ? cfg.r2_sigma0_diag[i] : 0.0;
    double cd = 0.0, ch = 0.0;
    for (int s = 0; s < N; s++) {
        if (s < cfg.uph) {
            if (s >= 1) {
                double x[8], g4[4];
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = m.X[((size_t)b * NS + s) * 8 + i];
                h_eval(cfg, x, g4);
                const int ix[4] = {3, 4, 5, 7};
                double q = 0.0;
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) q += g4[a] * Sg[ix[a]][ix[c]] * g4[c];
                cd = sqrt(Sg[6][6]);
                ch = sqrt(q);
            }
            const double *A = m.stg + ((size_t)b * NS + s) * STG + L_A;
            // T = A Sg  (rows 0..5 structured, rows 6,7 copies), Sg = T A'
            double T[8][8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    double v = (r < 2) ? Sg[r][j] : 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) v += A[c * 6 + r] * Sg[2 + c][j];
                    T[r][j] = v;
                }
                T[6][j] = Sg[6][j]; T[7][j] = Sg[7][j];
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    double v = (r < 2) ? T[i][r] : 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) v += T[i][2 + c] * A[c * 6 + r];
                    Sg[i][r] = v;
                }
                Sg[i][6] = T[i][6]; Sg[i][7] = T[i][7];
            }
#pragma unroll
            for (int i = 0; i < 4; i++) Sg[2 + i][2 + i] += cfg.r2_wd[i];
        }
        if (s >= 1) {
            m.lb_i[((size_t)b * NS + s) * 3 + 1] = -cfg.delta_max + cd;
            m.ub_i[((size_t)b * NS + s) * 3 + 1] = cfg.delta_max - cd;
            m.ub_i[((size_t)b * NS + s) * 3 + 2] = 1.0 - ch;
        }
    }
}
void launch_r2_tighten(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    k_r2_tighten<<<(cfg.B + 63) / 64, 64, 0, st>>>(cfg, m);
}

}  // namespace b200
