// k_misc.cu -- layout conversion and bookkeeping kernels of libb200nmpc.
#include "nmpc_kernels.cuh"

namespace b200 {

#define AT(p, idx) (p)[(size_t)(idx) * B + b]

// Instance-major host order [B][ld] (columns off..off+n) -> SoA [n][B], through a 32x33 shared tile so
// that both the read (along n) and the write (along B) are coalesced.
__global__ void k_pack(const double *__restrict__ src, int ld, int off, double *__restrict__ dst, int B, int n)
{
    __shared__ double tile[32][33];
    const int b0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int b = b0 + r, i = i0 + threadIdx.x;
        if (b < B && i < n) tile[r][threadIdx.x] = src[(size_t)b * ld + off + i];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, b = b0 + threadIdx.x;
        if (b < B && i < n) dst[(size_t)i * B + b] = tile[threadIdx.x][r];
    }
}

__global__ void k_unpack(const double *__restrict__ src, double *__restrict__ dst, int ld, int off, int B, int n)
{
    __shared__ double tile[32][33];
    const int b0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, b = b0 + threadIdx.x;
        if (b < B && i < n) tile[r][threadIdx.x] = src[(size_t)i * B + b];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int b = b0 + r, i = i0 + threadIdx.x;
        if (b < B && i < n) dst[(size_t)b * ld + off + i] = tile[threadIdx.x][r];
    }
}

void launch_pack_strided(const double *src, int ld, int off, double *dst_nb, int B, int n, cudaStream_t st)
{
    dim3 block(32, 8), grid((B + 31) / 32, (n + 31) / 32);
    k_pack<<<grid, block, 0, st>>>(src, ld, off, dst_nb, B, n);
}
void launch_unpack_strided(const double *src_nb, double *dst, int ld, int off, int B, int n, cudaStream_t st)
{
    dim3 block(32, 8), grid((B + 31) / 32, (n + 31) / 32);
    k_unpack<<<grid, block, 0, st>>>(src_nb, dst, ld, off, B, n);
}
void launch_pack(const double *src_bn, double *dst_nb, int B, int n, cudaStream_t st)
{
    launch_pack_strided(src_bn, n, 0, dst_nb, B, n, st);
}
void launch_unpack(const double *src_nb, double *dst_bn, int B, int n, cudaStream_t st)
{
    launch_unpack_strided(src_nb, dst_bn, n, 0, B, n, st);
}

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

// x_k = x0 for all k, u = 0: acados' initial guess after construction / the reference's reset(x0)
// (NMPC_class.py:250-254)
__global__ void k_init_iterate(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int B = cfg.B, N = cfg.N;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (b >= B) return;
#pragma unroll
    for (int i = 0; i < 8; i++) AT(m.X, k * 8 + i) = AT(m.x0, i);
    if (k < N) { AT(m.U, k * 2) = 0.0; AT(m.U, k * 2 + 1) = 0.0; }
}
void launch_init_iterate(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    dim3 block(128), grid((cfg.B + 127) / 128, cfg.N + 1);
    k_init_iterate<<<grid, block, 0, st>>>(cfg, m);
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
// and zl = zu = p5, Zl = Zu = p6 on every soft row.  params is SoA [7][B] (per_instance) or [7].
__global__ void k_update_weights(const __grid_constant__ DevCfg cfg, const DevMem m, const double *params, int per_instance)
{
    const int B = cfg.B, N = cfg.N;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (per_instance) {
        if (b >= B) return;
        double p[7];
#pragma unroll
        for (int i = 0; i < 7; i++) p[i] = params[(size_t)i * B + b];
        AT(m.W_i, k * 6 + 0) = p[0]; AT(m.W_i, k * 6 + 1) = p[0]; AT(m.W_i, k * 6 + 2) = p[1]; AT(m.W_i, k * 6 + 3) = p[2];
        AT(m.W_i, k * 6 + 4) = (k < N) ? p[3] : 0.0; AT(m.W_i, k * 6 + 5) = (k < N) ? p[4] : 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            AT(m.zl_i, k * 3 + j) = p[5]; AT(m.zu_i, k * 3 + j) = p[5];
            AT(m.Zl_i, k * 3 + j) = p[6]; AT(m.Zu_i, k * 3 + j) = p[6];
        }
    } else {
        if (b != 0) return;
        const double *p = params;
        m.W_s[k * 6 + 0] = p[0]; m.W_s[k * 6 + 1] = p[0]; m.W_s[k * 6 + 2] = p[1]; m.W_s[k * 6 + 3] = p[2];
        m.W_s[k * 6 + 4] = (k < N) ? p[3] : 0.0; m.W_s[k * 6 + 5] = (k < N) ? p[4] : 0.0;
        for (int j = 0; j < 3; j++) {
            m.zl_s[k * 3 + j] = p[5]; m.zu_s[k * 3 + j] = p[5];
            m.Zl_s[k * 3 + j] = p[6]; m.Zu_s[k * 3 + j] = p[6];
        }
    }
}
void launch_update_weights(const DevCfg &cfg, const DevMem &m, const double *params_nb, int per_instance, cudaStream_t st)
{
    dim3 block(128), grid(per_instance ? (cfg.B + 127) / 128 : 1, cfg.N + 1);
    k_update_weights<<<grid, block, 0, st>>>(cfg, m, params_nb, per_instance);
}

// K5: R2NMPC back-off computation (Reduced_Robustified_NMPC_class.py:287-365,
// Robust_NMPC_pred_model_utils.py:221-223), fused behind the solve, one thread per instance:
//   Sigma_{s+1} = A_s Sigma_s A_s' + B W_d B'      (A_s: pre-step linearisation of the solve just done)
//   1 <= s < UPH: c_delta = sqrt(Sigma_s[6,6]), c_h = sqrt(grad h(x_s)' Sigma_s grad h(x_s)) with x_s the
//   POST-step iterate; lbx/ubx[s] = -+delta_max +- c_delta, uh[s] = 1 - c_h; stages UPH..N-1 reuse the last c.
__global__ void __launch_bounds__(32) k_r2_tighten(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int B = cfg.B, N = cfg.N;
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
                for (int i = 0; i < 8; i++) x[i] = AT(m.X, s * 8 + i);
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
            double A[36];
#pragma unroll
            for (int i = 0; i < 36; i++) A[i] = AT(m.A36, s * 36 + i);
            // T = A Sg  (rows 0..5 structured, rows 6,7 copies), Sg = T A'
            double T[8][8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    double v = (r < 2) ? Sg[r][j] : 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) v += A[r * 6 + c] * Sg[2 + c][j];
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
                    for (int c = 0; c < 6; c++) v += T[i][2 + c] * A[r * 6 + c];
                    Sg[i][r] = v;
                }
                Sg[i][6] = T[i][6]; Sg[i][7] = T[i][7];
            }
#pragma unroll
            for (int i = 0; i < 4; i++) Sg[2 + i][2 + i] += cfg.r2_wd[i];
        }
        if (s >= 1) {
            AT(m.lb_i, s * 3 + 1) = -cfg.delta_max + cd;
            AT(m.ub_i, s * 3 + 1) = cfg.delta_max - cd;
            AT(m.ub_i, s * 3 + 2) = 1.0 - ch;
        }
    }
}
void launch_r2_tighten(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    k_r2_tighten<<<(cfg.B + 31) / 32, 32, 0, st>>>(cfg, m);
}

}  // namespace b200
