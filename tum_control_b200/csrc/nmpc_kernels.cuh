// nmpc_kernels.cuh -- device memory map and kernel declarations of libb200nmpc.
//
// HBM layout (all float64).  Everything is INSTANCE-MAJOR: the arrays a host caller owns
// (x0 [B][8], yref [B][N+1][6], X [B][N+1][8], U [B][N][2]) are stored exactly in their C order, so the
// bulk set/get entry points are single cudaMemcpyAsync calls with no layout kernel, and everything one
// instance needs for one stage is one contiguous record:
//   LIN  part (K1 -> K2), LREC doubles of the stage record of an (instance, stage): q_x | r_u | Qv | h | b | grad h (state order) | grad h | A36 | B12
//   WORK block  (K2 private), per instance: QP iterate, Riccati gains, interior-point state per soft side.
// K2 runs one warp per instance; the 8 lanes of an octet read 8 consecutive doubles (one 64-byte
// segment) for every 8-vector, and the stage-parallel phases read the per-side arrays with 32
// consecutive lanes (256-byte segments).
#pragma once
#include "nmpc_model.cuh"

namespace b200 {

// ---- stage record: STG doubles (992 bytes) per (instance, stage); all offsets are relative to the record start.
// The fields are ordered by the sweeps that read them, so that every sweep of the QP kernel stages only a contiguous
// sub-range of the record in shared memory (62 -> 38..50 of the 16-byte chunks per stage visit; the copies were a quarter of
// all shared-memory wavefronts of the kernel, whose busiest unit is the load/store pipe):
//   [ dx ddx du ddu | q_x r_u Qv h . | b | grad h (state order) | grad h | A B | K | Hinv gus | kff . ]
//     0               20               32  40                     48       52     100 116        121
//   backward sweeps (factorisation, residual check)  [0, 100)      corrector sweep  [40, 122)
//   forward roll-out of a direction                  [48, 124)     phase-0 roll-out [30, 124)
// LIN part (K1 -> K2): the LREC doubles from LIN_OFF; WORK part (K2 private): the 20 doubles before and the 24 after it.
constexpr int W_DX = 0;       // 8 : QP iterate dx_k
constexpr int W_DDX = 8;      // 8 : Newton direction ddx_k of the current iteration
constexpr int W_DU = 16;      // 2 : QP iterate du_k
constexpr int W_DDU = 18;     // 2 : Newton direction ddu_k (read by the lazy primal update of the next backward sweep; long-horizon
                              //     layout only, short horizons keep it in shared memory)
constexpr int LIN_OFF = 20;
constexpr int L_QX = 20;      // 5 : cost gradient w.r.t. (px,py,yaw,vl,vt)
constexpr int L_RG = 25;      // 2 : cost gradient w.r.t. u
constexpr int L_QV = 27;      // 3 : (vl,vt) Gauss-Newton block [Q33,Q34,Q44] (SNMPC cost only)
constexpr int L_H = 30;       // 1 : h(x_k)                                                          (31: pad)
constexpr int L_b = 32;       // 8 : defect b_k = Phi(x_k,u_k) - x_{k+1}
constexpr int L_GF = 40;      // 8 : grad h scattered to state order [0,0,0,g_vl,g_vt,g_r,0,g_a] (lane j reads entry j)
constexpr int L_GH = 48;      // 4 : grad h w.r.t. (vl,vt,r,a)
constexpr int L_A = 52;       // 36: rows 0..5 x cols 2..7 of A_k, stored TRANSPOSED: rec[L_A + 6 c + r] = A[r][2 + c] (every
                              //     product of the Riccati sweeps walks a column of A: contiguous, 128-bit loads, no stride)
constexpr int L_B = 88;       // 12: rows 0..5 of B_k, transposed: rec[L_B + 6 c + r] = B[r][c]
constexpr int LREC = 80;      //     LIN part = [LIN_OFF, LIN_OFF + LREC)
constexpr int W_K = 100;      // 16: feedback gain K_k (2x8 row-major)
constexpr int W_HI = 116;     // 3 : (R + B'PB)^-1 (symmetric 2x2)
constexpr int W_GUS = 119;    // 2 : u-stationarity residual of the backward sweep (read by the corrector sweep; long-horizon layout)
constexpr int W_KFF = 121;    // 2 : feed-forward input kff_k of the factorisation / corrector (read by the roll-outs; long-horizon
                              //     layout)                                                          (123: pad)
constexpr int STG = 124;
static_assert(L_B == L_A + 36 && LIN_OFF + LREC == W_K && LIN_OFF % 2 == 0 && L_A % 2 == 0 && L_GF % 2 == 0 && STG % 2 == 0, "record layout");
// sub-ranges [lo, hi) of the record (doubles, lo even) the sweeps of the QP kernel stage in shared memory
constexpr int R_BACK_LO = 0, R_BACK_HI = 100, R_CORR_LO = 40, R_CORR_HI = 122, R_FWD_LO = 48, R_FWD_HI = 124, R_ROLL_LO = 30, R_ROLL_HI = 124;
// per-side arrays, index = k*6 + row*2 + side  (row: 0 bu, 1 bx, 2 h; side: 0 lower, 1 upper)
// S_PR1/S_PR2: second-order products of the predictor, then (with S_D3/S_D4) the final directions ds, dt, dlam, dmu.
// (Caching the per-iterate reciprocals in three more arrays was measured SLOWER: the side arrays miss the small L1 left
// beside 199 KB of shared memory, and an L2 round trip costs more than the FP64 divisions it saves.)
// Quantities that are always read and written together are interleaved as double2 pairs (one 128-bit access instead of
// two 64-bit ones per item): P_ST = (s, t), P_LM = (lam, mu), P_PR = (PR1, PR2), P_D = (D3, D4), P_Z = (z, Z); pair p
// occupies the doubles [2 p n, 2 (p + 1) n) of the instance's side block (n = 6 (N + 1) items), the two single arrays
// S_BND and S_ACT follow at 10 n and 11 n.
enum { P_ST = 0, P_LM, P_PR, P_D, P_Z, P_COUNT };
enum { S_BND = 2 * P_COUNT, S_ACT, S_CSCP, S_COUNT };      // S_CSCP: per (stage, row) cs [3 (N+1)] | cp [3 (N+1)] (see k_qp_warp.cu)

// ---- SNMPC sample record (K4a -> K4b / step), SREC doubles per (instance, stage < UPH, sample):
//      A36 | B12 | xn8 (RK4 image of the sample iterate) | h | grad h (vl, vt, r, a)
constexpr int SR_A = 0, SR_B = 36, SR_XN = 48, SR_H = 56, SR_GH = 57, SREC = 64;

// ---- SNMPC condensed head (K4b -> K2), per instance, H = UPH stages, nh = 2H head inputs (du_0 .. du_{H-1}):
//      the sample states of stages <= H are affine in the head inputs only (all sample initial states are fixed,
//      SNMPC_class.py:259-264), so stages 0..H-1 are eliminated onto nh variables and the Riccati recursion
//      starts at stage H with dx_H = G_H u~ + c_H.
//   HD_G   [H][8][nh]  G_k, k = 1..H (index k-1): nominal (PCE-mean) state maps
//   HD_C   [H][8]      c_k, k = 1..H
//   HD_ROW [H][nh]     chance-constraint rows of stages 0..H-1 (row 0 unused)
//   HD_V0  [H]         their values at u~ = 0
//   HD_HH  [nh][nh]    head Hessian sum_k G_k' Q_k G_k + R;   HD_GH [nh]  head gradient
struct HeadLayout {
    int H, nh, G, C, ROW, V0, HH, GH, stride;
    __host__ __device__ explicit HeadLayout(int H_) : H(H_), nh(2 * H_)
    {
        G = 0; C = G + H * 8 * nh; ROW = C + H * 8; V0 = ROW + H * nh; HH = V0 + H; GH = HH + nh * nh;
        stride = (GH + nh + 1) & ~1;
    }
};

struct DevMem {
    // caller-visible state, instance-major
    double *X;      // [B][N+1][8]
    double *U;      // [B][N][2]
    double *x0;     // [B][8]
    double *yref;   // [B][N+1][6]
    // stage data: shared [(N+1)][n] and optional per-instance overrides [B][N+1][n] (nullptr = shared)
    double *W_s, *lb_s, *ub_s, *zl_s, *zu_s, *Zl_s, *Zu_s;
    double *W_i, *lb_i, *ub_i, *zl_i, *zu_i, *Zl_i, *Zu_i;
    double *stg;    // [B][N+1][STG]: the stage records (WORK and LIN parts interleaved by reader, see above)
    double *work;   // [B][work_stride]: the per-side arrays
    // results
    double *sl, *su;   // [B][N+1][3]
    double *cost;      // [B]
    double *kkt;       // [B]
    int *qp_iter;      // [B]
    int *status;       // [B]
    long long work_stride;
    // SNMPC
    double *XS;        // [B][HA+1][ns][8] sample iterates (HA = allocated propagation horizon)
    double *x0s;       // [B][ns][8] sample initial states (compute_x0dist)
    double *slin;      // [B][HA][ns][SREC]
    double *head;      // [B][head_stride]
    double *sn_A;      // [L][ns] PCE matrix
    double *sn_woff;   // [ns][8] x0 sample offsets
    int HA;            // allocated head stages (cfg.head_H <= HA)
    int head_stride;
    // closed loop with stop conditions: instances with halt[b] >= 0 are skipped by the per-step kernels (or null)
    const int *halt;
    // closed loop: launch slot -> instance, heaviest expected instance first (or null: identity)
    const int *order;
};

__host__ __device__ inline long long work_stride_for(int N)
{
    return (long long)S_COUNT * (N + 1) * 6;
}
// dynamic shared memory of k_qp_warp, bytes per warp
__host__ __device__ inline int qp_smem_per_warp(int N, bool longh);

void launch_linearize(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
int launch_qp_solve_warp(const DevCfg &cfg, const DevMem &m, cudaStream_t st);      // returns cudaError_t as int
// SNMPC: x0 sample matrix / initial guess (per-sample linearisation: work items of launch_linearize; head condensing and
// sample iterate update: prologue / epilogue of the QP kernel, snmpc_head.cuh)
void launch_snmpc_x0_samples(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
void launch_snmpc_init_samples(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
void launch_init_iterate(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
void launch_r2_tighten(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
void launch_fill(double *dst, double v, size_t n, cudaStream_t st);
// dst[b*ld + off + i] = values[i] for all b < B, i < n (values: device pointer)
void launch_fill_rows(double *dst, long long ld, int off, const double *values_dev, int B, int n, cudaStream_t st);
void launch_solver_debug(const DevCfg &cfg, const DevMem &m, double time_tot, double *out_b5, cudaStream_t st);
// launch slots of the next QP kernel: instances in descending order of their last interior-point iteration count
void launch_launch_order(const DevMem &m, int *order, int B, cudaStream_t st);
void launch_update_weights(const DevCfg &cfg, const DevMem &m, const double *params_b7, int per_instance, cudaStream_t st);

}  // namespace b200
