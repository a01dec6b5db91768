// nmpc_kernels.cuh -- device memory map and kernel declarations of libb200nmpc.
#pragma once
#include "nmpc_model.cuh"

namespace b200 {

// All arrays are float64 "batch-contiguous SoA": element (idx, b) lives at idx * B + b, so that the 32
// lanes of a warp (32 consecutive instances) read one 256-byte line per scalar field.
struct DevMem {
    // iterate and inputs
    double *X;      // [(N+1)*8][B]
    double *U;      // [N*2][B]
    double *x0;     // [8][B]
    double *yref;   // [(N+1)*6][B]
    // stage data: shared [(N+1)*n] and optional per-instance overrides [(N+1)*n][B] (nullptr = shared)
    double *W_s, *lb_s, *ub_s, *zl_s, *zu_s, *Zl_s, *Zu_s;
    double *W_i, *lb_i, *ub_i, *zl_i, *zu_i, *Zl_i, *Zu_i;
    // linearisation (K1 output): the QP data of SURVEY A.7
    double *A36;    // [N*36][B]  rows 0..5 x cols 2..7 of A_k
    double *B12;    // [N*12][B]  rows 0..5 of B_k
    double *bd;     // [N*8][B]   defect b_k = Phi(x_k,u_k) - x_{k+1}
    double *qx;     // [(N+1)*5][B] gradient w.r.t. (px,py,yaw,vl,vt)
    double *rg;     // [N*2][B]   gradient w.r.t. u
    double *Qv;     // [(N+1)*3][B] (vl,vt) Hessian block [Q33,Q34,Q44] (SNMPC cost only)
    double *hval;   // [(N+1)][B]
    double *gh;     // [(N+1)*4][B] grad h w.r.t. (vl,vt,r,a)
    // QP workspace
    double *K;      // [N*16][B]
    double *Hi;     // [N*3][B]
    double *kff;    // [N*2][B]
    double *dx;     // [(N+1)*8][B]
    double *du;     // [N*2][B]
    double *ddu;    // [N*2][B]
    double *gus;    // [N*2][B]
    double *ips, *iplam, *ipt, *ipmu;        // [(N+1)*6][B]   index (k*3+c)*2+side
    double *ids, *idlam, *idt, *idmu;        // directions
    double *pr1, *pr2;                       // Mehrotra second-order products
    // results
    double *sl, *su;   // [(N+1)*3][B]
    double *cost;      // [B]
    double *kkt;       // [B]
    int *qp_iter;      // [B]
    int *status;       // [B]
};

void launch_linearize(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
void launch_qp_solve(const DevCfg &cfg, const DevMem &m, cudaStream_t st);          // one thread per instance (v1)
void launch_qp_solve_octet(const DevCfg &cfg, const DevMem &m, cudaStream_t st);    // eight lanes per instance (v2)
void launch_init_iterate(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
void launch_r2_tighten(const DevCfg &cfg, const DevMem &m, cudaStream_t st);
// instance-major [B][n] (host order)  <->  SoA [n][B]
void launch_pack(const double *src_bn, double *dst_nb, int B, int n, cudaStream_t st);
void launch_unpack(const double *src_nb, double *dst_bn, int B, int n, cudaStream_t st);
// strided variants: rows of width n taken from / written to rows of width ld at column offset off
void launch_pack_strided(const double *src, int ld, int off, double *dst_nb, int B, int n, cudaStream_t st);
void launch_unpack_strided(const double *src_nb, double *dst, int ld, int off, int B, int n, cudaStream_t st);
void launch_fill(double *dst, double v, size_t n, cudaStream_t st);
void launch_solver_debug(const DevCfg &cfg, const DevMem &m, double time_tot, double *out_b5, cudaStream_t st);
void launch_update_weights(const DevCfg &cfg, const DevMem &m, const double *params_nb, int per_instance, cudaStream_t st);

}  // namespace b200
