// b200nmpc.cu -- C-ABI of libb200nmpc (see include/b200nmpc.h): handle, device memory, stream-ordered
// set/get plumbing.  No PyTorch, no CPU fallback: every entry point fails with B200NMPC_E_NODEVICE /
// B200NMPC_E_CUDA when no sm_100 device is usable.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nmpc_kernels.cuh"

using namespace b200;

static thread_local char g_err[512] = "";
static int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return set_err(B200NMPC_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                           __FILE__, __LINE__);                                                    \
    } while (0)

// stage-data groups that can be shared (one value set for the whole batch) or per-instance
enum { G_W = 0, G_LB, G_UB, G_ZL, G_ZU, G_QL, G_QU, G_COUNT };
static const int G_NCOMP[G_COUNT] = {6, 3, 3, 3, 3, 3, 3};

struct b200nmpc_handle {
    DevCfg cfg;
    DevMem m;
    b200nmpc_config user_cfg;
    int B, N, device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1, ev_mid, ev_t0, ev_t1;
    bool timed;
    std::vector<void *> allocs;
    std::vector<double> host_shared[G_COUNT];      // host mirrors of the shared stage data
    double *staging;                               // device, instance-major scratch for pack/unpack
    size_t staging_elems;
    long long launches;
    // SNMPC host mirrors of the stage parameter p = [A, gamma, stop_flag] (SNMPC_class.py:120-125)
    std::vector<double> sn_A;                      // [L][ns]
    std::vector<int> sn_stop;                      // [N+1]
    double sn_gamma;
    bool sn_dirty;
    // launch order of the QP kernel: instances that iterated longest in the previous solve start first (the batch is a
    // few waves of warps at most, so the instances that start last should be the short ones); owned by the closed-loop
    // simulation while one runs (order_external)
    int *order;
    bool order_valid, order_hint, order_external;
    // pinned host ring: the per-stage set / constraints_set / cost_set calls copy the caller's values here and return
    // without waiting for the device (the reference's loop makes 2N + 4 such calls per solve, NMPC_class.py:169-180)
    double *ring;
    size_t ring_elems, ring_pos;
    // host mirror of the iterate for the per-stage get(j, "x") / get(j, "u") calls of small batches (the reference reads N
    // stages one by one, NMPC_class.py:193-201): the first get after a solve fetches X and U in one copy
    std::vector<double> host_xu;
    bool xu_cached;
};
static const size_t XU_CACHE_MAX_DOUBLES = 1u << 17;      // 1 MiB: beyond that the per-stage strided copies are cheaper

static double **shared_ptr(b200nmpc_handle *h, int g)
{
    DevMem &m = h->m;
    switch (g) {
        case G_W: return &m.W_s; case G_LB: return &m.lb_s; case G_UB: return &m.ub_s;
        case G_ZL: return &m.zl_s; case G_ZU: return &m.zu_s; case G_QL: return &m.Zl_s; default: return &m.Zu_s;
    }
}
static double **inst_ptr(b200nmpc_handle *h, int g)
{
    DevMem &m = h->m;
    switch (g) {
        case G_W: return &m.W_i; case G_LB: return &m.lb_i; case G_UB: return &m.ub_i;
        case G_ZL: return &m.zl_i; case G_ZU: return &m.zu_i; case G_QL: return &m.Zl_i; default: return &m.Zu_i;
    }
}

template <typename T>
static int dev_alloc(b200nmpc_handle *h, T **p, size_t n)
{
    CK(cudaMalloc((void **)p, n * sizeof(T)));
    CK(cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
    h->allocs.push_back(*p);
    return 0;
}

static int use_device(const b200nmpc_handle *h)
{
    CK(cudaSetDevice(h->device));
    return 0;
}

static int upload_shared(b200nmpc_handle *h, int g)
{
    CK(cudaMemcpyAsync(*shared_ptr(h, g), h->host_shared[g].data(), h->host_shared[g].size() * sizeof(double),
                       cudaMemcpyHostToDevice, h->stream));
    return 0;
}

// make group g per-instance: allocate [B][N+1][ncomp] and broadcast the shared values into it
static int ensure_inst(b200nmpc_handle *h, int g)
{
    double **ip = inst_ptr(h, g);
    if (*ip) return 0;
    const int n = (h->N + 1) * G_NCOMP[g];
    double *p = nullptr;
    int rc = dev_alloc(h, &p, (size_t)n * h->B);
    if (rc) return rc;
    launch_fill_rows(p, n, 0, *shared_ptr(h, g), h->B, n, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    *ip = p;
    return 0;
}

static void fold_config(const b200nmpc_config *c, int B, DevCfg *d)
{
    memset(d, 0, sizeof *d);
    d->lf = c->lf; d->lr = c->lr; d->m = c->m; d->inv_m = 1.0 / c->m; d->inv_Iz = 1.0 / c->Iz;
    d->Bf = c->Bf; d->Cf = c->Cf; d->Df = c->Df; d->Ef = c->Ef;
    d->Br = c->Br; d->Cr = c->Cr; d->Dr = c->Dr; d->Er = c->Er;
    d->fr0 = c->fr0; d->fr1 = c->fr1; d->fr4 = c->fr4;
    d->Fz_f = c->m * c->lr * c->g / (c->lf + c->lr);                 // pred_model_dynamic_stm_pacejka.py:120
    d->Fz_r = c->m * c->lf * c->g / (c->lf + c->lr);                 // :121
    d->inv_Fmax_f = 1.0 / std::sqrt(d->Fz_f * d->Fz_f + (c->Cf * d->Fz_f) * (c->Cf * d->Fz_f));   // :156
    d->inv_Fmax_r = 1.0 / std::sqrt(d->Fz_r * d->Fz_r + (c->Cr * d->Fz_r) * (c->Cr * d->Fz_r));   // :157
    d->kaero = 0.5 * c->ro * c->S * c->Cd;                           // :134
    d->ax_brake = c->ax_brake;
    for (int i = 0; i < B200NMPC_MAX_GGV; i++) { d->ggv_v[i] = c->ggv_v[i]; d->ggv_ax[i] = c->ggv_ax[i]; d->ggv_ay[i] = c->ggv_ay[i]; }
    d->dt = c->dt; d->ipm_tol = c->ipm_tol; d->ipm_reg = c->ipm_reg;
    d->ipm_eta_max = (c->ipm_eta_max >= 0.5 && c->ipm_eta_max <= 1.0) ? c->ipm_eta_max : 0.99;
    d->ipm_corr_mode = c->ipm_corr_mode;
    d->ipm_s0 = (c->ipm_s0 > 0.0) ? c->ipm_s0 : 0.1;
    for (int i = 0; i < 8; i++) d->r2_sigma0_diag[i] = c->r2_sigma0_diag[i];
    for (int i = 0; i < 4; i++) d->r2_wd[i] = c->r2_wd[i];
    d->delta_max = c->ubx;
    d->n_ggv = c->n_ggv; d->N = c->N; d->rk_steps = c->rk_steps; d->variant = c->variant;
    d->ipm_max_iter = c->ipm_max_iter; d->qp_no_shortcut = c->qp_no_shortcut; d->uph = c->uph;
    d->n_samples = c->n_samples;
    d->cost_vel_abs = (c->variant == B200NMPC_SNMPC);
    if (c->variant == B200NMPC_SNMPC) {
        d->rk_steps = 1;                          // integrator_type DISCRETE: one RK4 step of dt (pred_model_dynamic_disc.py:187-192)
        d->head_H = c->uph;
        d->sn_L = c->sn_L;
        d->sn_kappa = std::sqrt((1.0 - c->sn_gamma) / c->sn_gamma);
    }
    d->B = B;
}

// SNMPC: push the host mirrors of p (PCE matrix, gamma, stop flags) to the device configuration
static int snmpc_sync_params(b200nmpc_handle *h)
{
    if (h->cfg.variant != B200NMPC_SNMPC || !h->sn_dirty) return 0;
    const int N = h->N;
    int H = 0;
    while (H <= N && h->sn_stop[H] == 0) H++;
    for (int k = H; k <= N; k++)
        if (h->sn_stop[k] == 0)
            return set_err(B200NMPC_E_CONFIG, "SNMPC stop flags must be 0 on a prefix of stages and 1 afterwards (stage %d)", k);
    if (H > h->m.HA || H >= N)
        return set_err(B200NMPC_E_CONFIG, "uncertainty propagation horizon %d exceeds the handle's capacity %d (uph at creation, "
                       "<= %d and < N)", H, h->m.HA, B200NMPC_MAX_UPH);
    h->cfg.head_H = H;
    h->cfg.sn_kappa = std::sqrt((1.0 - h->sn_gamma) / h->sn_gamma);
    CK(cudaMemcpyAsync(h->m.sn_A, h->sn_A.data(), h->sn_A.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->sn_dirty = false;
    return 0;
}

extern "C" {

const char *b200nmpc_last_error(void) { return g_err; }
int b200nmpc_version(void) { return 100; }
int b200nmpc_sizeof_config(void) { return (int)sizeof(b200nmpc_config); }
int b200nmpc_batch(const b200nmpc_handle *h) { return h ? h->B : B200NMPC_E_ARG; }
int b200nmpc_horizon(const b200nmpc_handle *h) { return h ? h->N : B200NMPC_E_ARG; }
void *b200nmpc_stream(const b200nmpc_handle *h) { return h ? (void *)h->stream : nullptr; }
long long b200nmpc_launch_count(const b200nmpc_handle *h) { return h ? h->launches : 0; }

int b200nmpc_create(const b200nmpc_config *cfg, int batch, int device, b200nmpc_handle **out)
{
    if (!cfg || !out) return set_err(B200NMPC_E_ARG, "null argument");
    *out = nullptr;
    if (batch < 1) return set_err(B200NMPC_E_ARG, "batch must be >= 1");
    if (cfg->N < 1 || cfg->N > B200NMPC_MAX_N) return set_err(B200NMPC_E_CONFIG, "N out of range [1,%d]", B200NMPC_MAX_N);
    if (cfg->n_ggv < 2 || cfg->n_ggv > B200NMPC_MAX_GGV) return set_err(B200NMPC_E_CONFIG, "n_ggv out of range");
    if (!(cfg->dt > 0.0) || cfg->rk_steps < 1) return set_err(B200NMPC_E_CONFIG, "dt / rk_steps invalid");
    if (cfg->variant != B200NMPC_NOMINAL && cfg->variant != B200NMPC_R2 && cfg->variant != B200NMPC_SNMPC)
        return set_err(B200NMPC_E_CONFIG, "unknown variant %d", cfg->variant);
    if (cfg->variant == B200NMPC_SNMPC) {
        if (cfg->sn_L < 1 || cfg->sn_L > B200NMPC_PCE_L) return set_err(B200NMPC_E_CONFIG, "sn_L out of range [1,%d]", B200NMPC_PCE_L);
        if (cfg->n_samples < cfg->sn_L || cfg->n_samples > B200NMPC_MAX_NS)
            return set_err(B200NMPC_E_CONFIG, "n_samples must be in [sn_L,%d]", B200NMPC_MAX_NS);
        if (cfg->uph < 0 || cfg->uph > B200NMPC_MAX_UPH || cfg->uph >= cfg->N)
            return set_err(B200NMPC_E_CONFIG, "uph must be in [0,%d] and < N (the condensed head handles the propagation horizon)",
                           B200NMPC_MAX_UPH);
        if (!(cfg->sn_gamma > 0.0 && cfg->sn_gamma <= 1.0)) return set_err(B200NMPC_E_CONFIG, "sn_gamma must be in (0,1]");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(B200NMPC_E_NODEVICE, "no CUDA device: libb200nmpc has no CPU fallback");
    if (device < 0 || device >= ndev) return set_err(B200NMPC_E_ARG, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_err(B200NMPC_E_NODEVICE, "device %d is sm_%d%d; libb200nmpc is built for sm_100a only", device,
                       prop.major, prop.minor);

    b200nmpc_handle *h = new b200nmpc_handle();
    h->B = batch; h->N = cfg->N; h->device = device; h->launches = 0; h->timed = false;
    h->user_cfg = *cfg;
    fold_config(cfg, batch, &h->cfg);
    memset(&h->m, 0, sizeof h->m);
    // from here on a failure releases the partly built handle (b200nmpc_destroy tolerates null stream / events)
#define CKH(call)                                                                                            \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            b200nmpc_destroy(h);                                                                             \
            return set_err(B200NMPC_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                    \
    } while (0)
    CKH(cudaSetDevice(device));
    CKH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKH(cudaEventCreate(&h->ev0));
    CKH(cudaEventCreate(&h->ev1));
    CKH(cudaEventCreate(&h->ev_mid));
    CKH(cudaEventCreate(&h->ev_t0));
    CKH(cudaEventCreate(&h->ev_t1));
    const size_t B = batch, N = cfg->N, N1 = N + 1;
    DevMem &m = h->m;
    int rc = 0;
#define DA(ptr, n) if (!rc) rc = dev_alloc(h, &(ptr), (size_t)(n))
    DA(m.X, N1 * 8 * B); DA(m.U, N * 2 * B); DA(m.x0, 8 * B); DA(m.yref, N1 * 6 * B);
    DA(m.W_s, N1 * 6); DA(m.lb_s, N1 * 3); DA(m.ub_s, N1 * 3);
    DA(m.zl_s, N1 * 3); DA(m.zu_s, N1 * 3); DA(m.Zl_s, N1 * 3); DA(m.Zu_s, N1 * 3);
    m.work_stride = work_stride_for((int)N);
    DA(m.stg, N1 * STG * B); DA(m.work, (size_t)m.work_stride * B);
    DA(m.sl, N1 * 3 * B); DA(m.su, N1 * 3 * B); DA(m.cost, B); DA(m.kkt, B);
    DA(m.qp_iter, B); DA(m.status, B); DA(h->order, B);
    h->order_valid = false; h->order_external = false;
    h->ring = nullptr; h->ring_pos = 0; h->xu_cached = false;
    h->ring_elems = 32 * B > 8192 ? 32 * B : 8192;
    if (!rc && cudaHostAlloc((void **)&h->ring, h->ring_elems * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
        h->ring = nullptr;
        rc = set_err(B200NMPC_E_CUDA, "cudaHostAlloc of the host staging ring failed");
    }
    h->order_hint = getenv("B200NMPC_NO_LPT") == nullptr;
    h->staging_elems = 64 * B + 1024;
    DA(h->staging, h->staging_elems);
    if (cfg->variant == B200NMPC_SNMPC) {
        const size_t ns = cfg->n_samples, HA = cfg->uph > 0 ? cfg->uph : 1;
        m.HA = (int)HA;
        m.head_stride = HeadLayout((int)HA).stride;
        DA(m.XS, B * (HA + 1) * ns * 8); DA(m.x0s, B * ns * 8); DA(m.slin, B * HA * ns * SREC);
        DA(m.head, B * (size_t)m.head_stride); DA(m.sn_A, (size_t)cfg->sn_L * ns); DA(m.sn_woff, ns * 8);
    }
#undef DA
    if (rc) { b200nmpc_destroy(h); return rc; }
    // default stage data (NMPC_STM_acados_settings.py:55-56,107-139,164-224)
    for (int g = 0; g < G_COUNT; g++) h->host_shared[g].assign(N1 * G_NCOMP[g], 0.0);
    for (size_t k = 0; k <= N; k++) {
        for (int i = 0; i < 6; i++) h->host_shared[G_W][k * 6 + i] = (k < N) ? cfg->W[i] : (i < 4 ? cfg->W_e[i] : 0.0);
        const double lb[3] = {cfg->lbu, cfg->lbx, cfg->lh}, ub[3] = {cfg->ubu, cfg->ubx, cfg->uh};
        for (int j = 0; j < 3; j++) {
            h->host_shared[G_LB][k * 3 + j] = lb[j]; h->host_shared[G_UB][k * 3 + j] = ub[j];
            h->host_shared[G_ZL][k * 3 + j] = cfg->zl; h->host_shared[G_ZU][k * 3 + j] = cfg->zu;
            h->host_shared[G_QL][k * 3 + j] = cfg->Zl; h->host_shared[G_QU][k * 3 + j] = cfg->Zu;
        }
    }
    for (int g = 0; g < G_COUNT; g++)
        if ((rc = upload_shared(h, g))) { b200nmpc_destroy(h); return rc; }
    if (cfg->variant == B200NMPC_R2) {           // tightened bounds are per instance
        if ((rc = ensure_inst(h, G_LB)) || (rc = ensure_inst(h, G_UB))) { b200nmpc_destroy(h); return rc; }
    }
    h->sn_dirty = false;
    if (cfg->variant == B200NMPC_SNMPC) {
        const int ns = cfg->n_samples;
        h->sn_A.assign((size_t)cfg->sn_L * ns, 0.0);
        for (int j = 0; j < cfg->sn_L; j++)
            for (int i = 0; i < ns; i++) h->sn_A[(size_t)j * ns + i] = cfg->sn_A[j][i];
        h->sn_stop.assign(N1, 1);
        for (int k = 0; k < cfg->uph; k++) h->sn_stop[k] = 0;
        h->sn_gamma = cfg->sn_gamma;
        h->sn_dirty = true;
        std::vector<double> woff((size_t)ns * 8);
        for (int i = 0; i < ns; i++)
            for (int c = 0; c < 8; c++) woff[(size_t)i * 8 + c] = cfg->sn_w_off[i][c];
        if (cudaMemcpyAsync(m.sn_woff, woff.data(), woff.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) {
            b200nmpc_destroy(h);
            return set_err(B200NMPC_E_CUDA, "upload of the SNMPC sample offsets failed");
        }
        if ((rc = snmpc_sync_params(h))) { b200nmpc_destroy(h); return rc; }
    }
    CKH(cudaStreamSynchronize(h->stream));
#undef CKH
    *out = h;
    return 0;
}

int b200nmpc_destroy(b200nmpc_handle *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *p : h->allocs) cudaFree(p);
    if (h->ring) cudaFreeHost(h->ring);
    for (cudaEvent_t e : {h->ev0, h->ev1, h->ev_mid, h->ev_t0, h->ev_t1})
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return 0;
}

int b200nmpc_synchronize(b200nmpc_handle *h)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_timer_start(b200nmpc_handle *h)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev_t0, h->stream));
    return 0;
}

int b200nmpc_timer_stop(b200nmpc_handle *h, double *seconds)
{
    if (!h || !seconds) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev_t1, h->stream));
    CK(cudaEventSynchronize(h->ev_t1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
    *seconds = 1e-3 * ms;
    return 0;
}

void *b200nmpc_alloc_host(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void b200nmpc_free_host(void *p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"

// ---- internal helpers ------------------------------------------------------------------------------

// Packs rows x n doubles (row stride ld) into the pinned ring and returns the packed copy, or nullptr if the block is too
// large for the ring (the caller then copies straight from the user's buffer and waits).  A wrap waits for the stream
// once, so that no slot is overwritten before its copy has run.
static int ring_pack(b200nmpc_handle *h, const double *value, size_t rows, size_t n, size_t ld, const double **out)
{
    *out = nullptr;
    const size_t need = rows * n;
    if (!h->ring || need > h->ring_elems / 2) return 0;
    if (h->ring_pos + need > h->ring_elems) {
        CK(cudaStreamSynchronize(h->stream));
        h->ring_pos = 0;
    }
    double *dst = h->ring + h->ring_pos;
    if (ld == n) memcpy(dst, value, need * sizeof(double));
    else
        for (size_t r = 0; r < rows; r++) memcpy(dst + r * n, value + r * ld, n * sizeof(double));
    h->ring_pos += need;
    *out = dst;
    return 0;
}

// host [B][ld_batch] (or [n] broadcast when ld_batch == 0) -> field of n doubles at dst + b*pitch (doubles).
// The caller may reuse `value` as soon as we return: small blocks travel through the pinned ring and the call does not
// wait for the device (stream order keeps set -> solve -> get consistent); large blocks are copied directly and waited for.
static int put_field(b200nmpc_handle *h, const double *value, int ld_batch, int n, double *dst, long long pitch)
{
    if (!value) return set_err(B200NMPC_E_ARG, "null value");
    const double *staged = nullptr;
    int rc;
    if (ld_batch == 0) {
        if ((size_t)n > h->staging_elems) return set_err(B200NMPC_E_ARG, "row too long for staging");
        if ((rc = ring_pack(h, value, 1, (size_t)n, (size_t)n, &staged))) return rc;
        CK(cudaMemcpyAsync(h->staging, staged ? staged : value, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        launch_fill_rows(dst, pitch, 0, h->staging, h->B, n, h->stream);
        h->launches++;
        CK(cudaGetLastError());
    } else {
        if (ld_batch < n) return set_err(B200NMPC_E_ARG, "ld_batch %d < row length %d", ld_batch, n);
        if ((rc = ring_pack(h, value, (size_t)h->B, (size_t)n, (size_t)ld_batch, &staged))) return rc;
        CK(cudaMemcpy2DAsync(dst, (size_t)pitch * sizeof(double), staged ? staged : value,
                             (size_t)(staged ? n : ld_batch) * sizeof(double), (size_t)n * sizeof(double), (size_t)h->B,
                             cudaMemcpyHostToDevice, h->stream));
    }
    if (!staged) CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// field of n doubles at src + b*pitch -> host [B][n]
static int get_field(b200nmpc_handle *h, const double *src, long long pitch, int n, double *out)
{
    if (!out) return set_err(B200NMPC_E_ARG, "null output");
    CK(cudaMemcpy2DAsync(out, (size_t)n * sizeof(double), src, (size_t)pitch * sizeof(double), (size_t)n * sizeof(double),
                         (size_t)h->B, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// Small batches: fetch the whole iterate once per solve; returns false when the cache does not apply (large batch, SNMPC)
static bool xu_cache_fill(b200nmpc_handle *h, int *rc)
{
    *rc = 0;
    const size_t B = h->B, N = h->N, nx = B * (N + 1) * 8, nu = B * N * 2;
    if (h->cfg.variant == B200NMPC_SNMPC || nx + nu > XU_CACHE_MAX_DOUBLES) return false;
    if (h->xu_cached) return true;
    h->host_xu.resize(nx + nu);
    auto fail = [&](cudaError_t e) { *rc = set_err(B200NMPC_E_CUDA, "iterate read-back failed: %s", cudaGetErrorString(e)); return true; };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(h->host_xu.data(), h->m.X, nx * sizeof(double), cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(h->host_xu.data() + nx, h->m.U, nu * sizeof(double), cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(e);
    h->xu_cached = true;
    return true;
}

// SNMPC: stages inside the propagation horizon are linearised per sample and condensed onto the head inputs; they have no
// single (A, B, b) of the 8-state model to hand out (the reference's QP there is over the 8 (1 + n_s) augmented states)
static bool snmpc_head_stage(const b200nmpc_handle *h, int stage)
{
    return h->cfg.variant == B200NMPC_SNMPC && stage < h->cfg.head_H;
}
#define SNMPC_HEAD_MSG "get_from_qp_in: stage %d lies inside the SNMPC propagation horizon (%d stages, condensed per sample)"

// slack-vector entry -> soft-row component for a stage: stage 0: [bu]; 1..N-1: [bu,bx,h]; N: [bx,h]
static int slack_layout(int N, int stage, int *first)
{
    if (stage == 0) { *first = 0; return 1; }
    if (stage == N) { *first = 1; return 2; }
    *first = 0;
    return 3;
}

// write components [first, first+n) of a stage-data group; value layout as put_field
static int set_group(b200nmpc_handle *h, int g, int stage, int first, int n, const double *value, int ld_batch)
{
    if (!value) return set_err(B200NMPC_E_ARG, "null value");
    const int nc = G_NCOMP[g];
    const long long pitch = (long long)(h->N + 1) * nc;
    if (ld_batch == 0) {
        for (int i = 0; i < n; i++) h->host_shared[g][(size_t)stage * nc + first + i] = value[i];
        const double *staged = nullptr;
        int rc = ring_pack(h, value, 1, (size_t)n, (size_t)n, &staged);      // n <= 6: always fits
        if (rc) return rc;
        if (!staged) return set_err(B200NMPC_E_CUDA, "host staging ring unavailable");
        CK(cudaMemcpyAsync(*shared_ptr(h, g) + (size_t)stage * nc + first, staged, (size_t)n * sizeof(double),
                           cudaMemcpyHostToDevice, h->stream));
        double *ip = *inst_ptr(h, g);
        if (ip) {
            launch_fill_rows(ip + (size_t)stage * nc + first, pitch, 0, *shared_ptr(h, g) + (size_t)stage * nc + first, h->B, n,
                             h->stream);
            h->launches++;
            CK(cudaGetLastError());
        }
        return 0;      // stream-ordered: no wait (the values were copied into the pinned ring)
    }
    int rc = ensure_inst(h, g);
    if (rc) return rc;
    return put_field(h, value, ld_batch, n, *inst_ptr(h, g) + (size_t)stage * nc + first, pitch);
}

extern "C" {

int b200nmpc_set(b200nmpc_handle *h, int stage, int field, const double *value, int ld_batch)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    h->xu_cached = false;
    const int N = h->N;
    switch (field) {
        case B200NMPC_F_X:
            if (stage < 0 || stage > N) return set_err(B200NMPC_E_STAGE, "set('x'): stage %d not in [0,%d]", stage, N);
            if (h->cfg.variant == B200NMPC_SNMPC) {
                // augmented state [x; x^1..x^ns]: the samples of stages beyond the propagation horizon are frozen
                // copies that enter nothing (pred_model_dynamic_disc.py:204-210) and are not stored
                const int ns = h->cfg.n_samples, na = 8 * (1 + ns);
                if (!value) return set_err(B200NMPC_E_ARG, "null value");
                if (ld_batch != 0 && ld_batch < na) return set_err(B200NMPC_E_ARG, "ld_batch %d < row length %d", ld_batch, na);
                rc = put_field(h, value, ld_batch, 8, h->m.X + (size_t)stage * 8, (long long)(N + 1) * 8);
                if (rc || stage > h->m.HA) return rc;
                return put_field(h, value + 8, ld_batch, ns * 8, h->m.XS + (size_t)stage * ns * 8, (long long)(h->m.HA + 1) * ns * 8);
            }
            return put_field(h, value, ld_batch, 8, h->m.X + (size_t)stage * 8, (long long)(N + 1) * 8);
        case B200NMPC_F_U:
            if (stage < 0 || stage >= N) return set_err(B200NMPC_E_STAGE, "set('u'): stage %d not in [0,%d)", stage, N);
            return put_field(h, value, ld_batch, 2, h->m.U + (size_t)stage * 2, (long long)N * 2);
        case B200NMPC_F_YREF:
            if (stage < 0 || stage > N) return set_err(B200NMPC_E_STAGE, "set('yref'): stage %d not in [0,%d]", stage, N);
            return put_field(h, value, ld_batch, stage == N ? 4 : 6, h->m.yref + (size_t)stage * 6, (long long)(N + 1) * 6);
        case B200NMPC_F_P: {
            if (h->cfg.variant != B200NMPC_SNMPC) return set_err(B200NMPC_E_FIELD, "set('p') is only meaningful for the SNMPC variant");
            if (stage < 0 || stage > N) return set_err(B200NMPC_E_STAGE, "set('p'): stage %d not in [0,%d]", stage, N);
            if (!value) return set_err(B200NMPC_E_ARG, "null value");
            if (ld_batch != 0) return set_err(B200NMPC_E_ARG, "set('p'): one parameter set for all instances (ld_batch must be 0)");
            // p = [A.flatten(), gamma, stop_flag]; A and gamma are the same for every stage in the reference
            const size_t na = h->sn_A.size();
            if (!(value[na] > 0.0 && value[na] <= 1.0)) return set_err(B200NMPC_E_ARG, "set('p'): gamma must be in (0,1]");
            for (size_t i = 0; i < na; i++) h->sn_A[i] = value[i];
            h->sn_gamma = value[na];
            h->sn_stop[stage] = value[na + 1] != 0.0;
            h->sn_dirty = true;
            return 0;
        }
        default:
            return set_err(B200NMPC_E_FIELD, "set: unknown field %d", field);
    }
}

int b200nmpc_constraints_set(b200nmpc_handle *h, int stage, int field, const double *value, int ld_batch)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    const int N = h->N;
    if (stage < 0 || stage > N) return set_err(B200NMPC_E_STAGE, "constraints_set: stage %d not in [0,%d]", stage, N);
    switch (field) {
        case B200NMPC_F_LBX:
        case B200NMPC_F_UBX:
            if (stage == 0) {     // lbx_0 = ubx_0 = x0 (NMPC_class.py:111-112,243-246)
                if (h->cfg.variant == B200NMPC_SNMPC) {      // the x0 sample matrix (SNMPC_class.py:120-121,259-264)
                    const int ns = h->cfg.n_samples, na = 8 * (1 + ns);
                    if (!value) return set_err(B200NMPC_E_ARG, "null value");
                    if (ld_batch != 0 && ld_batch < na) return set_err(B200NMPC_E_ARG, "ld_batch %d < row length %d", ld_batch, na);
                    rc = put_field(h, value, ld_batch, 8, h->m.x0, 8);
                    if (rc) return rc;
                    return put_field(h, value + 8, ld_batch, ns * 8, h->m.x0s, (long long)ns * 8);
                }
                return put_field(h, value, ld_batch, 8, h->m.x0, 8);
            }
            return set_group(h, field == B200NMPC_F_LBX ? G_LB : G_UB, stage, 1, 1, value, ld_batch);
        case B200NMPC_F_LBU:
        case B200NMPC_F_UBU:
            if (stage >= N) return set_err(B200NMPC_E_STAGE, "constraints_set('lbu'/'ubu'): stage %d has no input", stage);
            return set_group(h, field == B200NMPC_F_LBU ? G_LB : G_UB, stage, 0, 1, value, ld_batch);
        case B200NMPC_F_LH:
        case B200NMPC_F_UH:
            if (stage == 0) return set_err(B200NMPC_E_STAGE, "constraints_set('lh'/'uh'): stage 0 has no h row (nh_0 = 0)");
            return set_group(h, field == B200NMPC_F_LH ? G_LB : G_UB, stage, 2, 1, value, ld_batch);
        default:
            return set_err(B200NMPC_E_FIELD, "constraints_set: unknown field %d", field);
    }
}

int b200nmpc_cost_set(b200nmpc_handle *h, int stage, int field, const double *value, int ld_batch)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    const int N = h->N;
    if (stage < 0 || stage > N) return set_err(B200NMPC_E_STAGE, "cost_set: stage %d not in [0,%d]", stage, N);
    switch (field) {
        case B200NMPC_F_W:
            return set_group(h, G_W, stage, 0, stage == N ? 4 : 6, value, ld_batch);
        case B200NMPC_F_ZL_LIN:
        case B200NMPC_F_ZU_LIN:
        case B200NMPC_F_ZL_QUAD:
        case B200NMPC_F_ZU_QUAD: {
            int first;
            const int n = slack_layout(N, stage, &first);
            const int g = field == B200NMPC_F_ZL_LIN ? G_ZL : field == B200NMPC_F_ZU_LIN ? G_ZU
                        : field == B200NMPC_F_ZL_QUAD ? G_QL : G_QU;
            return set_group(h, g, stage, first, n, value, ld_batch);
        }
        default:
            return set_err(B200NMPC_E_FIELD, "cost_set: unknown field %d", field);
    }
}

int b200nmpc_solve(b200nmpc_handle *h, int *status)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    h->xu_cached = false;
    if ((rc = snmpc_sync_params(h))) return rc;
    CK(cudaEventRecord(h->ev0, h->stream));
    // SNMPC: the sample copies are extra work items of this launch; condensing and the sample step ride in the QP kernel
    launch_linearize(h->cfg, h->m, h->stream);
    CK(cudaEventRecord(h->ev_mid, h->stream));
    if (!h->order_external) h->m.order = (h->order_hint && h->order_valid) ? h->order : nullptr;
    const int e = launch_qp_solve_warp(h->cfg, h->m, h->stream);
    if (e != 0) return set_err(B200NMPC_E_CUDA, "k_qp_solve_warp launch failed: %s", cudaGetErrorString((cudaError_t)e));
    h->launches += 2;
    if (h->cfg.variant == B200NMPC_R2) { launch_r2_tighten(h->cfg, h->m, h->stream); h->launches++; }
    if (!h->order_external) {
        h->m.order = nullptr;
        if (h->order_hint && h->B > 1) { launch_launch_order(h->m, h->order, h->B, h->stream); h->launches++; h->order_valid = true; }
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    CK(cudaGetLastError());
    if (status) return b200nmpc_get_status(h, status);
    return 0;
}

int b200nmpc_options_set(b200nmpc_handle *h, const char *field, double value)
{
    if (!h || !field) return set_err(B200NMPC_E_ARG, "null argument");
    if (!strcmp(field, "launch_order_hint")) {
        h->order_hint = value != 0.0;
        h->order_valid = false;
        return 0;
    }
    return set_err(B200NMPC_E_FIELD, "options_set: unknown field '%s'", field);
}

int b200nmpc_get_status(b200nmpc_handle *h, int *status)
{
    if (!h || !status) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(status, h->m.status, (size_t)h->B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_get(b200nmpc_handle *h, int stage, int field, double *out)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    const int N = h->N;
    const size_t B = h->B;
    const long long lpitch = (long long)(N + 1) * STG;
    switch (field) {
        case B200NMPC_F_X:
            if (stage < 0 || stage > N) return set_err(B200NMPC_E_STAGE, "get('x'): stage %d not in [0,%d]", stage, N);
            if (h->cfg.variant == B200NMPC_SNMPC) {
                // [x; x^1..x^ns]; stages beyond the propagation horizon hold the frozen samples of stage UPH
                if (!out) return set_err(B200NMPC_E_ARG, "null output");
                const size_t ns = h->cfg.n_samples, na = 8 * (1 + ns);
                const int ks = stage < h->cfg.head_H ? stage : h->cfg.head_H;
                CK(cudaMemcpy2DAsync(out, na * sizeof(double), h->m.X + (size_t)stage * 8, (size_t)(N + 1) * 8 * sizeof(double),
                                     8 * sizeof(double), B, cudaMemcpyDeviceToHost, h->stream));
                CK(cudaMemcpy2DAsync(out + 8, na * sizeof(double), h->m.XS + (size_t)ks * ns * 8,
                                     (size_t)(h->m.HA + 1) * ns * 8 * sizeof(double), ns * 8 * sizeof(double), B,
                                     cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                return 0;
            }
            if (xu_cache_fill(h, &rc)) {
                if (rc) return rc;
                if (!out) return set_err(B200NMPC_E_ARG, "null output");
                for (size_t b = 0; b < B; b++) memcpy(out + b * 8, h->host_xu.data() + (b * (N + 1) + stage) * 8, 8 * sizeof(double));
                return 0;
            }
            return get_field(h, h->m.X + (size_t)stage * 8, (long long)(N + 1) * 8, 8, out);
        case B200NMPC_F_U:
            if (stage < 0 || stage >= N) return set_err(B200NMPC_E_STAGE, "get('u'): stage %d not in [0,%d)", stage, N);
            if (xu_cache_fill(h, &rc)) {
                if (rc) return rc;
                if (!out) return set_err(B200NMPC_E_ARG, "null output");
                const double *U = h->host_xu.data() + B * (N + 1) * 8;
                for (size_t b = 0; b < B; b++) memcpy(out + b * 2, U + (b * N + stage) * 2, 2 * sizeof(double));
                return 0;
            }
            return get_field(h, h->m.U + (size_t)stage * 2, (long long)N * 2, 2, out);
        case B200NMPC_F_SL:
        case B200NMPC_F_SU: {
            if (stage < 0 || stage > N) return set_err(B200NMPC_E_STAGE, "get('sl'/'su'): stage %d not in [0,%d]", stage, N);
            int first;
            const int n = slack_layout(N, stage, &first);
            const double *src = (field == B200NMPC_F_SL ? h->m.sl : h->m.su) + (size_t)stage * 3 + first;
            return get_field(h, src, (long long)(N + 1) * 3, n, out);
        }
        case B200NMPC_F_QP_A: {
            if (stage < 0 || stage >= N) return set_err(B200NMPC_E_STAGE, "get_from_qp_in('A'): stage %d not in [0,%d)", stage, N);
            if (snmpc_head_stage(h, stage)) return set_err(B200NMPC_E_STAGE, SNMPC_HEAD_MSG, stage, h->cfg.head_H);
            if (!out) return set_err(B200NMPC_E_ARG, "null output");
            // rebuild the dense 8x8 from the structural 6x6 block
            std::vector<double> a36(B * 36);
            rc = get_field(h, h->m.stg + (size_t)stage * STG + L_A, lpitch, 36, a36.data());
            if (rc) return rc;
            for (size_t b = 0; b < B; b++) {
                double *A = out + b * 64;
                for (int i = 0; i < 64; i++) A[i] = 0.0;
                A[0] = 1.0; A[9] = 1.0; A[6 * 8 + 6] = 1.0; A[7 * 8 + 7] = 1.0;
                for (int r = 0; r < 6; r++)
                    for (int c = 0; c < 6; c++) A[r * 8 + 2 + c] = a36[b * 36 + c * 6 + r];      // stored transposed
            }
            return 0;
        }
        case B200NMPC_F_QP_B: {
            if (stage < 0 || stage >= N) return set_err(B200NMPC_E_STAGE, "get_from_qp_in('B'): stage %d not in [0,%d)", stage, N);
            if (snmpc_head_stage(h, stage)) return set_err(B200NMPC_E_STAGE, SNMPC_HEAD_MSG, stage, h->cfg.head_H);
            if (!out) return set_err(B200NMPC_E_ARG, "null output");
            std::vector<double> b12(B * 12);
            rc = get_field(h, h->m.stg + (size_t)stage * STG + L_B, lpitch, 12, b12.data());
            if (rc) return rc;
            for (size_t b = 0; b < B; b++) {
                double *Bm = out + b * 16;
                for (int r = 0; r < 6; r++) { Bm[r * 2] = b12[b * 12 + r]; Bm[r * 2 + 1] = b12[b * 12 + 6 + r]; }      // stored transposed
                Bm[12] = 0.0; Bm[13] = h->cfg.dt; Bm[14] = h->cfg.dt; Bm[15] = 0.0;
            }
            return 0;
        }
        case B200NMPC_F_QP_b:
            if (stage < 0 || stage >= N) return set_err(B200NMPC_E_STAGE, "get_from_qp_in('b'): stage %d not in [0,%d)", stage, N);
            if (snmpc_head_stage(h, stage)) return set_err(B200NMPC_E_STAGE, SNMPC_HEAD_MSG, stage, h->cfg.head_H);
            return get_field(h, h->m.stg + (size_t)stage * STG + L_b, lpitch, 8, out);
        case B200NMPC_F_LBX: case B200NMPC_F_UBX: case B200NMPC_F_LBU: case B200NMPC_F_UBU: case B200NMPC_F_LH: case B200NMPC_F_UH: {
            // the bound in force for every instance, [B][1]: the per-instance value where one exists (R2NMPC writes its
            // tightened lbx / ubx / uh on the device, Reduced_Robustified_NMPC_class.py:335-365), else the shared value
            if (!out) return set_err(B200NMPC_E_ARG, "null output");
            const bool lower = field == B200NMPC_F_LBX || field == B200NMPC_F_LBU || field == B200NMPC_F_LH;
            const int row = (field == B200NMPC_F_LBU || field == B200NMPC_F_UBU) ? 0 : (field == B200NMPC_F_LBX || field == B200NMPC_F_UBX) ? 1 : 2;
            const int lo = row == 0 ? 0 : 1, hi = row == 0 ? N - 1 : N;
            if (stage < lo || stage > hi) return set_err(B200NMPC_E_STAGE, "get(bound): stage %d not in [%d,%d]", stage, lo, hi);
            const int g = lower ? G_LB : G_UB;
            const double *ip = *inst_ptr(h, g);
            if (ip) return get_field(h, ip + (size_t)stage * 3 + row, (long long)(N + 1) * 3, 1, out);
            for (size_t b = 0; b < B; b++) out[b] = h->host_shared[g][(size_t)stage * 3 + row];
            return 0;
        }
        default:
            return set_err(B200NMPC_E_FIELD, "get: unknown field %d", field);
    }
}

int b200nmpc_get_cost(b200nmpc_handle *h, double *out)
{
    if (!h || !out) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, h->m.cost, (size_t)h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

static int batch_seconds(b200nmpc_handle *h, double *sec)
{
    *sec = 0.0;
    if (!h->timed) return 0;
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *sec = 1e-3 * ms;
    return 0;
}

int b200nmpc_get_stats(b200nmpc_handle *h, int stat, double *out)
{
    if (!h || !out) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = use_device(h);
    if (rc) return rc;
    const int B = h->B;
    double sec;
    switch (stat) {
        case B200NMPC_S_TIME_TOT:
            if ((rc = batch_seconds(h, &sec))) return rc;
            for (int b = 0; b < B; b++) out[b] = sec / B;
            return 0;
        case B200NMPC_S_TIME_BATCH:
            if ((rc = batch_seconds(h, &sec))) return rc;
            out[0] = sec;
            return 0;
        case B200NMPC_S_TIME_LINEARIZE:
        case B200NMPC_S_TIME_QP: {
            out[0] = 0.0;
            if (!h->timed) return 0;
            CK(cudaEventSynchronize(h->ev1));
            float ms = 0.f;
            if (stat == B200NMPC_S_TIME_LINEARIZE) CK(cudaEventElapsedTime(&ms, h->ev0, h->ev_mid));
            else CK(cudaEventElapsedTime(&ms, h->ev_mid, h->ev1));
            out[0] = 1e-3 * ms;
            return 0;
        }
        case B200NMPC_S_SQP_ITER:
            for (int b = 0; b < B; b++) out[b] = 1.0;
            return 0;
        case B200NMPC_S_QP_ITER: {
            std::vector<int> it(B);
            CK(cudaMemcpyAsync(it.data(), h->m.qp_iter, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            for (int b = 0; b < B; b++) out[b] = it[b];
            return 0;
        }
        case B200NMPC_S_KKT:
            CK(cudaMemcpyAsync(out, h->m.kkt, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            return 0;
        default:
            return set_err(B200NMPC_E_FIELD, "get_stats: unknown stat %d", stat);
    }
}

int b200nmpc_reset(b200nmpc_handle *h)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    h->xu_cached = false;
    const size_t B = h->B, N = h->N;
    CK(cudaMemsetAsync(h->m.X, 0, (N + 1) * 8 * B * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->m.U, 0, N * 2 * B * sizeof(double), h->stream));
    if (h->m.XS) CK(cudaMemsetAsync(h->m.XS, 0, B * (h->m.HA + 1) * h->cfg.n_samples * 8 * sizeof(double), h->stream));
    return 0;
}

// The caller-visible arrays are stored in their host order, so the bulk entry points are plain copies.
static int copy_in(b200nmpc_handle *h, double *dst, const double *src, size_t n, cudaMemcpyKind kind)
{
    if (!h || !src) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dst, src, n * sizeof(double), kind, h->stream));
    return 0;
}

// SNMPC: the x0 sample matrix follows x0 (compute_x0dist, SNMPC_class.py:259-264)
static int after_x0(b200nmpc_handle *h)
{
    if (h->cfg.variant != B200NMPC_SNMPC) return 0;
    launch_snmpc_x0_samples(h->cfg, h->m, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}
int b200nmpc_set_x0_dev(b200nmpc_handle *h, const double *x0_dev)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = copy_in(h, h->m.x0, x0_dev, (size_t)h->B * 8, cudaMemcpyDeviceToDevice);
    return rc ? rc : after_x0(h);
}
int b200nmpc_set_x0(b200nmpc_handle *h, const double *x0)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = copy_in(h, h->m.x0, x0, (size_t)h->B * 8, cudaMemcpyHostToDevice);
    return rc ? rc : after_x0(h);
}
int b200nmpc_set_yref_all_dev(b200nmpc_handle *h, const double *yref_dev)
{
    return h ? copy_in(h, h->m.yref, yref_dev, (size_t)h->B * (h->N + 1) * 6, cudaMemcpyDeviceToDevice)
             : set_err(B200NMPC_E_ARG, "null handle");
}
int b200nmpc_set_yref_all(b200nmpc_handle *h, const double *yref)
{
    return h ? copy_in(h, h->m.yref, yref, (size_t)h->B * (h->N + 1) * 6, cudaMemcpyHostToDevice)
             : set_err(B200NMPC_E_ARG, "null handle");
}

int b200nmpc_get_yref_all(b200nmpc_handle *h, double *yref)
{
    if (!h || !yref) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(yref, h->m.yref, (size_t)h->B * (h->N + 1) * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_set_iterate(b200nmpc_handle *h, const double *X, const double *U)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    h->xu_cached = false;
    const size_t B = h->B, N = h->N;
    if (X) CK(cudaMemcpyAsync(h->m.X, X, B * (N + 1) * 8 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (U) CK(cudaMemcpyAsync(h->m.U, U, B * N * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_init_iterate_from_x0(b200nmpc_handle *h)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    h->xu_cached = false;
    launch_init_iterate(h->cfg, h->m, h->stream);
    h->launches++;
    if (h->cfg.variant == B200NMPC_SNMPC) { launch_snmpc_init_samples(h->cfg, h->m, h->stream); h->launches++; }
    CK(cudaGetLastError());
    return 0;
}

int b200nmpc_get_iterate_dev(b200nmpc_handle *h, double *X_dev, double *U_dev)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    const size_t B = h->B, N = h->N;
    if (X_dev) CK(cudaMemcpyAsync(X_dev, h->m.X, B * (N + 1) * 8 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (U_dev) CK(cudaMemcpyAsync(U_dev, h->m.U, B * N * 2 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

int b200nmpc_get_results_async(b200nmpc_handle *h, double *X, double *U, double *solver_debug)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    const size_t B = h->B, N = h->N;
    if (X) CK(cudaMemcpyAsync(X, h->m.X, B * (N + 1) * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (U) CK(cudaMemcpyAsync(U, h->m.U, B * N * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (solver_debug) {
        launch_solver_debug(h->cfg, h->m, 0.0, h->staging, h->stream);
        h->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(solver_debug, h->staging, B * 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    return 0;
}

int b200nmpc_get_iterate(b200nmpc_handle *h, double *X, double *U)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    const size_t B = h->B, N = h->N;
    if (X) CK(cudaMemcpyAsync(X, h->m.X, B * (N + 1) * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (U) CK(cudaMemcpyAsync(U, h->m.U, B * N * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_get_solver_debug_dev(b200nmpc_handle *h, double *out_dev)
{
    if (!h || !out_dev) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = use_device(h);
    if (rc) return rc;
    double sec;
    if ((rc = batch_seconds(h, &sec))) return rc;
    launch_solver_debug(h->cfg, h->m, sec / h->B, out_dev, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int b200nmpc_get_solver_debug(b200nmpc_handle *h, double *out)
{
    if (!h || !out) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = b200nmpc_get_solver_debug_dev(h, h->staging);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, h->staging, (size_t)h->B * 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_update_cost_function_weights(b200nmpc_handle *h, const double *params, int ld_batch)
{
    if (!h || !params) return set_err(B200NMPC_E_ARG, "null argument");
    int rc = use_device(h);
    if (rc) return rc;
    const int N = h->N;
    const int gs[5] = {G_W, G_ZL, G_ZU, G_QL, G_QU};
    if (ld_batch == 0) {
        const double *p = params;
        for (int k = 0; k <= N; k++) {
            double *W = &h->host_shared[G_W][(size_t)k * 6];
            W[0] = p[0]; W[1] = p[0]; W[2] = p[1]; W[3] = p[2]; W[4] = k < N ? p[3] : 0.0; W[5] = k < N ? p[4] : 0.0;
            for (int j = 0; j < 3; j++) {
                h->host_shared[G_ZL][(size_t)k * 3 + j] = p[5]; h->host_shared[G_ZU][(size_t)k * 3 + j] = p[5];
                h->host_shared[G_QL][(size_t)k * 3 + j] = p[6]; h->host_shared[G_QU][(size_t)k * 3 + j] = p[6];
            }
        }
        for (int g : gs) {
            if ((rc = upload_shared(h, g))) return rc;
            double *ip = *inst_ptr(h, g);
            if (ip) {
                const int n = (N + 1) * G_NCOMP[g];
                launch_fill_rows(ip, n, 0, *shared_ptr(h, g), h->B, n, h->stream);
                h->launches++;
            }
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (ld_batch != 7) return set_err(B200NMPC_E_ARG, "per-instance params must be [B][7] (ld_batch 7), got %d", ld_batch);
    for (int g : gs)
        if ((rc = ensure_inst(h, g))) return rc;
    CK(cudaMemcpyAsync(h->staging, params, (size_t)h->B * 7 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_update_weights(h->cfg, h->m, h->staging, 1, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

static int need_snmpc(b200nmpc_handle *h)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    if (h->cfg.variant != B200NMPC_SNMPC) return set_err(B200NMPC_E_FIELD, "only meaningful for the SNMPC variant");
    return use_device(h);
}

int b200nmpc_set_x0_samples(b200nmpc_handle *h, const double *x0s)
{
    int rc = need_snmpc(h);
    if (rc) return rc;
    if (!x0s) return set_err(B200NMPC_E_ARG, "null argument");
    const size_t ns = h->cfg.n_samples, na = 8 * (1 + ns), B = h->B;
    CK(cudaMemcpy2DAsync(h->m.x0, 8 * sizeof(double), x0s, na * sizeof(double), 8 * sizeof(double), B, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpy2DAsync(h->m.x0s, ns * 8 * sizeof(double), x0s + 8, na * sizeof(double), ns * 8 * sizeof(double), B,
                         cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_set_samples(b200nmpc_handle *h, const double *XS)
{
    int rc = need_snmpc(h);
    if (rc) return rc;
    if (!XS) return set_err(B200NMPC_E_ARG, "null argument");
    CK(cudaMemcpyAsync(h->m.XS, XS, (size_t)h->B * (h->m.HA + 1) * h->cfg.n_samples * 8 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_get_samples(b200nmpc_handle *h, double *XS)
{
    int rc = need_snmpc(h);
    if (rc) return rc;
    if (!XS) return set_err(B200NMPC_E_ARG, "null argument");
    CK(cudaMemcpyAsync(XS, h->m.XS, (size_t)h->B * (h->m.HA + 1) * h->cfg.n_samples * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_snmpc_uph(const b200nmpc_handle *h)
{
    if (!h || h->cfg.variant != B200NMPC_SNMPC) return B200NMPC_E_ARG;
    int H = 0;
    while (H <= h->N && h->sn_stop[H] == 0) H++;
    return H;
}

int b200nmpc_snmpc_get_head(b200nmpc_handle *h, double *out, int *stride)
{
    int rc = need_snmpc(h);
    if (rc) return rc;
    if (stride) *stride = h->m.head_stride;
    if (!out) return 0;
    CK(cudaMemcpyAsync(out, h->m.head, (size_t)h->B * h->m.head_stride * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_r2_tighten(b200nmpc_handle *h)
{
    if (!h) return set_err(B200NMPC_E_ARG, "null handle");
    int rc = use_device(h);
    if (rc) return rc;
    if ((rc = ensure_inst(h, G_LB)) || (rc = ensure_inst(h, G_UB))) return rc;
    launch_r2_tighten(h->cfg, h->m, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"

// ---- closed loop on the device (include/b200nmpc.h, b200nmpc_sim_*) -----------------------------------------
#include "nmpc_sim.cuh"

struct b200nmpc_sim {
    b200nmpc_handle *h;
    SimMem s;
    std::vector<void *> allocs;
    double *x0_stage;            // [B][8] device staging of X0_MPC
    double *dist_deriv, *dist_est;
    double *play[2];             // device copies of playback rows
    int *end_idx;                // [B] device copy of the segment end indices
    int *mask;                   // [B] device copy of a per-instance reset mask
    int *order;                  // [B] launch order of the next solve (heaviest instance first)
    bool order_valid;
    int steps;
    double run_seconds;          // device time of all runs since reset
    long long run_steps;
    bool time_pending;           // the last run was asynchronous: its e0 -> e1 time is not in run_seconds yet
    size_t play_cap[2];          // doubles allocated behind play[ch]
    cudaEvent_t e0, e1;
};

// adds the device time of an asynchronous run (b200nmpc_sim_run with seconds == NULL) to the totals; blocks on its end event
static int sim_collect_time(b200nmpc_sim *s)
{
    if (!s->time_pending) return 0;
    CK(cudaEventSynchronize(s->e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, s->e0, s->e1));
    s->run_seconds += 1e-3 * ms;
    s->time_pending = false;
    return 0;
}

template <typename T>
static int sim_alloc(b200nmpc_sim *s, T **p, size_t n)
{
    CK(cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)));
    CK(cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(T), s->h->stream));
    s->allocs.push_back(*p);
    return 0;
}

extern "C" {

int b200nmpc_sim_create(b200nmpc_handle *h, const b200nmpc_sim_config *c, const double *ref_px, const double *ref_py,
                        const double *ref_yaw, const double *ref_v, b200nmpc_sim **out)
{
    if (!h || !c || !ref_px || !ref_py || !ref_yaw || !ref_v || !out) return set_err(B200NMPC_E_ARG, "null argument");
    *out = nullptr;
    if (c->n_ref < 2 || !(c->Ts > 0.0) || !(c->Tp > 0.0) || c->plant_substeps < 1 || c->log_capacity < 1)
        return set_err(B200NMPC_E_CONFIG, "invalid closed-loop configuration");
    for (int i = 0; i < 8; i++)
        if (c->windows[i] < 1 || c->windows[i] > 4) return set_err(B200NMPC_E_CONFIG, "state-estimation windows must be in [1,4]");
    int rc = use_device(h);
    if (rc) return rc;
    const int n = c->n_ref;
    // travel time from every point to its successor, and the longest time-horizon window over all start indices
    std::vector<double> seg(n);
    for (int i = 0; i < n; i++) {
        const int nx = (i + 1 < n) ? i + 1 : 0;
        seg[i] = std::hypot(ref_px[nx] - ref_px[i], ref_py[nx] - ref_py[i]) / ref_v[nx];
    }
    int maxw = 2;
    for (int c0 = 0; c0 < n; c0++) {
        double T = 0.0;
        int cur = c0, L = 1;
        while (T <= c->Tp && L < 4096) { T += seg[cur]; cur = (cur + 1 < n) ? cur + 1 : 0; L++; }
        if (L > maxw) maxw = L;
    }
    maxw = (maxw + 2) & ~1;
    if (planner_smem_bytes(maxw) > 48 * 1024)
        return set_err(B200NMPC_E_CONFIG, "planner window of %d points does not fit shared memory", maxw);
    b200nmpc_sim *s = new b200nmpc_sim();
    s->h = h; s->steps = 0; s->run_seconds = 0.0; s->run_steps = 0; s->dist_deriv = s->dist_est = nullptr;
    s->time_pending = false; s->play_cap[0] = s->play_cap[1] = 0;
    s->play[0] = s->play[1] = nullptr;
    memset(&s->s, 0, sizeof s->s);
    SimMem &m = s->s;
    const size_t B = h->B, cap = c->log_capacity;
    double *rp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const double *src[5] = {ref_px, ref_py, ref_yaw, ref_v, seg.data()};
    for (int i = 0; i < 5 && !rc; i++) {
        rc = sim_alloc(s, &rp[i], (size_t)n);
        if (!rc && cudaMemcpyAsync(rp[i], src[i], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
            rc = set_err(B200NMPC_E_CUDA, "upload of the reference trajectory failed");
    }
    m.ref_px = rp[0]; m.ref_py = rp[1]; m.ref_yaw = rp[2]; m.ref_v = rp[3]; m.seg_dt = rp[4];
    m.n_ref = n; m.maxw = maxw; m.Ts = c->Ts; m.Tp = c->Tp; m.nsub = c->plant_substeps; m.cap = (int)cap;
    m.jump_thresh = 250.0 * (M_PI / 180.0);        // np.deg2rad(250)
    for (int i = 0; i < 8; i++) m.win[i] = c->windows[i];
#define SA(ptr, cnt) if (!rc) rc = sim_alloc(s, &(ptr), (size_t)(cnt))
    SA(m.pose, B * 4); SA(m.cilx, B * 7); SA(m.pred_x1, B * 8); SA(m.fbuf, B * 32); SA(m.fcount, B); SA(m.ref_idx, B);
    SA(m.log_mpcx, B * (cap + 1) * 8); SA(m.log_cilx, B * (cap + 1) * 7); SA(m.log_u, B * cap * 2);
    SA(m.log_ref, B * cap * 4); SA(m.log_dbg, B * cap * 5); SA(m.log_idx, B * cap);
    SA(m.log_distx, B * (cap + 1) * 7); SA(m.log_dist[0], B * cap * 7); SA(m.log_dist[1], B * cap * 7);
    SA(s->x0_stage, B * 8); SA(s->dist_deriv, B * 7); SA(s->dist_est, B * 7);
    SA(m.stop_step, B); SA(m.stop_flags, B); SA(s->end_idx, B); SA(s->mask, B); SA(s->order, B);
    if (!rc && cudaMemsetAsync(m.stop_step, 0xFF, (size_t)B * sizeof(int), h->stream) != cudaSuccess)      // -1: running
        rc = set_err(B200NMPC_E_CUDA, "closed-loop set-up failed");
#undef SA
    if (!rc && (cudaEventCreate(&s->e0) != cudaSuccess || cudaEventCreate(&s->e1) != cudaSuccess))
        rc = set_err(B200NMPC_E_CUDA, "event creation failed");
    if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = set_err(B200NMPC_E_CUDA, "closed-loop set-up failed");
    if (rc) { for (void *p : s->allocs) cudaFree(p); delete s; return rc; }
    *out = s;
    return 0;
}

int b200nmpc_sim_destroy(b200nmpc_sim *s)
{
    if (!s) return 0;
    cudaSetDevice(s->h->device);
    cudaStreamSynchronize(s->h->stream);
    for (void *p : s->allocs) cudaFree(p);
    cudaEventDestroy(s->e0);
    cudaEventDestroy(s->e1);
    delete s;
    return 0;
}

int b200nmpc_sim_reset(b200nmpc_sim *s, const double *X0_MPC)
{
    if (!s || !X0_MPC) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    h->xu_cached = false;
    CK(cudaMemcpyAsync(s->x0_stage, X0_MPC, (size_t)h->B * 8 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_sim_reset(s->s, s->x0_stage, h->B, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    if ((rc = b200nmpc_set_x0_dev(h, s->x0_stage))) return rc;
    if ((rc = b200nmpc_init_iterate_from_x0(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    s->steps = 0; s->run_seconds = 0.0; s->run_steps = 0; s->order_valid = false; s->time_pending = false;
    return 0;
}

int b200nmpc_sim_reset_instances(b200nmpc_sim *s, const int *mask, const double *X0_MPC)
{
    if (!s || !mask || !X0_MPC) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    h->xu_cached = false;
    CK(cudaMemcpyAsync(s->x0_stage, X0_MPC, (size_t)h->B * 8 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(s->mask, mask, (size_t)h->B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    launch_sim_reset_masked(h->cfg, h->m, s->s, s->x0_stage, s->mask, s->steps, h->stream);
    h->launches++;
    if (h->cfg.variant == B200NMPC_SNMPC) {
        if ((rc = snmpc_sync_params(h))) return rc;
        launch_snmpc_x0_samples(h->cfg, h->m, h->stream);                     // all instances: a pure function of x0
        launch_snmpc_init_samples_masked(h->cfg, h->m, s->mask, h->stream);
        h->launches += 2;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_sim_run(b200nmpc_sim *s, int nsteps, double *seconds)
{
    if (!s || nsteps < 0) return set_err(B200NMPC_E_ARG, "bad argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    if ((rc = snmpc_sync_params(h))) return rc;
    // the logs are linear arrays of `log_capacity` steps and the device-side stop predicates read the log row of the step:
    // a run past the capacity would silently stop logging and stop evaluating them
    if ((long long)s->steps + nsteps > (long long)s->s.cap)
        return set_err(B200NMPC_E_ARG, "sim_run: %d + %d steps exceed the log capacity %d (b200nmpc_sim_rebase_logs() restarts the "
                       "log at the current state, or create the simulation with a larger log_capacity)", s->steps, nsteps, s->s.cap);
    if ((rc = sim_collect_time(s))) return rc;
    const bool snmpc = h->cfg.variant == B200NMPC_SNMPC;
    CK(cudaEventRecord(s->e0, h->stream));
    h->m.halt = s->s.stop_on ? s->s.stop_step : nullptr;      // stopped instances skip the solve kernels of this run
    h->order_external = true;                                 // the loop keeps its own launch order (it knows the halted instances)
    const bool lpt = h->order_hint;
    for (int i = 0; i < nsteps; i++) {
        launch_planner(s->s, h->m.yref, h->B, h->N, s->steps, h->stream);
        h->launches++;
        h->m.order = (lpt && s->order_valid) ? s->order : nullptr;
        if ((rc = b200nmpc_solve(h, nullptr))) { h->m.halt = nullptr; h->m.order = nullptr; h->order_external = false; return rc; }
        if (snmpc) { launch_snmpc_reinit_failed(h->cfg, h->m, h->stream); h->launches++; }
        launch_sim_advance(h->cfg, h->m, s->s, s->steps, h->stream);
        h->launches++;
        if (lpt) { launch_launch_order(h->m, s->order, h->B, h->stream); h->launches++; s->order_valid = true; }
        if (snmpc) { launch_snmpc_x0_samples(h->cfg, h->m, h->stream); h->launches++; }
        s->steps++;
    }
    h->m.halt = nullptr;
    h->m.order = nullptr;
    h->order_external = false;
    h->order_valid = false;                                   // the next plain solve starts without a hint
    CK(cudaEventRecord(s->e1, h->stream));
    CK(cudaGetLastError());
    s->run_steps += nsteps;
    s->time_pending = true;
    if (seconds) {
        const double before = s->run_seconds;
        if ((rc = sim_collect_time(s))) return rc;
        *seconds = s->run_seconds - before;
    }
    return 0;
}

// Restarts the log at the current state: row 0 of MPC_SimX / CiLX / DisturbedX becomes the state after the last step and
// the step counter returns to 0, so that a caller that only reads the newest rows (RL_WMPC/environment.py:112-240 inspects
// one 20-step chunk at a time) can run for ever with a bounded log.  Vehicles, controller and filter state are untouched.
int b200nmpc_sim_rebase_logs(b200nmpc_sim *s)
{
    if (!s) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    if (s->s.stop_on) return set_err(B200NMPC_E_ARG, "sim_rebase_logs: stop conditions count steps from the last reset; disable them first");
    if ((rc = sim_collect_time(s))) return rc;
    const int row = s->steps < s->s.cap ? s->steps : s->s.cap;
    if (row > 0) {
        const size_t B = h->B, cap1 = (size_t)s->s.cap + 1;
        auto move = [&](double *base, size_t w) -> int {
            if (!base) return 0;
            CK(cudaMemcpy2DAsync(base, cap1 * w * sizeof(double), base + (size_t)row * w, cap1 * w * sizeof(double), w * sizeof(double), B,
                                 cudaMemcpyDeviceToDevice, h->stream));
            return 0;
        };
        if ((rc = move(s->s.log_cilx, 7)) || (rc = move(s->s.log_mpcx, 8)) || (rc = move(s->s.log_distx, 7))) return rc;
    }
    s->steps = 0;
    return 0;
}

int b200nmpc_sim_steps_done(const b200nmpc_sim *s) { return s ? s->steps : B200NMPC_E_ARG; }

int b200nmpc_sim_get_logs(b200nmpc_sim *s, double *MPC_SimX, double *CiLX, double *simU, double *simREF, double *simSolverDebug)
{
    if (!s) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    const size_t B = h->B, cap = s->s.cap;
    const size_t n = (size_t)s->steps < cap ? (size_t)s->steps : cap;
    auto pull = [&](double *dst, const double *src, size_t rows_have, size_t rows_cap, size_t w) -> int {
        if (!dst || rows_have == 0) return 0;
        CK(cudaMemcpy2DAsync(dst, rows_have * w * sizeof(double), src, rows_cap * w * sizeof(double), rows_have * w * sizeof(double),
                             B, cudaMemcpyDeviceToHost, h->stream));
        return 0;
    };
    if ((rc = pull(MPC_SimX, s->s.log_mpcx, n + 1, cap + 1, 8))) return rc;
    if ((rc = pull(CiLX, s->s.log_cilx, n + 1, cap + 1, 7))) return rc;
    if ((rc = pull(simU, s->s.log_u, n, cap, 2))) return rc;
    if ((rc = pull(simREF, s->s.log_ref, n, cap, 4))) return rc;
    if ((rc = pull(simSolverDebug, s->s.log_dbg, n, cap, 5))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    if (simSolverDebug && (rc = sim_collect_time(s))) return rc;
    if (simSolverDebug && s->run_steps > 0) {      // time_tot: mean device time per closed-loop step and instance
        const double t = s->run_seconds / ((double)s->run_steps * (double)B);
        for (size_t i = 0; i < B * n; i++) simSolverDebug[i * 5 + 1] = t;
    }
    return 0;
}

int b200nmpc_sim_get_logs_range(b200nmpc_sim *s, int first, int count, double *MPC_SimX, double *CiLX, double *simU, double *simREF,
                                double *simSolverDebug, int *ref_idx)
{
    if (!s) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    const size_t B = h->B, cap = s->s.cap;
    const size_t n = (size_t)s->steps < cap ? (size_t)s->steps : cap;
    if (first < 0 || count < 0 || (size_t)first + (size_t)count > n) return set_err(B200NMPC_E_ARG, "log range outside the steps done");
    const size_t f = (size_t)first, c = (size_t)count;
    auto pull = [&](void *dst, const void *src, size_t rows, size_t rows_cap, size_t wbytes) -> int {
        if (!dst || rows == 0) return 0;
        CK(cudaMemcpy2DAsync(dst, rows * wbytes, (const char *)src + f * wbytes, rows_cap * wbytes, rows * wbytes, B,
                             cudaMemcpyDeviceToHost, h->stream));
        return 0;
    };
    if ((rc = pull(MPC_SimX, s->s.log_mpcx, c + 1, cap + 1, 8 * sizeof(double)))) return rc;
    if ((rc = pull(CiLX, s->s.log_cilx, c + 1, cap + 1, 7 * sizeof(double)))) return rc;
    if ((rc = pull(simU, s->s.log_u, c, cap, 2 * sizeof(double)))) return rc;
    if ((rc = pull(simREF, s->s.log_ref, c, cap, 4 * sizeof(double)))) return rc;
    if ((rc = pull(simSolverDebug, s->s.log_dbg, c, cap, 5 * sizeof(double)))) return rc;
    if ((rc = pull(ref_idx, s->s.log_idx, c, cap, sizeof(int)))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    if (simSolverDebug && (rc = sim_collect_time(s))) return rc;
    if (simSolverDebug && s->run_steps > 0) {
        const double t = s->run_seconds / ((double)s->run_steps * (double)B);
        for (size_t i = 0; i < B * c; i++) simSolverDebug[i * 5 + 1] = t;
    }
    return 0;
}

int b200nmpc_sim_set_stop(b200nmpc_sim *s, int enable, const int *end_idx, double max_lat_dev, double max_a_comb)
{
    if (!s) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    if (enable && end_idx) {
        CK(cudaMemcpyAsync(s->end_idx, end_idx, (size_t)h->B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    s->s.stop_on = enable ? 1 : 0;
    s->s.end_idx = (enable && end_idx) ? s->end_idx : nullptr;
    s->s.max_lat_dev = max_lat_dev;
    s->s.max_a_comb = max_a_comb;
    return 0;
}

int b200nmpc_sim_get_stop(b200nmpc_sim *s, int *stop_step, int *flags)
{
    if (!s) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    if (stop_step) CK(cudaMemcpyAsync(stop_step, s->s.stop_step, (size_t)h->B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (flags) CK(cudaMemcpyAsync(flags, s->s.stop_flags, (size_t)h->B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_sim_get_ref_idx(b200nmpc_sim *s, int *out)
{
    if (!s || !out) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    const size_t B = h->B, cap = s->s.cap;
    const size_t n = (size_t)s->steps < cap ? (size_t)s->steps : cap;
    if (n == 0) return 0;
    CK(cudaMemcpy2DAsync(out, n * sizeof(int), s->s.log_idx, cap * sizeof(int), n * sizeof(int), B, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_sim_set_disturbances(b200nmpc_sim *s, const double *deriv, const double *est)
{
    if (!s) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    const size_t n = (size_t)h->B * 7 * sizeof(double);
    if (deriv) CK(cudaMemcpyAsync(s->dist_deriv, deriv, n, cudaMemcpyHostToDevice, h->stream));
    if (est) CK(cudaMemcpyAsync(s->dist_est, est, n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    s->s.dist_deriv = deriv ? s->dist_deriv : nullptr;
    s->s.dist_est = est ? s->dist_est : nullptr;
    return 0;
}

int b200nmpc_sim_config_disturbances(b200nmpc_sim *s, int type_deriv, const double *bounds_deriv, int type_est,
                                     const double *bounds_est, unsigned long long seed)
{
    if (!s) return set_err(B200NMPC_E_ARG, "null argument");
    const int types[2] = {type_deriv, type_est};
    const double *bnds[2] = {bounds_deriv, bounds_est};
    for (int ch = 0; ch < 2; ch++) {
        if (types[ch] < 0 || types[ch] > 3) return set_err(B200NMPC_E_ARG, "disturbance type must be 0 (off) .. 3 (absolute)");
        if (types[ch] != 0 && !bnds[ch]) return set_err(B200NMPC_E_ARG, "disturbance bounds missing");
        s->s.gen_type[ch] = types[ch];
        for (int i = 0; i < 7; i++) {
            const double v = types[ch] ? bnds[ch][i] : 0.0;
            if (types[ch] == 1 && v < 0.0) return set_err(B200NMPC_E_ARG, "ellipsoid bounds must be non-negative");
            s->s.gen_bound[ch][i] = v;
        }
    }
    s->s.seed = seed;
    return 0;
}

int b200nmpc_sim_set_playback(b200nmpc_sim *s, const double *deriv, const double *est, int n)
{
    if (!s || n < 1) return set_err(B200NMPC_E_ARG, "bad argument");
    if (n > s->s.cap) return set_err(B200NMPC_E_ARG, "sim_set_playback: %d rows exceed the log capacity %d", n, s->s.cap);
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    const double *src[2] = {deriv, est};
    const size_t cnt = (size_t)h->B * n * 7;
    for (int ch = 0; ch < 2; ch++) {
        if (!src[ch]) { if (s->s.gen_type[ch] == 4) s->s.gen_type[ch] = 0; continue; }
        double *d = s->play[ch];
        if (!d || s->play_cap[ch] < cnt) {         // grow only: repeated playback set-ups (one per evaluated run) reuse the buffer
            if (d) {
                CK(cudaStreamSynchronize(h->stream));
                for (void *&a : s->allocs) if (a == (void *)d) { cudaFree(a); a = nullptr; }
            }
            d = nullptr;
            if ((rc = sim_alloc(s, &d, cnt))) return rc;
            s->play_cap[ch] = cnt;
        }
        CK(cudaMemcpyAsync(d, src[ch], cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        s->play[ch] = d; s->s.play[ch] = d; s->s.gen_type[ch] = 4;
    }
    s->s.play_n = n;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_sim_get_disturbance_logs(b200nmpc_sim *s, double *DisturbedX, double *deriv, double *est)
{
    if (!s) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    const size_t B = h->B, cap = s->s.cap;
    const size_t n = (size_t)s->steps < cap ? (size_t)s->steps : cap;
    if (DisturbedX)
        CK(cudaMemcpy2DAsync(DisturbedX, (n + 1) * 7 * sizeof(double), s->s.log_distx, (cap + 1) * 7 * sizeof(double),
                             (n + 1) * 7 * sizeof(double), B, cudaMemcpyDeviceToHost, h->stream));
    double *dst[2] = {deriv, est};
    for (int ch = 0; ch < 2; ch++)
        if (dst[ch] && n)
            CK(cudaMemcpy2DAsync(dst[ch], n * 7 * sizeof(double), s->s.log_dist[ch], cap * 7 * sizeof(double), n * 7 * sizeof(double), B,
                                 cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int b200nmpc_sim_plan(b200nmpc_sim *s, const double *pose, int *closest)
{
    if (!s || !pose) return set_err(B200NMPC_E_ARG, "null argument");
    b200nmpc_handle *h = s->h;
    int rc = use_device(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s->s.pose, pose, (size_t)h->B * 4 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    SimMem tmp = s->s;
    tmp.log_ref = nullptr; tmp.log_idx = nullptr;
    launch_planner(tmp, h->m.yref, h->B, h->N, 0, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    if (closest) CK(cudaMemcpyAsync(closest, s->s.ref_idx, (size_t)h->B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
