// k_sim.cu -- the closed loop around the solver, on the device: planner emulator,
// plant step, state-estimation filter and per-step logging, so that B closed-loop simulations advance without
// touching the host between solves.
//
// Reference being replaced (paths relative to /root/reference):
//   planner window    Utils/MPC_sim_utils.py:137-194 (PlannerEmulator), :126-135 (postprocess_yaw)
//   plant             Vehicle_Simulator/VehicleSimulator.py:60-94 (RK4, 4 finite elements over Ts),
//                     Vehicle_Simulator/sim_model_dynamic_stm_pacejka.py:137-196 (7-state STM, inputs a, omega_f)
//   loop glue         main.py:48-78, Utils/SimulationMode_main_class.py:106-156 (sim_step, StateEstimation)
//   log arrays        Utils/Logging_Plotting.py:124-145 (MPC_SimX, CiLX, simU, simREF, simSolverDebug)
#include "nmpc_sim.cuh"

namespace b200 {

namespace {

constexpr int PW = 4;      // planner warps (instances) per block
#define FULLMASK 0xffffffffu

// numpy float modulo (result takes the sign of the divisor)
__device__ __forceinline__ double np_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((b < 0.0) != (r < 0.0))) r += b;
    return r;
}

// np.interp on the index grid xp = 0..L-1 without FMA contraction (bitwise the host arithmetic)
__device__ __forceinline__ double interp_idx(const double *vals, int L, double xq)
{
    int j = (int)floor(xq);
    if (j >= L - 1) return vals[L - 1];
    if (j < 0) return vals[0];
    const double slope = __ddiv_rn(__dsub_rn(vals[j + 1], vals[j]), 1.0);
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(xq, (double)j)), vals[j]);
}

// ---- counter-based random numbers: Philox4x32-10 (Salmon et al., SC'11), one call = 4 x 32 random bits
__device__ __forceinline__ void philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned *out)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// two uniforms in (0, 1) from 4 x 32 bits (53-bit mantissas, never exactly 0)
__device__ __forceinline__ void uniform2(const unsigned *r, double &u0, double &u1)
{
    const double scale = 1.0 / 9007199254740992.0;      // 2^-53
    u0 = ((double)((((unsigned long long)r[0] << 32) | r[1]) >> 11) + 0.5) * scale;
    u1 = ((double)((((unsigned long long)r[2] << 32) | r[3]) >> 11) + 0.5) * scale;
}
// 7 standard normals + 1 uniform for (instance, step, channel): 4 Philox calls, Box-Muller
__device__ __forceinline__ void normals7(unsigned long long seed, int b, int step, int channel, double *z, double &u)
{
    double g[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        unsigned r[4];
        philox4x32((unsigned)b, (unsigned)step, (unsigned)(channel * 4 + j), 0x5eedu, (unsigned)seed, (unsigned)(seed >> 32), r);
        double u0, u1;
        uniform2(r, u0, u1);
        if (j < 4) {
            const double rad = sqrt(-2.0 * log(u0));
            double sn, cs;
            sincospi(2.0 * u1, &sn, &cs);
            g[2 * j] = rad * cs; g[2 * j + 1] = rad * sn;
        }
    }
#pragma unroll
    for (int i = 0; i < 7; i++) z[i] = g[i];
    // the 8th normal is mapped back to a uniform for the radial position of the ellipsoid sample
    u = 0.5 * erfc(-g[7] * 0.70710678118654752440);
}
// one disturbance row of closed-loop step `step` (generate_disturbances, Utils/MPC_sim_utils.py:54-84):
//   uniform  -> sampleFromEllipsoid(0, diag(bounds)): x ~ N(0,I) / |x| * U^(1/7), y_i = sqrt(bound_i) x_i  (:68-84)
//   gaussian -> N(0, bound_i^2) per component (:60-62);   absolute -> the bounds themselves (:63-64)
__device__ __forceinline__ bool disturbance_row(const SimMem &s, int ch, int b, int step, double *w)
{
    const int type = s.gen_type[ch];
    if (type == 0) return false;
    if (type == 4) {
        if (step >= s.play_n || !s.play[ch]) return false;
        const double *p = s.play[ch] + ((size_t)b * s.play_n + step) * 7;
#pragma unroll
        for (int i = 0; i < 7; i++) w[i] = p[i];
        return true;
    }
    if (type == 3) {
#pragma unroll
        for (int i = 0; i < 7; i++) w[i] = s.gen_bound[ch][i];
        return true;
    }
    double z[7], u;
    normals7(s.seed, b, step, ch, z, u);
    if (type == 2) {
#pragma unroll
        for (int i = 0; i < 7; i++) w[i] = s.gen_bound[ch][i] * z[i];
        return true;
    }
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < 7; i++) nrm += z[i] * z[i];
    const double sc = pow(u, 1.0 / 7.0) / sqrt(nrm);
#pragma unroll
    for (int i = 0; i < 7; i++) w[i] = sqrt(s.gen_bound[ch][i]) * z[i] * sc;
    return true;
}

}  // namespace

// One warp per instance.  Nearest reference point (warp arg-min), time-horizon window walk, resampling of every
// key to N+1 points uniformly in index space, yaw unwrapped first when the window straddles 0/2pi.  Writes
// yref rows [x_ref, y_ref, yaw_ref, v_ref, 0, 0] (NMPC_class.py:169-180) straight into the solver's buffer.
__global__ void __launch_bounds__(PW * 32) k_planner(const SimMem s, double *yref, int B, int N, int step)
{
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * PW + warp;
    if (b >= B) return;
    if (s.stop_on && s.stop_step[b] >= 0) return;      // stopped instance (whole warp)
    // per-warp scratch: window values of the four keys [4][maxw] and the window indices
    double *vals = smem + (size_t)warp * (4 * s.maxw + (s.maxw + 1) / 2 * 1);
    int *widx = reinterpret_cast<int *>(vals + 4 * s.maxw);
    const double x = s.pose[(size_t)b * 4], y = s.pose[(size_t)b * 4 + 1];

    // ---- 1. closest point (first index on ties, like np.argmin)
    double best = 1e300;
    int bi = 0x7fffffff;
    for (int i = lane; i < s.n_ref; i += 32) {
        const double d = hypot(s.ref_px[i] - x, s.ref_py[i] - y);
        if (d < best) { best = d; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(FULLMASK, best, o);
        const int oi = __shfl_xor_sync(FULLMASK, bi, o);
        if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
    }
    // ---- 2. indices inside the time horizon
    int L = 0;
    if (lane == 0) {
        int cur = bi;
        double T = 0.0;
        widx[0] = cur; L = 1;
        while (T <= s.Tp && L < s.maxw) {
            T += s.seg_dt[cur];
            cur = (cur + 1 < s.n_ref) ? cur + 1 : 0;
            widx[L++] = cur;
        }
    }
    L = __shfl_sync(FULLMASK, L, 0);
    __syncwarp();
    for (int i = lane; i < L; i += 32) {
        const int r = widx[i];
        vals[i] = s.ref_px[r]; vals[s.maxw + i] = s.ref_py[r]; vals[2 * s.maxw + i] = s.ref_yaw[r]; vals[3 * s.maxw + i] = s.ref_v[r];
    }
    __syncwarp();
    const int npts = N + 1;
    double *yr = yref + (size_t)b * npts * 6;
    const bool resample = (L != npts);
    bool unwrap = false;
    if (resample) {
        const double *yw = vals + 2 * s.maxw;
        bool jump = false;
        for (int i = lane; i + 1 < L; i += 32) jump |= fabs(yw[i + 1] - yw[i]) > s.jump_thresh;
        unwrap = __any_sync(FULLMASK, jump);
        if (unwrap) {               // np.unwrap(fp, period=2pi), sequential cumulative correction
            __syncwarp();
            if (lane == 0) {
                double *w = vals + 2 * s.maxw;
                const double two_pi = 6.283185307179586476925286766559, pi = 3.14159265358979323846;
                double corr = 0.0, prev = w[0];
                for (int i = 1; i < L; i++) {
                    const double cur = w[i];
                    const double dd = cur - prev;
                    double ddmod = np_mod(dd + pi, two_pi) - pi;
                    if (ddmod == -pi && dd > 0.0) ddmod = pi;
                    double pc = ddmod - dd;
                    if (fabs(dd) < pi) pc = 0.0;
                    corr += pc;
                    w[i] = cur + corr;
                    prev = cur;
                }
            }
            __syncwarp();
        }
    }
    // ---- 3. resample (np.linspace(0, L-1, N+1) / np.interp) and emit yref
    const double stepq = __ddiv_rn((double)(L - 1), (double)(npts - 1));
    for (int j = lane; j < npts; j += 32) {
        double o0, o1, o2, o3;
        if (resample) {
            const double xq = (j == npts - 1) ? (double)(L - 1) : __dmul_rn((double)j, stepq);
            o0 = interp_idx(vals, L, xq);
            o1 = interp_idx(vals + s.maxw, L, xq);
            o2 = interp_idx(vals + 2 * s.maxw, L, xq);
            if (unwrap) o2 = np_mod(o2, 6.283185307179586476925286766559);
            o3 = interp_idx(vals + 3 * s.maxw, L, xq);
        } else {
            o0 = vals[j]; o1 = vals[s.maxw + j]; o2 = vals[2 * s.maxw + j]; o3 = vals[3 * s.maxw + j];
        }
        yr[j * 6 + 0] = o0; yr[j * 6 + 1] = o1; yr[j * 6 + 2] = o2; yr[j * 6 + 3] = o3; yr[j * 6 + 4] = 0.0; yr[j * 6 + 5] = 0.0;
        if (j == 0 && s.log_ref && step < s.cap) {
            double *lr = s.log_ref + ((size_t)b * s.cap + step) * 4;
            lr[0] = o0; lr[1] = o1; lr[2] = o2; lr[3] = o3;
        }
    }
    if (lane == 0 && s.ref_idx) s.ref_idx[b] = bi;
    if (lane == 0 && s.log_idx && s.log_ref && step < s.cap) s.log_idx[(size_t)b * s.cap + step] = bi;
}

int planner_smem_bytes(int maxw) { return PW * (4 * maxw + (maxw + 1) / 2) * (int)sizeof(double); }

void launch_planner(const SimMem &s, double *yref, int B, int N, int step, cudaStream_t st)
{
    k_planner<<<(B + PW - 1) / PW, PW * 32, planner_smem_bytes(s.maxw), st>>>(s, yref, B, N, step);
}

// RK4 with nsub finite elements over Ts of the 7-state plant (VehicleSimulator.py:60-94; disturbed variant :79-94: the
// additive disturbance w enters every stage derivative)
__device__ __forceinline__ void plant_rk4(const DevCfg &cfg, double *X, double a_in, double u1, const double *wd, double Ts, int nsub)
{
    const double h = Ts / nsub;
    for (int sub = 0; sub < nsub; sub++) {
        double acc[7], kk[7], xe[8];
#pragma unroll
        for (int st = 0; st < 4; st++) {
            const double c = (st == 0) ? 0.0 : (st == 3) ? h : 0.5 * h;
#pragma unroll
            for (int i = 0; i < 7; i++) xe[i] = (st == 0) ? X[i] : X[i] + c * kk[i];
            xe[7] = a_in;
            double f[6];
            stm_eval<false>(cfg, xe, f, nullptr);
#pragma unroll
            for (int i = 0; i < 6; i++) kk[i] = f[i];
            kk[6] = u1;
            if (wd) {
#pragma unroll
                for (int i = 0; i < 7; i++) kk[i] += wd[i];
            }
            const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
#pragma unroll
            for (int i = 0; i < 7; i++) acc[i] = (st == 0) ? kk[i] : acc[i] + wgt * kk[i];
        }
#pragma unroll
        for (int i = 0; i < 7; i++) X[i] += h / 6.0 * acc[i];
    }
}

// One thread per instance, after the solve of closed-loop step `step` (main.py:58-78):
//   u0, x_1 of the new iterate (kept from the last success otherwise, NMPC_class.py:195-201)
//   plant RK4 with inputs [a = x_1[7], omega_f = u0[1]] (SimulationMode_main_class.py:113-118)
//   x_next = [plant state (+ estimation noise), a] -> moving-average filter -> next x0 (:144-156)
//   failed solve: iterate re-initialised to the current x0 (main.py:59-61)
//   logs: simU[i], simSolverDebug[i], CiLX[i+1], MPC_SimX[i+1]; pose for the next planner call from the
//   undisturbed plant state (:146-149).
__global__ void __launch_bounds__(64) k_sim_advance(const __grid_constant__ DevCfg cfg, const DevMem m, const SimMem s, int step)
{
    const int B = cfg.B, N = cfg.N, NS = N + 1;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (s.stop_on && s.stop_step[b] >= 0) return;      // stopped instance: state and logs stay as they were
    const int status = m.status[b];
    double *Xb = m.X + (size_t)b * NS * 8, *Ub = m.U + (size_t)b * N * 2;
    const double u0 = Ub[0], u1 = Ub[1];
    double *p1 = s.pred_x1 + (size_t)b * 8;
    if (status == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) p1[i] = Xb[8 + i];
    }
    const double a_in = p1[7];
    // ---- plant: 7 states [x, y, yaw, vl, vt, r, delta], inputs (a, omega_f), nsub RK4 finite elements.
    // With a state-derivative disturbance the reference integrates twice from the same state
    // (SimulationMode_main_class.py:119-131): the undisturbed result continues as the plant state and pose
    // (Logger.CiLX, main.py:69), the disturbed one only feeds the state estimate of the controller (DisturbedX).
    double X[7], Xd[7];
#pragma unroll
    for (int i = 0; i < 7; i++) X[i] = Xd[i] = s.cilx[(size_t)b * 7 + i];
    // ---- stop conditions on the quantities Logger.logging_step records for this step: state before the step
    // against the reference point of the step (Logging_Plotting.py:147-179), objective_function.py:153,186-200
    if (s.stop_on && step < s.cap) {
        const double *lr = s.log_ref + ((size_t)b * s.cap + step) * 4;
        const double cy = cos(-X[2]), sy = sin(-X[2]);
        const double lat = sy * (lr[0] - X[0]) + cy * (lr[1] - X[1]);
        const double v = X[3], alat = v * X[5];
        double ax, ay, d_unused;
        interp_lin(cfg.n_ggv, cfg.ggv_v, cfg.ggv_ax, v, ax, d_unused);
        interp_lin(cfg.n_ggv, cfg.ggv_v, cfg.ggv_ay, v, ay, d_unused);
        const double alon = a_in > 0.0 ? a_in / ax : fabs(a_in) / cfg.ax_brake;
        const double acomb = sqrt(alon * alon + (alat / ay) * (alat / ay));
        const int done = (s.end_idx && s.ref_idx[b] == s.end_idx[b]) ? 1 : 0;
        const int crash = (lat > s.max_lat_dev || acomb > s.max_a_comb) ? 2 : 0;
        if (done | crash) { s.stop_step[b] = step + 1; s.stop_flags[b] = done | crash; }
    }
    double wgen[7];
    const double *wd = disturbance_row(s, 0, b, step, wgen) ? wgen : (s.dist_deriv ? s.dist_deriv + (size_t)b * 7 : nullptr);
    plant_rk4(cfg, X, a_in, u1, nullptr, s.Ts, s.nsub);
    if (wd) plant_rk4(cfg, Xd, a_in, u1, wd, s.Ts, s.nsub);
    else {
#pragma unroll
        for (int i = 0; i < 7; i++) Xd[i] = X[i];
    }
#pragma unroll
    for (int i = 0; i < 7; i++) s.cilx[(size_t)b * 7 + i] = X[i];
    s.pose[(size_t)b * 4 + 0] = X[0]; s.pose[(size_t)b * 4 + 1] = X[1]; s.pose[(size_t)b * 4 + 2] = X[2];
    s.pose[(size_t)b * 4 + 3] = sqrt(X[3] * X[3] + X[4] * X[4]);
    // ---- logs
    if (step < s.cap) {
        double *lu = s.log_u + ((size_t)b * s.cap + step) * 2;
        lu[0] = u0; lu[1] = u1;
        double *ld = s.log_dbg + ((size_t)b * s.cap + step) * 5;
        ld[0] = m.cost[b]; ld[1] = 0.0; ld[2] = 1.0; ld[3] = (double)m.qp_iter[b]; ld[4] = (double)status;
        double *lc = s.log_cilx + ((size_t)b * (s.cap + 1) + step + 1) * 7;
#pragma unroll
        for (int i = 0; i < 7; i++) lc[i] = X[i];
        double *lm = s.log_mpcx + ((size_t)b * (s.cap + 1) + step + 1) * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) lm[i] = p1[i];
    }
    // ---- failed solve: fresh solver at the current initial state (main.py:59-61)
    double *x0 = m.x0 + (size_t)b * 8;
    if (status != 0) {
        for (int k = 0; k <= N; k++) {
#pragma unroll
            for (int i = 0; i < 8; i++) Xb[k * 8 + i] = x0[i];
        }
        for (int k = 0; k < 2 * N; k++) Ub[k] = 0.0;
    }
    // ---- state estimation: per-state truncated moving average over the last w samples
    double egen[7];
    const double *we = disturbance_row(s, 1, b, step, egen) ? egen : (s.dist_est ? s.dist_est + (size_t)b * 7 : nullptr);
    if (step < s.cap && s.log_dist[0]) {
        double *l0 = s.log_dist[0] + ((size_t)b * s.cap + step) * 7, *l1 = s.log_dist[1] + ((size_t)b * s.cap + step) * 7;
#pragma unroll
        for (int i = 0; i < 7; i++) { l0[i] = wd ? wd[i] : 0.0; l1[i] = we ? we[i] : 0.0; }
    }
    double xn[8];
#pragma unroll
    for (int i = 0; i < 7; i++) xn[i] = Xd[i] + (we ? we[i] : 0.0);
    xn[7] = a_in;
    if (step < s.cap && s.log_distx) {
        double *lx = s.log_distx + ((size_t)b * (s.cap + 1) + step + 1) * 7;
#pragma unroll
        for (int i = 0; i < 7; i++) lx[i] = xn[i];
    }
    const int cnt = s.fcount[b];
    double *fb = s.fbuf + (size_t)b * 32;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        fb[i * 4 + (cnt & 3)] = xn[i];
        const int w = s.win[i];
        const int have = (cnt + 1 < w) ? cnt + 1 : w;
        double sum = 0.0;
        for (int q = have - 1; q >= 0; q--) sum += fb[i * 4 + ((cnt - q) & 3)];      // oldest first, like np.mean
        x0[i] = sum / (double)have;
    }
    s.fcount[b] = cnt + 1;
}

void launch_sim_advance(const DevCfg &cfg, const DevMem &m, const SimMem &s, int step, cudaStream_t st)
{
    k_sim_advance<<<(cfg.B + 63) / 64, 64, 0, st>>>(cfg, m, s, step);
}

// start of a run: plant state, pose, logs row 0, empty filter buffers (SimulationMode_main_class.py:44-66,100-104)
__global__ void k_sim_reset(const SimMem s, const double *x0_mpc, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double *x = x0_mpc + (size_t)b * 8;
    for (int i = 0; i < 7; i++) s.cilx[(size_t)b * 7 + i] = x[i];
    s.pose[(size_t)b * 4 + 0] = x[0]; s.pose[(size_t)b * 4 + 1] = x[1]; s.pose[(size_t)b * 4 + 2] = x[2];
    s.pose[(size_t)b * 4 + 3] = sqrt(x[3] * x[3] + x[4] * x[4]);
    for (int i = 0; i < 8; i++) s.pred_x1[(size_t)b * 8 + i] = x[i];
    for (int i = 0; i < 32; i++) s.fbuf[(size_t)b * 32 + i] = 0.0;
    s.fcount[b] = 0;
    s.stop_step[b] = -1; s.stop_flags[b] = 0;
    for (int i = 0; i < 7; i++) s.log_cilx[(size_t)b * (s.cap + 1) * 7 + i] = x[i];
    if (s.log_distx) for (int i = 0; i < 7; i++) s.log_distx[(size_t)b * (s.cap + 1) * 7 + i] = x[i];
    for (int i = 0; i < 8; i++) s.log_mpcx[(size_t)b * (s.cap + 1) * 8 + i] = x[i];
}

void launch_sim_reset(const SimMem &s, const double *x0_mpc_dev, int B, cudaStream_t st)
{
    k_sim_reset<<<(B + 127) / 128, 128, 0, st>>>(s, x0_mpc_dev, B);
}

// Per-instance reset while the batch keeps running (one environment of a vectorised RL run ends its episode,
// RL_WMPC/environment.py reset()): plant, pose, filter and stop state as in k_sim_reset, the controller as in
// MPC.reset(X0_MPC) (NMPC_class.py:243-254: x0 and the initial guess x_k = x0, u = 0); the instance's log rows go on at
// the batch's current step.
__global__ void k_sim_reset_masked(const __grid_constant__ DevCfg cfg, const DevMem m, const SimMem s, const double *x0_mpc,
                                   const int *mask, int step)
{
    const int B = cfg.B, N = cfg.N, NS = N + 1;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || mask[b] == 0) return;
    const double *x = x0_mpc + (size_t)b * 8;
    for (int i = 0; i < 7; i++) s.cilx[(size_t)b * 7 + i] = x[i];
    s.pose[(size_t)b * 4 + 0] = x[0]; s.pose[(size_t)b * 4 + 1] = x[1]; s.pose[(size_t)b * 4 + 2] = x[2];
    s.pose[(size_t)b * 4 + 3] = sqrt(x[3] * x[3] + x[4] * x[4]);
    for (int i = 0; i < 8; i++) s.pred_x1[(size_t)b * 8 + i] = x[i];
    for (int i = 0; i < 32; i++) s.fbuf[(size_t)b * 32 + i] = 0.0;
    s.fcount[b] = 0;
    s.stop_step[b] = -1; s.stop_flags[b] = 0;
    if (step <= s.cap) {
        for (int i = 0; i < 7; i++) s.log_cilx[((size_t)b * (s.cap + 1) + step) * 7 + i] = x[i];
        if (s.log_distx) for (int i = 0; i < 7; i++) s.log_distx[((size_t)b * (s.cap + 1) + step) * 7 + i] = x[i];
        for (int i = 0; i < 8; i++) s.log_mpcx[((size_t)b * (s.cap + 1) + step) * 8 + i] = x[i];
    }
    double *Xb = m.X + (size_t)b * NS * 8, *Ub = m.U + (size_t)b * N * 2;
    for (int i = 0; i < 8; i++) m.x0[(size_t)b * 8 + i] = x[i];
    for (int k = 0; k <= N; k++)
        for (int i = 0; i < 8; i++) Xb[k * 8 + i] = x[i];
    for (int k = 0; k < 2 * N; k++) Ub[k] = 0.0;
    m.status[b] = 0;
}
void launch_sim_reset_masked(const DevCfg &cfg, const DevMem &m, const SimMem &s, const double *x0_mpc_dev, const int *mask_dev,
                             int step, cudaStream_t st)
{
    k_sim_reset_masked<<<(cfg.B + 63) / 64, 64, 0, st>>>(cfg, m, s, x0_mpc_dev, mask_dev, step);
}

// Launch order of the next solve: instances sorted by the interior-point iterations they just needed, heaviest first
// (longest-processing-time-first: in a warm closed loop two thirds of the instances stop after phase 0 and the others
// run ~14 iterations; started in index order the late heavy ones run alone on a draining GPU).  The iteration count of
// the previous step is the predictor; results do not depend on the order.  One block: shared-memory histogram over
// 64 keys, descending offsets, scatter (the order inside a key is whatever the atomics give).
// Measured alternatives (profiles/r02_ab_launch_order_classes.jsonl): a STABLE order with the histogram cut into 2 / 3 / 4 / 8
// coarse classes, or stable with all 64 keys, is within +1.3 % / -3.6 % of this one depending on the horizon and no better
// overall.  Any longest-first order doubles the DRAM traffic of the QP kernel against index order (0.79 -> 1.6 GB per launch,
// L2 hit rate 88 -> 79 %), the coarse stable ones included: not because the resident set is scattered, but because the SMs
// then stay full to the end -- with 2 368 resident instances the working set (115 MB) sits at the capacity of the 126 MB L2,
// while index order ends in a long under-filled tail whose few instances hit in L2 all the time.
__global__ void __launch_bounds__(1024) k_launch_order(const int *qp_iter, const int *halt, int *order, int B)
{
    __shared__ int cnt[64], cur[64];
    const int t = threadIdx.x;
    if (t < 64) cnt[t] = 0;
    __syncthreads();
    for (int b = t; b < B; b += blockDim.x) {
        const int key = (halt && halt[b] >= 0) ? 0 : min(max(qp_iter[b], 0), 63);
        atomicAdd(&cnt[key], 1);
    }
    __syncthreads();
    if (t == 0) {
        int off = 0;
        for (int k = 63; k >= 0; k--) { cur[k] = off; off += cnt[k]; }
    }
    __syncthreads();
    for (int b = t; b < B; b += blockDim.x) {
        const int key = (halt && halt[b] >= 0) ? 0 : min(max(qp_iter[b], 0), 63);
        order[atomicAdd(&cur[key], 1)] = b;
    }
}
void launch_launch_order(const DevMem &m, int *order, int B, cudaStream_t st)
{
    k_launch_order<<<1, 1024, 0, st>>>(m.qp_iter, m.halt, order, B);
}

// SNMPC: sample iterates of the masked instances = their x0 sample matrix (set(j, "x", x0_samples), SNMPC_class.py:266-272)
__global__ void k_snmpc_init_samples_masked(const __grid_constant__ DevCfg cfg, const DevMem m, const int *mask)
{
    const int ns = cfg.n_samples, HS = m.HA + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)cfg.B * HS * ns * 8) return;
    const int b = (int)(g / ((long long)HS * ns * 8));
    if (mask[b] == 0) return;
    const int r = (int)(g - (long long)b * HS * ns * 8) % (ns * 8);
    m.XS[g] = m.x0s[(size_t)b * ns * 8 + r];
}
void launch_snmpc_init_samples_masked(const DevCfg &cfg, const DevMem &m, const int *mask_dev, cudaStream_t st)
{
    const long long items = (long long)cfg.B * (m.HA + 1) * cfg.n_samples * 8;
    if (items == 0) return;
    k_snmpc_init_samples_masked<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(cfg, m, mask_dev);
}

// SNMPC: instances whose solve failed restart their sample iterates from the x0 sample matrix
__global__ void k_snmpc_reinit_failed(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int ns = cfg.n_samples, HS = m.HA + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)cfg.B * HS * ns * 8) return;
    const int b = (int)(g / ((long long)HS * ns * 8));
    if (m.status[b] == 0 || (m.halt && m.halt[b] >= 0)) return;
    const int r = (int)(g - (long long)b * HS * ns * 8) % (ns * 8);
    m.XS[g] = m.x0s[(size_t)b * ns * 8 + r];
}
void launch_snmpc_reinit_failed(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    const long long items = (long long)cfg.B * (m.HA + 1) * cfg.n_samples * 8;
    if (items == 0) return;
    k_snmpc_reinit_failed<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(cfg, m);
}

}  // namespace b200
