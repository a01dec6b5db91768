// k_qp_warp.cu -- K2 + K3: per-instance structured QP solve and full step, ONE WARP PER INSTANCE.
//
// Replaces acados' feedback phase for this OCP: full condensing + HPIPM dense
// interior point + `x += dx, u += du` + get_cost().  The QP optimum is unique (R > 0, Z > 0), so the
// method is free: a Riccati recursion over the N stages inside a Mehrotra predictor-corrector primal-dual
// interior point iteration started from the equality-constrained optimum (the algorithm restated in
// oracle/nmpc_oracle.c, orc_qp_solve; tests compare the two).
//
// Mapping to the machine (profiles/r01_* hold the ncu evidence for every step; DESIGN.md section 4 the table):
//   v1  one thread per instance: < 1 warp per SM at B = 4096, 255 registers + spills, IPC 0.09.
//   v2  eight lanes per instance, four instances per warp: 37 % address arithmetic, coupled iteration counts.
//   v3  one warp per instance, instance-major records, octet recursions (lane j owns row j of P), exchanges through
//       shared memory instead of shuffles, 128 registers / 16 warps per SM.
//   v4+ (this file):
//       * every sweep visits the stages in order and needs exactly ONE contiguous sub-range of the 992-byte stage record
//         (fields ordered by reader, nmpc_kernels.cuh; LIN part written by k_linearize, the rest private to this kernel);
//         the sub-range of the next stage is copied into the other of two shared-memory slots while the recursion works on
//         the current one (cp.async + wait_group by default, TMA bulk copy + mbarrier behind -DPIPE_TMA: measured 6-10 %
//         slower for a record consumed by one warp);
//       * the per-iteration full backward sweep and phase 0 run on ALL 32 lanes with ONE instruction stream
//         (sweep_backward_full: lane = row x column pair of the 8x8 products, per-lane column / vector descriptors);
//       * the vector sweeps (corrector, roll-outs, residual check) keep the octet decomposition: lane j owns component j,
//         the four octets compute redundantly, 8-vector gathers through 16-double shared-memory exchange slots;
//       * the interior-point "side" arithmetic (6 soft sides per stage) has no stage coupling and runs STAGE-PARALLEL over
//         all 32 lanes on flat per-side arrays with warp-shuffle reductions; three passes per iteration (predictor
//         directions, corrector directions, step + next residuals): the corrector right-hand side is a per-row linear
//         update of the predictor coefficients;
//       * per-stage quantities that connect the two kinds of phases (row values, barrier weights, right-hand-side
//         coefficients, direction row values, feed-forward terms, cost diagonal) live in shared memory;
//       * the early exit (equality-constrained optimum feasible) and the iteration count are warp-uniform:
//         no divergence, no tail coupling between instances;
//       * v10-v12: the side arrays are interleaved double2 pairs read one round ahead of their use, the soft-side
//         activity is a per-lane bitmask register, the primal step of an iteration is applied by the next backward sweep
//         from the shared-memory stage copy, A and B sit TRANSPOSED in the stage record (every Riccati product walks a
//         contiguous column), and the sweeps carry no divergent branch: lane-specific work is a select or a predicated
//         store, reciprocals and reciprocal square roots are a hardware seed + two Newton steps (the library sequences
//         bring a divergent slow path), the stage refill is a predicated cp.async.
//       * round 2 (v13, v14): the launch slot -> instance map comes from the previous solve's iteration counts (longest first);
//         the full backward sweep of the nominal kernel runs its three 8x8x8 products per stage on the FP64 TENSOR CORES
//         (sweep_backward_tensor, mma.sync.m8n8k4.f64: +11 %); the per-stage shared block shrank to 27 doubles; the stage
//         pipeline is generalised to groups of records per copy transaction (A/B builds: tools/build_variants.sh).
//       * round 2 (v15-v20): QP tolerance 1e-9; per-horizon layout of the per-stage shared block (LONGH); the stage
//         pipeline keeps its addresses loop-carried; ncu then showed the LOAD/STORE PIPE as the busiest unit (72 %), which
//         delivers 128 bytes per cycle into the registers per ACTIVE lane whether or not lanes read the same data: K ddx of
//         the forward roll-outs became a shuffle butterfly, the stage record was re-ordered so that a sweep stages only the
//         sub-range it reads, and the corrector sweep / forward roll-out of the throughput instantiation run on octet 0
//         only (Mem Busy 61 %, 2.44 -> 2.07 ms at B = 4096, N = 30).
//
// SNMPC (template parameter HEAD): stages 0..H-1 of the Stochastic NMPC OCP carry the sample states; they are
// eliminated onto the nh = 2H head inputs by snmpc_condense_warp (snmpc_head.cuh, the prologue of this kernel; the sample
// iterates take their step in its epilogue).  The stage recursions below then
// stop at stage H and the head is one dense nh x nh system per interior-point iteration:
//   M = HH + sum_rows om_r row_r row_r' + G_H' P_H G_H      (Cholesky in shared memory, lane a owns row a)
// solved for the head inputs, from which dx_k = G_k u~ + c_k (k <= H) and the head row values follow.  The
// stage-parallel interior-point phases are unchanged: head rows use the same (stage, row) slots.
#include <cstdlib>

#include "nmpc_kernels.cuh"
#include "snmpc_head.cuh"

namespace b200 {

#define FULLMASK 0xffffffffu

namespace {

// ---- per-warp shared memory map (doubles)
constexpr int SM_X = 0;         // 3 exchange slots of 16 doubles (8-vector gathers of the vector sweeps)
constexpr int SBUF = STG;            // one stage slot: the record of the stage being visited (or the sub-range a sweep reads)
#ifndef PIPE_NSLOT
#define PIPE_NSLOT 2
#endif
#ifndef PIPE_GROUP
#define PIPE_GROUP 1
#endif
constexpr int NSLOT = PIPE_NSLOT;        // stage groups in flight: the copies run NSLOT - 1 groups ahead of the recursion
constexpr int PRING = NSLOT * PIPE_GROUP;   // stage records in the ring
constexpr int SM_BUF = 48;      // PRING record slots
constexpr int SM_MBAR = SM_BUF + PRING * SBUF;  // NSLOT mbarriers (8 bytes each), padded to an even count
// 32-lane Riccati sweep: P_{k+1} and P A as 8 x 8 tiles (row stride 10: 16-byte aligned pairs), small products, vectors
constexpr int TLD = 10;
constexpr int SM_PT = SM_MBAR + ((NSLOT + 1) & ~1);   // 80: P tile
constexpr int SM_PA = SM_PT + 80;               // 80: P A tile
constexpr int SM_SML = SM_PA + 80;              // 56: PB16 | Hux16 | hx8 | H4 | hu2 | tpi2 | pad8
constexpr int SM_PV = SM_SML + 56;              // 16: p8 | pi8
constexpr int SM_STAGE = SM_PV + 16;            // per-stage block starts here: (N+1) x PS_N
// per-stage block: rowv3, om3, cf3, ld3, dv3, w6 (sigma-scaled cost diagonal), kff2, gus2, ddu2 = 27 (odd stride: conflict-free);
// cs3, cp3 (sensitivities of the corrector right-hand side to the centring parameter / second-order term) live in the side block.
// LONG horizons (Ctx::longh, N > 38: 16 warps of 27 (N + 1) doubles no longer fit an SM): kff, gus, ddu -- sweep-to-sweep
// quantities, written by one sequential sweep and read by the next at the same stage -- travel in the WORK part of the stage
// record (W_KFF, W_GUS, W_DDU) and the block shrinks to 21 doubles per stage: N = 80 10 -> 12 resident warps per SM, +7.5 %;
// short horizons keep them in shared memory (the record round trip costs 1-3 % there; profiles/r02_ab_stage_block.jsonl)
constexpr int PS_ROWV = 0, PS_OM = 3, PS_CF = 6, PS_LD = 9, PS_DV = 12, PS_W = 15, PS_KFF = 21, PS_GUS = 23, PS_DDU = 25;
constexpr int PS_N_SHORT = 27, PS_N_LONG = 21;
static_assert((SM_BUF * 8) % 16 == 0 && (SBUF * 8) % 16 == 0, "bulk copies need 16-byte alignment");

// ---- TMA bulk copy (global -> shared, completion on an mbarrier) and the cross-proxy fence.  One elected lane arms
// the barrier with the byte count and issues the copies; every lane of the warp waits on the phase parity.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// generic-proxy writes (st.global of the WORK block) before async-proxy reads of the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// max that propagates NaN (fmax would drop it)
__device__ __forceinline__ double nmax(double a, double b) { return (a != a || b != b) ? nan("") : fmax(a, b); }
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULLMASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_nmax(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nmax(v, __shfl_xor_sync(FULLMASK, v, o));
    return v;
}

struct Ctx {
    int b, N, NS, lane32, lane;      // lane = lane32 & 7 (position in the octet)
    bool st;                         // this lane belongs to octet 0: the one that stores octet results
    double dt, reg, s0;
    const double *lin;               // LIN records of this instance
    double *wk;                      // WORK block of this instance
    double *side;                    // per-side arrays of this instance: side + arr*NS*6
    double *sm;                      // per-warp shared memory
    double *ps;                      // per-stage shared block
    mutable uint32_t phases;         // phase parity of the two stage-slot mbarriers (bit s = slot s)
    uint32_t actmask;                // bit r: this lane's item of round r (item r*32 + lane32) is an active soft side
    uint32_t sm32;                   // shared-window address of sm (mbarrier / bulk-copy operands)
    mutable const char *psrc;        // stage pipeline: this lane's first 16-byte chunk of the record the next refill copies
    uint32_t pdst;                   // stage pipeline: shared-window address of this lane's first chunk in record slot 0
    int psn;                         // doubles per stage of the per-stage shared block (PS_N_SHORT / PS_N_LONG)
    bool longh;                      // long-horizon layout (a compile-time constant of the kernel instantiation)
    unsigned vmask;                  // lanes that run the corrector sweep / forward roll-out: all, or octet 0 (see OCTET0)
    // SNMPC head (HEAD kernels only)
    int H, nh, ldm, hG, hC, hROW, hV0, hHH, hGH;
    const double *hd;                // condensed head block of this instance (global, read-only)
    double *hs;                      // per-warp shared head scratch: M | 1/diag | r_u | solution | P G
    __device__ __forceinline__ double *S(int arr) const { return side + (size_t)arr * NS * 6; }      // single arrays: S_BND, S_ACT
    __device__ __forceinline__ double2 *P(int pair) const { return reinterpret_cast<double2 *>(side + (size_t)pair * NS * 12); }
    __device__ __forceinline__ double &PS(int arr, int k, int c) const { return ps[k * psn + arr + c]; }
    __device__ __forceinline__ double *PSK(int k) const { return ps + k * psn; }     // stage block: psk[arr + c]
    // sweep-to-sweep pairs (kff, gus, ddu) of stage k: where a sequential sweep reads them (w = WORK part of the stage's
    // shared-memory record copy, psk = the stage block) and where it stores them (the record in HBM / the stage block)
    __device__ __forceinline__ const double *SWL(const double *w, const double *psk, int wo, int po) const { return longh ? w + wo : psk + po; }
    __device__ __forceinline__ double *SWS(int k, int wo, int po) const { return longh ? wk + (size_t)k * STG + wo : ps + k * psn + po; }
};

__device__ __forceinline__ double sval(const double *shared, const double *inst, const Ctx &c, int k, int ncomp, int comp)
{
    return inst ? inst[((size_t)c.b * c.NS + k) * ncomp + comp] : shared[k * ncomp + comp];
}

// Exchange of 8-vectors inside the warp through shared memory: lane j of every octet writes component j
// (the four octets hold identical values), then every lane reads the whole vector with 128-bit loads.
// This replaces 16 SHFL per gathered vector and every butterfly reduction of the second-generation kernel.
// SYNCED: the caller has passed a warp barrier since the last read of the slot (no second one here)
// MASK: the lanes that take part.  The corrector sweep and the forward roll-out of the nominal kernel run on octet 0 ONLY
// (OCTET0): the other three octets would repeat the same arithmetic, which costs nothing in issue slots but every
// shared-memory access occupies the load/store pipe -- the kernel's busiest unit -- per ACTIVE lane (128 bytes per cycle
// into the registers): a 128-bit gather is 4 pipe cycles for 32 lanes and 1 for 8.  B = 4096: 2.100 -> 2.066 ms
// (profiles/r02_ab_vector_sweeps_octet0.jsonl; -DVEC_ALL_OCTETS: the redundant form).  Ctx::vmask selects it per kernel
// instantiation: the throughput kernel only.  The small-batch instantiation is pure dependent latency, where the
// divergence costs more than the pipe cycles (B = 1 cold: 0.70 -> 0.73 ms), and the head stages of the SNMPC kernel read
// the recursion state of every lane.
#ifndef VEC_ALL_OCTETS
constexpr unsigned OCTET0 = 0xffu;
#else
constexpr unsigned OCTET0 = FULLMASK;
#endif
template <bool SYNCED = false>
__device__ __forceinline__ void xchg1(const Ctx &c, int slot, double a, double *A8, unsigned MASK = FULLMASK)
{
    double *buf = c.sm + SM_X + slot * 16;
    if (!SYNCED) __syncwarp(MASK);
    buf[c.lane] = a;
    __syncwarp(MASK);
    const double2 *b2 = reinterpret_cast<const double2 *>(buf);
#pragma unroll
    for (int i = 0; i < 4; i++) { const double2 v = b2[i]; A8[2 * i] = v.x; A8[2 * i + 1] = v.y; }
}
__device__ __forceinline__ double sum8(const double *v)
{
    return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}

// Branch-free FP64 reciprocal and reciprocal square root: hardware seed (relative error 2^-23) + two Newton steps, ~1 ulp.
// `1.0 / x` and `rsqrt(x)` compile to the same seed + refinement PLUS a divergent slow-path call for denormal / huge
// operands (BSSY / BRA / BSYNC around every use: 54 divisions per lane and iteration in the stage-parallel passes); the
// operands here are barrier variables and Cholesky pivots, a non-positive or non-finite one yields NaN / inf that the
// status checks catch.
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx, y * y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx, y * y, 0.5);
    return fma(y, e, y);
}

// FP64 tensor-core product D(8x8) += A(8x4) B(4x8) (DMMA): lane (g = lane / 4, t = lane % 4) holds A[g][t], B[t][g] and
// D[g][2t], D[g][2t+1].  Measured on B200 (tools/fp64_microbench.cu): 16 cycles issue interval per SM sub-partition = the
// rate of the vector pipe (8 DFMA warp instructions), 26 cycles dependent latency.
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Elimination of one soft side (derivation: oracle/nmpc_oracle.c above orc_qp_solve)
struct SideK { double th, iDp, rho, cc, om, itt, is; };
__device__ __forceinline__ SideK side_coeffs(double reg, double Z, double s, double t, double lam, double mu,
                                            double rt, double rs, double r1, double r2)
{
    SideK o;
    o.itt = fast_rcp(t + reg * lam);
    o.is = fast_rcp(s);
    o.th = lam * o.itt;
    const double ms = mu * o.is;
    o.iDp = fast_rcp(Z + o.th + ms);
    o.rho = -rs + r1 * o.itt + r2 * o.is + o.th * rt;
    o.cc = r1 * o.itt + o.th * rt - o.th * o.rho * o.iDp;
    o.om = o.th * (Z + ms) * o.iDp;
    return o;
}
// Stage pipeline: the LIN record and the WORK record of stage k are one contiguous stage record in HBM (960 bytes); the
// sweeps visit the stages in order, so the record of a later stage is copied into another shared-memory slot while the
// recursion works on the current one.  Two interchangeable copy engines, both measured on B200 (B = 4096, N = 30):
//   cp.async (default): every lane issues 16-byte cp.async.cg copies, one commit group per stage visit,
//       cp.async.wait_group + __syncwarp to consume                                   -> 1.10 M solves/s
//   TMA (-DPIPE_TMA): one elected lane arms an mbarrier and issues ONE cp.async.bulk of the whole record, the warp waits
//       on the phase parity; needs fence.proxy.async after this warp's st.global of the WORK part -> 1.03 M solves/s
// For a 960-byte record consumed by a single warp the bulk engine loses: mbarrier.try_wait costs 60-90 cycles even when
// the data has landed, the elected-lane issue serialises the warp, and the proxy fence is paid in every sweep.
// `on`: warp-uniform "stage k is inside the sweep"; the copies are predicated instructions, not a branch (a branch
// around them costs BSSY / BRA / BSYNC and a branch-resolve stall in every stage visit)
// PIPE_GROUP consecutive stage visits share ONE copy transaction (one commit group / one mbarrier phase): the records of
// neighbouring stages are contiguous in HBM, so a group is a single 960 * PIPE_GROUP byte block.  The ring holds
// NSLOT groups; `slot` (the visit counter modulo PRING) selects group slot / PIPE_GROUP and the record inside it.
template <bool SYNC = true>
__device__ __forceinline__ void pipe_fetch(const Ctx &c, int k, int dir, int klast, int group, bool on = true)
{
    if (SYNC) __syncwarp();                        // every lane is done with the group that is being refilled
#if PIPE_GROUP == 1 && !defined(PIPE_TMA)
    {   // one record per group: nothing to clip or to place (the general form below costs 2.4 % more instructions)
        on = on && (klast - k) * dir >= 0;
        const uint32_t dst = c.sm32 + (SM_BUF + group * SBUF) * 8;
        const char *src = reinterpret_cast<const char *>(c.lin + (size_t)(on ? k : 0) * STG);
#pragma unroll
        for (int j = 0; j < (SBUF * 8 / 16 + 31) / 32; j++) {
            const int e = j * 32 + c.lane32;
            const int go = (on && e < SBUF * 8 / 16) ? 1 : 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                         ::"r"(dst + e * 16), "l"(src + e * 16), "r"(go) : "memory");
        }
        return;
    }
#endif
    // stages k, k + dir, ... clipped at klast; in memory order the block starts at the lowest stage of the group
    int count = (klast - k) * dir + 1;
    count = count > PIPE_GROUP ? PIPE_GROUP : count;
    on = on && count > 0;
    const int klow = (dir > 0) ? k : k - count + 1;
    const int first = (dir > 0) ? 0 : PIPE_GROUP - count;      // backward sweeps fill the group from its top entry down
#ifndef PIPE_TMA
    const uint32_t dst = c.sm32 + (SM_BUF + (group * PIPE_GROUP + first) * SBUF) * 8;
    const char *src = reinterpret_cast<const char *>(c.lin + (size_t)(on ? klow : 0) * STG);
    const int chunks = on ? count * (SBUF * 8 / 16) : 0;
#pragma unroll
    for (int j = 0; j < (PIPE_GROUP * SBUF * 8 / 16 + 31) / 32; j++) {
        const int e = j * 32 + c.lane32;
        const int go = (e < chunks) ? 1 : 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                     ::"r"(dst + e * 16), "l"(src + e * 16), "r"(go) : "memory");
    }
#else
    if (!on) return;
    if (c.lane32 == 0) {
        const uint32_t bar = c.sm32 + (SM_MBAR + group) * 8;
        const uint32_t dst = c.sm32 + (SM_BUF + (group * PIPE_GROUP + first) * SBUF) * 8;
        mbar_expect_tx(bar, count * SBUF * 8);
        bulk_g2s(dst, c.lin + (size_t)klow * STG, count * SBUF * 8, bar);
    }
#endif
}
__device__ __forceinline__ void pipe_commit()
{
#ifndef PIPE_TMA
    asm volatile("cp.async.commit_group;" ::: "memory");      // one commit group per stage group (possibly empty)
#endif
}
// The default engine (cp.async, one record per group, two slots) keeps its addresses loop-carried: `psrc` walks the stage
// records in sweep direction (one 64-bit add per visit), the two slots are `pdst` and `pdst + 992`, and both 16-byte copies
// of a lane (chunks lane and 32 + lane of the 62) are predicated instructions of one asm block.  The general form above
// recomputes stage index, clipping, source and destination in every visit: 26 instructions per stage visit, 12 % of all
// executed instructions (profiles/r02_k_qp_solve_warp_final_by_function.txt).
#if PIPE_GROUP == 1 && PIPE_NSLOT == 2 && !defined(PIPE_TMA)
#define PIPE_FAST 1
// LO, HI: the sub-range [LO, HI) of the record (doubles, LO even) the sweep reads: chunks LO / 2 ... (HI + 1) / 2 - 1
template <int LO, int HI>
struct PipeRange {
    static constexpr int first = LO / 2, count = (HI + 1) / 2 - LO / 2;
    static_assert(LO % 2 == 0 && count > 32 && count <= 64 && HI <= SBUF, "two chunks per lane");
};
__device__ __forceinline__ void pipe_issue(uint32_t dst, const char *src, bool on, bool second)
{
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %2, 0;\n\tsetp.ne.s32 q, %3, 0;\n\t"
                 "@p cp.async.cg.shared.global [%0], [%1], 16;\n\t"
                 "@q cp.async.cg.shared.global [%0+512], [%1+512], 16;\n\t}"
                 ::"r"(dst), "l"(src), "r"((int)on), "r"((int)(on && second)) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
#endif
// first stage k0, direction dir (+1 / -1), last stage klast: fills NSLOT - 1 groups
template <int LO = 0, int HI = SBUF>
__device__ __forceinline__ void pipe_begin(const Ctx &c, int k0, int dir, int klast)
{
#ifdef PIPE_FAST
    using R = PipeRange<LO, HI>;
    __syncwarp();                                  // every lane is done with the slot that is being refilled
    const char *src = reinterpret_cast<const char *>(c.lin + (size_t)k0 * STG) + (R::first + c.lane32) * 16;
    pipe_issue(c.pdst + R::first * 16, src, (klast - k0) * dir >= 0, c.lane32 < R::count - 32);
    c.psrc = src + dir * (STG * 8);
    return;
#endif
#ifdef PIPE_TMA
    fence_proxy_async();
#endif
#pragma unroll
    for (int j = 0; j < NSLOT - 1; j++) {
        pipe_fetch(c, k0 + dir * j * PIPE_GROUP, dir, klast, j);
        pipe_commit();
    }
}
// inside the loop at stage k (visit number `slot` modulo PRING): at the first visit of a group, refill the group of the
// previous visits with the stages NSLOT - 1 groups ahead
// SYNCED: the caller's previous stage visit ended with a warp barrier after its last read of the record (no second one here)
template <bool SYNCED = false, int LO = 0, int HI = SBUF>
__device__ __forceinline__ void pipe_advance(const Ctx &c, int k, int dir, int klast, int slot)
{
#ifdef PIPE_FAST
    using R = PipeRange<LO, HI>;
    if (!SYNCED) __syncwarp();
    pipe_issue(c.pdst + R::first * 16 + (slot ^ 1) * (SBUF * 8), c.psrc, k != klast, c.lane32 < R::count - 32);      // stage k + dir
    c.psrc += dir * (STG * 8);
    return;
#endif
    if (PIPE_GROUP > 1 && (slot % PIPE_GROUP) != 0) return;      // (warp-uniform)
    const int kn = k + dir * PIPE_GROUP * (NSLOT - 1);
    if (!SYNCED) __syncwarp();                     // every lane is done with the group that is being refilled (kept outside
                                                   // the branch: inside it the compiler emits a divergence-safe barrier)
    pipe_fetch<false>(c, kn, dir, klast, (slot / PIPE_GROUP + NSLOT - 1) % NSLOT);
    pipe_commit();
}
// blocks until the group of this visit has landed (first visit of a group only); returns the shared-memory copy
// [LIN record | WORK record] of the visit's stage
// SYNC = false: the caller waits BEFORE it calls pipe_advance for this visit (so the groups in flight are one fewer) and
// follows up with its own warp barrier before any lane reads the record (cp.async: every lane has waited for its own copies
// only); that barrier then also frees the slot pipe_advance<true> refills and the caller's exchange slots
template <bool SYNC = true>
__device__ __forceinline__ const double *pipe_wait(const Ctx &c, int slot, int dir)
{
    const int group = slot / PIPE_GROUP, sub = slot % PIPE_GROUP;
    if (PIPE_GROUP == 1 || sub == 0) {
#ifndef PIPE_TMA
        // all but the newest groups in flight have landed: NSLOT - 1 of them after this visit's pipe_advance, NSLOT - 2 before it
        asm volatile("cp.async.wait_group %0;" ::"n"(SYNC ? NSLOT - 1 : NSLOT - 2) : "memory");
        if (SYNC) __syncwarp();
#else
        mbar_wait(c.sm32 + (SM_MBAR + group) * 8, (c.phases >> group) & 1u);
        c.phases ^= 1u << group;
#endif
    }
    return c.sm + SM_BUF + (group * PIPE_GROUP + (dir > 0 ? sub : PIPE_GROUP - 1 - sub)) * SBUF;
}

// B_k rows 0..5 into registers (the record holds B transposed: 6 x 128-bit broadcast loads); rows 6,7 are [0 dt], [dt 0]
__device__ __forceinline__ void load_B(const double *AB, double *B0, double *B1)
{
    const double2 *b2 = reinterpret_cast<const double2 *>(AB + L_B);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const double2 v = b2[r], w = b2[3 + r];
        B0[2 * r] = v.x; B0[2 * r + 1] = v.y; B1[2 * r] = w.x; B1[2 * r + 1] = w.y;
    }
}
// six contiguous doubles (a column of A or B in the transposed record, or the six zeros) with 128-bit loads
__device__ __forceinline__ void load6(const double *p, double *o)
{
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int r = 0; r < 3; r++) { const double2 v = p2[r]; o[2 * r] = v.x; o[2 * r + 1] = v.y; }
}
// (B' v), v gathered
__device__ __forceinline__ void Bt_dot(const double *B0, const double *B1, double dt, const double *v, double &y0, double &y1)
{
    double s0 = dt * v[7], s1 = dt * v[6];
#pragma unroll
    for (int r = 0; r < 6; r++) { s0 += B0[r] * v[r]; s1 += B1[r] * v[r]; }
    y0 = s0; y1 = s1;
}
// (A' v)_j, v gathered (vj: this lane's own component): column (j-2) of A36 plus the unit parts
__device__ __forceinline__ double At_comp(const Ctx &c, const double *AB, const double *v, double vj)
{
    // every lane runs the same stream (lanes 0, 1 read column 0 and discard the result: a branch costs more than the
    // six DFMA); two chains of three: the dot product sits on the critical path of every stage of the backward sweeps
    const int cc = (c.lane < 2) ? 0 : c.lane - 2;
    const double v6 = v[6], v7 = v[7];
    const double s6 = (c.lane == 6) ? v6 : 0.0;
    double a[6];
    load6(AB + L_A + cc * 6, a);                   // column cc of A: contiguous in the transposed record
    double s = (c.lane == 7) ? v7 : s6, t = a[3] * v[3];
    s += a[0] * v[0]; t += a[4] * v[4];
    s += a[1] * v[1]; t += a[5] * v[5];
    s += a[2] * v[2];
    return (c.lane < 2) ? vj : s + t;
}
// (A x)_j + (B u)_j, x gathered (xj: this lane's own component): row j of A36 plus the unit parts
__device__ __forceinline__ double dyn_comp(const Ctx &c, const double *AB, const double *x, double xj, double u0, double u1)
{
    // every lane runs the same stream (lanes 6, 7 redo row 5 and discard it); A x in two chains of three (independent
    // of u: overlaps the reduction that produces u), then B u
    const int r = (c.lane < 6) ? c.lane : 5;
    const double *Ar = AB + L_A + r;               // row r of A: stride 6 in the transposed record (immediate offsets)
    double s = (c.lane < 2) ? xj : 0.0, t = Ar[18] * x[5];
    s += Ar[0] * x[2]; t += Ar[24] * x[6];
    s += Ar[6] * x[3]; t += Ar[30] * x[7];
    s += Ar[12] * x[4];
    s += t;
    s += AB[L_B + r] * u0 + AB[L_B + 6 + r] * u1;
    return (c.lane < 6) ? s : xj + c.dt * ((c.lane == 6) ? u1 : u0);
}


// ------------------------------------------------------------------------------------------------------
// SNMPC head: dense nh x nh system of the eliminated stages (lane a < nh owns head input a)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int hs_id(const Ctx &c) { return c.nh * c.ldm; }
__device__ __forceinline__ int hs_ru(const Ctx &c) { return c.nh * c.ldm + c.nh; }
__device__ __forceinline__ int hs_dd(const Ctx &c) { return c.nh * c.ldm + 2 * c.nh; }
__device__ __forceinline__ int hs_pg(const Ctx &c) { return c.nh * c.ldm + 3 * c.nh; }
__host__ __device__ inline int head_smem_doubles(int H) { const int nh = 2 * H; return (nh * (nh + 1) + 3 * nh + 8 * nh + 1) & ~1; }

// sum over the head rows of coef_r * row_r[a]; coefficients are the per-(stage,row) values PS(arr, k, row)
__device__ __forceinline__ double head_rows_dot(const Ctx &c, int arr, int a)
{
    if (a >= c.nh) return 0.0;
    double s = (a & 1) ? c.PS(arr, a >> 1, 0) : 0.0;                    // bu row of stage a/2 is e_a
    for (int k = 1; k < c.H; k++)
        s += c.PS(arr, k, 1) * c.hd[c.hG + ((k - 1) * 8 + 6) * c.nh + a]    // bx row: delta row of G_k
           + c.PS(arr, k, 2) * c.hd[c.hROW + k * c.nh + a];                 // chance-constraint row
    return s;
}
// (G_H' v)_a, v gathered
__device__ __forceinline__ double head_GtV(const Ctx &c, const double *v, int a)
{
    if (a >= c.nh) return 0.0;
    const double *G = c.hd + c.hG + (size_t)(c.H - 1) * 8 * c.nh + a;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += G[i * c.nh] * v[i];
    return s;
}
// Up to HF_NH head inputs (the committed configuration has UPH = 5: nh = 10) the head matrix is assembled and factored in
// REGISTERS: lane a keeps row a of M, the Cholesky columns travel by warp shuffles (a pivot broadcast, a branch-free
// reciprocal square root every lane computes redundantly, one shuffle per row below the pivot), and only the finished factor
// goes to shared memory for the two solves of the iteration.  The shared-memory version below (three warp barriers and an
// IEEE sqrt + division per column, read-modify-write loops over the rows) cost ~11 k cycles per interior-point iteration of
// a kernel that is pure dependent latency at the BASELINE size; it remains the path for longer propagation horizons.
constexpr int HF_NH = 10;
// right-looking Cholesky, column J (a template recursion so that every register index is a compile-time constant): pivot
// from lane J, l_aJ = m_aJ / sqrt(m_JJ) for a >= J (a == J: the pivot's root), then m_al -= l_aJ l_lJ for J < l <= a with
// l_lJ shuffled from lane l
template <int J>
__device__ __forceinline__ void head_chol_cols(double (&mr)[HF_NH], int a, int ar, int nh, double *ID)
{
    if constexpr (J < HF_NH) {
        if (J < nh) {                               // (warp-uniform)
            const double piv = __shfl_sync(FULLMASK, mr[J], J);
            const double id = fast_rsqrt(piv);
            const double lij = mr[J] * id;
            mr[J] = lij;
            if (a == 0) ID[J] = id;
#pragma unroll
            for (int l = J + 1; l < HF_NH; l++) {
                if (l < nh) {
                    const double llj = __shfl_sync(FULLMASK, lij, l);
                    mr[l] -= (l <= ar) ? lij * llj : 0.0;
                }
            }
            head_chol_cols<J + 1>(mr, a, ar, nh, ID);
        }
    }
}
__device__ __forceinline__ void head_factor_regs(const Ctx &c, bool with_rows)
{
    double *tile = c.sm + SM_PT, *M = c.hs, *ID = c.hs + hs_id(c), *PG = c.hs + hs_pg(c);
    const int a = c.lane32, nh = c.nh, ldm = c.ldm;
    const int ar = a < nh ? a : 0;                  // lanes beyond nh shadow row 0 (their results are never stored)
    __syncwarp();
    double gcol[8];
    {
        const double *G = c.hd + c.hG + (size_t)(c.H - 1) * 8 * nh + ar;
#pragma unroll
        for (int i = 0; i < 8; i++) gcol[i] = G[i * nh];
        if (a < nh) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                double sacc = 0.0;
#pragma unroll
                for (int l = 0; l < 8; l++) sacc += tile[i * TLD + l] * gcol[l];
                PG[i * nh + a] = sacc;
            }
        }
    }
    __syncwarp();
    double mr[HF_NH];                               // row `ar` of M (columns >= nh stay zero)
#pragma unroll
    for (int b = 0; b < HF_NH; b++) {
        double sacc = 0.0;
        if (b < nh) {                               // (warp-uniform)
            sacc = c.hd[c.hHH + ar * nh + b];
#pragma unroll
            for (int i = 0; i < 8; i++) sacc += gcol[i] * PG[i * nh + b];
        }
        mr[b] = sacc;
    }
    if (with_rows) {
        for (int k = 1; k < c.H; k++) {
            const double *gk = c.hd + c.hG + ((k - 1) * 8 + 6) * nh, *rk = c.hd + c.hROW + k * nh;
            const double wx = c.PS(PS_OM, k, 1) * gk[ar], wh = c.PS(PS_OM, k, 2) * rk[ar];
#pragma unroll
            for (int b = 0; b < HF_NH; b++)
                if (b < nh) mr[b] += wx * gk[b] + wh * rk[b];
        }
        const double obu = (ar & 1) ? c.PS(PS_OM, ar >> 1, 0) : 0.0;      // bu row of stage a/2 is e_a
#pragma unroll
        for (int b = 1; b < HF_NH; b += 2) mr[b] = fma(obu, (b == ar) ? 1.0 : 0.0, mr[b]);
    }
    head_chol_cols<0>(mr, a, ar, nh, ID);
    if (a < nh) {
#pragma unroll
        for (int b = 0; b < HF_NH; b++)
            if (b <= a) M[a * ldm + b] = mr[b];
    }
    __syncwarp();
}

// M = HH (+ barrier terms of the head rows) + G_H' P_H G_H, then in-place Cholesky (lower) and 1/diag
// P_H is in the P tile of the 32-lane sweep
__device__ void head_factor(const Ctx &c, bool with_rows)
{
    if (c.nh <= HF_NH) { head_factor_regs(c, with_rows); return; }
    double *tile = c.sm + SM_PT, *M = c.hs, *ID = c.hs + hs_id(c), *PG = c.hs + hs_pg(c);
    const int a = c.lane32, nh = c.nh, ldm = c.ldm;
    __syncwarp();
    double gcol[8];
    if (a < nh) {
        const double *G = c.hd + c.hG + (size_t)(c.H - 1) * 8 * nh + a;
#pragma unroll
        for (int i = 0; i < 8; i++) gcol[i] = G[i * nh];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            double sacc = 0.0;
#pragma unroll
            for (int l = 0; l < 8; l++) sacc += tile[i * TLD + l] * gcol[l];
            PG[i * nh + a] = sacc;
        }
    }
    __syncwarp();
    if (a < nh) {
        for (int b = 0; b <= a; b++) {
            double sacc = c.hd[c.hHH + a * nh + b];
#pragma unroll
            for (int i = 0; i < 8; i++) sacc += gcol[i] * PG[i * nh + b];
            M[a * ldm + b] = sacc;
        }
        if (with_rows) {
            for (int k = 1; k < c.H; k++) {
                const double *gk = c.hd + c.hG + ((k - 1) * 8 + 6) * nh, *rk = c.hd + c.hROW + k * nh;
                const double wx = c.PS(PS_OM, k, 1) * gk[a], wh = c.PS(PS_OM, k, 2) * rk[a];
                for (int b = 0; b <= a; b++) M[a * ldm + b] += wx * gk[b] + wh * rk[b];
            }
            if (a & 1) M[a * ldm + a] += c.PS(PS_OM, a >> 1, 0);
        }
    }
    __syncwarp();
    for (int j = 0; j < nh; j++) {
        const double d = sqrt(M[j * ldm + j]);
        const double id = 1.0 / d;
        const bool below = a > j && a < nh;
        double lij = 0.0;
        if (below) lij = M[a * ldm + j] * id;
        __syncwarp();
        if (a == j) { M[j * ldm + j] = d; ID[j] = id; }
        if (below) M[a * ldm + j] = lij;
        __syncwarp();
        if (below)
            for (int l = j + 1; l <= a; l++) M[a * ldm + l] -= lij * M[l * ldm + j];
        __syncwarp();
    }
}
// solves M x = rhs with the stored factor (lane a holds rhs_a) and leaves -x in the shared solution vector
__device__ void head_solve(const Ctx &c, double rhs)
{
    const double *M = c.hs, *ID = c.hs + hs_id(c);
    double *DD = c.hs + hs_dd(c);
    const int a = c.lane32, nh = c.nh, ldm = c.ldm;
    double r = (a < nh) ? rhs : 0.0;
    for (int j = 0; j < nh; j++) {
        const double xj = __shfl_sync(FULLMASK, r, j) * ID[j];
        if (a == j) r = xj;
        else if (a > j && a < nh) r -= M[a * ldm + j] * xj;
    }
    for (int j = nh - 1; j >= 0; j--) {
        const double xj = __shfl_sync(FULLMASK, r, j) * ID[j];
        if (a == j) r = xj;
        else if (a < j) r -= M[j * ldm + a] * xj;
    }
    __syncwarp();
    if (a < nh) DD[a] = -r;
    __syncwarp();
}
// Expands the head solution in the shared solution vector to the eliminated stages.
//   AFFINE (phase 0): u~ itself -> QP iterate dx_k = G_k u~ + c_k, du_k, row values of stages < H.
//   else            : Newton direction -> ddx_k (kept if STORE), ddu_k, direction row values.
// Returns component `lane` of the stage-H vector (dx_H or ddx_H).
template <bool AFFINE, bool STORE>
__device__ double head_apply(const DevMem &m, const Ctx &c)
{
    const double *uu = c.hs + hs_dd(c);
    const int H = c.H, nh = c.nh;
    double *xb = c.sm + SM_X;
    __syncwarp();
    for (int item = c.lane32; item < H * 8 + H; item += 32) {
        if (item < H * 8) {
            const int k = (item >> 3) + 1, j = item & 7;
            const double *Gr = c.hd + c.hG + (size_t)item * nh;
            double sacc = AFFINE ? c.hd[c.hC + item] : 0.0;
            for (int a = 0; a < nh; a++) sacc += Gr[a] * uu[a];
            if (k < H) {
                double *w = c.wk + (size_t)k * STG;
                if (AFFINE) w[W_DX + j] = sacc;
                else if (STORE) w[W_DDX + j] = sacc;
                if (j == 6) {
                    if (AFFINE) c.PS(PS_ROWV, k, 1) = m.X[((size_t)c.b * c.NS + k) * 8 + 6] + sacc;
                    else c.PS(PS_DV, k, 1) = sacc;
                }
            } else {
                xb[j] = sacc;
            }
        } else {
            const int k = item - H * 8;
            double sacc = 0.0;
            if (k >= 1) {
                const double *rk = c.hd + c.hROW + k * nh;
                sacc = AFFINE ? c.hd[c.hV0 + k] : 0.0;
                for (int a = 0; a < nh; a++) sacc += rk[a] * uu[a];
            }
            const double u0 = uu[2 * k], u1 = uu[2 * k + 1];
            double *w = c.wk + (size_t)k * STG;
            if (AFFINE) {
                w[W_DU] = u0; w[W_DU + 1] = u1;
                c.PS(PS_ROWV, k, 0) = m.U[((size_t)c.b * c.N + k) * 2 + 1] + u1;
                c.PS(PS_ROWV, k, 2) = sacc;
                if (k == 0) c.PS(PS_ROWV, 0, 1) = m.X[(size_t)c.b * c.NS * 8 + 6];
            } else {
                { double *q = c.SWS(k, W_DDU, PS_DDU); q[0] = u0; q[1] = u1; }
                c.PS(PS_DV, k, 0) = u1;
                c.PS(PS_DV, k, 2) = sacc;
                if (k == 0) c.PS(PS_DV, 0, 1) = 0.0;
            }
        }
    }
    if (c.st) {       // stage 0: dx_0 = x0 - x_0 is fixed, its direction is zero
        if (AFFINE) c.wk[W_DX + c.lane] = m.x0[(size_t)c.b * 8 + c.lane] - m.X[(size_t)c.b * c.NS * 8 + c.lane];
        else if (STORE) c.wk[W_DDX + c.lane] = 0.0;
    }
    __syncwarp();
    return xb[c.lane];
}

// sigma-scaled cost diagonal of stage k from the per-stage shared block (filled once per solve by phase_prep)
__device__ __forceinline__ void cost_diag(const DevCfg &cfg, const Ctx &c, int k, double &qdj, double &rd0, double &rd1)
{
    const int nw = cfg.cost_vel_abs ? 3 : 4;
    qdj = (c.lane < nw) ? c.PS(PS_W, k, c.lane & 3) : 0.0;
    rd0 = c.PS(PS_W, k, 4);
    rd1 = c.PS(PS_W, k, 5);
}

// ------------------------------------------------------------------------------------------------------
// Sequential sweeps (stage recursion, octet decomposition)
// ------------------------------------------------------------------------------------------------------

// Backward vector sweep at the current iterate: costates (defined by the x-stationarity rows) and the u-stationarity
// residual; returns max |r_u|.  Runs once per solve, for the convergence check after the complementarity products and
// the linear residuals of the sides have reached the tolerance (octet decomposition: lane j owns component j).
template <bool HEAD>
__device__ double sweep_residual(const DevCfg &cfg, const DevMem &m, const Ctx &c)
{
    const int N = c.N;
    double pij = 0.0, rstat = 0.0;
    const int kstop = HEAD ? c.H : 0;
    int slot = 0;
    pipe_begin<R_BACK_LO, R_BACK_HI>(c, N, -1, kstop);
    for (int k = N; k >= kstop; k--, slot = (slot + 1) % PRING) {
        pipe_advance<false, R_BACK_LO, R_BACK_HI>(c, k, -1, kstop, slot);
        const double *rec = pipe_wait(c, slot, -1);      // LIN record of stage k (A36 | B12 | b8 first)
        const double *w = rec;                    // WORK record of stage k
        double qdj, rd0, rd1;
        cost_diag(cfg, c, k, qdj, rd0, rd1);
        const double gj = rec[L_GF + c.lane];
        const double dxj = w[W_DX + c.lane];
        double du0 = 0.0, du1 = 0.0;
        if (k < N) { du0 = w[W_DU]; du1 = w[W_DU + 1]; }
        double gxj = qdj * dxj + ((c.lane < 5) ? rec[L_QX + (c.lane < 5 ? c.lane : 0)] : 0.0);
        if (cfg.cost_vel_abs) {
            const double Qv0 = rec[L_QV], Qv1 = rec[L_QV + 1], Qv2 = rec[L_QV + 2], dx3 = w[W_DX + 3], dx4 = w[W_DX + 4];
            if (c.lane == 3) gxj += Qv0 * dx3 + Qv1 * dx4;
            if (c.lane == 4) gxj += Qv1 * dx3 + Qv2 * dx4;
        }
        double gu0 = 0.0, gu1 = 0.0;
        if (k < N) { gu0 = rec[L_RG] + rd0 * du0; gu1 = rec[L_RG + 1] + rd1 * du1; }
        const double ld_bu = c.PS(PS_LD, k, 0), ld_bx = c.PS(PS_LD, k, 1), ld_h = c.PS(PS_LD, k, 2);
        gu1 -= ld_bu;
        gxj -= ((c.lane == 6) ? ld_bx : 0.0) + ld_h * gj;
        if (k < N) {
            double piall[8];
            xchg1(c, 0, pij, piall);
            gxj += At_comp(c, rec, piall, pij);
            double B0[6], B1[6], t0, t1;
            load_B(rec, B0, B1);
            Bt_dot(B0, B1, c.dt, piall, t0, t1);
            gu0 += t0; gu1 += t1;
            rstat = nmax(rstat, nmax(fabs(gu0), fabs(gu1)));
        }
        pij = gxj;
    }
    if (HEAD) {
        const int a = c.lane32;
        double piall[8];
        xchg1(c, 0, pij, piall);
        double ru = 0.0;
        if (a < c.nh) {
            ru = c.hd[c.hGH + a];
            for (int b = 0; b < c.nh; b++) ru += c.hd[c.hHH + a * c.nh + b] * c.wk[(size_t)(b >> 1) * STG + W_DU + (b & 1)];
            ru += head_GtV(c, piall, a) - head_rows_dot(c, PS_LD, a);
        }
        rstat = nmax(rstat, warp_nmax(fabs(ru)));
    }
    __syncwarp();
    return rstat;
}

// The per-iteration full backward sweep (barrier Hessian, factorisation, predictor right-hand side, costates and
// u-stationarity residual) on ALL 32 lanes: lane (i, q) = (row i = lane & 7, column pair q = lane >> 3) owns
// P[i][2q..2q+1].  P_{k+1}, P A and the small products live in shared-memory tiles; three warp syncs per stage:
//   A: row i of P from the tile -> (P A)[i][2q..2q+1] (q = 0: the unit columns, so those lanes do P B instead);
//      lanes q = 1 add (A' pi)_i to the x-stationarity row, two lanes form B' pi
//   B: (A' P A)[i][2q..2q+1] from columns of the P A tile; designated lanes: B' P A (16), A' p + q~ (8), R~ + B' P B (3),
//      B' p + r~ (2)
//   C: 2x2 Cholesky (every lane), K, P_k = A'PA - Y'Y + Q~ back into the tile, p_k, pi_k
// Compared with the octet recursion of v3/v4 (lane j owned row j of P in registers, four redundant octets) each lane
// issues ~45 instead of ~160 DFMA per stage and no register-to-register exchange is needed.
// PHASE0: the equality-constrained (no barrier terms) AFFINE solve of phase 0: v = p + P b replaces p in B'v, A'v
template <bool HEAD, bool PHASE0>
__device__ double sweep_backward_full(const DevCfg &cfg, const DevMem &m, const Ctx &c, double ap = 0.0)
{
    const int N = c.N, i = c.lane, q = c.lane32 >> 3, L = c.lane32;
    double *Pt = c.sm + SM_PT, *PAt = c.sm + SM_PA, *sml = c.sm + SM_SML, *pv = c.sm + SM_PV, *piv = pv + 8;
    double *PBs = sml, *Huxs = sml + 16, *hxs = sml + 32, *Hs = sml + 40, *hus = sml + 44, *tpi = sml + 46;
    double *vv = PHASE0 ? c.sm + SM_X : pv;                 // phase 0: v = p + P b of the stage (exchange slots are free here)
    const double *zero6 = sml + 48;                        // six zeros: the "column" of a unit row / an idle lane
    double *dummy = sml + 54;
    const double dt = c.dt;
    double rstat = 0.0;
    const int kstop = HEAD ? c.H : 0;
    const int c0 = 2 * q;                                  // first of this lane's two columns
    const bool unit_i = (i < 2) || (i >= 6);               // rows / columns of A with a unit entry on the diagonal

    // ---- loop-invariant per-lane descriptors: every product below is  w * Z[x] + sum_{r<6} M[r * ms] * Z[r * zs]
    // with the column M of [A36 | B12] (or six zeros) and the vector Z chosen per lane, so that all lanes run ONE
    // instruction stream (the first version of this sweep serialised four designated-lane branches per phase).
    // phase A, products 1-2: row i of P times two columns: q = 0 -> P B (the unit columns 0, 1 of A need no product),
    // q >= 1 -> columns c0, c0+1 of A (q = 3: plus the unit entries of columns 6, 7)
    const int a_off = (q == 0) ? L_B : L_A + (c0 - 2) * 6;      // two adjacent columns = 12 contiguous doubles
    const int a_x0 = (q == 0) ? 7 : 6, a_x1 = (q == 0) ? 6 : 7;
    const double a_w = (q == 0) ? dt : (q == 3) ? 1.0 : 0.0;
    double *a_out = (q == 0) ? PBs + 2 * i : PAt + i * TLD + c0;
    // phase A, product 3 (lanes 0..9): [A | B]' pi
    const bool p3_zero = (L < 2) || (L >= 10);
    const int p3_off = (L < 8) ? L_A + (L - 2) * 6 : L_B + (L - 8) * 6;
    const int p3_x = (L < 8) ? L : (L < 10) ? 15 - L : 0;                 // B' pi: rows 7, 6 carry dt
    const double p3_w = (L < 8) ? (((L < 2) || (L >= 6)) ? 1.0 : 0.0) : (L < 10) ? dt : 0.0;
    // phase B, products 1-2: (A' P A)[i][c0..c0+1]
    const bool b_zero = i < 2;
    const int b_off = L_A + (i - 2) * 6;
    const double b_w = unit_i ? 1.0 : 0.0;
    // phase B, product 3: lanes 0..15 B' P A, 16..23 A' p + q~, 24..26 R~ + B' P B, 27..28 B' p + r~
    bool g_zero = false;
    int g_off = L_B, g_zs = 1, g_x = 0;
    double g_w = 0.0, f_qt = 0.0, f_rd0 = 0.0, f_h11 = 0.0, f_gu0 = 0.0, f_gu1 = 0.0;
    const double *g_Z = vv;
    double *g_out = dummy;
    if (L < 16) { g_off = L_B + q * 6; g_Z = PAt + i; g_zs = TLD; g_x = 7 - q; g_w = dt; g_out = Huxs + q * 8 + i; }
    else if (L < 24) { g_zero = b_zero; g_off = b_off; g_x = i; g_w = b_w; f_qt = 1.0; g_out = hxs + i; }
    else if (L < 27) {
        const int e = L - 24, ca = (e == 2) ? 1 : 0, cb = (e == 0) ? 0 : 1;
        g_off = L_B + ca * 6; g_Z = PBs + cb; g_zs = 2; g_x = 7 - ca; g_w = dt; g_out = Hs + e;
        f_rd0 = (e == 0) ? 1.0 : 0.0; f_h11 = (e == 2) ? 1.0 : 0.0;
    } else if (L < 29) {
        const int e = L - 27;
        g_off = L_B + e * 6; g_x = 7 - e; g_w = dt; g_out = hus + e; f_gu0 = (e == 0) ? 1.0 : 0.0; f_gu1 = (e == 1) ? 1.0 : 0.0;
    } else { g_zero = true; }
    // Q~ selectors of this lane's two entries
    const double s_da = (i == c0) ? 1.0 : 0.0, s_db = (i == c0 + 1) ? 1.0 : 0.0, s_bx = (i == 6 && q == 3) ? 1.0 : 0.0;
    const double s_v0b = (i == 3 && q == 1) ? 1.0 : 0.0, s_v1a = (i == 3 && q == 2) ? 1.0 : 0.0;
    const double s_v1b = (i == 4 && q == 1) ? 1.0 : 0.0, s_v2a = (i == 4 && q == 2) ? 1.0 : 0.0;
    const double s_g3 = (i == 3) ? 1.0 : 0.0, s_g4 = (i == 4) ? 1.0 : 0.0, s_i6 = (i == 6) ? 1.0 : 0.0;
    const int qx_i = (i < 5) ? i : 0;
    const double s_qx = (i < 5) ? 1.0 : 0.0;

    int slot = 0;
    pipe_begin<R_BACK_LO, R_BACK_HI>(c, N, -1, kstop);
    for (int k = N; k >= kstop; k--, slot = (slot + 1) % PRING) {
        pipe_advance<true, R_BACK_LO, R_BACK_HI>(c, k, -1, kstop, slot);      // (every stage visit below ends with a warp barrier)
        const double *rec = pipe_wait(c, slot, -1);
        double *const psk = c.PSK(k);
        const double *w = rec;
        double qdj, rd0, rd1;
        cost_diag(cfg, c, k, qdj, rd0, rd1);
        const double gj = rec[L_GF + i];
        const double2 gc = *reinterpret_cast<const double2 *>(rec + L_GF + c0);
        double om_bu = 0.0, om_bx = 0.0, om_h = 0.0, cf_bu = 0.0, cf_bx = 0.0, cf_h = 0.0, ld_bu = 0.0, ld_bx = 0.0, ld_h = 0.0;
        if (!PHASE0) {
            om_bu = psk[PS_OM + 0]; om_bx = psk[PS_OM + 1]; om_h = psk[PS_OM + 2];
            cf_bu = psk[PS_CF + 0]; cf_bx = psk[PS_CF + 1]; cf_h = psk[PS_CF + 2];
            ld_bu = psk[PS_LD + 0]; ld_bx = psk[PS_LD + 1]; ld_h = psk[PS_LD + 2];
        }
        // x-stationarity row i at the current iterate, without the costate term
        // pending primal step of the previous iteration (length ap along the direction the last forward sweep stored): applied
        // here from the shared-memory copy of the record and written back, instead of a stage-parallel pass over HBM
        const bool lazy = !PHASE0 && ap != 0.0;
        double dxi = PHASE0 ? 0.0 : w[W_DX + i];
        if (lazy) {
            dxi += ap * w[W_DDX + i];
            if (q == 0) c.wk[(size_t)k * STG + W_DX + i] = dxi;
        }
        double gxj = PHASE0 ? 0.0 : qdj * dxi + s_qx * rec[L_QX + qx_i] - s_i6 * ld_bx - ld_h * gj;
        // Q~[i][c0], Q~[i][c0+1]: cost diagonal, soft-row barrier terms, |v| block
        const double ohg = om_h * gj;
        double qa = ohg * gc.x + s_da * qdj + s_bx * om_bx, qb = ohg * gc.y + s_db * qdj;
        if (cfg.cost_vel_abs) {
            const double Qv0 = rec[L_QV], Qv1 = rec[L_QV + 1], Qv2 = rec[L_QV + 2];
            const double dx3 = w[W_DX + 3] + (lazy ? ap * w[W_DDX + 3] : 0.0), dx4 = w[W_DX + 4] + (lazy ? ap * w[W_DDX + 4] : 0.0);
            if (!PHASE0) gxj += s_g3 * (Qv0 * dx3 + Qv1 * dx4) + s_g4 * (Qv1 * dx3 + Qv2 * dx4);
            qa += s_v1a * Qv1 + s_v2a * Qv2; qb += s_v0b * Qv0 + s_v1b * Qv1;
        }
        const double qtj = PHASE0 ? s_qx * rec[L_QX + qx_i] : -(s_i6 * cf_bx + cf_h * gj);
        if (k == N) {
            *reinterpret_cast<double2 *>(Pt + i * TLD + c0) = make_double2(qa, qb);
            if (q == 0) { pv[i] = qtj; piv[i] = gxj; }
            __syncwarp();
            continue;
        }
        double du0 = PHASE0 ? 0.0 : w[W_DU], du1 = PHASE0 ? 0.0 : w[W_DU + 1];
        if (lazy) {
            const double *q = c.SWL(w, psk, W_DDU, PS_DDU);
            du0 += ap * q[0]; du1 += ap * q[1];
            if (L == 0) { c.wk[(size_t)k * STG + W_DU] = du0; c.wk[(size_t)k * STG + W_DU + 1] = du1; }
        }
        // ---------------- phase A
        double Prow[8];
        {
            const double2 *p2 = reinterpret_cast<const double2 *>(Pt + i * TLD);
#pragma unroll
            for (int j = 0; j < 4; j++) { const double2 v = p2[j]; Prow[2 * j] = v.x; Prow[2 * j + 1] = v.y; }
        }
        {
            double Ma[12];
            load6(rec + a_off, Ma); load6(rec + a_off + 6, Ma + 6);
            double o0 = a_w * Pt[i * TLD + a_x0], o1 = a_w * Pt[i * TLD + a_x1];
#pragma unroll
            for (int r = 0; r < 6; r++) { o0 += Prow[r] * Ma[r]; o1 += Prow[r] * Ma[6 + r]; }
            *reinterpret_cast<double2 *>(a_out) = make_double2(o0, o1);
            if (q == 0) *reinterpret_cast<double2 *>(PAt + i * TLD) = make_double2(Prow[0], Prow[1]);
        }
        if (PHASE0) {                                        // v_i = p_i + (P b)_i
            double vi = pv[i];
#pragma unroll
            for (int l = 0; l < 8; l++) vi += Prow[l] * rec[L_b + l];
            if (q == 1) vv[i] = vi;
        } else {
            double Mp[6];
            load6(p3_zero ? zero6 : rec + p3_off, Mp);
            double t = p3_w * piv[p3_x];
#pragma unroll
            for (int r = 0; r < 6; r++) t += Mp[r] * piv[r];
            gxj += (L < 8) ? t : 0.0;                       // lanes (i, q = 0) own the costate row i
            if (L >= 8 && L < 10) tpi[L - 8] = t;           // B' pi from lanes 8, 9 (a lone predicated store)
        }
        __syncwarp();
        // ---------------- phase B
        const double gu0 = PHASE0 ? rec[L_RG] : rec[L_RG] + rd0 * du0 + tpi[0];
        const double gu1 = PHASE0 ? rec[L_RG + 1] : rec[L_RG + 1] + rd1 * du1 - ld_bu + tpi[1];
        rstat = fmax(rstat, fmax(fabs(gu0), fabs(gu1)));    // a NaN here also poisons the side variables and is caught there
        double hxx0, hxx1;
        {
            double Mb[6];
            load6(b_zero ? zero6 : rec + b_off, Mb);
            const double2 ex = *reinterpret_cast<const double2 *>(PAt + i * TLD + c0);
            hxx0 = b_w * ex.x; hxx1 = b_w * ex.y;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                const double2 pr = *reinterpret_cast<const double2 *>(PAt + r * TLD + c0);
                const double a = Mb[r];
                hxx0 += a * pr.x; hxx1 += a * pr.y;
            }
        }
        {
            double Mg[6];
            load6(g_zero ? zero6 : rec + g_off, Mg);
            double sacc = g_w * g_Z[g_x * g_zs] + f_qt * qtj + f_rd0 * rd0 + f_h11 * (rd1 + om_bu) + f_gu0 * gu0 + f_gu1 * (gu1 - cf_bu);
#pragma unroll
            for (int r = 0; r < 6; r++) sacc += Mg[r] * g_Z[r * g_zs];
            *g_out = sacc;
        }
        __syncwarp();
        // ---------------- phase C: 2x2 Cholesky through two reciprocal square roots, gains, P_k, p_k, pi_k.
        // P_k = A'PA - Y'Y with Y = L^-1 (B'PA): the subtraction is bitwise symmetric in (i, c), which keeps P symmetric
        // under the 1e9 barrier weights (a gain-form update P + Hux' K drifts and loses definiteness there).
        const double h00 = Hs[0], h01 = Hs[1], h11 = Hs[2], hu0 = hus[0], hu1 = hus[1];
        const double i11 = fast_rsqrt(h00), l21 = h01 * i11;
        const double i22 = fast_rsqrt(h11 - l21 * l21);
        const double2 ua = *reinterpret_cast<const double2 *>(Huxs + c0), ub = *reinterpret_cast<const double2 *>(Huxs + 8 + c0);
        const double y0a = ua.x * i11, y0b = ua.y * i11;
        const double y1a = (ub.x - l21 * y0a) * i22, y1b = (ub.y - l21 * y0b) * i22;
        const double y0i = Huxs[i] * i11, y1i = (Huxs[8 + i] - l21 * y0i) * i22;
        const double pn0 = hxx0 - (y0i * y0a + y1i * y1a) + qa, pn1 = hxx1 - (y0i * y0b + y1i * y1b) + qb;
        const double K1 = -y1i * i22;
        const double K0 = (-y0i - l21 * K1) * i11;
        const double pnew = hxs[i] + K0 * hu0 + K1 * hu1;
        // (the P tile and the vectors of stage k+1 were last read before the phase-B barrier)
        *reinterpret_cast<double2 *>(Pt + i * TLD + c0) = make_double2(pn0, pn1);
        if (q == 0) {
            pv[i] = pnew; piv[i] = gxj;
            double *wg = c.wk + (size_t)k * STG + W_K;
            wg[i] = K0; wg[8 + i] = K1;
        }
        {   // every lane computes, lane 31 stores (predicated stores instead of a divergent block)
            const double m21 = -l21 * i11 * i22;
            const double Hi0 = i11 * i11 + m21 * m21, Hi1 = m21 * i22, Hi3 = i22 * i22;
            const double kf0 = -(Hi0 * hu0 + Hi1 * hu1), kf1 = -(Hi1 * hu0 + Hi3 * hu1);
            if (L == 31) {
                double *wg = c.wk + (size_t)k * STG;
                wg[W_HI] = Hi0; wg[W_HI + 1] = Hi1; wg[W_HI + 2] = Hi3;
                double *qk = c.SWS(k, W_KFF, PS_KFF), *qg = c.SWS(k, W_GUS, PS_GUS);
                qk[0] = kf0; qk[1] = kf1;
                qg[0] = gu0; qg[1] = gu1;
            }
        }
        __syncwarp();
    }
    if (HEAD && PHASE0) {
        // gradient of the cost-to-go at dx_H = c_H: v = p_H + P_H c_H
        const int a = c.lane32;
        double vj = pv[i];
#pragma unroll
        for (int l = 0; l < 8; l++) vj += Pt[i * TLD + l] * c.hd[c.hC + (c.H - 1) * 8 + l];
        double vall[8];
        xchg1(c, 0, vj, vall);
        head_factor(c, false);
        head_solve(c, (a < c.nh ? c.hd[c.hGH + a] : 0.0) + head_GtV(c, vall, a));
    } else if (HEAD) {
        const int a = c.lane32;
        double piall[8], pall[8];
#pragma unroll
        for (int l = 0; l < 8; l++) { piall[l] = piv[l]; pall[l] = pv[l]; }
        double ru = 0.0;
        if (a < c.nh) {
            ru = c.hd[c.hGH + a];
            for (int b = 0; b < c.nh; b++) ru += c.hd[c.hHH + a * c.nh + b] * c.wk[(size_t)(b >> 1) * STG + W_DU + (b & 1)];
            ru += head_GtV(c, piall, a) - head_rows_dot(c, PS_LD, a);
            c.hs[hs_ru(c) + a] = ru;
        }
        rstat = nmax(rstat, warp_nmax(fabs(ru)));
        head_factor(c, true);
        head_solve(c, ru - head_rows_dot(c, PS_CF, a) + head_GtV(c, pall, a));
    }
    __syncwarp();
    return rstat;
}

#ifndef QP_DMMA
#define QP_DMMA 1      // 0: the full backward sweep on the FP64 vector pipe (sweep_backward_full); 1: on the FP64 tensor cores
#endif                 // (sweep_backward_tensor).  A/B on B200: profiles/r02_ab_dmma*.jsonl

// The same sweep with its three matrix products on the FP64 TENSOR CORES (mma.sync m8n8k4, DMMA).  With
// M = [A(:, 2..7) | B] (8 x 8: exactly the 48 stored doubles of the record plus the unit / dt rows 6, 7; the unit columns
// 0, 1 of A need no product) one stage is
//   phase A   T1 = P M              -> P A (columns 2..7) and P B         (PHASE0: + P b for v = p + P b)
//   phase B   T2 = M' (P A)         -> rows 2..7 of A'PA and B'PA
//             T3 = M' [P B | v | pi] -> B'PB, A'v, B'v, A'pi, B'pi          (v = p, or p + P b in phase 0)
//   phase C   as in sweep_backward_full (2x2 Cholesky, gains, P_k = A'PA - Y'Y + Q~, p_k, pi_k) on the (row, column pair)
//             lanes, which pick their entries of A'PA up from the P tile (dead between phase A and the end of phase C).
// Fragment layout of DMMA: lane (fg = lane / 4, ft = lane % 4) holds A[fg][ft], B[ft][fg], D[fg][2 ft .. 2 ft + 1]; the operand
// M[k][fg] is the B operand of T1 AND the A operand of T2, T3 (M' [m = fg][k] = M[k][fg]): loaded once per stage.  Each
// product is two chained DMMA (k = 0..3, 4..7).  Against the vector-pipe version this removes per lane and stage 38 DFMA, 13
// 128-bit and 16 64-bit shared-memory loads and every per-lane operand descriptor (6 DMMA, 8 64-bit loads instead); the
// results differ from it in the last bits only (summation order).  Measured: +4.7 % solves/s at B = 4096, N = 30.
template <bool HEAD, bool PHASE0>
__device__ double sweep_backward_tensor(const DevCfg &cfg, const DevMem &m, const Ctx &c, double ap = 0.0)
{
    const int N = c.N, i = c.lane, q = c.lane32 >> 3, L = c.lane32;
    double *Pt = c.sm + SM_PT, *PAt = c.sm + SM_PA, *sml = c.sm + SM_SML, *pv = c.sm + SM_PV, *piv = pv + 8;
    double *PBs = sml, *Huxs = sml + 16, *hxs = sml + 32, *Hs = sml + 40, *hus = sml + 44;
    double *vv = PHASE0 ? c.sm + SM_X : pv;                 // phase 0: v = p + P b of the stage (exchange slots are free here)
    double *atp = c.sm + SM_X + 16;                        // A' pi of the stage (second exchange slot)
    const double *zero6 = sml + 48;
    const double dt = c.dt;
    double rstat = 0.0;
    const int kstop = HEAD ? c.H : 0;
    const int c0 = 2 * q;                                  // first of this lane's two columns (phase C)
    // ---- loop-invariant fragment descriptors
    const int fg = L >> 2, ft = L & 3;
    const int m_off = L_A + fg * 6 + ft;                   // M[k = ft][n = fg] (k-step 0: rows 0..3 are stored rows)
    const bool m1_ld = ft < 2;                             // k-step 1: rows 4, 5 stored; rows 6, 7 structural
    const double m1_c = (ft == 2) ? ((fg == 4) ? 1.0 : (fg == 7) ? dt : 0.0) : ((fg == 5) ? 1.0 : (fg == 6) ? dt : 0.0);
    double *t1_out = (ft < 3) ? PAt + fg * TLD + 2 + 2 * ft : PBs + 2 * fg;
    double *t2_out = (fg < 6) ? Pt + (2 + fg) * TLD + 2 * ft : Huxs + (fg - 6) * 8 + 2 * ft;
    // B operand of T3: column fg of [P B | v | pi | 0 ...]
    const double *z_ptr = (fg < 2) ? PBs + fg : (fg == 2) ? vv : (fg == 3 && !PHASE0) ? piv : zero6;
    const int z_s = (fg < 2) ? 2 : (fg == 2 || (fg == 3 && !PHASE0)) ? 1 : 0;
    const int e = fg - 6;                                  // input index of the lanes that hold rows 6, 7 of T3
    // Q~ selectors of this lane's two entries (phase C)
    const double s_da = (i == c0) ? 1.0 : 0.0, s_db = (i == c0 + 1) ? 1.0 : 0.0, s_bx = (i == 6 && q == 3) ? 1.0 : 0.0;
    const double s_v0b = (i == 3 && q == 1) ? 1.0 : 0.0, s_v1a = (i == 3 && q == 2) ? 1.0 : 0.0;
    const double s_v1b = (i == 4 && q == 1) ? 1.0 : 0.0, s_v2a = (i == 4 && q == 2) ? 1.0 : 0.0;
    const double s_g3 = (i == 3) ? 1.0 : 0.0, s_g4 = (i == 4) ? 1.0 : 0.0, s_i6 = (i == 6) ? 1.0 : 0.0;
    const int qx_i = (i < 5) ? i : 0;
    const double s_qx = (i < 5) ? 1.0 : 0.0;

    int slot = 0;
    pipe_begin<R_BACK_LO, R_BACK_HI>(c, N, -1, kstop);
    for (int k = N; k >= kstop; k--, slot = (slot + 1) % PRING) {
        pipe_advance<true, R_BACK_LO, R_BACK_HI>(c, k, -1, kstop, slot);      // (every stage visit below ends with a warp barrier)
        const double *rec = pipe_wait(c, slot, -1);
        double *const psk = c.PSK(k);
        const double *w = rec;
        double qdj, rd0, rd1;
        cost_diag(cfg, c, k, qdj, rd0, rd1);
        const double gj = rec[L_GF + i];
        const double2 gc = *reinterpret_cast<const double2 *>(rec + L_GF + c0);
        double om_bu = 0.0, om_bx = 0.0, om_h = 0.0, cf_bu = 0.0, cf_bx = 0.0, cf_h = 0.0, ld_bu = 0.0, ld_bx = 0.0, ld_h = 0.0;
        if (!PHASE0) {
            om_bu = psk[PS_OM + 0]; om_bx = psk[PS_OM + 1]; om_h = psk[PS_OM + 2];
            cf_bu = psk[PS_CF + 0]; cf_bx = psk[PS_CF + 1]; cf_h = psk[PS_CF + 2];
            ld_bu = psk[PS_LD + 0]; ld_bx = psk[PS_LD + 1]; ld_h = psk[PS_LD + 2];
        }
        // x-stationarity row i at the current iterate, without the costate term; the pending primal step of the previous
        // iteration is applied here from the shared-memory copy of the record and written back (see sweep_backward_full)
        const bool lazy = !PHASE0 && ap != 0.0;
        double dxi = PHASE0 ? 0.0 : w[W_DX + i];
        if (lazy) {
            dxi += ap * w[W_DDX + i];
            if (q == 0) c.wk[(size_t)k * STG + W_DX + i] = dxi;
        }
        double gxj = PHASE0 ? 0.0 : qdj * dxi + s_qx * rec[L_QX + qx_i] - s_i6 * ld_bx - ld_h * gj;
        // Q~[i][c0], Q~[i][c0+1]: cost diagonal, soft-row barrier terms, |v| block
        const double ohg = om_h * gj;
        double qa = ohg * gc.x + s_da * qdj + s_bx * om_bx, qb = ohg * gc.y + s_db * qdj;
        if (cfg.cost_vel_abs) {
            const double Qv0 = rec[L_QV], Qv1 = rec[L_QV + 1], Qv2 = rec[L_QV + 2];
            const double dx3 = w[W_DX + 3] + (lazy ? ap * w[W_DDX + 3] : 0.0), dx4 = w[W_DX + 4] + (lazy ? ap * w[W_DDX + 4] : 0.0);
            if (!PHASE0) gxj += s_g3 * (Qv0 * dx3 + Qv1 * dx4) + s_g4 * (Qv1 * dx3 + Qv2 * dx4);
            qa += s_v1a * Qv1 + s_v2a * Qv2; qb += s_v0b * Qv0 + s_v1b * Qv1;
        }
        const double qtj = PHASE0 ? s_qx * rec[L_QX + qx_i] : -(s_i6 * cf_bx + cf_h * gj);
        if (k == N) {
            *reinterpret_cast<double2 *>(Pt + i * TLD + c0) = make_double2(qa, qb);
            if (q == 0) { pv[i] = qtj; piv[i] = gxj; }
            __syncwarp();
            continue;
        }
        double du0 = PHASE0 ? 0.0 : w[W_DU], du1 = PHASE0 ? 0.0 : w[W_DU + 1];
        if (lazy) {
            const double *q = c.SWL(w, psk, W_DDU, PS_DDU);
            du0 += ap * q[0]; du1 += ap * q[1];
            if (L == 0) { c.wk[(size_t)k * STG + W_DU] = du0; c.wk[(size_t)k * STG + W_DU + 1] = du1; }
        }
        // ---------------- phase A: T1 = P M
        const double mb0 = rec[m_off], mb1 = m1_ld ? rec[m_off + 4] : m1_c;
        {
            const double pa0 = Pt[fg * TLD + ft], pa1 = Pt[fg * TLD + 4 + ft];
            double d0 = 0.0, d1 = 0.0;
            dmma884(d0, d1, pa0, mb0);
            dmma884(d0, d1, pa1, mb1);
            *reinterpret_cast<double2 *>(t1_out) = make_double2(d0, d1);
            if (ft == 0) *reinterpret_cast<double2 *>(PAt + fg * TLD) = *reinterpret_cast<const double2 *>(Pt + fg * TLD);
            if (PHASE0) {                                    // v = p + P b: b as column 0 of a second B operand
                const double bb0 = (fg == 0) ? rec[L_b + ft] : 0.0, bb1 = (fg == 0) ? rec[L_b + 4 + ft] : 0.0;
                double v0 = 0.0, v1 = 0.0;
                dmma884(v0, v1, pa0, bb0);
                dmma884(v0, v1, pa1, bb1);
                if (ft == 0) vv[fg] = pv[fg] + v0;
            }
        }
        __syncwarp();
        // ---------------- phase B: T2 = M' (P A), T3 = M' [P B | v | pi]
        {
            const double pb0 = PAt[ft * TLD + fg], pb1 = PAt[(4 + ft) * TLD + fg];
            const double zb0 = z_ptr[ft * z_s], zb1 = z_ptr[(4 + ft) * z_s];
            double d0 = 0.0, d1 = 0.0, z0 = 0.0, z1 = 0.0;
            dmma884(d0, d1, mb0, pb0);
            dmma884(z0, z1, mb0, zb0);
            dmma884(d0, d1, mb1, pb1);
            dmma884(z0, z1, mb1, zb1);
            // rows 0..5 of T2: rows 2..7 of A'PA into the (dead) P tile; rows 6, 7: B'PA
            *reinterpret_cast<double2 *>(t2_out) = make_double2(d0, d1);
            if (ft == 1) {
                if (fg < 6) {                                // (A'v)_{2+fg}, (A'pi)_{2+fg}
                    hxs[2 + fg] = z0; atp[2 + fg] = z1;
                } else {                                     // (B'v)_e, (B'pi)_e: u-stationarity residual and feed-forward right-hand side
                    const double rde = e ? rd1 : rd0, due = e ? du1 : du0;
                    const double gu = PHASE0 ? rec[L_RG + e] : rec[L_RG + e] + rde * due + z1 - (e ? ld_bu : 0.0);
                    hus[e] = z0 + gu - (e ? cf_bu : 0.0);
                    c.SWS(k, W_GUS, PS_GUS)[e] = gu;
                    rstat = fmax(rstat, fabs(gu));           // a NaN here also poisons the side variables and is caught there
                }
            } else if (ft == 0 && fg >= 6) {                 // B'PB + R~
                if (e == 0) { Hs[0] = z0 + rd0; Hs[1] = z1; } else { Hs[2] = z1 + rd1 + om_bu; }
            } else if (ft == 2 && fg >= 6) {                 // unit rows 0, 1 of A': (A'v)_e = v_e, (A'pi)_e = pi_e
                hxs[e] = vv[e]; atp[e] = PHASE0 ? 0.0 : piv[e];
            }
        }
        __syncwarp();
        // ---------------- phase C: 2x2 Cholesky through two reciprocal square roots, gains, P_k, p_k, pi_k.
        // P_k = A'PA - Y'Y with Y = L^-1 (B'PA): the subtraction is bitwise symmetric in (i, c), which keeps P symmetric
        // under the 1e9 barrier weights (a gain-form update P + Hux' K drifts and loses definiteness there).
        const double2 hx = *reinterpret_cast<const double2 *>(((i < 2) ? PAt : Pt) + i * TLD + c0);   // rows 0, 1 of A'PA = rows 0, 1 of P A
        const double h00 = Hs[0], h01 = Hs[1], h11 = Hs[2], hu0 = hus[0], hu1 = hus[1];
        const double i11 = fast_rsqrt(h00), l21 = h01 * i11;
        const double i22 = fast_rsqrt(h11 - l21 * l21);
        const double2 ua = *reinterpret_cast<const double2 *>(Huxs + c0), ub = *reinterpret_cast<const double2 *>(Huxs + 8 + c0);
        const double y0a = ua.x * i11, y0b = ua.y * i11;
        const double y1a = (ub.x - l21 * y0a) * i22, y1b = (ub.y - l21 * y0b) * i22;
        const double y0i = Huxs[i] * i11, y1i = (Huxs[8 + i] - l21 * y0i) * i22;
        const double pn0 = hx.x - (y0i * y0a + y1i * y1a) + qa, pn1 = hx.y - (y0i * y0b + y1i * y1b) + qb;
        const double K1 = -y1i * i22;
        const double K0 = (-y0i - l21 * K1) * i11;
        const double pnew = hxs[i] + qtj + K0 * hu0 + K1 * hu1;
        if (!PHASE0) gxj += atp[i];
        // (each lane overwrites exactly the entry of the P tile it has just read; the vectors of stage k+1 were last read
        // before the phase-B barrier)
        *reinterpret_cast<double2 *>(Pt + i * TLD + c0) = make_double2(pn0, pn1);
        if (q == 0) {
            pv[i] = pnew; piv[i] = gxj;
            double *wg = c.wk + (size_t)k * STG + W_K;
            wg[i] = K0; wg[8 + i] = K1;
        }
        {   // every lane computes, lane 31 stores (predicated stores instead of a divergent block)
            const double m21 = -l21 * i11 * i22;
            const double Hi0 = i11 * i11 + m21 * m21, Hi1 = m21 * i22, Hi3 = i22 * i22;
            const double kf0 = -(Hi0 * hu0 + Hi1 * hu1), kf1 = -(Hi1 * hu0 + Hi3 * hu1);
            if (L == 31) {
                double *wg = c.wk + (size_t)k * STG;
                wg[W_HI] = Hi0; wg[W_HI + 1] = Hi1; wg[W_HI + 2] = Hi3;
                double *qk = c.SWS(k, W_KFF, PS_KFF);
                qk[0] = kf0; qk[1] = kf1;
            }
        }
        __syncwarp();
    }
    rstat = warp_nmax(rstat);                                // (held by the two lanes of rows 6, 7 of T3)
    if (HEAD && PHASE0) {
        // gradient of the cost-to-go at dx_H = c_H: v = p_H + P_H c_H
        const int a = c.lane32;
        double vj = pv[i];
#pragma unroll
        for (int l = 0; l < 8; l++) vj += Pt[i * TLD + l] * c.hd[c.hC + (c.H - 1) * 8 + l];
        double vall[8];
        xchg1(c, 0, vj, vall);
        head_factor(c, false);
        head_solve(c, (a < c.nh ? c.hd[c.hGH + a] : 0.0) + head_GtV(c, vall, a));
    } else if (HEAD) {
        const int a = c.lane32;
        double piall[8], pall[8];
#pragma unroll
        for (int l = 0; l < 8; l++) { piall[l] = piv[l]; pall[l] = pv[l]; }
        double ru = 0.0;
        if (a < c.nh) {
            ru = c.hd[c.hGH + a];
            for (int b = 0; b < c.nh; b++) ru += c.hd[c.hHH + a * c.nh + b] * c.wk[(size_t)(b >> 1) * STG + W_DU + (b & 1)];
            ru += head_GtV(c, piall, a) - head_rows_dot(c, PS_LD, a);
            c.hs[hs_ru(c) + a] = ru;
        }
        rstat = nmax(rstat, warp_nmax(fabs(ru)));
        head_factor(c, true);
        head_solve(c, ru - head_rows_dot(c, PS_CF, a) + head_GtV(c, pall, a));
    }
    __syncwarp();
    return rstat;
}

// the backward sweep the kernel runs: the tensor-core version in both instantiations of the nominal kernel, so that a batch
// gives bit-identical results whatever its size and however it is sharded (the small-batch latency instantiation alone would
// be 5 % faster on the vector pipe: one partial wave is pure dependent latency, two chained DMMA = 52 cycles + fragment round
// trips against 6 chained DFMA; profiles/r02_batch_sweep.jsonl: B = 1 cold 0.80 vs 0.85 ms, from one full wave on the tensor
// version wins: B = 2368 1.42 vs 1.32 M solves/s).  The SNMPC head kernel, a different problem class that never runs more than
// one partial wave at the BASELINE size, keeps the vector-pipe version (2.5 % faster there).
template <bool HEAD, bool PHASE0, bool TENSOR>
__device__ __forceinline__ double sweep_backward(const DevCfg &cfg, const DevMem &m, const Ctx &c, double ap = 0.0)
{
#if QP_DMMA >= 1
    if (TENSOR) return sweep_backward_tensor<HEAD, PHASE0>(cfg, m, c, ap);
#endif
    return sweep_backward_full<HEAD, PHASE0>(cfg, m, c, ap);
}

// Corrector: backward vector sweep with the stored factorisation (cf per stage from shared memory).
template <bool HEAD>
__device__ void sweep_corrector(const Ctx &c)
{
    const int N = c.N;
    double pj = 0.0;
    const int kstop = HEAD ? c.H : 0;
    int slot = 0;
    pipe_begin<R_CORR_LO, R_CORR_HI>(c, N, -1, kstop);
    for (int k = N; k >= kstop; k--, slot = (slot + 1) % PRING) {
        // (the barrier inside pipe_wait also frees the exchange slot of the previous stage: xchg1<true> below.  Waiting first
        // and refilling the other slot after ONE barrier per stage was measured too: +5 % for the SNMPC head kernel, -1 % at
        // B = 4096 where the later prefetch costs more than the barrier)
        pipe_advance<false, R_CORR_LO, R_CORR_HI>(c, k, -1, kstop, slot);
        const double *rec = pipe_wait(c, slot, -1);
        const unsigned OM = c.vmask;
        if (OM != FULLMASK && !c.st) continue;             // octet 0 only (see OCTET0); the others wait at the next stage's barrier
        double *const psk = c.PSK(k);
        const double *w = rec;
        const double gj = rec[L_GF + c.lane];
        const double cf_bu = psk[PS_CF + 0], cf_bx = psk[PS_CF + 1], cf_h = psk[PS_CF + 2];
        const double qtj = -(((c.lane == 6) ? cf_bx : 0.0) + cf_h * gj);
        if (k == N) { pj = qtj; continue; }
        double pall[8], hu0, hu1;
        double B0[6], B1[6];
        xchg1<true>(c, 0, pj, pall, OM);
        load_B(rec, B0, B1);
        Bt_dot(B0, B1, c.dt, pall, hu0, hu1);      // (B'p by a shuffle butterfly like the forward sweeps' K ddx was measured too:
                                                   // -0.7 % at B = 4096, -1.4 % for the SNMPC head kernel -- the longer dependent
                                                   // path costs more than the 8 pipe cycles it saves; profiles/r02_ab_forward_shuffle.jsonl)
        { const double *q = c.SWL(w, psk, W_GUS, PS_GUS); hu0 += q[0]; hu1 += q[1] - cf_bu; }
        const double Hi0 = w[W_HI], Hi1 = w[W_HI + 1], Hi3 = w[W_HI + 2];
        const double kf0 = -(Hi0 * hu0 + Hi1 * hu1), kf1 = -(Hi1 * hu0 + Hi3 * hu1);
        if (c.lane32 == 0) { double *qk = c.SWS(k, W_KFF, PS_KFF); qk[0] = kf0; qk[1] = kf1; }
        // (at k = 0 of the nominal kernel the costate is not needed any more: computed anyway, a branch costs more)
        pj = At_comp(c, rec, pall, pj) + qtj + w[W_K + c.lane] * hu0 + w[W_K + 8 + c.lane] * hu1;
    }
    if (HEAD) {
        const int a = c.lane32;
        double pall[8];
        xchg1(c, 0, pj, pall);
        const double ru = (a < c.nh) ? c.hs[hs_ru(c) + a] : 0.0;
        head_solve(c, ru - head_rows_dot(c, PS_CF, a) + head_GtV(c, pall, a));
    }
    __syncwarp();
}

// Forward roll-out of the Newton direction: ddu_k = kff_k + K_k ddx_k, ddx_{k+1} = A ddx + B ddu; the
// direction row values dv (bu, bx, h) and ddu go to shared memory for the stage-parallel phases; with
// STORE the direction ddx_k is kept in the WORK block for the stage-parallel update.
template <bool STORE, bool HEAD>
__device__ void sweep_forward_dir(const DevMem &m, const Ctx &c)
{
    const int N = c.N;
    double ddxj = 0.0;
    int k0 = 0;
    if (HEAD) { ddxj = head_apply<false, STORE>(m, c); k0 = c.H; }
    int slot = 0;
    pipe_begin<R_FWD_LO, R_FWD_HI>(c, k0, 1, N);
    for (int k = k0; k <= N; k++, slot = (slot + 1) % PRING) {
        pipe_advance<false, R_FWD_LO, R_FWD_HI>(c, k, 1, N, slot);
        const double *rec = pipe_wait(c, slot, 1);          // (its barrier also frees the exchange slots of the previous stage)
        const unsigned OM = c.vmask;
        if (OM != FULLMASK && !c.st) continue;             // octet 0 only (see OCTET0)
        double *const psk = c.PSK(k);
        const double *w = rec;
        if (STORE && c.st) c.wk[(size_t)k * STG + W_DDX + c.lane] = ddxj;
        double x[8], ddu0 = 0.0, ddu1 = 0.0;
        if (k < N && k > 0) {
#ifndef FWD_XCHG_PRODUCTS
            // K ddx by a shuffle butterfly inside the octet (same summation tree as sum8: bit-identical), only ddx goes through
            // shared memory.  The kernel is bound by the load/store pipe (72 % busy), which delivers 128 bytes per cycle to the
            // registers whether or not the four octets read the same data: gathering the two product vectors cost 32 of
            // the ~90 pipe cycles of a stage visit, the six 64-bit shuffles cost 12 (-DFWD_XCHG_PRODUCTS: the exchange form;
            // B = 4096: 2.154 -> 2.100 ms, profiles/r02_ab_forward_shuffle.jsonl)
            double p0 = w[W_K + c.lane] * ddxj, p1 = w[W_K + 8 + c.lane] * ddxj;
            xchg1<true>(c, 0, ddxj, x, OM);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { p0 += __shfl_xor_sync(OM, p0, o); p1 += __shfl_xor_sync(OM, p1, o); }
            const double *qk = c.SWL(w, psk, W_KFF, PS_KFF);
            ddu0 = qk[0] + p0;
            ddu1 = qk[1] + p1;
#else
            double k0[8], k1[8];
            // one exchange carries ddx and the two partial-product vectors of K ddx
            double *buf = c.sm + SM_X;
            buf[c.lane] = ddxj; buf[8 + c.lane] = w[W_K + c.lane] * ddxj; buf[16 + c.lane] = w[W_K + 8 + c.lane] * ddxj;
            __syncwarp();
            const double2 *b2 = reinterpret_cast<const double2 *>(buf);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double2 a = b2[i], b = b2[4 + i], cc = b2[8 + i];
                x[2 * i] = a.x; x[2 * i + 1] = a.y; k0[2 * i] = b.x; k0[2 * i + 1] = b.y; k1[2 * i] = cc.x; k1[2 * i + 1] = cc.y;
            }
            const double *qk = c.SWL(w, psk, W_KFF, PS_KFF);
            ddu0 = qk[0] + sum8(k0);
            ddu1 = qk[1] + sum8(k1);
#endif
        } else {
            xchg1<true>(c, 0, ddxj, x, OM);
            if (k < N) { const double *qk = c.SWL(w, psk, W_KFF, PS_KFF); ddu0 = qk[0]; ddu1 = qk[1]; }
        }
        {
            const double *g = rec + L_GH;
            const double dvh = g[0] * x[3] + g[1] * x[4] + g[2] * x[5] + g[3] * x[7];      // every lane computes, lane 0 stores
            if (c.lane32 == 0) {
                psk[PS_DV + 0] = ddu1; psk[PS_DV + 1] = x[6]; psk[PS_DV + 2] = dvh;
                double *qd = c.SWS(k, W_DDU, PS_DDU); qd[0] = ddu0; qd[1] = ddu1;
            }
        }
        if (k < N) ddxj = dyn_comp(c, rec, x, ddxj, ddu0, ddu1);
    }
    __syncwarp();
}

// Phase 0 roll-out: dx_0 = x0 - x_0, du_k = kff_k + K_k dx_k, dx_{k+1} = A dx + B du + b; stores the QP
// iterate and the row values v (bu, bx, h) of every stage.
template <bool HEAD>
__device__ void sweep_rollout0(const DevMem &m, const Ctx &c)
{
    const int N = c.N;
    double dxj;
    int k0 = 0;
    if (HEAD) { dxj = head_apply<true, false>(m, c); k0 = c.H; }
    else dxj = m.x0[(size_t)c.b * 8 + c.lane] - m.X[(size_t)c.b * c.NS * 8 + c.lane];
    int slot = 0;
    pipe_begin<R_ROLL_LO, R_ROLL_HI>(c, k0, 1, N);
    for (int k = k0; k <= N; k++, slot = (slot + 1) % PRING) {
        pipe_advance<false, R_ROLL_LO, R_ROLL_HI>(c, k, 1, N, slot);
        // the iterate rows used by the row values: issued before the wait so that both latencies overlap
        const double u1k = (k < N) ? m.U[((size_t)c.b * N + k) * 2 + 1] : 0.0;
        const double x6k = m.X[((size_t)c.b * c.NS + k) * 8 + 6];
        const double *rec = pipe_wait(c, slot, 1);
        double *const psk = c.PSK(k);
        const double *w = rec;
        double *wg = c.wk + (size_t)k * STG;
        if (c.st) wg[W_DX + c.lane] = dxj;
        double x[8], du0 = 0.0, du1 = 0.0;
        if (k < N) {
#ifndef FWD_XCHG_PRODUCTS
            double p0 = w[W_K + c.lane] * dxj, p1 = w[W_K + 8 + c.lane] * dxj;
            xchg1(c, 0, dxj, x);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { p0 += __shfl_xor_sync(FULLMASK, p0, o); p1 += __shfl_xor_sync(FULLMASK, p1, o); }
            const double *qk = c.SWL(w, psk, W_KFF, PS_KFF);
            du0 = qk[0] + p0;
            du1 = qk[1] + p1;
#else
            double k0[8], k1[8];
            double *buf = c.sm + SM_X;
            __syncwarp();
            buf[c.lane] = dxj; buf[8 + c.lane] = w[W_K + c.lane] * dxj; buf[16 + c.lane] = w[W_K + 8 + c.lane] * dxj;
            __syncwarp();
            const double2 *b2 = reinterpret_cast<const double2 *>(buf);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double2 a = b2[i], b = b2[4 + i], cc = b2[8 + i];
                x[2 * i] = a.x; x[2 * i + 1] = a.y; k0[2 * i] = b.x; k0[2 * i + 1] = b.y; k1[2 * i] = cc.x; k1[2 * i + 1] = cc.y;
            }
            const double *qk = c.SWL(w, psk, W_KFF, PS_KFF);
            du0 = qk[0] + sum8(k0);
            du1 = qk[1] + sum8(k1);
#endif
            if (c.lane32 == 0) { wg[W_DU] = du0; wg[W_DU + 1] = du1; }
        } else {
            xchg1(c, 0, dxj, x);
        }
        if (c.lane32 == 0) {
            const double *g = rec + L_GH;
            psk[PS_ROWV + 0] = (k < N) ? u1k + du1 : 0.0;
            psk[PS_ROWV + 1] = x6k + x[6];
            psk[PS_ROWV + 2] = rec[L_H] + g[0] * x[3] + g[1] * x[4] + g[2] * x[5] + g[3] * x[7];
        }
        if (k < N) dxj = dyn_comp(c, rec, x, dxj, du0, du1) + rec[L_b + c.lane];
    }
    __syncwarp();
}

// (The vector sweeps -- forward roll-outs, corrector -- were also built on the tensor cores: a matrix-vector product as a DMMA
// with the vector as the row-replicated A operand and the contraction ordered even / odd components, which keeps the state of
// the recursion in two registers per lane with no exchange at all.  Measured on B200 it does NOT pay: 5 DMMA per stage are 80
// cycles of the FP64 pipe for one warp against ~40 for the octet version, and three chained DMMA are a longer dependent path
// than the gather + 7 DFMA they replace: same 1.49 M solves/s at B = 4096, N = 30, -4 % at N = 20, -8 % for the SNMPC
// head kernel (profiles/r02_ab_dmma_vector_sweeps.jsonl, variant tensor2).  Removed again; the full backward sweep, where the
// products are 8x8x8, is where the tensor cores win.)

// ------------------------------------------------------------------------------------------------------
// Stage-parallel phases: item i = k*6 + row*2 + side over all 32 lanes
// ------------------------------------------------------------------------------------------------------
struct Item { int k, row; double sgn; bool valid; };
__device__ __forceinline__ Item item_of(int i, int nitems)
{
    Item it;
    it.valid = i < nitems;
    const int ii = it.valid ? i : 0;
    it.k = ii / 6;
    const int rs = ii - it.k * 6;
    it.row = rs >> 1;
    it.sgn = (rs & 1) ? -1.0 : 1.0;
    return it;
}

// soft-row constants per side: bound, sigma-scaled penalties, activity (stage 0: bu only; 1..N-1: all; N: bx, h)
__device__ int phase_prep(const DevCfg &cfg, const DevMem &m, Ctx &c)
{
    static_assert((B200NMPC_MAX_N + 1) * 6 <= 32 * 32, "activity mask: one bit per round of 32 items");
    c.actmask = 0u;
    const int nitems = c.NS * 6;
    double nact2 = 0.0;
    // sigma-scaled cost diagonal of every stage (dt on the path stages, 1 on the terminal stage; no input at N)
    for (int i = c.lane32; i < nitems; i += 32) {
        const int k = i / 6, cc = i - k * 6;
        const double sig = (k < c.N) ? c.dt : 1.0;
        c.PS(PS_W, k, cc) = (k == c.N && cc >= 4) ? 0.0 : sig * sval(m.W_s, m.W_i, c, k, 6, cc);
    }
    for (int base = 0; base < nitems; base += 32) {
        const int i = base + c.lane32;
        const Item it = item_of(i, nitems);
        if (!it.valid) continue;
        const bool upper = it.sgn < 0.0;
        const double sig = (it.k < c.N) ? c.dt : 1.0;
        const double zl = sig * sval(m.zl_s, m.zl_i, c, it.k, 3, it.row), Zl = sig * sval(m.Zl_s, m.Zl_i, c, it.k, 3, it.row);
        const double zu = sig * sval(m.zu_s, m.zu_i, c, it.k, 3, it.row), Zu = sig * sval(m.Zu_s, m.Zu_i, c, it.k, 3, it.row);
        const bool exists = (it.row == 0) ? (it.k < c.N) : (it.k >= 1);
        const bool act = exists && (zl + Zl > 0.0) && (zu + Zu > 0.0);      // a row without penalty is vacuous
        c.S(S_BND)[i] = upper ? sval(m.ub_s, m.ub_i, c, it.k, 3, it.row) : sval(m.lb_s, m.lb_i, c, it.k, 3, it.row);
        c.P(P_Z)[i] = upper ? make_double2(zu, Zu) : make_double2(zl, Zl);
        c.S(S_ACT)[i] = act ? 1.0 : 0.0;
        if (act) c.actmask |= 1u << (base >> 5);
        nact2 += act ? 1.0 : 0.0;
    }
    __syncwarp();
    return (int)(warp_sum(nact2) * 0.5 + 0.25);
}

// feasibility of the equality-constrained optimum and the interior start (zero linear residuals:
// s_l = s_u = S, lam_l = lam_u, so the phase-0 point stays stationary)
__device__ bool phase_init(const Ctx &c)
{
    const int nitems = c.NS * 6;
    double infeas = 0.0;
    for (int base = 0; base < nitems; base += 32) {
        const int i = base + c.lane32;
        const Item it = item_of(i, nitems);
        const int ic = it.valid ? i : 0;
        const double2 zZ = c.P(P_Z)[ic];
        const double bnd = c.S(S_BND)[ic], z = zZ.x, Z = zZ.y;
        const bool act = (c.actmask >> (base >> 5)) & 1u;      // implies it.valid
        const double bo = __shfl_xor_sync(FULLMASK, bnd, 1), zo = __shfl_xor_sync(FULLMASK, z, 1), Zo = __shfl_xor_sync(FULLMASK, Z, 1);
        if (!act) continue;
        const double v = c.PS(PS_ROWV, it.k, it.row);
        const double lo = it.sgn > 0.0 ? bnd : bo, up = it.sgn > 0.0 ? bo : bnd;
        if (v < lo || v > up) infeas = 1.0;
        const double viol = fmax(fmax(lo - v, v - up), 0.0);
        const double S = c.s0 + viol;
        const double gme = z + Z * S, goth = zo + Zo * S;
        const double l0 = 0.5 * fmin(gme, goth);
        c.P(P_ST)[i] = make_double2(S, it.sgn * (v - bnd) + S);
        c.P(P_LM)[i] = make_double2(l0, gme - l0);
    }
    __syncwarp();
    return warp_sum(infeas) == 0.0;
}

// Side data of one item, loaded one round AHEAD of its use: the six rounds of a stage-parallel pass are independent, but
// the compiler cannot move the loads of round r+1 above the stores of round r, so every round used to expose one L2
// round trip.  Loads are unconditional at a clamped index (inactive items read valid memory and are discarded).
template <bool PR, bool D>
struct SideIn {
    double2 st, lm, zZ, pr, d;
    double bnd;
    __device__ __forceinline__ void load(const Ctx &c, int base, int nitems)
    {
        const int i = base + c.lane32, ic = (i < nitems) ? i : 0;
        st = c.P(P_ST)[ic]; lm = c.P(P_LM)[ic]; zZ = c.P(P_Z)[ic]; bnd = c.S(S_BND)[ic];
        if (PR) pr = c.P(P_PR)[ic];
        if (D) d = c.P(P_D)[ic];
    }
};

// MODE 0: residual measures + predictor coefficients (om, cf, ld per row -> shared memory)
// MODE 1: corrector coefficients (cf per row -> shared memory)
// MODE 2: apply the step of length alpha (= sigma_mu argument) with the stored directions, then MODE 0 at the new iterate
//         (one pass over the side arrays instead of two)
template <int MODE>
__device__ void phase_coeffs(const Ctx &c, double sigma_mu, double corr, int nact, double *gap, double *cmax, double *rmax)
{
    const int nitems = c.NS * 6;
    double g = 0.0, cm = 0.0, rm = 0.0;
    SideIn<MODE != 0, MODE == 2> nx;
    nx.load(c, 0, nitems);
    for (int base = 0; base < nitems; base += 32) {
        const int i = base + c.lane32;
        const Item it = item_of(i, nitems);
        const bool act = (c.actmask >> (base >> 5)) & 1u;      // implies it.valid
        const SideIn<MODE != 0, MODE == 2> in = nx;
        if (base + 32 < nitems) nx.load(c, base + 32, nitems);
        double om_s = 0.0, cc_s = 0.0, ld_s = 0.0;
        if (act) {
            const double2 st = in.st, lm = in.lm, zZ = in.zZ;
            double s = st.x, t = st.y, lam = lm.x, mu = lm.y;
            if (MODE == 2) {
                const double alpha = sigma_mu;
                const double2 d12 = in.pr, d34 = in.d;
                s += alpha * d12.x; t += alpha * d12.y; lam += alpha * d34.x; mu += alpha * d34.y;
                c.P(P_ST)[i] = make_double2(s, t); c.P(P_LM)[i] = make_double2(lam, mu);
            }
            const double bnd = in.bnd, z = zZ.x, Z = zZ.y;
            const double v = c.PS(PS_ROWV, it.k, it.row);
            const double rt = t - (it.sgn * (v - bnd) + s);
            const double rs = z + Z * s - lam - mu;
            double r1 = -lam * t, r2 = -mu * s;
            if (MODE != 1) {
                g += lam * t + mu * s;
                cm = nmax(cm, nmax(lam * t, mu * s));
                rm = nmax(rm, nmax(fabs(rt), fabs(rs)));
            } else {
                const double2 pr = in.pr;
                r1 += sigma_mu - corr * pr.x;
                r2 += sigma_mu - corr * pr.y;
            }
            const SideK sk = side_coeffs(c.reg, Z, s, t, lam, mu, rt, rs, r1, r2);
            om_s = sk.om; cc_s = it.sgn * sk.cc; ld_s = it.sgn * lam;
        }
        // the two sides of a row sit on adjacent lanes
        const double cf = cc_s + __shfl_xor_sync(FULLMASK, cc_s, 1);
        if (MODE != 1) {
            const double om = om_s + __shfl_xor_sync(FULLMASK, om_s, 1);
            const double ld = ld_s + __shfl_xor_sync(FULLMASK, ld_s, 1);
            if (it.valid && it.sgn > 0.0) { c.PS(PS_OM, it.k, it.row) = om; c.PS(PS_LD, it.k, it.row) = ld; }
        }
        if (it.valid && it.sgn > 0.0) c.PS(PS_CF, it.k, it.row) = cf;
    }
    if (MODE != 1) {
        *gap = warp_sum(g) / (4.0 * nact);
        *cmax = warp_nmax(cm);
        *rmax = warp_nmax(rm);
    }
    __syncwarp();
}

// slack / multiplier directions of every side from the direction row values dv.
// MODE 0 (predictor): step to the boundary, complementarity polynomial, second-order products stored.
// MODE 1 (corrector): step to the boundary only.      MODE 2: apply the step of length alpha.
template <int MODE>
__device__ double phase_dirs(const Ctx &c, double sigma_mu, double corr, double alpha, double *poly)
{
    const int nitems = c.NS * 6;
    double rnum = 0.0, rden = 1.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;      // largest ratio -d/x kept as a fraction: no division per item
    SideIn<MODE != 0, false> nx;
    nx.load(c, 0, nitems);
    for (int base = 0; base < nitems; base += 32) {
        const int i = base + c.lane32;
        const Item it = item_of(i, nitems);
        const bool act = (c.actmask >> (base >> 5)) & 1u;      // implies it.valid
        const SideIn<MODE != 0, false> in = nx;
        if (base + 32 < nitems) nx.load(c, base + 32, nitems);
        double cs_s = 0.0, cp_s = 0.0;
        if (act) {
            const double2 st = in.st, lm = in.lm, zZ = in.zZ;
            const double s = st.x, t = st.y, lam = lm.x, mu = lm.y;
            const double bnd = in.bnd, z = zZ.x, Z = zZ.y;
            const double v = c.PS(PS_ROWV, it.k, it.row), dv = c.PS(PS_DV, it.k, it.row);
            const double rt = t - (it.sgn * (v - bnd) + s);
            const double rs = z + Z * s - lam - mu;
            double r1 = -lam * t, r2 = -mu * s;
            if (MODE != 0) { const double2 pr = in.pr; r1 += sigma_mu - corr * pr.x; r2 += sigma_mu - corr * pr.y; }
            const SideK sk = side_coeffs(c.reg, Z, s, t, lam, mu, rt, rs, r1, r2);
            const double ds = (sk.rho - sk.th * it.sgn * dv) * sk.iDp;
            const double dl = (r1 - lam * (it.sgn * dv + ds - rt)) * sk.itt;
            const double dtt = it.sgn * dv + ds - rt + c.reg * dl;
            const double dm = (r2 - mu * ds) * sk.is;
            // step to the boundary: largest ratio -d/x over the four variables (alpha_max = 1/ratio), compared as fractions
            {
                const double n0 = -ds * sk.is;                                   // 1/s is at hand
                if (n0 * rden > rnum) { rnum = n0; rden = 1.0; }
                if (-dl * rden > rnum * lam) { rnum = -dl; rden = lam; }
                if (-dtt * rden > rnum * t) { rnum = -dtt; rden = t; }
                if (-dm * rden > rnum * mu) { rnum = -dm; rden = mu; }
            }
            if (MODE == 0) {
                const double p1 = dl * dtt, p2 = dm * ds;
                c.P(P_PR)[i] = make_double2(p1, p2);
                c0 += lam * t + mu * s;
                c1 += lam * dtt + t * dl + mu * ds + s * dm;
                c2 += p1 + p2;
                // the corrector right-hand side is linear in (r1, r2): cc = cc_aff + sigma_mu (a1 + a2) - corr (a1 p1 + a2 p2)
                const double a1 = sk.itt * (1.0 - sk.th * sk.iDp), a2 = -sk.th * sk.iDp * sk.is;
                cs_s = it.sgn * (a1 + a2); cp_s = it.sgn * (a1 * p1 + a2 * p2);
            } else {
                c.P(P_PR)[i] = make_double2(ds, dtt); c.P(P_D)[i] = make_double2(dl, dm);
            }
        }
        if (MODE == 0) {      // the two sides of a row sit on adjacent lanes
            const double cs = cs_s + __shfl_xor_sync(FULLMASK, cs_s, 1), cp = cp_s + __shfl_xor_sync(FULLMASK, cp_s, 1);
            // (cs, cp are touched by the stage-parallel passes only: they live in the instance's side block in HBM / L2, which
            // keeps the per-stage shared-memory block at 27 doubles and 20 warps resident per SM)
            if (it.valid && it.sgn > 0.0) { double *cscp = c.S(S_CSCP); cscp[it.k * 3 + it.row] = cs; cscp[c.NS * 3 + it.k * 3 + it.row] = cp; }
        }
    }
    double amax = 1e300;
    const double ratio = -warp_min(-(rnum / rden));
    if (ratio > 0.0) amax = 1.0 / ratio;
    if (MODE == 0) { poly[0] = warp_sum(c0); poly[1] = warp_sum(c1); poly[2] = warp_sum(c2); }
    __syncwarp();
    return amax;
}

}  // namespace

// doubles per warp, rounded up to an even count so that every warp's block stays 16-byte aligned
__host__ __device__ inline int qp_smem_doubles(int N, bool longh = false) { return (SM_STAGE + (longh ? PS_N_LONG : PS_N_SHORT) * (N + 1) + 1) & ~1; }
__host__ __device__ inline int qp_smem_per_warp(int N, bool longh) { return qp_smem_doubles(N, longh) * (int)sizeof(double); }
// SNMPC: + the head scratch; the condensing prologue aliases the whole block before it is initialised
__host__ __device__ inline int qp_smem_doubles_head(int N, int H, int ns)
{
    const int a = qp_smem_doubles(N) + head_smem_doubles(H), b = condense_smem_doubles(H, ns);
    return a > b ? a : b;
}

#ifndef QP_WARPS_N
#define QP_WARPS_N 4
#endif
#ifndef QP_MINB
#define QP_MINB (16 / QP_WARPS_N)
#endif
// instances (warps) per block, and the resident warps per SM the throughput instantiation is compiled for: 16 (128-register
// cap, 121 used).  Since the tensor-core backward sweep dropped the per-lane operand descriptors the kernel also compiles to 96
// registers without spills, and with 10.9 KB of shared memory per warp at N = 30 twenty warps fit an SM (-DQP_MINB=5) -- measured
// 7 % SLOWER (1.39 vs 1.50 M solves/s at B = 4096, profiles/r02_ab_occupancy.jsonl): at 16 warps the SM is already bound by
// its load/store and issue throughput, four more warps only shrink the L1 left beside the shared memory.
constexpr int QP_WARPS = QP_WARPS_N;
constexpr int QP_MAXW = QP_MINB * QP_WARPS_N;
constexpr int IPM_SAFE_ITER = 24;      // same constants as ORC_SAFE_ITER / ORC_POLISH_ITER of the oracle
constexpr int IPM_POLISH_ITER = 4;

// applies a pending primal step dx += pend ddx, du += pend ddu to the stage records of stages [k0, NS) in HBM (stage-parallel)
__device__ void flush_primal(const Ctx &c, double &pend, int k0 = 0)
{
    if (pend == 0.0) return;
    for (int j = k0 * 8 + c.lane32; j < c.NS * 8; j += 32) {
        double *w = c.wk + (size_t)(j >> 3) * STG;
        w[W_DX + (j & 7)] += pend * w[W_DDX + (j & 7)];
    }
    for (int j = k0 * 2 + c.lane32; j < 2 * c.N; j += 32) {
        const int k = j >> 1, cc = j & 1;
        c.wk[(size_t)k * STG + W_DU + cc] += pend * c.SWS(k, W_DDU, PS_DDU)[cc];
    }
    pend = 0.0;
    __syncwarp();
}
// SNMPC: the head reads the head inputs du_0 .. du_{H-1} from HBM in every iteration, so the step of the eliminated stages
// [0, H) is applied at once (H x 10 values); the stages the backward sweep visits take theirs lazily like in the nominal kernel
__device__ void flush_primal_head(const Ctx &c, double alpha)
{
    for (int j = c.lane32; j < c.H * 8; j += 32) {
        double *w = c.wk + (size_t)(j >> 3) * STG;
        w[W_DX + (j & 7)] += alpha * w[W_DDX + (j & 7)];
    }
    for (int j = c.lane32; j < 2 * c.H; j += 32) {
        const int k = j >> 1, cc = j & 1;
        c.wk[(size_t)k * STG + W_DU + cc] += alpha * c.SWS(k, W_DDU, PS_DDU)[cc];
    }
    __syncwarp();
}

// WIDE: compiled for 8 instead of 16 resident warps per SM (up to 255 registers): 5 % lower latency per instance, chosen
// by the launcher when the batch does not fill 8 warps per SM anyway (B = 1: 0.82 -> 0.78 ms; B = 2368: 1.78 -> 2.16 ms).
// LONGH: the long-horizon layout of the per-stage shared block (kff, gus, ddu in the stage record; see PS_N_LONG)
template <bool HEAD, bool WIDE = false, bool LONGH = false>
__global__ void __launch_bounds__(QP_WARPS * 32, (HEAD || WIDE) ? 8 / QP_WARPS : QP_MINB) k_qp_solve_warp(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5;
    const int lslot = blockIdx.x * (blockDim.x >> 5) + warp;     // 4, 2 or 1 instances per block (long horizons: see launch)
    if (lslot >= cfg.B) return;                          // whole warp exits together
    const int inst = m.order ? m.order[lslot] : lslot;   // closed loop: instances that iterated longest last step start first
    if (m.halt && m.halt[inst] >= 0) return;             // stopped closed-loop instance (results stay as they were)
    Ctx c;
    c.b = inst; c.N = cfg.N; c.NS = cfg.N + 1; c.lane32 = threadIdx.x & 31; c.lane = c.lane32 & 7; c.st = c.lane32 < 8;
    c.dt = cfg.dt; c.reg = cfg.ipm_reg; c.s0 = cfg.ipm_s0;
    c.wk = m.stg + (size_t)inst * c.NS * STG;            // stage records of this instance (L_* / W_* offsets are record-relative)
    c.lin = c.wk;
    c.side = m.work + (size_t)inst * m.work_stride;
    c.sm = smem + (size_t)warp * (HEAD ? qp_smem_doubles_head(cfg.N, cfg.head_H, cfg.n_samples) : qp_smem_doubles(cfg.N, LONGH));
    c.longh = LONGH; c.psn = LONGH ? PS_N_LONG : PS_N_SHORT;
    c.vmask = (HEAD || WIDE) ? FULLMASK : OCTET0;
    c.ps = c.sm + SM_STAGE;
    c.phases = 0u;
    c.sm32 = smem_u32(c.sm);
    c.pdst = c.sm32 + SM_BUF * 8 + c.lane32 * 16; c.psrc = nullptr;
    if (HEAD) {
        // prologue: eliminate the stages of the propagation horizon onto the head inputs (scratch = this warp's block,
        // initialised below); the head block goes through HBM and is read back by this warp
        snmpc_condense_warp(cfg, m, inst, c.lane32, c.sm);
        __syncwarp();
    }
    if (c.lane32 < 8) c.sm[SM_SML + 48 + c.lane32] = 0.0;        // the six zeros of sweep_backward_full (+ its dummy slot)
    if (c.lane32 == 0) {
#pragma unroll
        for (int j = 0; j < NSLOT; j++) mbar_init(smem_u32(c.sm + SM_MBAR + j), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (HEAD) {
        const HeadLayout hl(cfg.head_H);
        c.H = hl.H; c.nh = hl.nh; c.ldm = hl.nh + 1;
        c.hG = hl.G; c.hC = hl.C; c.hROW = hl.ROW; c.hV0 = hl.V0; c.hHH = hl.HH; c.hGH = hl.GH;
        c.hd = m.head + (size_t)inst * m.head_stride;
        c.hs = c.sm + qp_smem_doubles(cfg.N);
    }
    const int N = c.N, NS = c.NS;
    const double tol = cfg.ipm_tol;

    // ---------------- phase 0: soft-row constants, equality-constrained optimum, feasibility
    const int nact = phase_prep(cfg, m, c);
    constexpr bool TENSOR = !HEAD;
    sweep_backward<HEAD, true, TENSOR>(cfg, m, c);
    sweep_rollout0<HEAD>(m, c);
    const bool feasible = phase_init(c);

    int iters = 0, qp_status = 0;
    double kkt = 0.0;
    const bool run_ipm = !((feasible && !cfg.qp_no_shortcut) || nact == 0);
    if (run_ipm) {
        // ---------------- interior-point iterations (warp-uniform control flow)
        bool converged = false;
        int it, polish = 0;
        double rs_last = 0.0, pend = 0.0;      // pend: length of the primal step not yet applied to the stage records
        double gap, cmax, rmax;
        phase_coeffs<0>(c, 0.0, 0.0, nact, &gap, &cmax, &rmax);
        for (it = 0; it <= cfg.ipm_max_iter; it++) {
            if (!(cmax == cmax) || !(rmax == rmax)) { qp_status = 3; break; }
            if (cmax <= tol && rmax <= tol) {
                flush_primal(c, pend, HEAD ? c.H : 0);
                const double rs = sweep_residual<HEAD>(cfg, m, c);
                if (!(rs == rs)) { qp_status = 3; break; }
                if (rs <= tol) { converged = true; kkt = fmax(fmax(cmax, rmax), rs); break; }
                rs_last = rs;
            }
            if (it == cfg.ipm_max_iter) { kkt = fmax(cmax, rmax); break; }
            // complementarity converged but the linear residuals sit on their round-off floor (badly scaled, diverged
            // instances): a few polishing iterations, then stop
            if (cmax <= tol && ++polish > IPM_POLISH_ITER) { kkt = fmax(fmax(cmax, rmax), rs_last); break; }
            const double rstat = sweep_backward<HEAD, false, TENSOR>(cfg, m, c, pend);      // applies the pending primal step
            pend = 0.0;
            if (!(rstat == rstat)) { qp_status = 3; break; }
            rs_last = rstat;
            sweep_forward_dir<false, HEAD>(m, c);
            double poly[3];
            const double amax_aff = phase_dirs<0>(c, 0.0, 0.0, 0.0, poly);
            const double a = fmin(1.0, amax_aff);
            const double gap_aff = (poly[0] + a * poly[1] + a * a * poly[2]) / (4.0 * nact);
            const double sg = gap_aff / gap;
            const double sigma_mu = fmax(sg * sg * sg * gap, 0.1 * tol);   // Mehrotra centring with a floor
            // second-order term scaled by how far the predictor could actually go (0: full, 1: a, 2: a^2)
            // safe mode from iteration IPM_SAFE_ITER on: damped corrector and short steps break the limit cycle of a
            // de-centred iterate (rare: degenerate tail rows)
            const bool safe = it >= IPM_SAFE_ITER;
            const double corr = safe ? a * a : cfg.ipm_corr_mode == 1 ? a : cfg.ipm_corr_mode == 2 ? a * a : 1.0;
            // corrector right-hand side per row: linear in the centring parameter and the second-order term (no side pass)
            {
                const double *cscp = c.S(S_CSCP);
                for (int j = c.lane32; j < NS * 3; j += 32) {
                    const int k = j / 3, cc = j - k * 3;
                    c.PS(PS_CF, k, cc) += sigma_mu * cscp[j] - corr * cscp[NS * 3 + j];
                }
            }
            __syncwarp();
            sweep_corrector<HEAD>(c);
            sweep_forward_dir<true, HEAD>(m, c);
            const double amax = phase_dirs<1>(c, sigma_mu, corr, 0.0, poly);
            // fraction to the boundary: capped, or a blocking variable shrinks by the factor `gap` in one step and the
            // degenerate tail rows (linearised lh = 0 of neighbouring stages) de-centre into a limit cycle
            const double eta = safe ? 0.9 : fmin(fmax(0.995, 1.0 - gap), cfg.ipm_eta_max);
            const double alpha = fmin(1.0, eta * amax);
            // update of the QP iterate: left pending for the next backward sweep, which has every stage record in shared
            // memory anyway (the SNMPC head reads the head inputs from HBM: applied at once there); the row values
            // (linear in the direction) are updated here
            pend = alpha;
            if (HEAD) flush_primal_head(c, alpha);
            for (int j = c.lane32; j < NS * 3; j += 32) {
                const int k = j / 3, cc = j - k * 3;
                c.PS(PS_ROWV, k, cc) += alpha * c.PS(PS_DV, k, cc);
            }
            __syncwarp();
            // apply the step to the side variables and, in the same pass, the residuals / coefficients of the next iteration
            phase_coeffs<2>(c, alpha, 0.0, nact, &gap, &cmax, &rmax);
        }
        flush_primal(c, pend, HEAD ? c.H : 0);
        iters = it;
        if (!converged && qp_status == 0) qp_status = 1;
    }
    __syncwarp();

    // ---------------- K3: full step (nlp_solver_step_length 1.0), slack read-out, cost at the new iterate
    double badf = (qp_status == 3) ? 1.0 : 0.0;
    for (int j = c.lane32; j < NS * 10; j += 32) {
        const int k = j / 10, cc = j - k * 10;
        if (cc >= 8 && k == N) continue;
        const double d = c.wk[(size_t)k * STG + (cc < 8 ? W_DX + cc : W_DU + cc - 8)];
        if (!(d == d)) badf = 1.0;
    }
    const bool bad = warp_sum(badf) != 0.0;
    double *Xb = m.X + (size_t)inst * NS * 8, *Ub = m.U + (size_t)inst * N * 2;
    if (!bad) {
        for (int j = c.lane32; j < NS * 8; j += 32) Xb[j] += c.wk[(size_t)(j >> 3) * STG + W_DX + (j & 7)];
        for (int j = c.lane32; j < N * 2; j += 32) Ub[j] += c.wk[(size_t)(j >> 1) * STG + W_DU + (j & 1)];
    }
    __syncwarp();
    double Jl = 0.0;
    const double *yr = m.yref + (size_t)inst * NS * 6;
    for (int j = c.lane32; j < NS * 6; j += 32) {
        const int k = j / 6, cc = j - k * 6;
        const double sig = (k < N) ? c.dt : 1.0;
        // tracking term of output component cc at stage k
        if (cc < 4 || k < N) {
            double y;
            if (cc < 4) {
                y = Xb[k * 8 + cc];
                if (cc == 2) y = wrap_yaw(y);
                if (cc == 3 && cfg.cost_vel_abs) { const double x4 = Xb[k * 8 + 4]; y = sqrt(y * y + x4 * x4); }
            } else {
                y = Ub[k * 2 + cc - 4];
            }
            const double e = y - yr[k * 6 + cc];
            Jl += sig * 0.5 * sval(m.W_s, m.W_i, c, k, 6, cc) * e * e;
        }
        // slack term of side j (same flat index: k*6 + row*2 + side); penalties are already sigma-scaled
        const int row = cc >> 1;
        double sv = 0.0;
        if (run_ipm && c.S(S_ACT)[j] != 0.0) {
            sv = c.P(P_ST)[j].x;
            const double2 zZ = c.P(P_Z)[j];
            Jl += zZ.x * sv + 0.5 * zZ.y * sv * sv;
        }
        ((cc & 1) ? m.su : m.sl)[((size_t)inst * NS + k) * 3 + row] = sv;
    }
    if (HEAD) {
        // epilogue: full step of the sample iterates with the head inputs of this QP (K3 for the samples)
        __syncwarp();
        if (!bad) snmpc_step_warp(cfg, m, inst, c.lane32);
    }
    const double J = warp_sum(Jl);
    if (c.lane32 == 0) {
        m.cost[inst] = J;
        m.kkt[inst] = kkt;
        m.qp_iter[inst] = iters;
        m.status[inst] = bad ? 4 : 0;     // ACADOS_QP_FAILURE; RTI keeps status 0 when the QP hits max-iter (SURVEY B-9)
    }
}

// Instances (warps) per block: 4 while 16 warps fit the shared memory of an SM, fewer for long horizons so that the
// per-SM shared memory is used in finer units (N = 80: 8 -> 10 resident warps per SM).
static int pick_warps_per_block(int smem_per_warp, int max_warps_per_sm, int *resident = nullptr)
{
    const int smem_sm = 227 * 1024;
    int best_w = QP_WARPS, best_res = 0;
    for (int w = QP_WARPS; w >= 1; w >>= 1) {
        const int blocks = smem_sm / (w * smem_per_warp + 1024);
        int res = blocks * w;
        if (res > max_warps_per_sm) res = max_warps_per_sm;
        if (res > best_res) { best_res = res; best_w = w; }
    }
    if (resident) *resident = best_res;
    return best_w;
}

int launch_qp_solve_warp(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    if (cfg.head_H > 0) {
        const int per_warp = qp_smem_doubles_head(cfg.N, cfg.head_H, cfg.n_samples) * (int)sizeof(double);
        const int w = pick_warps_per_block(per_warp, 8);
        const int smem = w * per_warp;
        if (smem > 48 * 1024) {     // per device and cheap: no caching across handles / devices
            cudaError_t e = cudaFuncSetAttribute(k_qp_solve_warp<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
        }
        k_qp_solve_warp<true, false><<<(cfg.B + w - 1) / w, w * 32, smem, st>>>(cfg, m);
        return (int)cudaGetLastError();
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // (B200NMPC_FORCE_THROUGHPUT in the environment: the throughput instantiation for any batch size -- compute-sanitizer runs
    // of that instantiation on a small case, tools/gpu_sanitize.sh)
    static const bool force_throughput = getenv("B200NMPC_FORCE_THROUGHPUT") != nullptr;
    const bool wide = cfg.B <= 8 * sms && !force_throughput;   // small batch: the latency variant (8 resident warps per SM are enough)
    const int maxw = wide ? 8 : QP_MAXW;
    // long-horizon layout when it lets more warps stay resident (throughput kernel: N > 38)
    int w = QP_WARPS, res_short = 0, res_long = 0;
    const int w_short = pick_warps_per_block(qp_smem_per_warp(cfg.N, false), maxw, &res_short);
    const int w_long = pick_warps_per_block(qp_smem_per_warp(cfg.N, true), maxw, &res_long);
    static const char *env_long = getenv("B200NMPC_QP_LONGH");      // development knob: force either layout
    const bool longh = env_long ? atoi(env_long) != 0 : (!wide && res_long > res_short);
    w = longh ? w_long : w_short;
    const int per_warp = qp_smem_per_warp(cfg.N, longh);
    int smem = w * per_warp;
    // development knobs (tools/gpu_ab_env.sh): instances per block and a cap on the resident warps per SM (enforced by
    // padding the block's dynamic shared memory)
    static const int env_wpb = getenv("B200NMPC_QP_WPB") ? atoi(getenv("B200NMPC_QP_WPB")) : 0;
    static const int env_maxw = getenv("B200NMPC_QP_MAXW") ? atoi(getenv("B200NMPC_QP_MAXW")) : 0;
    if (env_wpb == 1 || env_wpb == 2 || env_wpb == 4) { w = env_wpb; smem = w * per_warp; }
    if (env_maxw > 0 && env_maxw < QP_MAXW) {
        const int blocks = (env_maxw + w - 1) / w;
        const int pad = 227 * 1024 / blocks - 1024;
        if (pad > smem) smem = pad & ~15;
    }
    auto launch = [&](auto kernel) -> int {
        if (smem > 48 * 1024) {     // per device and cheap: no caching across handles / devices
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
        }
        kernel<<<(cfg.B + w - 1) / w, w * 32, smem, st>>>(cfg, m);
        return (int)cudaGetLastError();
    };
    if (wide) return longh ? launch(k_qp_solve_warp<false, true, true>) : launch(k_qp_solve_warp<false, true, false>);
    return longh ? launch(k_qp_solve_warp<false, false, true>) : launch(k_qp_solve_warp<false, false, false>);
}

}  // namespace b200
