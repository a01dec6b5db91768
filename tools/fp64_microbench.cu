// FP64 anchors for the compute side of the roofline on this GPU, and the numbers behind the "tensor cores for the 8x8
// Riccati products?" decision:
//   * DFMA: dependent-issue latency and sustained rate
//   * DMMA (mma.sync.aligned.m8n8k4.row.col.f64): dependent-issue latency (accumulator chain), issue rate per SM
//     sub-partition with 1..8 independent accumulators, and sustained FLOP rate
//   * the two ways to get an accumulator fragment back into an operand fragment (what every stage of a Riccati
//     recursion needs: P_k comes out of one product and goes into the next): 64-bit warp shuffles vs a shared-memory
//     round trip (st.shared + bar.warp.sync + ld.shared)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_latency(double *out, long long *cyc, int n)
{
    double a = out[0], b = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < n; i++) a = fma(a, b, c);           // one dependent chain
    const long long t1 = clock64();
    out[0] = a;
    cyc[0] = t1 - t0;
}

template <int CH>
__global__ void k_throughput(double *out, int n)
{
    double a[CH];
    for (int j = 0; j < CH; j++) a[j] = out[(blockIdx.x * blockDim.x + threadIdx.x) % 64] + j;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int j = 0; j < CH; j++) a[j] = fma(a[j], b, c);
    double s = 0;
    for (int j = 0; j < CH; j++) s += a[j];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// CH independent accumulator pairs per warp; one warp per block measures latency (CH = 1) and the per-warp issue rate
template <int CH>
__global__ void k_dmma(double *out, long long *cyc, int n)
{
    double c0[CH], c1[CH];
    const double a = 1e-3 * (threadIdx.x % 7), b = 1e-3 * (threadIdx.x % 5);
    for (int j = 0; j < CH; j++) { c0[j] = out[threadIdx.x % 64] + j; c1[j] = c0[j] + 1.0; }
    const long long t0 = clock64();
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int j = 0; j < CH; j++) dmma884(c0[j], c1[j], a, b);
    const long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < CH; j++) s += c0[j] + c1[j];
    if (s == 12345.678) out[0] = s;
    if (cyc && threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// accumulator fragment (row g = lane / 4, columns 2 t, 2 t + 1, t = lane % 4) -> A-operand fragments of the two k-steps
// (row g, columns t and t + 4) through quad shuffles: the conversion a DMMA Riccati recursion needs once per product
__global__ void k_frag_shfl(double *out, long long *cyc, int n)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double c0 = out[lane], c1 = c0 + 1.0;
    const int s0 = g * 4 + (t >> 1), s1 = s0 + 2;
    const long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        const double x0 = __shfl_sync(0xffffffffu, c0, s0), x1 = __shfl_sync(0xffffffffu, c1, s0);
        const double y0 = __shfl_sync(0xffffffffu, c0, s1), y1 = __shfl_sync(0xffffffffu, c1, s1);
        const double a0 = (t & 1) ? x1 : x0, a1 = (t & 1) ? y1 : y0;
        c0 = a0 + 1e-9; c1 = a1 - 1e-9;                      // dependent: next round converts the new values
    }
    const long long t1 = clock64();
    out[lane] = c0 + c1;
    if (lane == 0) cyc[0] = t1 - t0;
}
// the same conversion through shared memory
__global__ void k_frag_smem(double *out, long long *cyc, int n)
{
    __shared__ double tile[8 * 10];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double c0 = out[lane], c1 = c0 + 1.0;
    const long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        *reinterpret_cast<double2 *>(tile + g * 10 + 2 * t) = make_double2(c0, c1);
        __syncwarp();
        const double a0 = tile[g * 10 + t], a1 = tile[g * 10 + t + 4];
        __syncwarp();
        c0 = a0 + 1e-9; c1 = a1 - 1e-9;
    }
    const long long t1 = clock64();
    out[lane] = c0 + c1;
    if (lane == 0) cyc[0] = t1 - t0;
}

template <typename F>
static double cycles_per_iter(F launch, long long *cyc, int n)
{
    launch(); launch();
    long long h;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    return (double)h / n;
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 64 * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
    cudaMemset(out, 0, 64 * sizeof(double));
    const int n = 100000;
    printf("{\"dfma_dependent_latency_cycles\": %.2f", cycles_per_iter([&] { k_latency<<<1, 1>>>(out, cyc, n); }, cyc, n));
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    float ms;
    k_throughput<8><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e0); k_throughput<8><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"dfma_tflops\": %.2f", 2.0 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12);
    // DMMA m8n8k4: 256 FMA per warp instruction
    printf(", \"dmma884_dependent_latency_cycles\": %.2f", cycles_per_iter([&] { k_dmma<1><<<1, 32>>>(out, cyc, n); }, cyc, n));
    printf(", \"dmma884_cycles_per_instr_one_warp_2acc\": %.2f", cycles_per_iter([&] { k_dmma<2><<<1, 32>>>(out, cyc, n); }, cyc, n) / 2);
    printf(", \"dmma884_cycles_per_instr_one_warp_4acc\": %.2f", cycles_per_iter([&] { k_dmma<4><<<1, 32>>>(out, cyc, n); }, cyc, n) / 4);
    printf(", \"dmma884_cycles_per_instr_one_warp_8acc\": %.2f", cycles_per_iter([&] { k_dmma<8><<<1, 32>>>(out, cyc, n); }, cyc, n) / 8);
    // four warps of one block land on the four sub-partitions of one SM: per-SM issue rate
    printf(", \"dmma884_cycles_per_instr_per_warp_4warps_8acc\": %.2f", cycles_per_iter([&] { k_dmma<8><<<1, 128>>>(out, cyc, n); }, cyc, n) / 8);
    printf(", \"dmma884_cycles_per_instr_per_warp_16warps_4acc\": %.2f", cycles_per_iter([&] { k_dmma<4><<<1, 512>>>(out, cyc, n); }, cyc, n) / 4);
    k_dmma<8><<<blocks, threads>>>(out, nullptr, iters);
    cudaEventRecord(e0); k_dmma<8><<<blocks, threads>>>(out, nullptr, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf(", \"dmma884_tflops\": %.2f", 2.0 * 256 * 8 * (double)iters * blocks * (threads / 32) / (ms * 1e-3) / 1e12);
    printf(", \"acc_to_operand_fragment_shfl_cycles\": %.2f", cycles_per_iter([&] { k_frag_shfl<<<1, 32>>>(out, cyc, n); }, cyc, n));
    printf(", \"acc_to_operand_fragment_smem_cycles\": %.2f", cycles_per_iter([&] { k_frag_smem<<<1, 32>>>(out, cyc, n); }, cyc, n));
    printf(", \"sms\": %d, \"sm_clock_mhz\": %d}\n", p.multiProcessorCount, p.clockRate / 1000);
    return cudaGetLastError() != cudaSuccess;
}
