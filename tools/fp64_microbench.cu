// FP64 anchors for the compute side of the roofline on this GPU: dependent-issue latency of DFMA and the sustained
// DFMA rate with enough independent chains and warps.   nvcc -arch=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
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

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 64 * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
    cudaMemset(out, 0, 64 * sizeof(double));
    const int n = 100000;
    k_latency<<<1, 1>>>(out, cyc, n); k_latency<<<1, 1>>>(out, cyc, n);
    long long h; cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("{\"dfma_dependent_latency_cycles\": %.2f", (double)h / n);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    k_throughput<8><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e0); k_throughput<8><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * (double)iters * blocks * threads;
    printf(", \"dfma_tflops\": %.2f, \"sms\": %d, \"sm_clock_mhz\": %d}\n", flops / (ms * 1e-3) / 1e12, p.multiProcessorCount, p.clockRate / 1000);
    return 0;
}
