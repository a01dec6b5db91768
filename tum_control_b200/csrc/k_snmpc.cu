// k_snmpc.cu -- set-up kernels of the Stochastic NMPC variant (x0 sample matrix, initial guess of the sample iterates).
// The per-sample linearisation lives in k_linearize.cu (extra work items of the integrator launch), head condensing and the
// sample-iterate step in snmpc_head.cuh (prologue / epilogue of the QP kernel).
#include "nmpc_kernels.cuh"

namespace b200 {

// compute_x0dist (stochastic_mpc_utils.py:78-91): x0_s = x0 + std (.) w_s
__global__ void k_snmpc_x0_samples(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int ns = cfg.n_samples;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)cfg.B * ns * 8) return;
    const int b = (int)(g / (ns * 8)), r = (int)(g - (long long)b * ns * 8);
    m.x0s[g] = m.x0[(size_t)b * 8 + (r & 7)] + m.sn_woff[r];
}
void launch_snmpc_x0_samples(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    const long long items = (long long)cfg.B * cfg.n_samples * 8;
    if (items == 0) return;
    k_snmpc_x0_samples<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(cfg, m);
}

// set(j, "x", x0_samples.flatten()) for every stage (SNMPC_class.py:127-130,266-272): sample iterates = x0 samples
__global__ void k_snmpc_init_samples(const __grid_constant__ DevCfg cfg, const DevMem m)
{
    const int ns = cfg.n_samples, HS = m.HA + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)cfg.B * HS * ns * 8) return;
    const int b = (int)(g / ((long long)HS * ns * 8));
    const int r = (int)(g - (long long)b * HS * ns * 8) % (ns * 8);
    m.XS[g] = m.x0s[(size_t)b * ns * 8 + r];
}
void launch_snmpc_init_samples(const DevCfg &cfg, const DevMem &m, cudaStream_t st)
{
    const long long items = (long long)cfg.B * (m.HA + 1) * cfg.n_samples * 8;
    if (items == 0) return;
    k_snmpc_init_samples<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(cfg, m);
}

}  // namespace b200
