// Micro-benchmarks behind the latency model in DESIGN.md: dependent-issue latency of the FP64 instructions the
// solvers are made of, FP64 issue rate against resident warps x independent chains, shared-memory and shuffle
// round trips.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_latency scripts/microbench/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

template <int ILP>
__global__ void dfma_chain(double *out, long long *cycles, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = threadIdx.x + k;
    const long long t0 = clk();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
    }
    const long long t1 = clk();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void rcp_chain(double *out, long long *cycles, int iters)
{
    double x = 1.5 + threadIdx.x;
    const long long t0 = clk();
    for (int i = 0; i < iters; i++) {
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        x = r + 1.25;
    }
    const long long t1 = clk();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void rsqrt_newton_chain(double *out, long long *cycles, int iters)
{
    double x = 1.5 + threadIdx.x;
    const long long t0 = clk();
    for (int i = 0; i < iters; i++) {
        double r;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        double hx = 0.5 * x;
        r = r * fma(-hx * r, r, 1.5);
        r = r * fma(-hx * r, r, 1.5);
        x = r + 1.25;
    }
    const long long t1 = clk();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void shfl_chain(double *out, long long *cycles, int iters)
{
    double x = 1.5 + threadIdx.x;
    const long long t0 = clk();
    for (int i = 0; i < iters; i++) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    const long long t1 = clk();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void lds_chain(double *out, long long *cycles, int iters)
{
    __shared__ double buf[1024];
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) buf[k] = (double)((k * 33 + 7) & 1023);
    __syncthreads();
    int idx = threadIdx.x;
    const long long t0 = clk();
    for (int i = 0; i < iters; i++) idx = (int)buf[idx];
    const long long t1 = clk();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void dsetp_chain(double *out, long long *cycles, int iters, double a)
{
    double x = 1.5 + threadIdx.x;
    const long long t0 = clk();
    for (int i = 0; i < iters; i++) x = fmax(x, a) + 1.0;
    const long long t1 = clk();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
void run_tp(int warps_per_sm, int sms, double *out, long long *cyc)
{
    // one CTA per SM with `warps_per_sm` warps: FP64 issue rate as a function of resident warps x chains
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    dfma_chain<ILP><<<sms, warps_per_sm * 32>>>(out, cyc, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dfma_chain<ILP><<<sms, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_sm_per_clk = (double)iters * ILP * warps_per_sm / (double)c;     // warp-DFMAs per clock per SM
    printf("warps/SM %2d  chains %2d : %6.3f warp-DFMA/clk/SM  (%.1f TFLOP/s at this clock count, %.3f ms)\n", warps_per_sm, ILP,
           per_sm_per_clk, 2.0 * 32 * iters * ILP * warps_per_sm * sms / (ms * 1e-3) / 1e12, ms);
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 100000;
    long long c;
#define LAT(name, call) call; cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s %7.2f cycles per dependent step\n", name, (double)c / iters);
    LAT("DFMA (1 warp, 1 chain)", (dfma_chain<1><<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9)));
    LAT("MUFU.RCP64H + DADD", (rcp_chain<<<1, 32>>>(out, cyc, iters)));
    LAT("rsqrt seed + 2 Newton + DADD", (rsqrt_newton_chain<<<1, 32>>>(out, cyc, iters)));
    LAT("SHFL.64 + DADD", (shfl_chain<<<1, 32>>>(out, cyc, iters)));
    LAT("LDS.64 + F2I", (lds_chain<<<1, 32>>>(out, cyc, iters)));
    LAT("DMNMX + DADD", (dsetp_chain<<<1, 32>>>(out, cyc, iters, 0.5)));
    for (int w : {1, 2, 4, 8, 12, 16, 32}) {
        run_tp<1>(w, p.multiProcessorCount, out, cyc);
        run_tp<2>(w, p.multiProcessorCount, out, cyc);
        run_tp<4>(w, p.multiProcessorCount, out, cyc);
        run_tp<8>(w, p.multiProcessorCount, out, cyc);
    }
    return 0;
}
