// Micro-benchmark of the FP64 tensor-core instruction (mma.sync.m8n8k4.f64 = DMMA.8x8x4 on sm_100a): dependent-issue
// latency, issue rate against resident warps x independent accumulators, and whether it shares the FP64 pipe with DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma scripts/microbench/dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ILP independent accumulator tiles per warp, MIX DFMAs (independent chains of their own) issued per DMMA
template <int ILP, int MIX>
__global__ void dmma_chain(double *out, long long *cycles, int iters, double a, double b)
{
    double c0[ILP], c1[ILP], x[MIX > 0 ? MIX : 1];
#pragma unroll
    for (int k = 0; k < ILP; k++) { c0[k] = threadIdx.x + k; c1[k] = 0.5 * k; }
#pragma unroll
    for (int k = 0; k < (MIX > 0 ? MIX : 1); k++) x[k] = threadIdx.x + k;
    const long long t0 = clk();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) dmma(c0[k], c1[k], a, b);
#pragma unroll
        for (int k = 0; k < MIX; k++) x[k] = fma(x[k], a, b);
    }
    const long long t1 = clk();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += c0[k] + c1[k];
#pragma unroll
    for (int k = 0; k < (MIX > 0 ? MIX : 1); k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP, int MIX>
static void rate(int warps, double *out, long long *cyc, int sms)
{
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dmma_chain<ILP, MIX><<<sms, warps * 32>>>(out, cyc, 64, 1e-9, 1e-9);
    cudaEventRecord(e0);
    dmma_chain<ILP, MIX><<<sms, warps * 32>>>(out, cyc, iters, 1e-9, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double per_clk = (double)warps * ILP * iters / (double)c;
    printf("warps/SM %2d  tiles %d  dfma/dmma %d : %.3f warp-DMMA/clk/SM = %.1f cycles per DMMA per scheduler (%.1f TFLOP/s tensor + %.1f TFLOP/s DFMA, %.3f ms)\n",
           warps, ILP, MIX, per_clk, 4.0 / per_clk, (double)sms * warps * ILP * iters * 512.0 / (ms * 1e-3) * 1e-12,
           (double)sms * warps * MIX * iters * 64.0 / (ms * 1e-3) * 1e-12, ms);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long));
    {
        dmma_chain<1, 0><<<1, 32>>>(out, cyc, 4096, 1e-9, 1e-9);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        printf("DMMA.8x8x4 (1 warp, accumulator chain)  %.2f cycles per dependent step\n", (double)c / 4096);
    }
    for (int w : {1, 2, 4, 8, 12, 16}) {
        rate<1, 0>(w, out, cyc, sms);
        rate<2, 0>(w, out, cyc, sms);
        rate<4, 0>(w, out, cyc, sms);
        rate<8, 0>(w, out, cyc, sms);
    }
    // FP64 pipe shared with DFMA?  4 independent DMMA tiles + 8 / 16 / 32 independent DFMAs per loop trip
    for (int w : {4, 8, 12}) {
        rate<4, 8>(w, out, cyc, sms);
        rate<4, 16>(w, out, cyc, sms);
        rate<4, 32>(w, out, cyc, sms);
    }
    return 0;
}
