// reset kernels and the stand-alone barrier-QP kernels
#include "launchers.h"
#include "step_thread.cuh"
#include "step_warp.cuh"

namespace mrb {

template <int SCN>
static cudaError_t launch_reset_scn(const Params &p, const uint8_t *mask, cudaStream_t s)
{
    const int tpb = 128;
    reset_kernel<SCN><<<(unsigned)((p.env_hi - p.env_lo + tpb - 1) / tpb), tpb, 0, s>>>(p, mask);
    return cudaGetLastError();
}

cudaError_t launch_reset(const Params &p, const uint8_t *mask, cudaStream_t s)
{
    switch (p.cfg.scenario) {
    case MRB_PCP: return launch_reset_scn<MRB_PCP>(p, mask, s);
    case MRB_WAREHOUSE: return launch_reset_scn<MRB_WAREHOUSE>(p, mask, s);
    case MRB_MATERIAL: return launch_reset_scn<MRB_MATERIAL>(p, mask, s);
    case MRB_ARCTIC: return launch_reset_scn<MRB_ARCTIC>(p, mask, s);
    default: return launch_reset_scn<MRB_SIMPLE>(p, mask, s);
    }
}

template <int N>
__global__ void __launch_bounds__(kThreadsPerBlock)
qp_thread_kernel(int64_t B, int barrier_default, const double *__restrict__ dxi, const double *__restrict__ xi,
                 double *__restrict__ u, int32_t *__restrict__ iters)
{
    const int64_t e = (int64_t)blockIdx.x * kThreadsPerBlock + threadIdx.x;
    if (e >= B) return;
    double xix[N], xiy[N], ux[N], uy[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        xix[i] = xi[i * B + e]; xiy[i] = xi[(N + i) * B + e];
        ux[i] = dxi[i * B + e]; uy[i] = dxi[(N + i) * B + e];
    }
    QpForTeam<N> qp;
    const int it = qp.run(xix, xiy, ux, uy, barrier_default != 0);
#pragma unroll
    for (int i = 0; i < N; i++) { u[i * B + e] = ux[i]; u[(N + i) * B + e] = uy[i]; }
    if (iters) iters[e] = it;
}

cudaError_t launch_barrier_qp(int N, int barrier_default, int64_t B, const double *dxi, const double *xi, double *u,
                              int32_t *iters, cudaStream_t s)
{
    const unsigned grid = (unsigned)((B + kThreadsPerBlock - 1) / kThreadsPerBlock);
    switch (N) {
    case 2: qp_thread_kernel<2><<<grid, kThreadsPerBlock, 0, s>>>(B, barrier_default, dxi, xi, u, iters); break;
    case 3: qp_thread_kernel<3><<<grid, kThreadsPerBlock, 0, s>>>(B, barrier_default, dxi, xi, u, iters); break;
    case 4: qp_thread_kernel<4><<<grid, kThreadsPerBlock, 0, s>>>(B, barrier_default, dxi, xi, u, iters); break;
    case 5: qp_thread_kernel<5><<<grid, kThreadsPerBlock, 0, s>>>(B, barrier_default, dxi, xi, u, iters); break;
    case 6: qp_thread_kernel<6><<<grid, kThreadsPerBlock, 0, s>>>(B, barrier_default, dxi, xi, u, iters); break;
    default: return launch_qp_warp(N, barrier_default, B, dxi, xi, u, iters, s);
    }
    return cudaGetLastError();
}

}  // namespace mrb
