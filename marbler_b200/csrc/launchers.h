// Launch entry points, one translation unit per scenario so that the library builds in parallel.
#pragma once
#include "common.cuh"

namespace mrb {
// *launched = false when no kernel exists for (scenario, num_robots)
cudaError_t launch_step_pcp(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched);
cudaError_t launch_step_warehouse(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched);
cudaError_t launch_step_material(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched);
cudaError_t launch_step_arctic(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched);
cudaError_t launch_step_simple(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched);
// team sizes that have a kernel of their own (compile-time team size, tensor-core solver; kern_team_*.cu + kern_teams.cu):
// *handled = false when (scenario, num_robots) has none and the caller falls back to the run-time team size kernels
cudaError_t launch_step_team(int scenario, const Params &p, const int32_t *actions, cudaStream_t s, bool *handled);
cudaError_t launch_qp_team(int N, int barrier_default, int64_t B, const double *dxi, const double *xi, double *u, int32_t *iters,
                           cudaStream_t s, bool *handled);
cudaError_t launch_reset(const Params &p, const uint8_t *mask, cudaStream_t s);
cudaError_t launch_barrier_qp(int N, int barrier_default, int64_t B, const double *dxi, const double *xi, double *u,
                              int32_t *iters, cudaStream_t s);
}  // namespace mrb
