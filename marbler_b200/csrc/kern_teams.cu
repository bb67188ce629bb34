// Dispatch to the team sizes that have a kernel of their own (kern_team_*.cu: compile-time team size, FP64 tensor-core
// solver).  Multiples of four robots that the scenarios' spawn grids can hold; every other team size of 7..32 robots runs
// the run-time team size kernels of step_warp.cuh.  Measured (PredatorCapturePrey, 32,768 envs, ms per step, run-time vs
// compile-time team size): 8 robots 5.81 vs 3.34, 12: 10.6 vs 5.81, 16: 15.1 vs 9.09, 20: 29.0 vs 13.7.
#include "launchers.h"

namespace mrb {
cudaError_t launch_step_team_pcp_8(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_pcp_12(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_pcp_16(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_pcp_20(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_pcp_24(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_pcp_28(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_warehouse_8(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_simple_8(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_simple_12(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_step_team_simple_16(const Params &p, const int32_t *actions, cudaStream_t s);
cudaError_t launch_qp_team_n_8(int barrier_default, int64_t B, const double *dxi, const double *xi, double *u, int32_t *iters, cudaStream_t s);
cudaError_t launch_qp_team_n_12(int barrier_default, int64_t B, const double *dxi, const double *xi, double *u, int32_t *iters, cudaStream_t s);
cudaError_t launch_qp_team_n_16(int barrier_default, int64_t B, const double *dxi, const double *xi, double *u, int32_t *iters, cudaStream_t s);
cudaError_t launch_qp_team_n_20(int barrier_default, int64_t B, const double *dxi, const double *xi, double *u, int32_t *iters, cudaStream_t s);
cudaError_t launch_qp_team_n_24(int barrier_default, int64_t B, const double *dxi, const double *xi, double *u, int32_t *iters, cudaStream_t s);
cudaError_t launch_qp_team_n_28(int barrier_default, int64_t B, const double *dxi, const double *xi, double *u, int32_t *iters, cudaStream_t s);

cudaError_t launch_step_team(int scenario, const Params &p, const int32_t *actions, cudaStream_t s, bool *handled)
{
    *handled = true;
    if (scenario == MRB_PCP) switch (p.cfg.num_robots) {
    case 8: return launch_step_team_pcp_8(p, actions, s);
    case 12: return launch_step_team_pcp_12(p, actions, s);
    case 16: return launch_step_team_pcp_16(p, actions, s);
    case 20: return launch_step_team_pcp_20(p, actions, s);
    case 24: return launch_step_team_pcp_24(p, actions, s);
    case 28: return launch_step_team_pcp_28(p, actions, s);
    default: break;
    }
    if (scenario == MRB_WAREHOUSE) switch (p.cfg.num_robots) {
    case 8: return launch_step_team_warehouse_8(p, actions, s);
    default: break;
    }
    if (scenario == MRB_SIMPLE) switch (p.cfg.num_robots) {
    case 8: return launch_step_team_simple_8(p, actions, s);
    case 12: return launch_step_team_simple_12(p, actions, s);
    case 16: return launch_step_team_simple_16(p, actions, s);
    default: break;
    }
    *handled = false;
    return cudaSuccess;
}

cudaError_t launch_qp_team(int N, int barrier_default, int64_t B, const double *dxi, const double *xi, double *u, int32_t *iters,
                           cudaStream_t s, bool *handled)
{
    *handled = true;
    switch (N) {
    case 8: return launch_qp_team_n_8(barrier_default, B, dxi, xi, u, iters, s);
    case 12: return launch_qp_team_n_12(barrier_default, B, dxi, xi, u, iters, s);
    case 16: return launch_qp_team_n_16(barrier_default, B, dxi, xi, u, iters, s);
    case 20: return launch_qp_team_n_20(barrier_default, B, dxi, xi, u, iters, s);
    case 24: return launch_qp_team_n_24(barrier_default, B, dxi, xi, u, iters, s);
    case 28: return launch_qp_team_n_28(barrier_default, B, dxi, xi, u, iters, s);
    default: break;
    }
    *handled = false;
    return cudaSuccess;
}
}  // namespace mrb
