// One translation unit per (scenario, compile-time team size): the warp kernel with the FP64 tensor-core solver
// (QpWarp<PPL, NC>::factor_tiles / solve_tiles).  Included by kern_team_*.cu, which define MRB_TEAM_SCN, MRB_TEAM_TAG and
// MRB_TEAM_N; the PredatorCapturePrey unit of a team size also carries the QP-alone kernel (mrb_barrier_qp).
#include "step_thread.cuh"     // ObsRow, reset_env
#include "step_warp.cuh"

#define MRB_CAT2(a, b, c, d) a##b##c##d
#define MRB_CAT(a, b, c, d) MRB_CAT2(a, b, c, d)

namespace mrb {
cudaError_t MRB_CAT(launch_step_team_, MRB_TEAM_TAG, _, MRB_TEAM_N)(const Params &p, const int32_t *actions, cudaStream_t s)
{
    return launch_step_warp_ppl<MRB_TEAM_SCN, team_ppl(MRB_TEAM_N), MRB_TEAM_N>(p, actions, s);
}
#ifdef MRB_TEAM_WITH_QP
cudaError_t MRB_CAT(launch_qp_team_, n, _, MRB_TEAM_N)(int barrier_default, int64_t B, const double *dxi, const double *xi, double *u,
                                                      int32_t *iters, cudaStream_t s)
{
    return launch_qp_warp_ppl<team_ppl(MRB_TEAM_N), MRB_TEAM_N>(MRB_TEAM_N, barrier_default, B, dxi, xi, u, iters, s);
}
#endif
}  // namespace mrb
