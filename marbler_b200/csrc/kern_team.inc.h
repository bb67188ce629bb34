// One translation unit per (scenario, compile-time size of the Newton system): the warp kernel with the FP64 tensor-core
// solver (QpWarp<PPL, NC, EXACT>::factor_tiles / solve_tiles) in two variants -- exactly NC robots (team size folded into
// the code) and N <= NC robots at run time (phantom robots pad the last tile).  Included by kern_team_*.cu, which define
// MRB_TEAM_SCN, MRB_TEAM_TAG and MRB_TEAM_N; the PredatorCapturePrey unit of a size also carries the QP-alone kernels
// (mrb_barrier_qp).
#include "step_thread.cuh"     // ObsRow, reset_env
#include "step_warp.cuh"

#define MRB_CAT2(a, b, c, d) a##b##c##d
#define MRB_CAT(a, b, c, d) MRB_CAT2(a, b, c, d)

namespace mrb {
cudaError_t MRB_CAT(launch_step_team_, MRB_TEAM_TAG, _, MRB_TEAM_N)(const Params &p, const int32_t *actions, cudaStream_t s, bool exact)
{
    if (exact) return launch_step_warp_ppl<MRB_TEAM_SCN, team_ppl(MRB_TEAM_N), MRB_TEAM_N, true>(p, actions, s);
    return launch_step_warp_ppl<MRB_TEAM_SCN, team_ppl(MRB_TEAM_N), MRB_TEAM_N, false>(p, actions, s);
}
#ifdef MRB_TEAM_WITH_QP
cudaError_t MRB_CAT(launch_qp_team_, n, _, MRB_TEAM_N)(int N, int barrier_default, int64_t B, const double *dxi, const double *xi, double *u,
                                                      int32_t *iters, cudaStream_t s)
{
    if (N == MRB_TEAM_N) return launch_qp_warp_ppl<team_ppl(MRB_TEAM_N), MRB_TEAM_N, true>(N, barrier_default, B, dxi, xi, u, iters, s);
    return launch_qp_warp_ppl<team_ppl(MRB_TEAM_N), MRB_TEAM_N, false>(N, barrier_default, B, dxi, xi, u, iters, s);
}
#endif
}  // namespace mrb
