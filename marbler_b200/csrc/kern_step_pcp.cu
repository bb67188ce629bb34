#include "step_dispatch.cuh"
namespace mrb {
cudaError_t launch_step_pcp(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched)
{
    return launch_step_generic<MRB_PCP>(p, actions, s, launched);
}
}  // namespace mrb
