#include "step_dispatch.cuh"
namespace mrb {
cudaError_t launch_step_warehouse(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched)
{
    return launch_step_generic<MRB_WAREHOUSE>(p, actions, s, launched);
}
}  // namespace mrb
