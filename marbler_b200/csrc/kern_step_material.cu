#include "step_dispatch.cuh"
namespace mrb {
cudaError_t launch_step_material(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched)
{
    return launch_step_generic<MRB_MATERIAL>(p, actions, s, launched);
}
}  // namespace mrb
