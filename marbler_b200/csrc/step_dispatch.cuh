// Shared body of the per-scenario launch translation units (kern_step_*.cu).
#pragma once
#include "launchers.h"
#include "step_thread.cuh"
#include "step_warp.cuh"

namespace mrb {

template <int SCN, int N>
inline cudaError_t launch_thread(const Params &p, const int32_t *actions, cudaStream_t s)
{
    const unsigned grid = (unsigned)((p.env_hi - p.env_lo + kThreadsPerBlock - 1) / kThreadsPerBlock);
    step_thread_kernel<SCN, N><<<grid, kThreadsPerBlock, 0, s>>>(p, actions);
    return cudaGetLastError();
}

// teams of 2..6 robots: one env per thread; anything else up to 32 robots: one env per warp
template <int SCN>
inline cudaError_t launch_step_generic(const Params &p, const int32_t *actions, cudaStream_t s, bool *launched)
{
    *launched = true;
    switch (p.cfg.num_robots) {
    case 2: if (SCN != MRB_MATERIAL) return launch_thread<SCN, (SCN != MRB_MATERIAL ? 2 : 4)>(p, actions, s); break;
    case 3: if (SCN != MRB_MATERIAL) return launch_thread<SCN, (SCN != MRB_MATERIAL ? 3 : 4)>(p, actions, s); break;
    case 4: return launch_thread<SCN, 4>(p, actions, s);
    case 5: return launch_thread<SCN, 5>(p, actions, s);
    case 6: return launch_thread<SCN, 6>(p, actions, s);
    default: break;
    }
    return launch_step_warp<SCN>(p, actions, s);
}

}  // namespace mrb
