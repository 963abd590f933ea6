// mr_kernels_fast.cu — MR_MATH_FAST instantiations of the trace kernel
// (restructured f64 stage, FMA contraction allowed; the f32 stages use
// explicit _rn intrinsics and stay value-identical to the reference).
#include "mr_trace_kernel.cuh"
#include "mr_launch.hpp"

namespace mr {
cudaError_t launch_trace_fast(const TraceArgs &a, cudaStream_t stream)
{
    return launch_trace_math<MR_MATH_FAST>(a, stream);
}
}  // namespace mr
