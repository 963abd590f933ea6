// mr_kernels_strict.cu — MR_MATH_STRICT instantiations of the trace kernel:
// the reference's expression tree operation by operation.  This file is
// compiled with -fmad=false so no multiply-add is ever contracted.
#include "mr_trace_kernel.cuh"
#include "mr_launch.hpp"

namespace mr {
cudaError_t launch_trace_strict(const TraceArgs &a, cudaStream_t stream)
{
    return launch_trace_math<MR_MATH_STRICT>(a, stream);
}
}  // namespace mr
