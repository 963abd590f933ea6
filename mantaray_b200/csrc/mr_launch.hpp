// mr_launch.hpp — host-callable launchers of the trace kernel.
#pragma once
#include <cuda_runtime.h>
namespace mr {
struct TraceArgs;
cudaError_t launch_trace_fast(const TraceArgs &a, cudaStream_t stream);
cudaError_t launch_trace_strict(const TraceArgs &a, cudaStream_t stream);
// register-only DFMA loop used by mr_measure_fp64_peak
cudaError_t launch_dfma_probe(double *sink, int iters, int blocks, cudaStream_t stream);
}  // namespace mr
