// mr_launch.hpp — host-callable launchers of the trace kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
namespace mr {
struct TraceArgs;
struct BathyDev;
struct CurrentDev;
cudaError_t launch_trace_fast(const TraceArgs &a, cudaStream_t stream);
cudaError_t launch_trace_strict(const TraceArgs &a, cudaStream_t stream);
// depth()/current() at rows x n points laid out [rows][ld] (mr_kernels_env.cu); outputs may be NULL
cudaError_t launch_sample(const BathyDev &b, const CurrentDev &c, int64_t rows, int64_t n, int64_t ld,
                          const double *x, const double *y, float *depth, double *u, double *v,
                          cudaStream_t stream);
}  // namespace mr
