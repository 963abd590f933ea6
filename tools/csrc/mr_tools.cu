// mr_tools.cu — measurement and self-test plumbing, built as tools/libmr_tools.so.
//
// NOT part of the product library or of its ABI (include/mantaray_b200.h): bench.py uses the DFMA probe as
// the denominator of the FP64 roofline, tests/test_gpu_api.py runs the exhaustive check of the fast path's
// exact f32 division.  Both link against the product's own device code (mr_device.cuh), so what is tested is
// what ships.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "../../mantaray_b200/csrc/mr_device.cuh"
#include "../../mantaray_b200/csrc/mr_internal.hpp"

namespace {

// ---- DFMA probe -------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(double *sink, int iters)
{
    // 8 independent chains per thread keep the FP64 pipe full at any occupancy
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999999, c = 1e-12;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s;     // never true; keeps the chains alive
}

// ---- exhaustive check of fdiv_const ---------------------------------------------
__global__ void fdiv_selftest_kernel(float s, float r, unsigned long long *bad)
{
    unsigned long long local = 0;
    // every non-negative finite float: bit patterns 0 .. 0x7f7fffff
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= 0x7f7fffffull;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float t = __uint_as_float((unsigned)b);
        const float want = __fdiv_rn(t, s), got = mr::fdiv_const(t, s, r);
        // Below |t| = 2^-100 the exact residual t - q*s underflows and the last bits of the
        // quotient may differ; there (s > 1e-30 is enforced) both quotients are in [0, 1): cell 0
        // and in bounds either way, which is all the caller derives from the index.
        const bool tiny = t > 0.0f && t < 7.8886090522101181e-31f;
        const bool same = (__float_as_uint(want) == __float_as_uint(got)) || (isinf(want) && !(got == got)) ||
                          (!(want == want) && !(got == got)) || (tiny && want < 1.0f && got >= 0.0f && got < 1.0f);
        local += same ? 0 : 1;
    }
    if (local) atomicAdd(bad, local);
}

int device_count_quiet()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

}  // namespace

#define TOOLS_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { std::fprintf(stderr, "mr_tools: %s: %s\n", #call, cudaGetErrorString(e__)); return MR_ERR_CUDA; } } while (0)

extern "C" {

/* Sustained FP64 FMA throughput of `device` in TFLOP/s (2 flop per DFMA), measured with a register-only DFMA
 * kernel timed by CUDA events over about `millis` ms of work: the denominator of the FP64 roofline. */
int mrt_measure_fp64_peak(int device, int millis, double *tflops)
{
    if (!tflops) return MR_ERR_BAD_ARG;
    *tflops = 0.0;
    if (device < 0 || device >= device_count_quiet()) return MR_ERR_CUDA;
    int cur = -1;
    TOOLS_CUDA(cudaGetDevice(&cur));
    TOOLS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TOOLS_CUDA(cudaGetDeviceProperties(&prop, device));
    double *sink = nullptr;
    TOOLS_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    TOOLS_CUDA(cudaEventCreate(&e0));
    TOOLS_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8;
    int iters = 2000;
    double best = 0.0;
    float ms = 0.f;
    // warm up, then size the loop for ~millis of work
    dfma_probe_kernel<<<blocks, 256>>>(sink, iters);
    TOOLS_CUDA(cudaDeviceSynchronize());
    for (int rep = 0; rep < 4; ++rep) {
        TOOLS_CUDA(cudaEventRecord(e0, 0));
        dfma_probe_kernel<<<blocks, 256>>>(sink, iters);
        TOOLS_CUDA(cudaEventRecord(e1, 0));
        TOOLS_CUDA(cudaEventSynchronize(e1));
        TOOLS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
        const double tf = flops / ((double)ms * 1e-3) / 1e12;
        if (rep > 0) best = std::max(best, tf);
        if (rep == 0 && ms > 0.f && millis > 0) {
            const double scale = (double)millis / ms;
            iters = (int)std::min(2e6, std::max(200.0, iters * scale));
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (cur != device) cudaSetDevice(cur);
    *tflops = best;
    return MR_OK;
}

/* Self-test of the fast path's exact f32 division by a launch constant (the fractional index of
 * src/bathymetry/cartesian_netcdf3.rs:289): compares mr::fdiv_const with the IEEE divide for EVERY non-negative
 * finite float t, for the divisor `spacing`, on `device`.  *mismatches receives the number of t whose quotients
 * differ in any bit, except that an infinite quotient may come out as NaN (both are out of bounds) and that for
 * 0 < t < 2^-100 (where the exact residual underflows) both quotients only have to lie in [0, 1), i.e. cell 0 and
 * in bounds either way; *usable receives 0 if the library would not use the shortcut for this spacing
 * (mr::recip_ok, the rule upload_fields applies). */
int mrt_selftest_fdiv(int device, float spacing, uint64_t *mismatches, int32_t *usable)
{
    if (!mismatches || !usable) return MR_ERR_BAD_ARG;
    *mismatches = 0;
    float r = 0.f;
    *usable = mr::recip_ok(spacing, &r) ? 1 : 0;
    if (!*usable) return MR_OK;
    if (device < 0 || device >= device_count_quiet()) return MR_ERR_CUDA;
    int cur = -1;
    TOOLS_CUDA(cudaGetDevice(&cur));
    TOOLS_CUDA(cudaSetDevice(device));
    unsigned long long *bad = nullptr, h = 0;
    TOOLS_CUDA(cudaMalloc(&bad, sizeof(*bad)));
    TOOLS_CUDA(cudaMemset(bad, 0, sizeof(*bad)));
    fdiv_selftest_kernel<<<148 * 32, 256>>>(spacing, r, bad);
    TOOLS_CUDA(cudaGetLastError());
    TOOLS_CUDA(cudaMemcpy(&h, bad, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(bad);
    if (cur != device) cudaSetDevice(cur);
    *mismatches = h;
    return MR_OK;
}

}  // extern "C"
