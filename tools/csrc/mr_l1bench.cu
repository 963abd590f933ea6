// L1 data-stage microbenchmark: what one 256-bit gather costs the L1TEX data pipe as a function of WHERE in their
// 128-byte lines the lanes' 32-byte sectors lie.  Measurement plumbing, not part of the product.
//
//   mr_l1bench            prints one line per pattern: ns per warp-load (CUDA events)
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,
//       l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ./mr_l1bench     wavefronts per request, per pattern
//
// Patterns (every lane reads 32 bytes; "line" = 128-byte line, chosen per lane and iteration by a hash):
//   0  all lanes the same sector                       (broadcast)
//   1  lane i sector i of 8 consecutive lines          (perfectly coalesced kilobyte)
//   2  distinct lines, every lane at offset 0          (one bank group)
//   3  distinct lines, offset 64*(lane&1)              (two bank groups: 64-byte records, first halves)
//   4  distinct lines, offset 32*(lane&3)              (four bank groups: 32-byte records side by side)
//   5  distinct lines, offset 32*(hash&3)              (four groups, random)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int P>
__global__ void __launch_bounds__(128, 7) gather(const char *base, uint32_t line_mask, int iters, float *out)
{
    const uint32_t lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const uint32_t h = mix(warp * 0x9e3779b9u + (uint32_t)it * 32u + (P == 0 ? 0u : lane));
        uint32_t line = h & line_mask, off = 0;
        if (P == 1) { line = (mix(warp * 0x9e3779b9u + (uint32_t)it) & line_mask & ~7u) + (lane >> 2); off = 32u * (lane & 3u); }
        if (P == 3) off = 64u * (lane & 1u);
        if (P == 4) off = 32u * (lane & 3u);
        if (P == 5) off = 32u * ((h >> 27) & 3u);
        float a0, a1, a2, a3, a4, a5, a6, a7;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7)
                     : "l"(base + (size_t)line * 128u + off));
        acc += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    }
    if (acc == 12345.678f) out[0] = acc;
}

template <int P>
static void run(const char *base, uint32_t mask, int iters, float *out, const char *what)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 7;
    gather<P><<<grid, 128>>>(base, mask, iters, out);
    cudaEventRecord(e0);
    gather<P><<<grid, 128>>>(base, mask, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double loads = (double)grid * 4 * iters;
    printf("{\"pattern\": %d, \"what\": \"%s\", \"footprint_MB\": %.1f, \"ms\": %.3f, \"ns_per_warp_load_per_sm\": %.2f}\n", P, what,
           (mask + 1.0) * 128 / 1e6, ms, ms * 1e6 / (loads / 148));
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 4096;
    for (uint32_t lines : {1u << 10, 1u << 19}) {       // 128 KB (L1-resident) and 64 MB (L2-resident)
        char *buf; float *out;
        cudaMalloc(&buf, (size_t)lines * 128); cudaMemset(buf, 0, (size_t)lines * 128); cudaMalloc(&out, 4);
        run<0>(buf, lines - 1, iters, out, "broadcast");
        run<1>(buf, lines - 1, iters, out, "coalesced 1 KB");
        run<2>(buf, lines - 1, iters, out, "distinct lines, offset 0");
        run<3>(buf, lines - 1, iters, out, "distinct lines, offset 64*(lane&1)");
        run<4>(buf, lines - 1, iters, out, "distinct lines, offset 32*(lane&3)");
        run<5>(buf, lines - 1, iters, out, "distinct lines, offset 32*random");
        cudaFree(buf); cudaFree(out);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
