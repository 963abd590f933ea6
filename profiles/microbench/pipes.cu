// pipes.cu — per-SM throughput of the instructions the ray-tracing kernel is made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Prints thread-instructions per clock per SM for each op (8 independent chains per thread,
// 1024 threads per block, 2 blocks per SM), clock from the SM cycle counter.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int OP>
__device__ __forceinline__ double op(double a, double b, double c)
{
    if (OP == 0) return fma(a, b, c);                         // DFMA
    if (OP == 1) return __dadd_rn(a, c);                      // DADD
    if (OP == 2) return __dmul_rn(a, b);                      // DMUL
    if (OP == 3) return (double)(float)a + 0.0;               // F2F.F32.F64 + F2F.F64.F32 (+DADD)
    if (OP == 4) return a > c ? b : a;                        // DSETP + select
    if (OP == 5) return (double)__double2int_rd(a);           // F2I.F64 + I2F.F64
    if (OP == 6) return 1.0 / a;                              // f64 divide (reciprocal)
    if (OP == 7) return rsqrt(a);                             // f64 rsqrt
    if (OP == 8) return sqrt(a);                              // f64 sqrt
    if (OP == 9) return exp(-a);                              // f64 exp
    if (OP == 10) return expm1(-a);                           // f64 expm1
    if (OP == 11) return a / c;                               // f64 divide (full)
    if (OP == 12) return tanh(a);
    if (OP == 13) return atan2(a, c);
    if (OP == 14) { double s, cc; sincos(a, &s, &cc); return s + cc; }
    if (OP == 15) return (double)__fdiv_rn((float)a, 3.0f);   // f32 IEEE divide (+2 cvt)
    if (OP == 16) return (double)__float_as_int(__fmaf_rn(__int_as_float((int)a), 1.0001f, 0.5f)); // FFMA (+cvt noise)
    return a;
}

template <int OP>
__global__ void __launch_bounds__(1024) k(double *out, long long *cycles, double seed)
{
    double v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = seed + 1e-3 * (threadIdx.x + i);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) v[i] = op<OP>(v[i], 0.999999, 1.0000001);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// DFMA / DMUL / DADD with every operand a distinct live register (register-file read bandwidth)
__global__ void __launch_bounds__(1024) k3reg(double *out, long long *cycles, double seed, int mode)
{
    double v[CHAINS], b[CHAINS], c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { v[i] = seed + 1e-3 * (threadIdx.x + i); b[i] = 0.999999 + 1e-9 * (threadIdx.x + i); c[i] = 1e-7 * (i + 1 + threadIdx.x); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) v[i] = fma(v[i], b[i], c[i]);
        } else if (mode == 1) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) v[i] = fma(v[i], b[i], v[(i + 1) % CHAINS]);
        } else {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) v[i] = __dmul_rn(v[i], b[i]);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += v[i] + b[i] + c[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
void run3(const char *name, int mode)
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 2;
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * blocks);
    k3reg<<<blocks, 1024>>>(out, cyc, 1.0, mode); cudaDeviceSynchronize();
    k3reg<<<blocks, 1024>>>(out, cyc, 1.0, mode); cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, cyc, 8 * (blocks < 1024 ? blocks : 1024), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks && i < 1024; ++i) avg += h[i]; avg /= blocks;
    printf("%-34s %8.2f thread-ops/clk/SM\n", name, 2.0 * 1024 * CHAINS * ITERS / avg);
    cudaFree(out); cudaFree(cyc);
}

template <int OP>
void run(const char *name, double seed)
{
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = sms * 2;
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * blocks);
    k<OP><<<blocks, 1024>>>(out, cyc, seed);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, 1024>>>(out, cyc, seed);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; cudaMemcpy(h, cyc, 8 * (blocks < 1024 ? blocks : 1024), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks && i < 1024; ++i) avg += h[i]; avg /= blocks;
    double ops_per_sm = 2.0 * 1024 * CHAINS * ITERS;      // 2 blocks resident per SM
    printf("%-34s %8.2f thread-ops/clk/SM   (%.3f ms, %.1f Gops/s)\n", name, ops_per_sm / avg, ms,
           (double)blocks * 1024 * CHAINS * ITERS / ms / 1e6);
    cudaFree(out); cudaFree(cyc);
}

template <int OP>
__global__ void lat_kernel(double *out, long long *cycles, double seed)
{
    double v = seed + 1e-3 * threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < 4096; ++it) v = op<OP>(v, 0.999999, 1.0000001);
    long long t1 = clock64();
    if (v == 12345.678) out[0] = v;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int OP>
void lat(const char *name, double seed)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    lat_kernel<OP><<<1, 32>>>(out, cyc, seed);
    cudaDeviceSynchronize();
    lat_kernel<OP><<<1, 32>>>(out, cyc, seed);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %8.2f cycles dependent-issue latency (1 warp)\n", name, (double)h / 4096);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run3("DFMA 3 distinct regs", 0);
    run3("DFMA 3 regs, cross-chain", 1);
    run3("DMUL 2 distinct regs", 2);
    lat<0>("DFMA latency", 1.0);
    lat<1>("DADD latency", 1.0);
    lat<2>("DMUL latency", 1.0);
    lat<6>("f64 1/x latency", 1.3);
    lat<7>("f64 rsqrt latency", 1.3);
    lat<9>("f64 exp latency", 0.7);
    lat<3>("cvt f64->f32->f64+DADD latency", 1.0);

    run<0>("DFMA", 1.0);
    run<1>("DADD", 1.0);
    run<2>("DMUL", 1.0);
    run<3>("cvt f64->f32->f64 (+DADD)", 1.0);
    run<4>("DSETP+select", 1.0);
    run<5>("F2I.F64 floor + I2F.F64", 100.5);
    run<6>("f64 reciprocal 1/x", 1.3);
    run<7>("f64 rsqrt", 1.3);
    run<8>("f64 sqrt", 1.3);
    run<9>("f64 exp(-x)", 0.7);
    run<10>("f64 expm1(-x)", 0.7);
    run<11>("f64 divide a/c", 1.3);
    run<12>("f64 tanh", 0.7);
    run<13>("f64 atan2", 0.7);
    run<14>("f64 sincos", 0.7);
    run<15>("f32 IEEE divide (+2 cvt)", 1.3);
    run<16>("FFMA (+cvt)", 100.0);
    return 0;
}
