// mr_trace_kernel.cuh — the batch driver as a CUDA kernel.
//
// Replaces ManyRays::trace_many + SingleRay::trace_individual (src/ray.rs:98-127,
// 198-213) and the ode_solvers 0.4.0 Rk4 stepper they call (external crate,
// Cargo.lock:653-656).  One ray per thread; the fixed-step RK4 loop runs entirely
// in registers; every row is written step-major, structure-of-arrays
// (out[field][row][ray]) so a warp stores 256 contiguous bytes per field per row.
//
// Stepper, restated from the published algorithm of ode_solvers::Rk4:
//   push(t0, y0); n = ceil((t_end - t0)/dt); half = dt/2
//   repeat n times:
//     k0 = f(y); k1 = f(y + k0*half); k2 = f(y + k1*half); k3 = f(y + k2*dt)
//     y  = y + (((k0 + k1*2) + k2*2) + k3) * (dt/6);  push(t, y)
//     stop if solout(y, k0): all four of y, or all four of k0, are NaN
//                                                   (src/wave_ray_path.rs:236-246)
// A stopped ray's last row is therefore always all-NaN, and rows it never
// reaches are NaN too (python/mantaray/core.py:115-119), so a stopped lane just
// keeps storing its NaN state; when a whole warp has stopped it leaves the RK4
// loop and only fills.
//
// The four stages run as ONE loop body (stage offset a_s in {0, dt/2, dt/2, dt},
// weight w_s in {1, 2, 2, 1}; multiplying by 1 or 2 is exact, so the accumulation
// order is the reference's) to keep the kernel inside the instruction cache.
#pragma once
#include "mr_device.cuh"

namespace mr {

struct TraceArgs {
    BathyDev   b;
    CurrentDev c;
    int64_t n;                 // rays
    const double *x0, *y0, *kx0, *ky0;
    double dt;
    int64_t nsteps;
    int32_t stride;
    double *x, *y, *kx, *ky;   // [rows][ld] or all NULL
    int64_t ld;
    int32_t *rows, *len;       // [n] or NULL
    double *fin;               // [4][n] or NULL
};

static constexpr int kBlock = kBlockThreads;

// build-time tuning knobs (see profiles/): unroll factor of the RK4 stage loop and the
// resident-blocks-per-SM target of the fast kernel
#ifndef MR_STAGE_UNROLL
#define MR_STAGE_UNROLL 1
#endif
#ifndef MR_MIN_BLOCKS
#define MR_MIN_BLOCKS 7
#endif
#ifndef MR_STREAM_STORES
#define MR_STREAM_STORES 1
#endif
static constexpr int kStageUnroll = MR_STAGE_UNROLL;

__device__ __forceinline__ bool any_nan4(const double y[4])
{
    return isnan(y[0]) || isnan(y[1]) || isnan(y[2]) || isnan(y[3]);
}
__device__ __forceinline__ bool all_nan4(const double y[4])
{
    return isnan(y[0]) && isnan(y[1]) && isnan(y[2]) && isnan(y[3]);
}

template <int BK, int CK, int MATH, bool UNI>
__global__ void __launch_bounds__(kBlock, (MATH == MR_MATH_FAST) ? MR_MIN_BLOCKS : 1)
trace_kernel(const __grid_constant__ TraceArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool valid = i < a.n;
    const bool store = valid && a.x != nullptr;
    const double dt = a.dt;
    const double half = dt / 2.0;
    const double sixth = dt / 6.0;

    double y[4];
    y[0] = valid ? a.x0[i]  : qnan();
    y[1] = valid ? a.y0[i]  : qnan();
    y[2] = valid ? a.kx0[i] : qnan();
    y[3] = valid ? a.ky0[i] : qnan();

    if (store) {
        a.x[i] = y[0]; a.y[i] = y[1]; a.kx[i] = y[2]; a.ky[i] = y[3];
    }

    bool alive = valid && a.nsteps > 0;
    bool clean = !any_nan4(y);             // no NaN seen yet: rows so far all count towards len
    int32_t rows = 1;
    int32_t len = clean ? 1 : 0;
    if (valid && a.fin && !clean) {        // no NaN-free row at all
        a.fin[i] = qnan(); a.fin[a.n + i] = qnan(); a.fin[2 * a.n + i] = qnan(); a.fin[3 * a.n + i] = qnan();
    }

    const int32_t nsteps = (int32_t)a.nsteps;      // < 2^31 (mr_num_steps)
    int32_t until_store = a.stride;        // counts down to the next stored row
    int64_t o = i;                         // offset of this ray in the last stored row
    int32_t rows_left = nsteps / a.stride; // stored rows still to write
    for (int32_t s = 1; s <= nsteps; ++s) {
        if (!__any_sync(0xffffffffu, alive)) break;
        if (alive) {
            double k[4] = {0.0, 0.0, 0.0, 0.0};
            double acc[4] = {-0.0, -0.0, -0.0, -0.0};     // -0 + k0 == k0 for every k0
            bool k0_nan = false;
#pragma unroll kStageUnroll
            for (int st = 0; st < 4; ++st) {
                const double as = (st == 0) ? 0.0 : (st == 3 ? dt : half);
                const double ws = (st == 1 || st == 2) ? 2.0 : 1.0;
                double yt[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double adv = (MATH == MR_MATH_STRICT) ? __dadd_rn(y[c], __dmul_rn(k[c], as)) : fma(k[c], as, y[c]);
                    // stage 0 evaluates f(y): k is still 0 there, and y + 0*0 == y (a -0 component
                    // would become +0, which the strict path must not allow)
                    yt[c] = (MATH == MR_MATH_STRICT && st == 0) ? y[c] : adv;
                }
                rhs<BK, CK, MATH, UNI>(a.b, a.c, yt[0], yt[1], yt[2], yt[3], k);
                if (st == 0) k0_nan = all_nan4(k);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    acc[c] = (MATH == MR_MATH_STRICT) ? __dadd_rn(acc[c], __dmul_rn(k[c], ws)) : fma(k[c], ws, acc[c]);
            }
            double yn[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                yn[c] = (MATH == MR_MATH_STRICT) ? __dadd_rn(y[c], __dmul_rn(acc[c], sixth)) : fma(acc[c], sixth, y[c]);
            rows = s + 1;
            const bool n0 = isnan(yn[0]), n1 = isnan(yn[1]), n2 = isnan(yn[2]), n3 = isnan(yn[3]);
            if (clean) {
                if (n0 || n1 || n2 || n3) {
                    clean = false;
                    if (a.fin) {           // y is the last NaN-free row
                        a.fin[i] = y[0]; a.fin[a.n + i] = y[1]; a.fin[2 * a.n + i] = y[2]; a.fin[3 * a.n + i] = y[3];
                    }
                } else {
                    len = s + 1;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) y[c] = yn[c];
            if (k0_nan || (n0 && n1 && n2 && n3)) alive = false;    // solout
        }
        if (--until_store == 0) {
            until_store = a.stride;
            o += a.ld;
            --rows_left;
            if (store) {
#if MR_STREAM_STORES
                __stcs(a.x + o, y[0]); __stcs(a.y + o, y[1]); __stcs(a.kx + o, y[2]); __stcs(a.ky + o, y[3]);
#else
                a.x[o] = y[0]; a.y[o] = y[1]; a.kx[o] = y[2]; a.ky[o] = y[3];
#endif
            }
        }
    }
    // whole warp stopped: rows it never reached are NaN
    if (store) {
        const double nan = qnan();
        for (; rows_left > 0; --rows_left) {
            o += a.ld;
            a.x[o] = nan; a.y[o] = nan; a.kx[o] = nan; a.ky[o] = nan;
        }
    }
    if (valid) {
        if (a.rows) a.rows[i] = rows;
        if (a.len)  a.len[i]  = len;
        if (a.fin && clean) {
            a.fin[i] = y[0]; a.fin[a.n + i] = y[1]; a.fin[2 * a.n + i] = y[2]; a.fin[3 * a.n + i] = y[3];
        }
    }
}

// One instantiation per (bathymetry kind, current kind[, uniform grids]); the kinds are
// uniform over a launch, so the dispatch is a host-side switch.
template <int MATH>
static cudaError_t launch_trace_math(const TraceArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((a.n + kBlock - 1) / kBlock);
    // the fast path's affine-coordinate specialisation needs every gridded field to qualify
    const bool uni = MATH == MR_MATH_FAST &&
                     (a.b.kind != MR_BATHY_GRID || a.b.uniform) && (a.c.kind != MR_CURRENT_GRID || a.c.uniform) &&
                     (a.b.kind == MR_BATHY_GRID || a.c.kind == MR_CURRENT_GRID);
#define MR_CASE(BKV, CKV)                                                                                      \
    if (a.b.kind == BKV && a.c.kind == CKV) {                                                                  \
        if (uni) trace_kernel<BKV, CKV, MATH, (MATH == MR_MATH_FAST)><<<grid, kBlock, 0, stream>>>(a);         \
        else     trace_kernel<BKV, CKV, MATH, false><<<grid, kBlock, 0, stream>>>(a);                          \
        return cudaGetLastError();                                                                             \
    }
    MR_CASE(MR_BATHY_CONSTANT, MR_CURRENT_CONSTANT)
    MR_CASE(MR_BATHY_CONSTANT, MR_CURRENT_GRID)
    MR_CASE(MR_BATHY_SLOPE,    MR_CURRENT_CONSTANT)
    MR_CASE(MR_BATHY_SLOPE,    MR_CURRENT_GRID)
    MR_CASE(MR_BATHY_GRID,     MR_CURRENT_CONSTANT)
    MR_CASE(MR_BATHY_GRID,     MR_CURRENT_GRID)
    MR_CASE(MR_BATHY_ARRAY,    MR_CURRENT_CONSTANT)
    MR_CASE(MR_BATHY_ARRAY,    MR_CURRENT_GRID)
#undef MR_CASE
    return cudaErrorInvalidValue;
}

}  // namespace mr
