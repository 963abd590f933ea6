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
// grids whose f32 coordinates are not affine keep the per-cell corner coordinates and basis live: more registers
#ifndef MR_MIN_BLOCKS_GENERIC
#define MR_MIN_BLOCKS_GENERIC 5
#endif
#ifndef MR_MIN_BLOCKS_NR2
#define MR_MIN_BLOCKS_NR2 4
#endif
// Rays per thread.  2 (adjacent rays, interleaved RHS phases, 16-byte row stores) was measured at
// 2.0e10 ray-steps/s against 2.8e10 for 1 on C4 (168 registers, 12 warps/SM): not compiled by default.
#ifndef MR_RAYS_PER_THREAD
#define MR_RAYS_PER_THREAD 1
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

// store one row of the NR rays of a thread (ray index o, o+1 adjacent in memory)
template <int NR>
__device__ __forceinline__ void store_row(const TraceArgs &a, int64_t o, const double (&y)[NR][4], const bool (&valid)[NR])
{
    if (NR == 2 && valid[1]) {             // both rays: one 16-byte store per field (host guarantees alignment)
        __stcs(reinterpret_cast<double2 *>(a.x + o),  make_double2(y[0][0], y[1][0]));
        __stcs(reinterpret_cast<double2 *>(a.y + o),  make_double2(y[0][1], y[1][1]));
        __stcs(reinterpret_cast<double2 *>(a.kx + o), make_double2(y[0][2], y[1][2]));
        __stcs(reinterpret_cast<double2 *>(a.ky + o), make_double2(y[0][3], y[1][3]));
    } else if (valid[0]) {
        __stcs(a.x + o, y[0][0]); __stcs(a.y + o, y[0][1]); __stcs(a.kx + o, y[0][2]); __stcs(a.ky + o, y[0][3]);
    }
}

// NR rays per thread (adjacent rays i0, i0+1): the RHS phases of the rays interleave (mr_device.cuh,
// rhs_fast_n), per-thread uniform work is shared, and rows are stored 16 bytes at a time.
template <int BK, int CK, int MATH, bool UNI, int NR>
__global__ void __launch_bounds__(kBlock, (MATH == MR_MATH_FAST) ? (NR == 2 ? MR_MIN_BLOCKS_NR2 : (UNI ? MR_MIN_BLOCKS : MR_MIN_BLOCKS_GENERIC)) : 1)
trace_kernel(const __grid_constant__ TraceArgs a)
{
    const int64_t i0 = ((int64_t)blockIdx.x * kBlock + threadIdx.x) * NR;
    const bool store = a.x != nullptr;
    const double dt = a.dt;
    const double half = dt / 2.0;
    const double sixth = dt / 6.0;

    bool valid[NR], alive[NR], clean[NR];
    int32_t rows[NR], len[NR];
    double y[NR][4];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int64_t i = i0 + r;
        valid[r] = i < a.n;
        y[r][0] = valid[r] ? a.x0[i]  : qnan();
        y[r][1] = valid[r] ? a.y0[i]  : qnan();
        y[r][2] = valid[r] ? a.kx0[i] : qnan();
        y[r][3] = valid[r] ? a.ky0[i] : qnan();
        alive[r] = valid[r] && a.nsteps > 0;
        clean[r] = !any_nan4(y[r]);        // no NaN seen yet: rows so far all count towards len
        rows[r] = 1;
        len[r] = clean[r] ? 1 : 0;
        if (valid[r] && a.fin && !clean[r]) {   // no NaN-free row at all
            a.fin[i] = qnan(); a.fin[a.n + i] = qnan(); a.fin[2 * a.n + i] = qnan(); a.fin[3 * a.n + i] = qnan();
        }
    }
    if (store) store_row<NR>(a, i0, y, valid);

    const int32_t nsteps = (int32_t)a.nsteps;      // < 2^31 (mr_num_steps)
    int32_t until_store = a.stride;        // counts down to the next stored row
    int64_t o = i0;                        // offset of this thread's rays in the last stored row
    int32_t rows_left = nsteps / a.stride; // stored rows still to write
    for (int32_t s = 1; s <= nsteps; ++s) {
        bool any_alive = alive[0];
#pragma unroll
        for (int r = 1; r < NR; ++r) any_alive = any_alive || alive[r];
        if (!__any_sync(0xffffffffu, any_alive)) break;
        if (any_alive) {
            double k[NR][4], acc[NR][4];
            bool k0_nan[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) { k[r][c] = 0.0; acc[r][c] = -0.0; }    // -0 + k0 == k0 for every k0
#pragma unroll kStageUnroll
            for (int st = 0; st < 4; ++st) {
                const double as = (st == 0) ? 0.0 : (st == 3 ? dt : half);
                const double ws = (st == 1 || st == 2) ? 2.0 : 1.0;
                double yt[NR][4];
#pragma unroll
                for (int r = 0; r < NR; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double adv = (MATH == MR_MATH_STRICT) ? __dadd_rn(y[r][c], __dmul_rn(k[r][c], as)) : fma(k[r][c], as, y[r][c]);
                        // stage 0 evaluates f(y): k is still 0 there, and y + 0*0 == y (a -0 component
                        // would become +0, which the strict path must not allow)
                        yt[r][c] = (MATH == MR_MATH_STRICT && st == 0) ? y[r][c] : adv;
                    }
                rhs<BK, CK, MATH, UNI, NR>(a.b, a.c, yt, k);
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    if (st == 0) k0_nan[r] = all_nan4(k[r]);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        acc[r][c] = (MATH == MR_MATH_STRICT) ? __dadd_rn(acc[r][c], __dmul_rn(k[r][c], ws)) : fma(k[r][c], ws, acc[r][c]);
                }
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                double yn[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    yn[c] = (MATH == MR_MATH_STRICT) ? __dadd_rn(y[r][c], __dmul_rn(acc[r][c], sixth)) : fma(acc[r][c], sixth, y[r][c]);
                const bool n0 = isnan(yn[0]), n1 = isnan(yn[1]), n2 = isnan(yn[2]), n3 = isnan(yn[3]);
                if (alive[r]) {
                    const int64_t i = i0 + r;
                    rows[r] = s + 1;
                    if (clean[r]) {
                        if (n0 || n1 || n2 || n3) {
                            clean[r] = false;
                            if (a.fin) {       // y is the last NaN-free row
                                a.fin[i] = y[r][0]; a.fin[a.n + i] = y[r][1]; a.fin[2 * a.n + i] = y[r][2]; a.fin[3 * a.n + i] = y[r][3];
                            }
                        } else {
                            len[r] = s + 1;
                        }
                    }
                    if (k0_nan[r] || (n0 && n1 && n2 && n3)) alive[r] = false;    // solout
                    // (a stopped ray's state is all-NaN from here on, whatever its sibling does)
#pragma unroll
                    for (int c = 0; c < 4; ++c) y[r][c] = yn[c];
                }
            }
        }
        if (--until_store == 0) {
            until_store = a.stride;
            o += a.ld;
            --rows_left;
            if (store) store_row<NR>(a, o, y, valid);
        }
    }
    // whole warp stopped: rows it never reached are NaN
    if (store) {
        double nanrow[NR][4];
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) nanrow[r][c] = qnan();
        for (; rows_left > 0; --rows_left) {
            o += a.ld;
            store_row<NR>(a, o, nanrow, valid);
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int64_t i = i0 + r;
        if (valid[r]) {
            if (a.rows) a.rows[i] = rows[r];
            if (a.len)  a.len[i]  = len[r];
            if (a.fin && clean[r]) {
                a.fin[i] = y[r][0]; a.fin[a.n + i] = y[r][1]; a.fin[2 * a.n + i] = y[r][2]; a.fin[3 * a.n + i] = y[r][3];
            }
        }
    }
}

// One instantiation per (bathymetry kind, current kind[, affine grids, rays per thread]); the
// kinds are uniform over a launch, so the dispatch is a host-side switch.
template <int MATH>
static cudaError_t launch_trace_math(const TraceArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    // the fast path's affine-coordinate specialisation needs every gridded field to qualify
    const bool uni = MATH == MR_MATH_FAST &&
                     (a.b.kind != MR_BATHY_GRID || a.b.uniform) && (a.c.kind != MR_CURRENT_GRID || a.c.uniform) &&
                     (a.b.kind == MR_BATHY_GRID || a.c.kind == MR_CURRENT_GRID);
    // two rays per thread need 16-byte aligned row pairs: even pitch and aligned planes
    const bool aligned = a.x == nullptr || (a.ld % 2 == 0 && ((uintptr_t)a.x | (uintptr_t)a.y | (uintptr_t)a.kx | (uintptr_t)a.ky) % 16 == 0);
#if MR_RAYS_PER_THREAD == 2
    const bool two = MATH == MR_MATH_FAST && aligned && a.n >= 2 * kBlock;
#else
    const bool two = false;
    (void)aligned;
#endif
    const int nr = two ? 2 : 1;
    const unsigned grid = (unsigned)((a.n + (int64_t)kBlock * nr - 1) / ((int64_t)kBlock * nr));
#define MR_LAUNCH(BKV, CKV, UNIV, NRV) trace_kernel<BKV, CKV, MATH, UNIV, NRV><<<grid, kBlock, 0, stream>>>(a)
    constexpr int kNr2 = (MATH == MR_MATH_FAST && MR_RAYS_PER_THREAD == 2) ? 2 : 1;
    constexpr bool kFast = MATH == MR_MATH_FAST;
#define MR_CASE(BKV, CKV)                                                                                      \
    if (a.b.kind == BKV && a.c.kind == CKV) {                                                                  \
        if (uni && two) MR_LAUNCH(BKV, CKV, kFast, kNr2);                                                      \
        else if (uni) MR_LAUNCH(BKV, CKV, kFast, 1);                                                           \
        else if (two) MR_LAUNCH(BKV, CKV, false, kNr2);                                                        \
        else MR_LAUNCH(BKV, CKV, false, 1);                                                                    \
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
#undef MR_LAUNCH
    return cudaErrorInvalidValue;
}

}  // namespace mr
