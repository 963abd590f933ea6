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
//     (k0_c enters component c of the sum, so "all four of k0 NaN" implies "all four of y NaN": one test)
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
    int32_t deep_map;          // use the bathymetry's depth-floor map where the grids allow it (MR_OPT_DEEP_MAP)
    int32_t same_grid;         // use the same-grid shortcut where the grids allow it (MR_OPT_SAME_GRID)
    int32_t current_map;       // use the uniform-current map where the current grid has one
    // filled by launch_trace_math: the trajectory planes as byte offsets from the x plane, the row pitch in bytes
    int64_t off_y, off_kx, off_ky, row_bytes;
    double sixth;              // dt / 6 (read from here by the depth-floor-map variant, which is short of registers)
    // the stage offsets {0, dt/2, dt/2, dt} and weights {1, 2, 2, 1} of RK4: the looped stages index them with the
    // stage number (one constant-bank load each) instead of deriving them with compares and branches
    double stage_a[4], stage_w[4];
};

static constexpr int kBlock = kBlockThreads;

// Which specialisations of the kernel a launch with these arguments runs (launch_trace_math dispatches on it,
// mr_trace_plan reports it).
struct TracePlan {
    bool uni;    // affine-coordinate fast path: every gridded field qualifies
    bool dmap;   // depth-floor map: the affine fast path on a gridded bathymetry that has one
    bool sg;     // same-grid shortcut: both fields gridded, affine, and on one grid (BathyDev::same_grid, set at upload)
    bool cmap;   // uniform-current map: the affine fast path on a gridded current that has one (not with the shortcut)
};
inline TracePlan plan_of(const TraceArgs &a, bool fast)
{
    TracePlan p;
    p.uni = fast && (a.b.kind != MR_BATHY_GRID || a.b.uniform) && (a.c.kind != MR_CURRENT_GRID || a.c.uniform) &&
            (a.b.kind == MR_BATHY_GRID || a.c.kind == MR_CURRENT_GRID);
    p.dmap = p.uni && a.deep_map && a.b.kind == MR_BATHY_GRID && a.b.dmap != nullptr;
    p.sg = p.uni && a.same_grid && a.b.kind == MR_BATHY_GRID && a.c.kind == MR_CURRENT_GRID && a.b.same_grid;
    p.cmap = p.uni && !p.sg && a.current_map && a.c.kind == MR_CURRENT_GRID && a.c.cmap != nullptr;
    return p;
}

// build-time tuning knobs (see profiles/): unroll factor of the RK4 stage loop and the
// resident-blocks-per-SM target of the fast kernel (28 warps per SM: 72 registers)
#ifndef MR_STAGE_UNROLL
#define MR_STAGE_UNROLL 1
#endif
#ifndef MR_STAGE_TABLE
#define MR_STAGE_TABLE 3          // bit 0: uniform-current-map kernels, bit 1: same-grid kernel (see the stage loop)
#endif
#ifndef MR_MIN_BLOCKS
#define MR_MIN_BLOCKS (kWarpsPerSM * 32 / kBlockThreads)
#endif
// experiment knob: the redundant all-NaN test of k0 (see solout below).  Without it ptxas (12.9) spills the
// pre-step state around the stage loop unless that state is parked in shared memory (MR_FIN_SHADOW): measured
// on C4, 51.0 ms with the test, 49.7 ms without it and with the shadow.
#ifndef MR_K0_TEST
#define MR_K0_TEST 0
#endif
#ifndef MR_FIN_SHADOW
#define MR_FIN_SHADOW 1
#endif

// grids whose f32 coordinates are not affine keep the per-cell corner coordinates and basis live: more registers
#ifndef MR_MIN_BLOCKS_GENERIC
#define MR_MIN_BLOCKS_GENERIC (20 * 32 / kBlockThreads)
#endif
// the same-grid kernel (no map) at 24 warps per SM: 80 registers, nothing spilled — with the stage table C5 70.2
// against 71.3 ms at 28 warps without it, C3 with both maps off 52.1 against 53.3 (profiles/r2/kbench_r2k16_same_grid_24_warps.txt)
#ifndef MR_MIN_BLOCKS_SG
#define MR_MIN_BLOCKS_SG (24 * 32 / kBlockThreads)
#endif
// the depth-floor-map variant carries a little more state per thread
#ifndef MR_MIN_BLOCKS_DMAP
#define MR_MIN_BLOCKS_DMAP (kWarpsPerSM * 32 / kBlockThreads)
#endif
static constexpr int kStageUnroll = MR_STAGE_UNROLL;
// (Two rays per thread — interleaved RHS phases, 16-byte row stores — was measured at 2.0e10 ray-steps/s
// against 2.8e10 for one on C4, 168 registers; the kernel carries one ray per thread.)

__device__ __forceinline__ bool any_nan4(const double y[4])
{
    return isnan(y[0]) || isnan(y[1]) || isnan(y[2]) || isnan(y[3]);
}
__device__ __forceinline__ bool all_nan4(const double y[4])
{
    return isnan(y[0]) && isnan(y[1]) && isnan(y[2]) && isnan(y[3]);
}

// Ray carried by this thread.  The threads of the last block that lie past the last ray repeat that
// ray instead of idling: they compute and store bit-identical values to the same addresses, which is
// harmless, and in exchange the kernel has no per-thread validity flag — every vote is over a full warp
// (so the step counter and the row bookkeeping stay warp-uniform) and every store is unconditional.
__device__ __forceinline__ int64_t ray_index(const TraceArgs &a)
{
    return min((int64_t)blockIdx.x * kBlock + threadIdx.x, a.n - 1);
}
// One row: p is this ray's element of the x plane; the other planes lie at fixed byte offsets from it.
__device__ __forceinline__ void store_row(const TraceArgs &a, char *p, const double y[4])
{
    __stcs((double *)p, y[0]);
    __stcs((double *)(p + a.off_y), y[1]);
    __stcs((double *)(p + a.off_kx), y[2]);
    __stcs((double *)(p + a.off_ky), y[3]);
}
__device__ __forceinline__ void store_fin(const TraceArgs &a, const double y[4])
{
    const int64_t i = ray_index(a);
    a.fin[i] = y[0]; a.fin[a.n + i] = y[1]; a.fin[2 * a.n + i] = y[2]; a.fin[3 * a.n + i] = y[3];
}
__device__ __forceinline__ void store_count(const TraceArgs &a, int32_t *dst, int32_t v)
{
    dst[ray_index(a)] = v;
}

// Per-ray bookkeeping is event-driven: `rows` is written when the ray stops and `len` when its first NaN
// appears (each at most once per ray, re-deriving the ray index on the spot), so the step loop carries two
// flags and one row pointer per thread; the step number and the store countdown are warp-uniform.
template <int BK, int CK, int MATH, bool UNI, bool DMAP, bool SG, bool CMAP>
__global__ void __launch_bounds__(kBlock, (MATH == MR_MATH_FAST) ? (UNI ? (DMAP ? MR_MIN_BLOCKS_DMAP : (SG ? MR_MIN_BLOCKS_SG : MR_MIN_BLOCKS)) : MR_MIN_BLOCKS_GENERIC) : 1)
trace_kernel(const __grid_constant__ TraceArgs a)
{
    const bool store = a.x != nullptr;
    const double dt = a.dt;
    const double sixth = DMAP ? a.sixth : dt / 6.0;
    const double half = dt / 2.0;
    const int32_t nsteps = (int32_t)a.nsteps;      // < 2^31 (mr_num_steps)

    double y[4];                           // the state (x, y, kx, ky) of the step being taken
    // MR_FIN_SHADOW: the state before the step is parked in shared memory (each thread its own slots, no
    // synchronisation) for the one event that needs it — the first NaN, when the last NaN-free state is written
    // out — instead of being held in registers, or spilled to local memory by the compiler, through the update
    constexpr bool kShadow = MR_FIN_SHADOW && MATH == MR_MATH_FAST;
    __shared__ double yprev[kShadow ? 4 : 1][kShadow ? kBlock : 1];
    char *p;                               // this ray's element of the last stored row of the x plane
    bool alive = nsteps > 0;
    bool clean;                            // no NaN seen yet: rows so far all count towards len
    {
        const int64_t i = ray_index(a);
        y[0] = a.x0[i]; y[1] = a.y0[i]; y[2] = a.kx0[i]; y[3] = a.ky0[i];
        p = (char *)(a.x + i);
        clean = !any_nan4(y);
        if (!clean) {                      // no NaN-free row at all
            if (a.len) store_count(a, a.len, 0);
            if (a.fin) { const double nanrow[4] = {qnan(), qnan(), qnan(), qnan()}; store_fin(a, nanrow); }
        }
        if (store) store_row(a, p, y);
    }

    int32_t until_store = a.stride;        // counts down to the next stored row
    for (int32_t s = 1; s <= nsteps; ++s) {
        if (!__any_sync(0xffffffffu, alive)) break;
        if (alive) {
            double k[1][4], acc[4];
#if MR_K0_TEST
            bool k0_nan;
#endif
#pragma unroll
            for (int c = 0; c < 4; ++c) { k[0][c] = 0.0; acc[c] = -0.0; }    // -0 + k0 == k0 for every k0
#pragma unroll kStageUnroll
            for (int st = 0; st < 4; ++st) {
                // the uniform-current-map kernels and the same-grid kernel read the stage constants from the parameter
                // bank (TraceArgs::stage_a): C2 46.7 against 47.4 ms, C3 40.5 against 41.3, C5 (with 24 warps per SM) 70.2
                // against 71.0; the others derive them (C4 the same either way)
                constexpr bool kTable = MATH == MR_MATH_FAST && (((MR_STAGE_TABLE & 1) && CMAP) || ((MR_STAGE_TABLE & 2) && SG && !DMAP));
                const double as = kTable ? a.stage_a[st] : ((st == 0) ? 0.0 : (st == 3 ? dt : half));
                const double ws = kTable ? a.stage_w[st] : ((st == 1 || st == 2) ? 2.0 : 1.0);
                double yt[1][4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double adv = (MATH == MR_MATH_STRICT) ? __dadd_rn(y[c], __dmul_rn(k[0][c], as)) : fma(k[0][c], as, y[c]);
                    // stage 0 evaluates f(y): k is 0 there, and y + 0*0 == y (a -0 component would become
                    // +0, which the strict path must not allow)
                    yt[0][c] = (MATH == MR_MATH_STRICT && st == 0) ? y[c] : adv;
                }
                rhs<BK, CK, MATH, UNI, 1, DMAP, SG, CMAP>(a.b, a.c, yt, k);
#if MR_K0_TEST
                if (st == 0) k0_nan = all_nan4(k[0]);
#endif
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    acc[c] = (MATH == MR_MATH_STRICT) ? __dadd_rn(acc[c], __dmul_rn(k[0][c], ws)) : fma(k[0][c], ws, acc[c]);
            }
            double yn[4];
            if (kShadow && a.fin) {
#pragma unroll
                for (int c = 0; c < 4; ++c) yprev[c][threadIdx.x] = y[c];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                yn[c] = (MATH == MR_MATH_STRICT) ? __dadd_rn(y[c], __dmul_rn(acc[c], sixth)) : fma(acc[c], sixth, y[c]);
            const bool n0 = isnan(yn[0]), n1 = isnan(yn[1]), n2 = isnan(yn[2]), n3 = isnan(yn[3]);
            if (clean && (n0 || n1 || n2 || n3)) {
                clean = false;
                if (a.len) store_count(a, a.len, s);
                if (a.fin) {                          // the row before this one is the last NaN-free state
                    double yo[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) yo[c] = kShadow ? yprev[c][threadIdx.x] : y[c];
                    store_fin(a, yo);
                }
            }
            // solout(y_new, k0) also stops on an all-NaN k0 — which needs no test of its own: k0 enters every
            // component of the sum, so an all-NaN k0 has made y_new all NaN
#if MR_K0_TEST
            if (k0_nan || (n0 && n1 && n2 && n3)) {
#else
            if (n0 && n1 && n2 && n3) {
#endif
                alive = false;
                if (a.rows) store_count(a, a.rows, s + 1);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) y[c] = yn[c];
        }
        if (--until_store == 0) {
            until_store = a.stride;
            p += a.row_bytes;
            if (store) store_row(a, p, y);
        }
    }
    // whole warp stopped early: the rows it never reached are NaN (how many is read off the row pointer,
    // so that nothing but the pointer is carried through the loop for it)
    if (store) {
        const double nanrow[4] = {qnan(), qnan(), qnan(), qnan()};
        char *const last = (char *)(a.x + ray_index(a)) + (int64_t)(nsteps / a.stride) * a.row_bytes;
        while (p != last) {
            p += a.row_bytes;
            store_row(a, p, nanrow);
        }
    }
    // a ray still integrating when the loop ends ran all nsteps (nsteps == 0: the initial row only);
    // one that never met a NaN is still integrating, so its len is its rows
    if (alive || nsteps <= 0) {
        if (a.rows) store_count(a, a.rows, nsteps + 1);
    }
    if (clean) {
        if (a.len) store_count(a, a.len, nsteps + 1);
        if (a.fin) store_fin(a, y);
    }
}

// One instantiation per (bathymetry kind, current kind[, affine grids]); the kinds are uniform over a
// launch, so the dispatch is a host-side switch.
template <int MATH>
static cudaError_t launch_trace_math(const TraceArgs &args, cudaStream_t stream)
{
    if (args.n <= 0) return cudaSuccess;
    TraceArgs a = args;
    a.off_y = (const char *)a.y - (const char *)a.x; a.off_kx = (const char *)a.kx - (const char *)a.x;
    a.off_ky = (const char *)a.ky - (const char *)a.x; a.row_bytes = a.ld * (int64_t)sizeof(double);
    a.sixth = a.dt / 6.0;
    a.stage_a[0] = 0.0; a.stage_a[1] = a.stage_a[2] = a.dt / 2.0; a.stage_a[3] = a.dt;
    a.stage_w[0] = a.stage_w[3] = 1.0; a.stage_w[1] = a.stage_w[2] = 2.0;
    const TracePlan plan = plan_of(a, MATH == MR_MATH_FAST);
    const bool uni = plan.uni, dmap = plan.dmap, sg = plan.sg, cmap = plan.cmap;
    const unsigned grid = (unsigned)((a.n + (int64_t)kBlock - 1) / (int64_t)kBlock);
    constexpr bool kFast = MATH == MR_MATH_FAST;
#define MR_LAUNCH(BKV, CKV, UNIV)                                                                              \
    do {                                                                                                       \
        constexpr bool kGG = kFast && UNIV && BKV == MR_BATHY_GRID && CKV == MR_CURRENT_GRID;                  \
        constexpr bool kBG = kFast && UNIV && BKV == MR_BATHY_GRID;                                            \
        constexpr bool kCG = kFast && UNIV && CKV == MR_CURRENT_GRID;                                          \
        if (dmap && sg) trace_kernel<BKV, CKV, MATH, UNIV, kBG, kGG, false><<<grid, kBlock, 0, stream>>>(a);   \
        else if (sg) trace_kernel<BKV, CKV, MATH, UNIV, false, kGG, false><<<grid, kBlock, 0, stream>>>(a);    \
        else if (dmap && cmap) trace_kernel<BKV, CKV, MATH, UNIV, kBG, false, kCG><<<grid, kBlock, 0, stream>>>(a); \
        else if (cmap) trace_kernel<BKV, CKV, MATH, UNIV, false, false, kCG><<<grid, kBlock, 0, stream>>>(a);  \
        else if (dmap) trace_kernel<BKV, CKV, MATH, UNIV, kBG, false, false><<<grid, kBlock, 0, stream>>>(a);  \
        else trace_kernel<BKV, CKV, MATH, UNIV, false, false, false><<<grid, kBlock, 0, stream>>>(a);          \
    } while (0)
#define MR_CASE(BKV, CKV)                                                                                      \
    if (a.b.kind == BKV && a.c.kind == CKV) {                                                                  \
        if (uni) MR_LAUNCH(BKV, CKV, kFast);                                                                   \
        else MR_LAUNCH(BKV, CKV, false);                                                                       \
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
