// mr_device.cuh — device-side field lookups and the ray-equation right-hand side.
//
// One ray per thread, FP64 state (x, y, kx, ky).  Everything the reference does
// per RHS evaluation (src/wave_ray_path.rs:118-150) is inlined here:
//   bathymetry lookup  src/bathymetry/cartesian_netcdf3.rs:98-135 (+ analytic kinds)
//   current lookup     src/current/cartesian_current.rs:487-542
//   f32 bilinear       src/interpolator.rs:39-84
//   group velocity     src/wave_ray_path.rs:177-188
//   dk/dt              src/wave_ray_path.rs:207-216
//
// Two arithmetic modes (include/mantaray_b200.h):
//   MR_MATH_STRICT  the reference's expression tree operation by operation, on the
//                   f64 node grids exactly as the reference stores them.
//   MR_MATH_FAST    the production path.  The f32 stages (position rounding,
//                   fractional index, cell choice, bilinear) are VALUE-IDENTICAL to
//                   the reference — every f32 operation is an explicit round-to-
//                   nearest intrinsic that nvcc cannot contract — but they read
//                   per-cell records precomputed at upload (corner values already
//                   cast to f32, finite-difference gradients already divided), so an
//                   RHS does no f64->f32 conversions of grid data and no divisions
//                   by grid constants.  The f64 stage is restructured around one
//                   exponential (see rhs_f64_fast).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mantaray_b200.h"

namespace mr {

// experiment knob: with the same-grid shortcut, hand the current's cell geometry to the bathymetry lookup
// (saves ~14 instructions per shallow evaluation, costs the registers that keep it alive)
// experiment knob: deep lanes of the depth-floor map take the fourth-root form (no k, no direction cosines)
// experiment knobs: one higher-order correction step instead of two second-order ones (sqrt/rsqrt, reciprocal), the
// exponential's polynomial as two interleaved Horner chains
// (MR_LEAN_SQRT bit mask: 1 = sqrt(G k tanh kh) of the general wave terms, 2 = sqrt(G k) of the deep-water branch,
// 4 = k = sqrt(k^2) of the kernels without the depth-floor map)
#ifndef MR_LEAN_SQRT
#define MR_LEAN_SQRT 3
#endif
#ifndef MR_LEAN_RCP
#define MR_LEAN_RCP 1
#endif
#ifndef MR_EXP_2CHAIN
#define MR_EXP_2CHAIN 1
#endif
#ifndef MR_DEEP_ROOT4
#define MR_DEEP_ROOT4 1
#endif
// experiment knob: fold the current's four gradients into the two advection sums -kx du/dx - ky dv/dx and
// -kx du/dy - ky dv/dy as soon as the record is there (they need nothing but the wavenumber), so that eight
// registers are not held through the bilinears and the wave terms
// experiment knob: with the depth-floor map, the lanes that are NOT proven deep fetch their bathymetry record only
// when they reach the depth lookup (one more exposed round trip for those lanes) instead of with the other loads
// (eight registers reserved through the current's bilinear in every lane)
#ifndef MR_DMAP_LATE_LOAD
#define MR_DMAP_LATE_LOAD 1
#endif
// (bit mask: 1 = the plain kernel, 2 = the depth-floor-map kernels, 4 = the same-grid kernel without the map)
#ifndef MR_EARLY_GRAD
#define MR_EARLY_GRAD 4
#endif
#ifndef MR_SG_SHARE_GEOM
#define MR_SG_SHARE_GEOM 1
#endif

typedef unsigned long long f32x2;    // two f32 in one 64-bit register (lo, hi), see the packed helpers below

// ---- device-resident field descriptors (passed by value as kernel params) ----

struct BathyDev {
    int32_t kind;
    int32_t nx, ny;
    const float  *x, *y;       // GRID coordinates (f32, as CartesianNetcdf3 holds them)
    const double *depth;       // GRID [ny*nx]                                  (strict path)
    const float  *array;       // ARRAY [nx*ny]
    float h0, x0, y0, dhdx, dhdy;
    // derived at upload (GRID)
    float  xf0, yf0;           // x[0], y[0]
    float  sx, sy;             // |x[1]-x[0]|, |y[1]-y[0]| in f32   (cartesian_netcdf3.rs:287)
    float  rsx, rsy;           // RN(1/sx), RN(1/sy); 0 when the exact-division shortcut does not apply
    int32_t fastdiv;           // fdiv_const(., sx, rsx) and (., sy, rsy) are usable (affine grid or not)
    double x_space, y_space;   // x[1]-x[0] in f64 of the f32 values (cartesian_netcdf3.rs:119-120)
    // per-cell records for the fast path: cell (x1,y1), x1 < nx-1, y1 < ny-1, 32 bytes each:
    //   float4 {z_sw, a10, a01, a11} (bilinear_coeffs of the corner depths as f32), double2 {dhdx, dhdy} (the f32 gradient of
    //   cartesian_netcdf3.rs:134, stored already widened back to f64 as wave_ray_path.rs:125-126 does)
    const float4 *cell;
    float nxm1f, nym1f;        // (nx-1) as f32, (ny-1) as f32: the bound of cartesian_netcdf3.rs:291
    int32_t zero;              // always 0, but opaque to the compiler (see rhs_fast, phase 4)
    // coordinates exactly affine in f32 (x[i] == fmaf(i, dxf, x[0]) for every i, same for y):
    // corner coordinates and the change-of-basis coefficients become launch constants
    int32_t uniform;
    float dxf, dyf, c01, c10;
    // the same constants as (x, y) pairs for the packed f32 path
    f32x2 p0, rs2, ns2, d2, c2;   // {x0,y0} {1/sx,1/sy} {-sx,-sy} {dx,dy} {c10,c01}
    // depth-floor map (see FastRay, DMAP): per block of kDeepBlock x kDeepBlock cells, the square of a lower
    // bound of every depth the f32 bilinear can return anywhere in the block (0: no bound, e.g. a dry or
    // non-finite node); [dmap_nby][dmap_nbx] floats, row-major
    const float *dmap;
    int32_t dmap_nbx;
    // Same-grid shortcut (FastRay, SG): the current lives on this very grid — same shape, same f32 coordinates,
    // and f64 coordinates that are exactly the f32 ones widened.  Then the current's cell (from its f64 index,
    // cartesian_current.rs:246) is the bathymetry's cell (from the f32 index, cartesian_netcdf3.rs:289) whenever
    // the f32 index is further than sg_delta from an integer; sg_lim = 0.5 - sg_delta.  0: not applicable.
    int32_t same_grid;
    float sg_lim;
};
static constexpr int kDeepShift = 3;                 // log2 of the block side, in cells
static constexpr int kDeepBlock = 1 << kDeepShift;

struct CurrentDev {
    int32_t kind;
    int32_t nx, ny;
    const double *x, *y, *u, *v;   // GRID, f64 nodes                                (strict path)
    double u0, v0;
    // derived at upload (GRID)
    double xd0, yd0;           // x[0], y[0]
    double sx, sy;             // |x[1]-x[0]|, |y[1]-y[0]|          (cartesian_current.rs:244)
    double inv_sx, inv_sy;     // RN(1/sx), RN(1/sy)
    double x_space, y_space;   // x[1]-x[0] (signed)                (cartesian_current.rs:515-516)
    // fast path: one 64-byte record per cell: the bilinear_coeffs of the u and of the v corners (as f32)
    // interleaved {u_sw,v_sw,u_a10,v_a10}, {u_a01,v_a01,u_a11,v_a11}, then double2 {dudx,dudy}, double2
    // {dvdx,dvdy} (the f64 finite differences, divided)
    const float4 *cell;
    const float *xf, *yf;      // coordinates cast to f32 (cartesian_current.rs:375-376)
    double nxm1d, nym1d;       // (nx-1), (ny-1) as f64: the bound of cartesian_current.rs:248
    int32_t uniform;
    float xf0, yf0, dxf, dyf, c01, c10;
    f32x2 p0, d2, c2;          // {x0,y0} {dx,dy} {c10,c01} as f32 pairs
    // uniform-current map (see FastRay, CMAP): per block of kDeepBlock x kDeepBlock cells, {u, v} as f32 where every
    // node the block's cells touch holds the same u and the same v (then the reference's bilinear returns exactly
    // that value anywhere in the block and all four finite differences are exactly 0), {NaN, NaN} elsewhere;
    // [cmap_nby][cmap_nbx] float2, row-major
    const float2 *cmap;
    int32_t cmap_nbx;
};

static constexpr double kG = 9.8;            // src/wave_ray_path.rs:23
// Threads per block of the trace kernel: ONE warp.  A block's slot on the SM is free again as soon as its last ray
// has stopped, without waiting for sibling warps; measured against 64 and 128 threads per block (the same 28 warps per
// SM): C4 49.1 / 49.5 / 49.7 ms, the other shapes unchanged (profiles/r2/kbench_r2k14_block_size.txt).
#ifndef MR_BLOCK_THREADS
#define MR_BLOCK_THREADS 32
#endif
static constexpr int kBlockThreads = MR_BLOCK_THREADS;
static constexpr int kWarpsPerSM = 28;                          // resident warps per SM the fast kernels are built for (72 registers)
static constexpr int kMachineRays = 148 * kWarpsPerSM * 32;     // rays that fill a B200 once

// Taylor coefficients 1/13! .. 1/2! of expm1 (see exp_expm1_neg).  In constant memory so
// that each DFMA reads its coefficient as a c[bank][offset] operand instead of the
// compiler materialising 64-bit immediates with two moves per use.
static __constant__ double kExpm1C[12] = {
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};
// {log2(e), 1.5*2^52, ln2_hi, ln2_lo, deep-water threshold on kh, G, G/2, sqrt(G)/2}
static __constant__ double kExpRed[8] = {1.4426950408889634074, 6755399441055744.0,
                                         6.93147180369123816490e-01, 1.90821492927058770002e-10,
                                         22.0, 9.8, 4.9,
                                         1.5652475842498528};      // sqrt(G) / 2: deep-water cg = (sqrt(G)/2) k^-1/2

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ float  qnanf() { return __int_as_float(0x7fc00000); }

// =============================================================================
// interpolator::bilinear on an axis-aligned cell
// =============================================================================
// points a=(xa,ya,z_sw) b=(xa,yb,z_nw) c=(xb,yb,z_ne) d=(xb,ya,z_se), target (tx,ty).
// With this corner order bt=(0,dy), dt=(dx,0), so (interpolator.rs:64-75)
//   det = 0*0 - dx*dy,  c00 = 0/det, c01 = -(dx/det), c10 = -(dy/det), c11 = 0/det,
//   X = c00*tt0 + c01*tt1,  Y = c10*tt0 + c11*tt1.
// c00*tt0 and c11*tt1 are signed zeros and adding a signed zero is exact, so
// X = RN(c01*tt1), Y = RN(c10*tt0) (only the sign of an exactly-zero result can
// differ, which no output can observe).

// strict form: every operation of the reference, branches and all
__device__ __forceinline__ bool bilinear_cell_strict(float xa, float xb, float ya, float yb,
                                                     float zsw, float znw, float zne, float zse,
                                                     float tx, float ty, float &out)
{
    // :46-50 coincidence with a corner, in the order a, b, c, d
    if (tx == xa && ty == ya) { out = zsw; return true; }
    if (tx == xa && ty == yb) { out = znw; return true; }
    if (tx == xb && ty == yb) { out = zne; return true; }
    if (tx == xb && ty == ya) { out = zse; return true; }
    float bt0 = __fsub_rn(xa, xa), bt1 = __fsub_rn(yb, ya);
    float dt0 = __fsub_rn(xb, xa), dt1 = __fsub_rn(ya, ya);
    float tt0 = __fsub_rn(tx, xa), tt1 = __fsub_rn(ty, ya);
    float det = __fsub_rn(__fmul_rn(bt0, dt1), __fmul_rn(dt0, bt1));
    if (det == 0.0f) return false;
    float c00 = __fdiv_rn(dt1, det);
    float c01 = -__fdiv_rn(dt0, det);
    float c10 = -__fdiv_rn(bt1, det);
    float c11 = __fdiv_rn(bt0, det);
    float X = __fadd_rn(__fmul_rn(c00, tt0), __fmul_rn(c01, tt1));
    float Y = __fadd_rn(__fmul_rn(c10, tt0), __fmul_rn(c11, tt1));
    float a10 = __fsub_rn(znw, zsw);
    float a01 = __fsub_rn(zse, zsw);
    float a11 = __fsub_rn(__fsub_rn(__fsub_rn(zne, zsw), a10), a01);
    float r = __fadd_rn(zsw, __fmul_rn(a10, X));
    r = __fadd_rn(r, __fmul_rn(a01, Y));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(a11, X), Y));
    out = r;
    return true;
}

// ---- packed f32 pairs (sm_100 FADD2 / FMUL2 / FFMA2) ------------------------------------------
// Two IEEE round-to-nearest f32 operations per instruction: each half is exactly the scalar _rn
// operation, so the reference's f32 arithmetic can be done two components at a time — (x, y)
// for indices and coordinates, (u, v) for the two current bilinears — at half the issue slots.
__device__ __forceinline__ f32x2 pk(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(f32x2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi_of(f32x2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// fast form.  The cell's fractional coordinates (X, Y) are resolved once per lookup and shared by
// every variable interpolated on that cell; the corner combinations a10 = z_nw - z_sw,
// a01 = z_se - z_sw, a11 = ((z_ne - z_sw) - a10) - a01 of interpolator.rs:78-81 depend on the cell
// only and are part of its record (computed once at upload with these same f32 operations).
__device__ __forceinline__ float4 bilinear_coeffs(float zsw, float znw, float zne, float zse)
{
    const float a10 = __fsub_rn(znw, zsw);
    const float a01 = __fsub_rn(zse, zsw);
    const float a11 = __fsub_rn(__fsub_rn(__fsub_rn(zne, zsw), a10), a01);
    return make_float4(zsw, a10, a01, a11);
}
__device__ __forceinline__ float bilinear_xy(float X, float Y, const float4 &c)    // interpolator.rs:83
{
    float r = __fadd_rn(c.x, __fmul_rn(c.y, X));
    r = __fadd_rn(r, __fmul_rn(c.z, Y));
    return __fadd_rn(r, __fmul_rn(__fmul_rn(c.w, X), Y));
}
// interpolator.rs:46-50: a target coincident with a corner returns that corner's value, tested in
// the order a, b, c, d (the later selects below take precedence).  Rare; the raw corner values
// are re-read from the f64 node grid.
__device__ __forceinline__ float corner_pick(float r, bool at_xa, bool at_xb, bool at_ya, bool at_yb,
                                             const double *node, int nx)
{
    const float zsw = (float)__ldg(node), zse = (float)__ldg(node + 1);
    const float znw = (float)__ldg(node + nx), zne = (float)__ldg(node + nx + 1);
    r = (at_xb && at_ya) ? zse : r;
    r = (at_xb && at_yb) ? zne : r;
    r = (at_xa && at_yb) ? znw : r;
    r = (at_xa && at_ya) ? zsw : r;
    return r;
}

// IEEE f32 quotient t/s for a launch-constant divisor s with r = RN(1/s) (computed on the
// host): q0 = RN(t r) is within 2 ulp; one exact-residual step makes it faithful, a second
// makes it the correctly rounded quotient (Markstein 1990; the host falls back to the true
// divide if s has an all-ones significand, the theorem's exception).  Verified exhaustively
// against __fdiv_rn over every float for a set of spacings by tests/test_gpu_api.py
// (mr_selftest_fdiv): bit-identical for every |t| >= 2^-100; below that the exact residual
// underflows and the last bits can differ, but both quotients are in [0, 1) (s > 1e-30), i.e.
// cell 0 and in bounds either way.  +-inf gives NaN instead of +-inf; both are out of bounds.
__device__ __forceinline__ float fdiv_const(float t, float s, float r)
{
    float q = __fmul_rn(t, r);
    q = __fmaf_rn(__fmaf_rn(-q, s, t), r, q);
    q = __fmaf_rn(__fmaf_rn(-q, s, t), r, q);
    return q;
}

// Read-only vector loads as volatile asm: the compiler keeps them where they are written
// (it otherwise sinks them into the conditional region of their first use, which exposes
// one full L2 round trip per load instead of one per RHS).
__device__ __forceinline__ float4 ldg_f4(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_f2(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldg_d2(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// One 32-byte record in ONE instruction (sm_100 LDG.256): a float4 followed by a double2, or
// two float4, or two double2.  What arrives together cannot be split into two round trips.
__device__ __forceinline__ void ldg_f4_d2(const float4 *p, float4 &a, double2 &b)
{
    unsigned long long q0, q1;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q0), "=l"(q1), "=d"(b.x), "=d"(b.y) : "l"(p));
    a.x = __uint_as_float((unsigned)q0); a.y = __uint_as_float((unsigned)(q0 >> 32));
    a.z = __uint_as_float((unsigned)q1); a.w = __uint_as_float((unsigned)(q1 >> 32));
}
// the same load, skipped by the lanes that do not need the record (a and b keep whatever their registers held)
__device__ __forceinline__ void ldg_f4_d2_unless(bool skip, const float4 *p, float4 &a, double2 &b)
{
    unsigned long long q0, q1;
    asm volatile("{ .reg .pred q; setp.eq.s32 q, %5, 0; @q ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4]; }"
                 : "=l"(q0), "=l"(q1), "=d"(b.x), "=d"(b.y) : "l"(p), "r"((int)skip));
    a.x = __uint_as_float((unsigned)q0); a.y = __uint_as_float((unsigned)(q0 >> 32));
    a.z = __uint_as_float((unsigned)q1); a.w = __uint_as_float((unsigned)(q1 >> 32));
}
__device__ __forceinline__ void ldg_f4_f4(const float4 *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void ldg_d2_d2(const double2 *p, double2 &a, double2 &b)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}

// The cell rule of four_corners (cartesian_netcdf3.rs:344-387, cartesian_current.rs
// :293-336) for an index already known to be in [0, n-1]: left edge -> (0,1), right
// edge -> (n-2,n-1), on a grid line -> (i,i+1), else (floor,ceil).  All four cases
// are i1 = min(floor(index), n-2), i2 = i1+1 (n >= 2 is enforced at upload).
// Out-of-range / NaN indices are clamped so the loads stay in bounds; the caller
// discards the result.
__device__ __forceinline__ int cell_of(float index, int n)
{
    int i = __float2int_rd(index);           // NaN -> 0, saturating
    return max(min(i, n - 2), 0);
}
__device__ __forceinline__ int cell_of(double index, int n)
{
    int i = __double2int_rd(index);
    return max(min(i, n - 2), 0);
}

// =============================================================================
// BathymetryData::depth_and_gradient
// =============================================================================
// Returns false for Err (the whole RHS then becomes NaN, wave_ray_path.rs:222-228).

__device__ __forceinline__ bool bathy_analytic(int kind, const BathyDev &b, float x, float y,
                                               float &h, float &gx, float &gy)
{
    if (kind == MR_BATHY_CONSTANT) {          // constant_depth.rs:39-45
        bool bad = isnan(x) || isnan(y);
        h = bad ? qnanf() : b.h0;
        gx = gy = bad ? qnanf() : 0.0f;
        return true;
    }
    if (kind == MR_BATHY_SLOPE) {             // constant_slope.rs:67-76
        bool bad = isnan(x) || isnan(y);
        float s = __fadd_rn(b.h0, __fmul_rn(b.dhdx, __fsub_rn(x, b.x0)));
        s = __fadd_rn(s, __fmul_rn(b.dhdy, __fsub_rn(y, b.y0)));
        h = bad ? qnanf() : s;
        gx = bad ? qnanf() : b.dhdx;
        gy = bad ? qnanf() : b.dhdy;
        return true;
    }
    // ARRAY: array_depth.rs:27-35 (`as usize` saturates)
    // negative -> 0 and huge -> max like Rust's cast; NaN -> 0 has to be spelled out (cvt gives 2^63)
    unsigned long long xi = isnan(x) ? 0ull : __float2ull_rz(x), yi = isnan(y) ? 0ull : __float2ull_rz(y);
    unsigned long long len = (unsigned long long)b.nx;
    bool oob = xi >= len || yi >= len;
    h = oob ? qnanf() : __ldg(b.array + (oob ? 0 : xi * (unsigned long long)b.ny + yi));
    gx = gy = oob ? qnanf() : 0.0f;
    return true;
}

// GRID, strict: cartesian_netcdf3.rs:98-135 on the f64 node grid
__device__ __forceinline__ bool bathy_grid_strict(const BathyDev &b, float x, float y,
                                                  float &h, float &gx, float &gy)
{
    if (isnan(x) || isnan(y)) { h = gx = gy = qnanf(); return true; }          // :101-103
    float ix = __fdiv_rn(__fsub_rn(x, b.xf0), b.sx);                          // :289
    float iy = __fdiv_rn(__fsub_rn(y, b.yf0), b.sy);
    // :291  index < 0 || index > n-1  (a NaN index, from inf-inf, passes in the
    // reference and then dies on det == 0; both end in Err)
    if (!(ix >= 0.0f && ix <= (float)(b.nx - 1))) return false;
    if (!(iy >= 0.0f && iy <= (float)(b.ny - 1))) return false;
    int x1 = cell_of(ix, b.nx), y1 = cell_of(iy, b.ny);
    const double *row0 = b.depth + (size_t)b.nx * y1 + x1;
    const double *row1 = row0 + b.nx;
    double dsw = __ldg(row0), dse = __ldg(row0 + 1);
    double dnw = __ldg(row1), dne = __ldg(row1 + 1);
    float xa = __ldg(b.x + x1), xb = __ldg(b.x + x1 + 1);
    float ya = __ldg(b.y + y1), yb = __ldg(b.y + y1 + 1);
    if (!bilinear_cell_strict(xa, xb, ya, yb, (float)dsw, (float)dnw, (float)dne, (float)dse, x, y, h))
        return false;
    gx = (float)__ddiv_rn(__dsub_rn(dse, dsw), b.x_space);                    // :126-134
    gy = (float)__ddiv_rn(__dsub_rn(dnw, dsw), b.y_space);
    return true;
}

// =============================================================================
// CurrentData::current_and_gradient
// =============================================================================
struct CurrentVal { double u, v, dudx, dudy, dvdx, dvdy; };

// GRID, strict: cartesian_current.rs:487-542 on the f64 node grids
__device__ __forceinline__ bool current_grid_strict(const CurrentDev &c, double x, double y, CurrentVal &o)
{
    double ix = __ddiv_rn(x - c.xd0, c.sx);                               // :246
    double iy = __ddiv_rn(y - c.yd0, c.sy);
    // :248  a NaN index passes this test in the reference; floor/ceil of NaN cast to 0
    // make x1 == x2, det == 0, Err (interpolator.rs:65).  Both ways: Err.
    if (!(ix >= 0.0 && ix <= (double)(c.nx - 1))) return false;
    if (!(iy >= 0.0 && iy <= (double)(c.ny - 1))) return false;
    int x1 = cell_of(ix, c.nx), y1 = cell_of(iy, c.ny);
    size_t o0 = (size_t)c.nx * y1 + x1;
    double usw = __ldg(c.u + o0), use_ = __ldg(c.u + o0 + 1);
    double unw = __ldg(c.u + o0 + c.nx), une = __ldg(c.u + o0 + c.nx + 1);
    double vsw = __ldg(c.v + o0), vse = __ldg(c.v + o0 + 1);
    double vnw = __ldg(c.v + o0 + c.nx), vne = __ldg(c.v + o0 + c.nx + 1);
    float xa = (float)__ldg(c.x + x1), xb = (float)__ldg(c.x + x1 + 1);   // :375-376
    float ya = (float)__ldg(c.y + y1), yb = (float)__ldg(c.y + y1 + 1);
    float xf = (float)x, yf = (float)y;                                   // :500
    float uf, vf;
    if (!bilinear_cell_strict(xa, xb, ya, yb, (float)usw, (float)unw, (float)une, (float)use_, xf, yf, uf)) return false;
    if (!bilinear_cell_strict(xa, xb, ya, yb, (float)vsw, (float)vnw, (float)vne, (float)vse, xf, yf, vf)) return false;
    o.u = (double)uf; o.v = (double)vf;
    o.dudx = __ddiv_rn(__dsub_rn(use_, usw), c.x_space);                  // :522-536
    o.dudy = __ddiv_rn(__dsub_rn(unw, usw), c.y_space);
    o.dvdx = __ddiv_rn(__dsub_rn(vse, vsw), c.x_space);
    o.dvdy = __ddiv_rn(__dsub_rn(vnw, vsw), c.y_space);
    return true;
}

// =============================================================================
// the f64 stage
// =============================================================================

// ---- reference expression tree (MR_MATH_STRICT) --------------------------------
// wave_ray_path.rs:132-147 with group_velocity :177-188 and dkdt_bathy :207-216.
// The strict translation unit is compiled with -fmad=false; the explicit _rn
// intrinsics make the intent visible as well.
__device__ __forceinline__ void rhs_f64_strict(double kx, double ky, double h, double dhdx, double dhdy,
                                               const CurrentVal &cv, double out[4])
{
    double k = sqrt(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)));      // :132
    double theta = atan2(ky, kx);                                          // :133
    double cg;
    if (h <= 0.0) {
        cg = qnan();                                                       // :178-180
    } else if (k <= 0.0) {
        out[0] = out[1] = out[2] = out[3] = qnan();                        // :181-183 Err
        return;
    } else {
        double kh = __dmul_rn(k, h);
        double ch = cosh(kh);
        double th = tanh(kh);
        double num = __dadd_rn(th, __ddiv_rn(kh, __dmul_rn(ch, ch)));
        double den = sqrt(__dmul_rn(__dmul_rn(k, kG), th));
        cg = __dmul_rn(kG / 2.0, __ddiv_rn(num, den));                     // :184-186
    }
    double sn, cs;
    sincos(theta, &sn, &cs);
    out[0] = __dadd_rn(__dmul_rn(cg, cs), cv.u);                           // :137
    out[1] = __dadd_rn(__dmul_rn(cg, sn), cv.v);                           // :138
    // :208-213  (-0.5)*k*1.0/sinh(kh)*1.0/cosh(kh)*sqrt(G*k*tanh(kh))*dh
    double kh = __dmul_rn(k, h);
    double a = __dmul_rn(-0.5, k);
    a = __ddiv_rn(a, sinh(kh));
    a = __ddiv_rn(a, cosh(kh));
    a = __dmul_rn(a, sqrt(__dmul_rn(__dmul_rn(kG, k), tanh(kh))));
    double bx = __dmul_rn(a, dhdx), by = __dmul_rn(a, dhdy);
    out[2] = __dsub_rn(__dsub_rn(bx, __dmul_rn(kx, cv.dudx)), __dmul_rn(ky, cv.dvdx));   // :146
    out[3] = __dsub_rn(__dsub_rn(by, __dmul_rn(kx, cv.dudy)), __dmul_rn(ky, cv.dvdy));   // :147
}

// ---- lean f64 primitives (MR_MATH_FAST) ------------------------------------------
// For normal, positive, finite arguments (what the path produces when `ok`); other
// inputs give NaN/garbage that the caller replaces.

// sqrt(x) and 1/sqrt(x) together: MUFU.RSQ64H seed y0 (within ~2^-20) and ONE fourth-order step.  With
// e = 1 - x y0^2 the exact value is y0 (1 - e)^(-1/2) = y0 (1 + e/2 + 3 e^2/8 + 5 e^3/16 + ...); the first omitted
// term is 0.27 e^4 < 2^-80.  1/sqrt within 1 ulp, sqrt = x / sqrt(x) within 1.5 ulp.  Eight operations, six deep
// (two coupled Goldschmidt steps were ten and seven).
template <bool LEAN = false>
__device__ __forceinline__ void sqrt_rsqrt(double x, double &s, double &rs)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  if (LEAN) {
    const double e = fma(-x, y0 * y0, 1.0);
    const double p = fma(e, fma(e, 0.3125, 0.375), 0.5);
    rs = fma(y0 * e, p, y0);
    s = x * rs;
  } else {
    double g = x * y0, hh = 0.5 * y0;              // two coupled Goldschmidt steps
    double r = fma(-g, hh, 0.5);
    g = fma(g, r, g); hh = fma(hh, r, hh);
    r = fma(-g, hh, 0.5);
    g = fma(g, r, g); hh = fma(hh, r, hh);
    s = g; rs = hh + hh;
  }
}

// 1/x: MUFU.RCP64H seed r0 and one third-order step: with e = 1 - x r0, 1/x = r0 (1 + e + e^2 + ...), e^3 < 2^-60
__device__ __forceinline__ double recip(double x)
{
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
#if MR_LEAN_RCP
    const double e = fma(-x, r0, 1.0);
    return fma(fma(e, e, e), r0, r0);
#else
    double e = fma(-x, r0, 1.0);                   // two Newton steps
    r0 = fma(r0, e, r0);
    e = fma(-x, r0, 1.0);
    return fma(r0, e, r0);
#endif
}

// x^(-1/4) for a normal, positive, finite x: three MUFU seeds (x^-1/2, its own inverse square root x^1/4, the
// reciprocal of that: within ~2^-18 together) and ONE fourth-order step.  With e = 1 - x t^4 the exact root is
// t (1 - e)^(-1/4) = t (1 + e/4 + 5 e^2/32 + 15 e^3/128 + ...); the first omitted term is 0.1 e^4 < 2^-67, and the
// rounding of e (an fma on t^4, itself two roundings) enters scaled by t/4: the result is within 1 ulp (checked
// against 80-bit arithmetic for seeds off by up to 2^-17).  Seven operations, six deep, against ten and eight for
// two Newton steps.  x = +inf gives NaN (0 * inf in the residual), x = 0 gives NaN (inf * 0), NaN gives NaN.
__device__ __forceinline__ double inv_root4(double x)
{
    double y, s, t;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(y));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(s));
    const double t2 = t * t;
    const double e = fma(-x, t2 * t2, 1.0);
    const double p = fma(e, fma(e, 0.1171875, 0.15625), 0.25);
    return fma(t * e, p, t);
}

// E = exp(z) and em = expm1(z) for z in [-700, 0] (the callers keep z there).
// z = n ln2 + r, |r| <= ln2/2;  expm1(r) = r + r^2 P(r) (Taylor through r^13, remainder
// < 2e-17 relative);  E = 2^n (1 + p),  em = 2^n p + (2^n - 1)  (2^n - 1 is exact).
__device__ __forceinline__ void exp_expm1_neg(double z, double &E, double &em)
{
    const double MAGIC = kExpRed[1];
    double t = fma(z, kExpRed[0], MAGIC);          // low word of t = rint(z*log2e) as int32
    int n = __double2loint(t);
    double nd = t - MAGIC;
    double r = fma(-nd, kExpRed[2], z);
    r = fma(-nd, kExpRed[3], r);
#if MR_EXP_2CHAIN
    // P(r) = c0 r^11 + ... + c11 as two Horner chains in r^2 (odd and even coefficients), half as deep as one
    const double r2 = r * r;
    double po = kExpm1C[0], pe = kExpm1C[1];
#pragma unroll
    for (int i = 2; i < 12; i += 2) { po = fma(po, r2, kExpm1C[i]); pe = fma(pe, r2, kExpm1C[i + 1]); }
    double p = fma(po, r, pe);
    p = fma(r2, p, r);                             // expm1(r)
#else
    double p = kExpm1C[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) p = fma(p, r, kExpm1C[i]);
    p = fma(r * r, p, r);                          // expm1(r)
#endif
    double s = __hiloint2double((n + 1023) << 20, 0);   // 2^n, n in [-1010, 0]
    E = fma(s, p, s);
    em = fma(s, p, s - 1.0);
}

// ---- restructured f64 stage (MR_MATH_FAST) ------------------------------------------
// Same functions of (k, h) as the reference, evaluated from ONE exponential:
//   E = exp(-2kh), m = 1-E = -expm1(-2kh), w = 1+E
//   tanh kh = m/w,  1/cosh^2 kh = 4E/w^2,  1/(sinh kh cosh kh) = 4E/(m w)
// and cos(theta) = kx/k, sin(theta) = ky/k instead of atan2 + sincos.
// Every case in which the reference's RHS is four NaNs (field lookup Err, h <= 0, h NaN,
// k == 0, k NaN) reaches this function as h = NaN and leaves it as four NaNs by plain NaN
// propagation, so there is no per-output select.  Large kh: E -> 0, tanh = 1, second cg term
// 0, bathymetric term -0 (the reference gets the same from cosh^2 -> inf and sinh -> inf).
// cg and the factor (bx, by) = Bc grad(h) of dk/dt, from (k, h, grad h).
//
// The general expressions.  UNDER: kh may be anything >= 22, so exp(-2kh) may underflow — it is then taken as
// exactly 0 (E = 0, expm1 = -1), which makes tanh = 1, the second cg term kh * 0 (NaN for an infinite kh, like the
// reference's inf/inf) and the bathymetric factor -0 (NaN for an infinite k), the reference's own limits.
template <bool UNDER>
__device__ __forceinline__ void wave_terms_general(double k, double kh, double dhdx, double dhdy,
                                                   double &cg, double &bx, double &by)
{
    // tanh kh = T, kh/cosh^2 kh = hs2, 1/(sinh kh cosh kh) = csch_sech
    double z = -2.0 * kh, E, em;
    const bool under = UNDER && z < -700.0;
    if (UNDER) z = under ? -700.0 : z;             // (a NaN stays a NaN)
    exp_expm1_neg(z, E, em);
    if (UNDER) { E = under ? 0.0 : E; em = under ? -1.0 : em; }
    const double m = -em, w = 2.0 + em;
    const double r = recip(m * w);
    const double invw = m * r;
    const double T = m * invw;
    const double E4 = 4.0 * E;
    const double hs2 = kh * ((E4 * invw) * invw);
    const double csch_sech = E4 * r;
    const double q = (k * kExpRed[5]) * T;
    double sq, rq;
    sqrt_rsqrt<(MR_LEAN_SQRT & 1) != 0>(q, sq, rq);
    cg = kExpRed[6] * ((T + hs2) * rq);
    const double Bc = ((-0.5 * k) * csch_sech) * sq;
    bx = Bc * dhdx; by = Bc * dhdy;
}

// Above this wavenumber the deep-water shortcut is not taken: what it drops from dk/dt, relative to k and per
// second, is 2 exp(-2kh) sqrt(G k) |grad h| — below 2e-18 |grad h| for kh >= 22 and k <= 16 rad/m (waves longer than
// 0.4 m, everything the model is about), but not for the k -> 1e15 of a ray creeping up to a shoreline node a few
// femtometres deep (C5 has them): there sqrt(G k) ~ 1e8 per second and the dropped term is 1e-13 per step.
static constexpr double kDeepMaxK = 16.0;

__device__ __forceinline__ void wave_terms(double k, double h, double dhdx, double dhdy,
                                           double &cg, double &bx, double &by)
{
    const double kh = k * h;
    if (!(kh >= kExpRed[4])) {
        wave_terms_general<false>(k, kh, dhdx, dhdy, cg, bx, by);
    } else if (k > kDeepMaxK) {
        wave_terms_general<true>(k, kh, dhdx, dhdy, cg, bx, by);
    } else {
        // Deep water, kh >= 22 (and k <= 16): exp(-2kh) < 8e-20, so in f64 tanh kh == 1 exactly, kh/cosh^2 kh < 2^-57
        // vanishes against it, and the bathymetric term changes k by less than 2e-18 |grad h| of itself per second —
        // below half an ulp per step, i.e. the reference's own sum rounds it away.  With T = 1 and the other two 0
        // the general expressions reduce, value for value, to cg = (G/2) rsqrt(G k) and Bc = -0.  The zeros are
        // computed, not written: kh * 0 is NaN when h is +inf, where the reference's kh/sinh(2kh) is
        // inf/inf and cg NaN, while its bathymetric term stays -0 * grad(h) for an infinite h (k * 0 is 0 then)
        // and propagates a non-finite gradient.
        const double z = kh * 0.0, zk = k * 0.0;
        double sq, rq;
        sqrt_rsqrt<(MR_LEAN_SQRT & 2) != 0>(k * kExpRed[5], sq, rq);
        cg = fma(kExpRed[6], rq, z);
        bx = -zk * dhdx; by = -zk * dhdy;
    }
}
// =============================================================================
// the fast RHS, in four phases so that every load of an evaluation is in flight
// before anything waits on one:
//   1. fractional indices and cell addresses of both fields
//   2. all record loads
//   3. the wavenumber-only f64 work (k, 1/k, direction cosines) under the loads
//   4. f32 bilinears, then the f64 stage
// A failed lookup (out of bounds, NaN index, degenerate cell) is not branched on: the cell
// index is clamped so the loads stay in bounds, and the depth handed to the f64 stage is
// replaced by NaN, which makes all four outputs NaN exactly as wave_ray_path.rs:222-228 does.
// A NaN position needs no special case either: the reference's Ok(NaN depth) also ends in four
// NaNs (cg and the bathymetric term are NaN), and a NaN index fails the bounds test here.
// =============================================================================
// One ray's evaluation in flight.  The four phases are separate functions so that a thread that
// carries NR rays runs each phase for all of them before the next (rhs_fast_n): the compiler then
// has NR independent instruction streams to interleave, and the uniform work of a phase (constant
// loads, loop control) is paid once per thread instead of once per ray.
// DMAP (affine gridded bathymetry only): the depth lookup is skipped where it cannot matter.  For kh >= 22 the
// f64 stage takes its deep-water branch, whose outputs do not depend on h or grad(h) at all (rhs_f64_fast).  A
// small map gives, per block of 8 x 8 cells, the square of a lower bound H of every depth the lookup could
// return there; k^2 H^2 >= 484 (with a margin that dwarfs the f32 roundings, 2e-5 relative) proves kh >= 22
// without the cell record: no 32-byte sector fetched (the map's floats are shared by whole blocks of lanes),
// no bilinear, no corner test.  Lanes that fail the test — shallow water, a dry or non-finite node in the
// block, a failed lookup — load the record and proceed exactly as without the map.
// SG (both fields gridded, affine, on the SAME grid; BathyDev::same_grid): one f32 fractional index serves both
// fields.  The reference computes two — f32 for the bathymetry (cartesian_netcdf3.rs:289), f64 for the current
// (cartesian_current.rs:246) — and they can name different cells only when the position is within a rounding
// error of a grid line: |ix32 - I| <= 2^-24 (3 n + |x0|/s) and |ix64 - I| <= 2^-51 n around the exact index I.
// So when the f32 index lies further than sg_delta (that bound plus slack) from every integer, both floors are
// equal and both bounds tests pass, and neither the f64 index nor any bounds test is evaluated; the cell
// geometry (corner coordinates, in-cell fractions, the corner-coincidence test of interpolator.rs:46-50) is then
// shared too.  Within sg_delta of a grid line (< 0.2 % of evaluations per axis on a 2048-point axis) the lane takes
// the two separate lookups exactly as without SG.  Every value is the one the separate lookups produce.
// CMAP (affine gridded current): zero-current files (the API always takes a current file) and piecewise-constant
// currents are common; where a block of the current grid is uniform the lookup degenerates — the reference's
// bilinear of four equal corners is that value, its finite differences are 0 — so a lane there loads 8 bytes of a
// small map (shared by all lanes of the block) instead of its 64-byte cell record and skips the bilinear and the
// corner test; the bounds test of the f64 index stays.  Lanes in other blocks load their record as without the map.
template <int BK, int CK, bool UNI, bool DMAP = false, bool SG = false, bool CMAP = false>
struct FastRay {
    static constexpr bool kDmap = DMAP && UNI && BK == MR_BATHY_GRID;
    static constexpr bool kCmap = CMAP && UNI && CK == MR_CURRENT_GRID && !SG;
    static constexpr bool kSame = SG && UNI && BK == MR_BATHY_GRID && CK == MR_CURRENT_GRID;
    // advection sums formed ahead of the bilinears: in the variants that need the registers (measured: the plain
    // kernel is 8 % slower with it, 63.9 -> 69.2 ms on C4)
    static constexpr bool kEarly = !(CMAP && !SG) && CK == MR_CURRENT_GRID && ((MR_EARLY_GRAD & 1) && !kDmap && !kSame || (MR_EARLY_GRAD & 2) && kDmap && !kSame || (MR_EARLY_GRAD & 4) && kSame);
    static constexpr bool kShare = kSame && !kDmap && MR_SG_SHARE_GEOM;      // one cell geometry for both lookups (with the map the depth lookup is the rare path)
    float xf, yf;
    bool ok;
    bool deep;
    bool cuni;                 // kCmap: this lane's block of the current grid is uniform
    float2 cm;                 // kCmap: the block's {u, v}, or NaNs
    float hsq;
    int bx1, by1, cx1, cy1;
    const float4 *brec;
    unsigned ccell;
    // (only the kinds selected by the template parameters touch these; no default initialisation)
    float4 Z, U, V;
    double2 gh, gu, gv;
    float bxa, bxb, bya, byb, cxa, cxb, cya, cyb;
    double k2, k, cs, sn;      // with the depth-floor map k holds t = k^-1/2 from phase 3 until phase 4
    double ax, ay;             // kEarly: -kx du/dx - ky dv/dx, -kx du/dy - ky dv/dy

    // f64 fractional index of the current and its cell (cartesian_current.rs:246-252).  The spacing is a launch
    // constant: q0 = t*RN(1/s), then q0 + (t - q0 s) RN(1/s) in one fma.  The argument of that last rounding is
    // t/s to within 2^-52 ulp (the residual is exact, only 1/s carries an error), so the result IS
    // RN(t/s) unless t/s lies within 2^-52 ulp of a rounding midpoint — and only the cell, floor(index),
    // is used: it can differ from the reference's only if that midpoint also neighbours an integer
    // (~1e-28 per evaluation).  Exact whenever t/s is representable (a ray sitting on a grid line).  An
    // infinite position turns into NaN here and fails the bounds test like the infinity.
    __device__ __forceinline__ bool current_index(const CurrentDev &c, double x, double y)
    {
        const double tx = x - c.xd0, ty = y - c.yd0;
        const double qx = tx * c.inv_sx, qy = ty * c.inv_sy;
        const double ix = fma(fma(-qx, c.sx, tx), c.inv_sx, qx);
        const double iy = fma(fma(-qy, c.sy, ty), c.inv_sy, qy);
        cx1 = cell_of(ix, c.nx); cy1 = cell_of(iy, c.ny);
        return ix >= 0.0 && ix <= c.nxm1d && iy >= 0.0 && iy <= c.nym1d;          // :248
    }

    // ---- phase 1: fractional indices and cell addresses ------------------------------------------
    // kSame: returns false — with nothing else decided — when the f32 index is too close to a grid line for one
    // cell to serve both fields; the caller then evaluates this point with the separate lookups (rhs_fast_n).
    // Otherwise (and always without kSame) returns true.
    __device__ __forceinline__ bool phase1(const BathyDev &b, const CurrentDev &c, double x, double y, double kx, double ky)
    {
        xf = (float)x; yf = (float)y;                                              // wave_ray_path.rs:122
        ok = true;
        deep = false;
        k2 = fma(kx, kx, ky * ky);
        if (BK == MR_BATHY_GRID) {
            float ix, iy;                                                          // cartesian_netcdf3.rs:289
            if (UNI) {
                // fdiv_const for x and y at once: t = p - p0; q = t*r; twice q += (t - q*s)*r
                const f32x2 t = sub2(pk(xf, yf), b.p0);
                f32x2 q = mul2(t, b.rs2);
                q = fma2(fma2(q, b.ns2, t), b.rs2, q);
                q = fma2(fma2(q, b.ns2, t), b.rs2, q);
                ix = lo_of(q); iy = hi_of(q);
            } else if (b.fastdiv) {     // the index divides by the constant |x[1]-x[0]| whatever the other spacings are
                ix = fdiv_const(__fsub_rn(xf, b.xf0), b.sx, b.rsx);
                iy = fdiv_const(__fsub_rn(yf, b.yf0), b.sy, b.rsy);
            } else {
                ix = __fdiv_rn(__fsub_rn(xf, b.xf0), b.sx);
                iy = __fdiv_rn(__fsub_rn(yf, b.yf0), b.sy);
            }
            bx1 = cell_of(ix, b.nx); by1 = cell_of(iy, b.ny);
            if (kSame) {
                // distance of the index from the middle of its (clamped) cell: below sg_lim = 0.5 - sg_delta on both
                // axes, the point is strictly inside the cell by more than the two indices can disagree — in bounds
                // for both fields, same cell.  (A clamped cell puts an out-of-range index at 0.5 or more; NaN fails.)
                const f32x2 mid = add2(pk((float)bx1, (float)by1), pk(0.5f, 0.5f));       // exact: cell numbers < 2^23
                const f32x2 off = sub2(pk(ix, iy), mid);
                if (!(fabsf(lo_of(off)) < b.sg_lim && fabsf(hi_of(off)) < b.sg_lim)) return false;
            } else {
                ok = ix >= 0.0f && ix <= b.nxm1f && iy >= 0.0f && iy <= b.nym1f;  // :291
            }
            brec = b.cell + 2ull * (unsigned)((b.nx - 1) * by1 + bx1);
            if (kDmap)
                asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(hsq)
                             : "l"(b.dmap + (unsigned)((by1 >> kDeepShift) * b.dmap_nbx + (bx1 >> kDeepShift))));
        }
        if (CK == MR_CURRENT_GRID) {
            if (kSame) {
                cx1 = bx1; cy1 = by1;
            } else {
                const bool okc = current_index(c, x, y);
                ok = ok && okc;
            }
            ccell = (unsigned)((c.nx - 1) * cy1 + cx1);
            if (kCmap)
                asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(cm.x), "=f"(cm.y)
                             : "l"(c.cmap + (unsigned)((cy1 >> kDeepShift) * c.cmap_nbx + (cx1 >> kDeepShift))));
        }
        return true;
    }

    // With the depth-floor map nothing of the bathymetry's cell is carried through the deep lanes' path: a lane that
    // does need the depth derives its cell again from (xf, yf) — the operations of phase 1
    __device__ __forceinline__ void bathy_cell_again(const BathyDev &b)
    {
        const f32x2 t = sub2(pk(xf, yf), b.p0);
        f32x2 q = mul2(t, b.rs2);
        q = fma2(fma2(q, b.ns2, t), b.rs2, q);
        q = fma2(fma2(q, b.ns2, t), b.rs2, q);
        bx1 = cell_of(lo_of(q), b.nx); by1 = cell_of(hi_of(q), b.ny);
    }

    // ---- phase 2: all record loads -----------------------------------------------------------------
    __device__ __forceinline__ void phase2(const BathyDev &b, const CurrentDev &c)
    {
        if (BK == MR_BATHY_GRID) {
            if (kDmap) {
                // (a NaN or infinite k^2 fails or passes harmlessly: NaN compares false; k^2 = inf makes k NaN,
                // and the deep-water branch turns that into four NaNs like the general one)
                if (!(MR_DEEP_ROOT4 && MR_DMAP_LATE_LOAD)) deep = ok && __fmul_rn((float)k2, hsq) >= 484.01f && (float)k2 <= (float)(kDeepMaxK * kDeepMaxK);
                if (!MR_DMAP_LATE_LOAD) ldg_f4_d2_unless(deep, brec, Z, gh);
            } else {
                ldg_f4_d2(brec, Z, gh);
            }
            if (!UNI) {
                bxa = __ldg(b.x + bx1); bxb = __ldg(b.x + bx1 + 1);
                bya = __ldg(b.y + by1); byb = __ldg(b.y + by1 + 1);
            }
        }
        if (CK == MR_CURRENT_GRID && kCmap) {
            // the lanes outside uniform blocks fetch their record when they get to the bilinear (current_part):
            // one more exposed round trip for them, sixteen registers fewer held by everybody
        } else if (CK == MR_CURRENT_GRID) {
            ldg_f4_f4(c.cell + 4ull * ccell, U, V);
            ldg_d2_d2((const double2 *)(c.cell + 4ull * ccell + 2), gu, gv);
            if (!UNI) {
                cxa = __ldg(c.xf + cx1); cxb = __ldg(c.xf + cx1 + 1);
                cya = __ldg(c.yf + cy1); cyb = __ldg(c.yf + cy1 + 1);
            }
        }
    }

    // ---- phase 3: wavenumber-only f64 work, under the loads --------------------------------------
    // With the depth-floor map everything is derived from ONE fourth root, t = (k^2)^-1/4 = k^-1/2, whichever way
    // the lane goes afterwards — so nothing here waits for the map's answer, whose load is still in flight:
    //   proven deep:  cg cos(theta) = (G/2) / sqrt(G k) * kx / k = (sqrt(G)/2) t^3 kx, likewise sin; k is never formed;
    //   otherwise:    1/k = t^2, k = k^2 t^2, and the general wave terms as without the map.
    __device__ __forceinline__ void phase3(double kx, double ky)
    {
        if (MR_DEEP_ROOT4 && kDmap) {
            k = inv_root4(k2);             // t, until phase 4
            // the map's answer is consumed only now, behind the fourth root's chain (a warp issues in order: the
            // first instruction that needs the loaded value is where it waits)
            if (MR_DMAP_LATE_LOAD) {
                const float k2f = (float)k2;           // (k <= kDeepMaxK as well: see wave_terms)
                deep = ok && __fmul_rn(k2f, hsq) >= 484.01f && k2f <= (float)(kDeepMaxK * kDeepMaxK);
            }
            return;
        }
        double rk;
        sqrt_rsqrt<(MR_LEAN_SQRT & 4) != 0>(k2, k, rk);
        cs = kx * rk; sn = ky * rk;
    }

    // ---- phase 4: f32 bilinears, then the f64 stage --------------------------------------------------
    // corner coordinates of a cell given as floats, and the in-cell fractions of p (UNI):
    // (xa,ya) = i*d + p0, (xb,yb) = (xa,ya) + d, (Y, X) = (c10, c01) * ((x,y) - (xa,ya)), two components per instruction
    struct Geom { float xa, xb, ya, yb, X, Y; };
    __device__ __forceinline__ static Geom geom_of(int x1, int y1, f32x2 p, f32x2 d2, f32x2 p0, f32x2 c2)
    {
        const f32x2 pa = fma2(pk((float)x1, (float)y1), d2, p0), pb = add2(pa, d2);
        const f32x2 yx = mul2(c2, sub2(p, pa));
        Geom g;
        g.xa = lo_of(pa); g.ya = hi_of(pa); g.xb = lo_of(pb); g.yb = hi_of(pb);
        g.Y = lo_of(yx); g.X = hi_of(yx);
        return g;
    }

    // depth and its gradient at (xf, yf) from the loaded record / the analytic kinds; `shared`: the affine cell
    // geometry already resolved for this very cell (kSame), else NULL
    __device__ __forceinline__ void bathy_part(const BathyDev &b, f32x2 p, float &h32, double &dhdx, double &dhdy,
                                               const Geom *shared = nullptr)
    {
        if (BK == MR_BATHY_GRID) {
            float X, Y;
            if (UNI) {
                const Geom g = shared ? *shared : geom_of(bx1, by1, p, b.d2, b.p0, b.c2);
                bxa = g.xa; bya = g.ya; bxb = g.xb; byb = g.yb;
                Y = g.Y; X = g.X;
            } else {
                const float dx = __fsub_rn(bxb, bxa), dy = __fsub_rn(byb, bya);
                const float det = __fsub_rn(0.0f, __fmul_rn(dx, dy));              // interpolator.rs:64
                ok = ok && det != 0.0f;                                            // :65-67
                const float c01 = -__fdiv_rn(dx, det), c10 = -__fdiv_rn(dy, det);  // :70-71
                X = __fmul_rn(c01, __fsub_rn(yf, bya)); Y = __fmul_rn(c10, __fsub_rn(xf, bxa));
            }
            h32 = bilinear_xy(X, Y, Z);
            if (xf == bxa || xf == bxb) {
                const bool at_ya = yf == bya, at_yb = yf == byb;
                if (at_ya || at_yb)
                    h32 = corner_pick(h32, xf == bxa, xf == bxb, at_ya, at_yb, b.depth + (size_t)b.nx * by1 + bx1, b.nx);
            }
            dhdx = gh.x; dhdy = gh.y;
        } else {
            float gx32, gy32;
            bathy_analytic(BK, b, xf, yf, h32, gx32, gy32);
            dhdx = (double)gx32; dhdy = (double)gy32;
        }
    }

    // current and its gradients at (xf, yf); *keep (kSame) receives the cell geometry for the bathymetry to reuse
    __device__ __forceinline__ void current_part(const CurrentDev &c, f32x2 p, CurrentVal &cv, Geom *keep = nullptr)
    {
        if (kCmap) cuni = cm.x == cm.x;             // not the NaN marker
        if (kCmap && cuni) {
            // four equal corners: interpolator.rs:78-83 gives a10 = a01 = a11 = 0 and returns the corner value (X and
            // Y are finite inside the grid, and a coinciding corner returns the same value); the finite differences
            // of cartesian_current.rs:522-536 are exactly 0
            cv.u = (double)cm.x; cv.v = (double)cm.y;
            cv.dudx = cv.dudy = cv.dvdx = cv.dvdy = 0.0;
            return;
        }
        if (kCmap) {
            ldg_f4_f4(c.cell + 4ull * ccell, U, V);
            ldg_d2_d2((const double2 *)(c.cell + 4ull * ccell + 2), gu, gv);
        }
        if (CK == MR_CURRENT_GRID) {
            float X, Y;
            if (UNI) {
                const Geom g = geom_of(cx1, cy1, p, c.d2, c.p0, c.c2);
                cxa = g.xa; cya = g.ya; cxb = g.xb; cyb = g.yb;
                Y = g.Y; X = g.X;
                if (keep) *keep = g;
            } else {
                const float dx = __fsub_rn(cxb, cxa), dy = __fsub_rn(cyb, cya);
                const float det = __fsub_rn(0.0f, __fmul_rn(dx, dy));
                ok = ok && det != 0.0f;
                const float c01 = -__fdiv_rn(dx, det), c10 = -__fdiv_rn(dy, det);
                X = __fmul_rn(c01, __fsub_rn(yf, cya)); Y = __fmul_rn(c10, __fsub_rn(xf, cxa));
            }
            // u and v together (interpolator.rs:83 on both): the record interleaves their coefficients,
            // U = {u_sw, v_sw, u_a10, v_a10}, V = {u_a01, v_a01, u_a11, v_a11}
            // The products go two at a time; the sums stay scalar add.rn: ptxas (12.9) contracts a
            // mul.rn.f32x2 feeding an add.rn.f32x2 into one FFMA2 — even with -fmad=false — which would
            // drop the rounding of the product that the reference performs.
            const f32x2 XX = pk(X, X), YY = pk(Y, Y);
            const f32x2 m10 = mul2(pk(U.z, U.w), XX);
            const f32x2 m01 = mul2(pk(V.x, V.y), YY);
            const f32x2 m11 = mul2(mul2(pk(V.z, V.w), XX), YY);
            float u32 = __fadd_rn(__fadd_rn(__fadd_rn(U.x, lo_of(m10)), lo_of(m01)), lo_of(m11));
            float v32 = __fadd_rn(__fadd_rn(__fadd_rn(U.y, hi_of(m10)), hi_of(m01)), hi_of(m11));
            if (xf == cxa || xf == cxb) {
                const bool at_ya = yf == cya, at_yb = yf == cyb;
                if (at_ya || at_yb) {
                    const size_t node = (size_t)c.nx * cy1 + cx1;
                    u32 = corner_pick(u32, xf == cxa, xf == cxb, at_ya, at_yb, c.u + node, c.nx);
                    v32 = corner_pick(v32, xf == cxa, xf == cxb, at_ya, at_yb, c.v + node, c.nx);
                }
            }
            cv.u = (double)u32; cv.v = (double)v32;
            if (!kEarly) { cv.dudx = gu.x; cv.dudy = gu.y; cv.dvdx = gv.x; cv.dvdy = gv.y; }
        } else {
            cv.u = c.u0; cv.v = c.v0; cv.dudx = cv.dudy = cv.dvdx = cv.dvdy = 0.0;   // constant_current.rs:69-77
        }
    }

    // the four outputs from cg cos, cg sin (or their deep-water form), the bathymetric terms and the current
    __device__ __forceinline__ void assemble(double kx, double ky, double cg, double bx, double by,
                                             const CurrentVal &cv, double out[4])
    {
        out[0] = fma(cg, cs, cv.u);
        out[1] = fma(cg, sn, cv.v);
        if (kEarly) {
            out[2] = ax + bx;
            out[3] = ay + by;
        } else {
            out[2] = fma(-ky, cv.dvdx, fma(-kx, cv.dudx, bx));
            out[3] = fma(-ky, cv.dvdy, fma(-kx, cv.dudy, by));
        }
    }

    __device__ __forceinline__ void phase4(const BathyDev &b, const CurrentDev &c, double kx, double ky, double out[4])
    {
        const f32x2 p = pk(xf, yf);
        if (kEarly) {
            ax = fma(-ky, gv.x, -kx * gu.x);
            ay = fma(-ky, gv.y, -kx * gu.y);
        }
        if (kDmap) {
            // With the depth-floor map the current goes first: its record is what every lane waits for, and
            // the lanes in proven deep water then go straight to the deep-water terms.
            CurrentVal cv;
            Geom g;
            current_part(c, p, cv, kShare ? &g : nullptr);
            if (!MR_DEEP_ROOT4 && deep) {
                const double zk = k * 0.0;
                double sq, rq;
                sqrt_rsqrt(k * kExpRed[5], sq, rq);
                assemble(kx, ky, fma(kExpRed[6], rq, zk), -zk, -zk, cv, out);
                return;
            }
            if (deep) {
                // cg cos = w kx, cg sin = w ky with w = (sqrt(G)/2) k^-3/2; the bathymetric term is -0, and k2 * 0
                // carries an infinite k^2 into the wavenumber derivatives as NaN, like the general branch (see
                // wave_terms; w itself is NaN then: inv_root4)
                const double t = k;
                const double w = (t * t) * (t * kExpRed[7]), z = k2 * 0.0;
                out[0] = fma(w, kx, cv.u);          // (an infinite k^2 has already turned w into NaN: inv_root4)
                out[1] = fma(w, ky, cv.v);
                if (kEarly) {
                    out[2] = ax - z;
                    out[3] = ay - z;
                } else {
                    out[2] = fma(-ky, cv.dvdx, fma(-kx, cv.dudx, -z));
                    out[3] = fma(-ky, cv.dvdy, fma(-kx, cv.dudy, -z));
                }
                return;
            }
            if (MR_DEEP_ROOT4) {               // not proven deep: 1/k = t^2, k = k^2 / k, direction cosines
                const double rk = k * k;
                k = k2 * rk;
                cs = kx * rk; sn = ky * rk;
            }
            double cg, bx, by;
            float h32;
            double dhdx, dhdy;
            if (MR_DMAP_LATE_LOAD) {
                // the lanes that need the depth after all: cell, record address and the record itself, now
                if (!kSame) bathy_cell_again(b);
                brec = b.cell + 2ull * (unsigned)((b.nx - 1) * by1 + bx1);
                ldg_f4_d2(brec, Z, gh);
            }
            bathy_part(b, p, h32, dhdx, dhdy, kShare ? &g : nullptr);
            ok = ok && h32 > 0.0f;
            wave_terms(k, (double)(ok ? h32 : qnanf()), dhdx, dhdy, cg, bx, by);
            assemble(kx, ky, cg, bx, by, cv, out);
            return;
        }
        if (kSame) {
            // one cell geometry for both fields wherever they share the cell (nearly always)
            // (scheduling fence, see below: the first consumer of the current record waits for the bathymetry record
            // too, so that every load of the evaluation is in flight before the first wait)
            U.x = __int_as_float(__float_as_int(U.x) | (__float_as_int(Z.x) & b.zero));
            CurrentVal cv;
            Geom g;
            current_part(c, p, cv, kShare ? &g : nullptr);
            float h32;
            double dhdx, dhdy;
            bathy_part(b, p, h32, dhdx, dhdy, kShare ? &g : nullptr);
            ok = ok && h32 > 0.0f;
            const double h = (double)(ok ? h32 : qnanf());
            double cg, bx, by;
            wave_terms(k, h, dhdx, dhdy, cg, bx, by);
            assemble(kx, ky, cg, bx, by, cv, out);
            return;
        }
        // Scheduling fence.  ptxas places the first consumer of the bathymetry record ahead of the
        // current record's loads (whose f64 address chain is longer), so a warp waited for one L2
        // round trip, issued the other loads, and waited again (profiles/r1/g_*).  OR-ing in
        // (bits of the current record) & 0 — a zero the compiler cannot see — changes no value but
        // makes the first bathymetry consumer depend on both loads, so both are in flight first.
        if (BK == MR_BATHY_GRID && CK == MR_CURRENT_GRID && !kCmap)
            Z.x = __int_as_float(__float_as_int(Z.x) | (__float_as_int(U.x) & b.zero));
        float h32;
        double dhdx, dhdy;
        bathy_part(b, p, h32, dhdx, dhdy);
        CurrentVal cv;
        current_part(c, p, cv);
        // h <= 0 -> cg = NaN and the bathymetric term is NaN too (inf*0 or sqrt of a negative): all four
        // NaN, like a failed lookup.  The depth is poisoned while still an f32 (one select, then the
        // conversion the path needs anyway).
        // k == 0 -> Err (wave_ray_path.rs:181-183), four NaN as well, needs no test: sqrt_rsqrt(0) is
        // 0 * rsqrt(0) = 0 * inf = NaN, so k, both direction cosines and kh are NaN, kh >= 22 is false, and
        // every output of the general branch carries one of them (cg through sqrt(G k T), the bathymetric
        // term through -k/2, the advection terms added to NaN stay NaN).
        ok = ok && h32 > 0.0f;
        const double h = (double)(ok ? h32 : qnanf());
        double cg, bx, by;
        wave_terms(k, h, dhdx, dhdy, cg, bx, by);
        assemble(kx, ky, cg, bx, by, cv, out);
    }
};

// The RHS of NR rays carried by one thread, phase by phase.
template <int BK, int CK, bool UNI, int NR, bool DMAP, bool SG, bool CMAP>
__device__ __forceinline__ void rhs_fast_n(const BathyDev &b, const CurrentDev &c,
                                           const double (&s)[NR][4], double (&out)[NR][4])
{
    if (SG && NR == 1) {
        // Same-grid shortcut: the evaluation with one cell for both fields, or — for a point within sg_delta of a
        // grid line — the complete evaluation with the two separate lookups.  Two code paths rather than one with
        // a join: the shortcut path then carries one cell, one geometry and no bounds state at all.
        FastRay<BK, CK, UNI, DMAP, true> ray;
        if (ray.phase1(b, c, s[0][0], s[0][1], s[0][2], s[0][3])) {
            ray.phase2(b, c);
            ray.phase3(s[0][2], s[0][3]);
            ray.phase4(b, c, s[0][2], s[0][3], out[0]);
        } else {
            // (without the depth-floor map: it only ever skips work whose result cannot matter)
            FastRay<BK, CK, UNI, false, false> sep;
            sep.phase1(b, c, s[0][0], s[0][1], s[0][2], s[0][3]);
            sep.phase2(b, c);
            sep.phase3(s[0][2], s[0][3]);
            sep.phase4(b, c, s[0][2], s[0][3], out[0]);
        }
        return;
    }
    FastRay<BK, CK, UNI, DMAP, false, CMAP> ray[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) ray[r].phase1(b, c, s[r][0], s[r][1], s[r][2], s[r][3]);
#pragma unroll
    for (int r = 0; r < NR; ++r) ray[r].phase2(b, c);
#pragma unroll
    for (int r = 0; r < NR; ++r) ray[r].phase3(s[r][2], s[r][3]);
#pragma unroll
    for (int r = 0; r < NR; ++r) ray[r].phase4(b, c, s[r][2], s[r][3], out[r]);
}

// =============================================================================
// System::system (wave_ray_path.rs:220-234): Err -> four NaN
// =============================================================================
template <int BK, int CK, int MATH, bool UNI, int NR, bool DMAP = false, bool SG = false, bool CMAP = false>
__device__ __forceinline__ void rhs(const BathyDev &b, const CurrentDev &c,
                                    const double (&s)[NR][4], double (&out)[NR][4])
{
    if (MATH == MR_MATH_STRICT) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const double x = s[r][0], y = s[r][1], kx = s[r][2], ky = s[r][3];
            const float xf = (float)x, yf = (float)y;                                  // :122
            float h32, gx32, gy32;
            CurrentVal cv;
            bool ok = (BK == MR_BATHY_GRID) ? bathy_grid_strict(b, xf, yf, h32, gx32, gy32)
                                            : bathy_analytic(BK, b, xf, yf, h32, gx32, gy32);   // :120-122
            if (ok) {
                if (CK == MR_CURRENT_GRID) ok = current_grid_strict(c, x, y, cv);              // :129
                else { cv.u = c.u0; cv.v = c.v0; cv.dudx = cv.dudy = cv.dvdx = cv.dvdy = 0.0; }
            }
            if (!ok) out[r][0] = out[r][1] = out[r][2] = out[r][3] = qnan();
            else rhs_f64_strict(kx, ky, (double)h32, (double)gx32, (double)gy32, cv, out[r]);
        }
    } else {
        rhs_fast_n<BK, CK, UNI, NR, DMAP, SG, CMAP>(b, c, s, out);
    }
}

}  // namespace mr
