// mr_kernels_env.cu — the environment along rays: depth and current at every stored state.
//
// The reference sketches, but never fills, a per-ray record Ray{time, state, depth: Vec<f32>,
// current: Vec<Current>} (src/datatype.rs:165-194).  This kernel produces those two columns for the
// step-major trajectory planes the trace kernel wrote: for every stored (x, y) it evaluates
// BathymetryData::depth(&Point<f32>) (src/bathymetry/mod.rs:38) and CurrentData::current(&Point<f64>)
// (src/current/mod.rs:24); an Err becomes NaN.
//
// Two forms of the same lookup, value-identical:
//   * affine grids (the usual case: coordinates i*step, detected at upload): the cell records of the
//     fast trace kernel — fractional index by the exact constant division, one record load per field,
//     the bilinear from the record's corner combinations;
//   * anything else: the reference's lookups operation by operation (the strict functions).
//
// It moves 16 B in and up to 20 B out per stored row: meant to be bound by HBM, not by arithmetic.
// Compiled with -fmad=false like the strict trace kernel.
#include <algorithm>
#include "mr_device.cuh"
#include "mr_launch.hpp"

namespace mr {

// depth() on an affine grid: cartesian_netcdf3.rs:65-78 through the cell record (FastRay::phase1/phase4)
__device__ __forceinline__ float depth_affine(const BathyDev &b, float xf, float yf)
{
    const f32x2 p = pk(xf, yf);
    const f32x2 t = sub2(p, b.p0);
    f32x2 q = mul2(t, b.rs2);                                                   // fdiv_const, x and y at once
    q = fma2(fma2(q, b.ns2, t), b.rs2, q);
    q = fma2(fma2(q, b.ns2, t), b.rs2, q);
    const float ix = lo_of(q), iy = hi_of(q);
    const bool ok = ix >= 0.0f && ix <= b.nxm1f && iy >= 0.0f && iy <= b.nym1f;    // :291 (a NaN fails too)
    const int x1 = cell_of(ix, b.nx), y1 = cell_of(iy, b.ny);
    const float4 Z = ldg_f4(b.cell + 2u * (unsigned)((b.nx - 1) * y1 + x1));
    const f32x2 pa = fma2(pk((float)x1, (float)y1), b.d2, b.p0), pb = add2(pa, b.d2);
    const f32x2 yx = mul2(b.c2, sub2(p, pa));
    float h = bilinear_xy(hi_of(yx), lo_of(yx), Z);
    const float xa = lo_of(pa), ya = hi_of(pa), xb = lo_of(pb), yb = hi_of(pb);
    if (xf == xa || xf == xb) {
        const bool at_ya = yf == ya, at_yb = yf == yb;
        if (at_ya || at_yb) h = corner_pick(h, xf == xa, xf == xb, at_ya, at_yb, b.depth + (size_t)b.nx * y1 + x1, b.nx);
    }
    return ok ? h : qnanf();
}

// current() on an affine grid: cartesian_current.rs:448-467 through the cell record
__device__ __forceinline__ void current_affine(const CurrentDev &c, double x, double y, double &u, double &v)
{
    // :246, as FastRay::phase1 does it: RN(t/s) up to rounding-midpoint ties of measure 2^-52 ulp, and only
    // the cell is taken from it
    const double tx = x - c.xd0, ty = y - c.yd0;
    const double qx = __dmul_rn(tx, c.inv_sx), qy = __dmul_rn(ty, c.inv_sy);
    const double ix = __fma_rn(__fma_rn(-qx, c.sx, tx), c.inv_sx, qx);
    const double iy = __fma_rn(__fma_rn(-qy, c.sy, ty), c.inv_sy, qy);
    const bool ok = ix >= 0.0 && ix <= c.nxm1d && iy >= 0.0 && iy <= c.nym1d;       // :248
    const int x1 = cell_of(ix, c.nx), y1 = cell_of(iy, c.ny);
    float4 U, V;
    ldg_f4_f4(c.cell + 4u * (unsigned)((c.nx - 1) * y1 + x1), U, V);
    const float xf = (float)x, yf = (float)y;                                       // :456-465 `as f32`
    const f32x2 p = pk(xf, yf);
    const f32x2 pa = fma2(pk((float)x1, (float)y1), c.d2, c.p0), pb = add2(pa, c.d2);
    const f32x2 yx = mul2(c.c2, sub2(p, pa));
    const float Y = lo_of(yx), X = hi_of(yx);
    // U = {u_sw, v_sw, u_a10, v_a10}, V = {u_a01, v_a01, u_a11, v_a11}
    float u32 = bilinear_xy(X, Y, make_float4(U.x, U.z, V.x, V.z));
    float v32 = bilinear_xy(X, Y, make_float4(U.y, U.w, V.y, V.w));
    const float xa = lo_of(pa), ya = hi_of(pa), xb = lo_of(pb), yb = hi_of(pb);
    if (xf == xa || xf == xb) {
        const bool at_ya = yf == ya, at_yb = yf == yb;
        if (at_ya || at_yb) {
            const size_t node = (size_t)c.nx * y1 + x1;
            u32 = corner_pick(u32, xf == xa, xf == xb, at_ya, at_yb, c.u + node, c.nx);
            v32 = corner_pick(v32, xf == xa, xf == xb, at_ya, at_yb, c.v + node, c.nx);
        }
    }
    u = ok ? (double)u32 : qnan();
    v = ok ? (double)v32 : qnan();
}

// x, y: [rows][ld]; one thread per ray column, rows walked with stride gridDim.y, so a warp reads and
// writes whole lines of one row.  BA / CA: the bathymetry / the current is a grid with affine coordinates.
template <bool BA, bool CA>
__global__ void __launch_bounds__(256)
sample_kernel(const __grid_constant__ BathyDev b, const __grid_constant__ CurrentDev c,
              int64_t rows, int64_t n, int64_t ld,
              const double *__restrict__ xs, const double *__restrict__ ys,
              float *__restrict__ depth, double *__restrict__ us, double *__restrict__ vs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const int64_t o = r * ld + i;
        const double x = __ldcs(xs + o), y = __ldcs(ys + o);
        if (depth) {
            const float xf = (float)x, yf = (float)y;           // wave_ray_path.rs:122
            float h;
            if (BA) {
                h = depth_affine(b, xf, yf);
            } else {
                float gx, gy;
                const bool ok = (b.kind == MR_BATHY_GRID) ? bathy_grid_strict(b, xf, yf, h, gx, gy)
                                                          : bathy_analytic(b.kind, b, xf, yf, h, gx, gy);
                h = ok ? h : qnanf();
            }
            __stcs(depth + o, h);
        }
        if (us || vs) {
            double u, v;
            if (CA) {
                current_affine(c, x, y, u, v);
            } else if (c.kind == MR_CURRENT_GRID) {
                CurrentVal cv;
                const bool ok = current_grid_strict(c, x, y, cv);
                u = ok ? cv.u : qnan(); v = ok ? cv.v : qnan();
            } else {
                u = c.u0; v = c.v0;                             // constant_current.rs:51-53: the point is ignored
            }
            if (us) __stcs(us + o, u);
            if (vs) __stcs(vs + o, v);
        }
    }
}

cudaError_t launch_sample(const BathyDev &b, const CurrentDev &c, int64_t rows, int64_t n, int64_t ld,
                          const double *x, const double *y, float *depth, double *u, double *v,
                          cudaStream_t stream)
{
    if (rows <= 0 || n <= 0) return cudaSuccess;
    const unsigned gx = (unsigned)((n + 255) / 256);
    // enough row-walkers to fill the machine a few times over, never more than there are rows
    const int64_t want = (148 * 8 * 4 + gx - 1) / gx;
    const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(rows, want), 65535));
    const dim3 grid(gx, gy);
    const bool ba = b.kind == MR_BATHY_GRID && b.uniform, ca = c.kind == MR_CURRENT_GRID && c.uniform;
    if (ba && ca)  sample_kernel<true, true><<<grid, 256, 0, stream>>>(b, c, rows, n, ld, x, y, depth, u, v);
    else if (ba)   sample_kernel<true, false><<<grid, 256, 0, stream>>>(b, c, rows, n, ld, x, y, depth, u, v);
    else if (ca)   sample_kernel<false, true><<<grid, 256, 0, stream>>>(b, c, rows, n, ld, x, y, depth, u, v);
    else           sample_kernel<false, false><<<grid, 256, 0, stream>>>(b, c, rows, n, ld, x, y, depth, u, v);
    return cudaGetLastError();
}

}  // namespace mr
