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
// The f32 stages (position rounding, fractional index, cell choice, bilinear)
// are VALUE-IDENTICAL to the reference in both math modes: every f32 operation
// is an explicit round-to-nearest intrinsic so nvcc can never contract it.
// Only the f64 stage differs between MR_MATH_STRICT and MR_MATH_FAST.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mantaray_b200.h"

namespace mr {

// ---- device-resident field descriptors (passed by value as kernel params) ----

struct BathyDev {
    int32_t kind;
    int32_t nx, ny;
    const float  *x, *y;       // GRID coordinates (f32, as CartesianNetcdf3 holds them)
    const double *depth;       // GRID [ny*nx]
    const float  *array;       // ARRAY [nx*ny]
    float h0, x0, y0, dhdx, dhdy;
    // derived at upload (GRID)
    float  xf0, yf0;           // x[0], y[0]
    float  sx, sy;             // |x[1]-x[0]|, |y[1]-y[0]| in f32   (cartesian_netcdf3.rs:287)
    double x_space, y_space;   // x[1]-x[0] in f64 of the f32 values (cartesian_netcdf3.rs:119-120)
    double inv_x_space, inv_y_space;
};

struct CurrentDev {
    int32_t kind;
    int32_t nx, ny;
    const double *x, *y, *u, *v;
    double u0, v0;
    // derived at upload (GRID)
    double xd0, yd0;           // x[0], y[0]
    double sx, sy;             // |x[1]-x[0]|, |y[1]-y[0]|          (cartesian_current.rs:244)
    double inv_sx, inv_sy;     // RN(1/sx), RN(1/sy)
    double x_space, y_space;   // x[1]-x[0] (signed)                (cartesian_current.rs:515-516)
    double inv_x_space, inv_y_space;
};

static constexpr double kG = 9.8;            // src/wave_ray_path.rs:23

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ float  qnanf() { return __int_as_float(0x7fc00000); }

// ---- interpolator::bilinear on an axis-aligned cell -------------------------
// points a=(xa,ya,z_sw) b=(xa,yb,z_nw) c=(xb,yb,z_ne) d=(xb,ya,z_se), target (tx,ty).
// With this corner order bt=(0,dy), dt=(dx,0), so (interpolator.rs:64-75)
//   det = 0 - dx*dy,  c01 = -(dx/det),  c10 = -(dy/det),  c00 = c11 = 0/det,
//   X = c00*tt0 + c01*tt1 = RN(c01*tt1),  Y = c10*tt0 + c11*tt1 = RN(c10*tt0)
// (adding a signed zero is exact).  Returns false for det == 0 (Err).
__device__ __forceinline__ bool bilinear_cell(float xa, float xb, float ya, float yb,
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

// The cell rule of four_corners (cartesian_netcdf3.rs:344-387, cartesian_current.rs
// :293-336) for an index already known to be in [0, n-1]: left edge -> (0,1), right
// edge -> (n-2,n-1), on a grid line -> (i,i+1), else (floor,ceil).  All four cases
// are i1 = min(floor(index), n-2), i2 = i1+1 (n >= 2 is enforced at upload).
__device__ __forceinline__ int cell_of(float index, int n)
{
    int i = __float2int_rd(index);
    return i < n - 2 ? i : n - 2;
}
__device__ __forceinline__ int cell_of(double index, int n)
{
    int i = __double2int_rd(index);
    return i < n - 2 ? i : n - 2;
}

// ---- BathymetryData::depth_and_gradient -------------------------------------
// Returns false for Err (the whole RHS then becomes NaN, wave_ray_path.rs:222-228).
template <int BK, int MATH>
__device__ __forceinline__ bool bathy_eval(const BathyDev &b, float x, float y,
                                           float &h, float &gx, float &gy)
{
    if (BK == MR_BATHY_CONSTANT) {            // constant_depth.rs:39-45
        bool bad = isnan(x) || isnan(y);
        h = bad ? qnanf() : b.h0;
        gx = gy = bad ? qnanf() : 0.0f;
        return true;
    }
    if (BK == MR_BATHY_SLOPE) {               // constant_slope.rs:67-76
        bool bad = isnan(x) || isnan(y);
        float s = __fadd_rn(b.h0, __fmul_rn(b.dhdx, __fsub_rn(x, b.x0)));
        s = __fadd_rn(s, __fmul_rn(b.dhdy, __fsub_rn(y, b.y0)));
        h = bad ? qnanf() : s;
        gx = bad ? qnanf() : b.dhdx;
        gy = bad ? qnanf() : b.dhdy;
        return true;
    }
    if (BK == MR_BATHY_ARRAY) {               // array_depth.rs:27-35 (`as usize` saturates)
        unsigned long long xi = __float2ull_rz(x), yi = __float2ull_rz(y);   // NaN -> 0, negative -> 0, huge -> max
        unsigned long long len = (unsigned long long)b.nx;
        bool oob = xi >= len || yi >= len;
        h = oob ? qnanf() : __ldg(b.array + (oob ? 0 : xi * (unsigned long long)b.ny + yi));
        gx = gy = oob ? qnanf() : 0.0f;
        return true;
    }
    // GRID: cartesian_netcdf3.rs:98-135
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
    if (!bilinear_cell(xa, xb, ya, yb, (float)dsw, (float)dnw, (float)dne, (float)dse, x, y, h))
        return false;
    if (MATH == MR_MATH_STRICT) {             // :126-134
        gx = (float)__ddiv_rn(__dsub_rn(dse, dsw), b.x_space);
        gy = (float)__ddiv_rn(__dsub_rn(dnw, dsw), b.y_space);
    } else {
        gx = (float)((dse - dsw) * b.inv_x_space);
        gy = (float)((dnw - dsw) * b.inv_y_space);
    }
    return true;
}

// ---- CurrentData::current_and_gradient --------------------------------------
struct CurrentVal { double u, v, dudx, dudy, dvdx, dvdy; };

template <int CK, int MATH>
__device__ __forceinline__ bool current_eval(const CurrentDev &c, double x, double y, CurrentVal &o)
{
    if (CK == MR_CURRENT_CONSTANT) {          // constant_current.rs:69-77
        o.u = c.u0; o.v = c.v0;
        o.dudx = o.dudy = o.dvdx = o.dvdy = 0.0;
        return true;
    }
    // GRID: cartesian_current.rs:487-542.  f64 fractional index (:246).
    double tx = x - c.xd0, ty = y - c.yd0;
    double ix, iy;
    if (MATH == MR_MATH_STRICT) {
        ix = __ddiv_rn(tx, c.sx);
        iy = __ddiv_rn(ty, c.sy);
    } else {
        // quotient by the loop-invariant spacing: q0 = t*RN(1/s), one residual
        // correction (exact whenever t/s is representable, e.g. on a grid line)
        double qx = tx * c.inv_sx, qy = ty * c.inv_sy;
        ix = fma(fma(-qx, c.sx, tx), c.inv_sx, qx);
        iy = fma(fma(-qy, c.sy, ty), c.inv_sy, qy);
        // inf - keeps inf (fma(-inf, s, inf) is NaN): restore the reference's inf -> OOB
        if (isinf(tx)) ix = tx;
        if (isinf(ty)) iy = ty;
    }
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
    if (!bilinear_cell(xa, xb, ya, yb, (float)usw, (float)unw, (float)une, (float)use_, xf, yf, uf)) return false;
    if (!bilinear_cell(xa, xb, ya, yb, (float)vsw, (float)vnw, (float)vne, (float)vse, xf, yf, vf)) return false;
    o.u = (double)uf; o.v = (double)vf;
    if (MATH == MR_MATH_STRICT) {             // :522-536
        o.dudx = __ddiv_rn(__dsub_rn(use_, usw), c.x_space);
        o.dudy = __ddiv_rn(__dsub_rn(unw, usw), c.y_space);
        o.dvdx = __ddiv_rn(__dsub_rn(vse, vsw), c.x_space);
        o.dvdy = __ddiv_rn(__dsub_rn(vnw, vsw), c.y_space);
    } else {
        o.dudx = (use_ - usw) * c.inv_x_space;
        o.dudy = (unw - usw) * c.inv_y_space;
        o.dvdx = (vse - vsw) * c.inv_x_space;
        o.dvdy = (vnw - vsw) * c.inv_y_space;
    }
    return true;
}

// ---- the f64 stage, reference expression tree (MR_MATH_STRICT) ---------------
// wave_ray_path.rs:132-147 with group_velocity :177-188 and dkdt_bathy :207-216.
// This translation unit is compiled with -fmad=false; the explicit _rn
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

// ---- the f64 stage, restructured (MR_MATH_FAST) ------------------------------
// Same functions of (k, h), evaluated from ONE exponential:
//   E = exp(-2kh), m = 1-E = -expm1(-2kh), w = 1+E
//   tanh kh = m/w,  1/cosh^2 kh = 4E/w^2,  1/(sinh kh cosh kh) = 4E/(m w)
// and cos(theta) = kx/k, sin(theta) = ky/k instead of atan2 + sincos.
// Special values follow the reference: h <= 0, h NaN, k == 0, k NaN -> four NaN;
// large kh: E underflows to 0 -> tanh = 1, second cg term 0, bathymetric term -0
// (the reference gets the same from cosh^2 -> inf and sinh -> inf).
__device__ __forceinline__ void rhs_f64_fast(double kx, double ky, double h, double dhdx, double dhdy,
                                             const CurrentVal &cv, double out[4])
{
    double k2 = fma(kx, kx, ky * ky);
    bool ok = (h > 0.0) && (k2 > 0.0);        // false for NaN h / NaN k as well
    double rk = rsqrt(k2);
    double k = k2 * rk;
    double cs = kx * rk, sn = ky * rk;
    double kh = k * h;
    double em = expm1(-2.0 * kh);             // E - 1, in [-1, 0)
    double E = 1.0 + em;
    double m = -em, w = 2.0 + em;
    double r = 1.0 / (m * w);
    double invw = m * r;
    double T = m * invw;                      // tanh(kh)
    double E4 = 4.0 * E;
    double sech2 = E4 * invw * invw;
    double csch_sech = E4 * r;
    double q = (k * kG) * T;
    double rq = rsqrt(q);
    double cg = (kG * 0.5) * ((T + kh * sech2) * rq);
    double Bc = (-0.5 * k) * csch_sech * (q * rq);
    double nan = qnan();
    out[0] = ok ? fma(cg, cs, cv.u) : nan;
    out[1] = ok ? fma(cg, sn, cv.v) : nan;
    out[2] = ok ? fma(-ky, cv.dvdx, fma(-kx, cv.dudx, Bc * dhdx)) : nan;
    out[3] = ok ? fma(-ky, cv.dvdy, fma(-kx, cv.dudy, Bc * dhdy)) : nan;
}

// ---- System::system (wave_ray_path.rs:220-234): Err -> four NaN ---------------
template <int BK, int CK, int MATH>
__device__ __forceinline__ void rhs(const BathyDev &b, const CurrentDev &c,
                                    double x, double y, double kx, double ky, double out[4])
{
    float h32, gx32, gy32;
    bool ok = bathy_eval<BK, MATH>(b, (float)x, (float)y, h32, gx32, gy32);   // :120-122
    CurrentVal cv;
    if (ok) ok = current_eval<CK, MATH>(c, x, y, cv);                          // :129
    if (!ok) {
        out[0] = out[1] = out[2] = out[3] = qnan();
        return;
    }
    if (MATH == MR_MATH_STRICT)
        rhs_f64_strict(kx, ky, (double)h32, (double)gx32, (double)gy32, cv, out);
    else
        rhs_f64_fast(kx, ky, (double)h32, (double)gx32, (double)gy32, cv, out);
}

}  // namespace mr
