/*
 * mr_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * An operation-for-operation restatement, in plain C, of the arithmetic on
 * mantaray's batch ray-tracing path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; nothing
 * under mantaray_b200/ does.
 *
 * Pinning: the reference is Rust and cannot be built in this image (no cargo /
 * rustc); its RK4 stepper is the un-vendored crate ode_solvers 0.4.0
 * (Cargo.toml:33, Cargo.lock:653-656) on nalgebra 0.32.6 (Cargo.lock:515-518).
 * The oracle is pinned against every known-answer value the reference's own
 * tests hold for this path (tests/test_oracle_kat.py lists them with
 * file:line); multi-step trajectories over gridded fields are NOT pinned by
 * any reference test, only by this restatement.
 *
 * Arithmetic rules followed here (so build with
 *   gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math):
 *   - Rust never contracts a*b+c into an FMA and evaluates left to right;
 *   - f32 expressions round to f32 after every operation;
 *   - `as usize` saturates (NaN -> 0, negative -> 0);
 *   - f64::{tanh,sinh,cosh,atan2,sin,cos} are the platform libm (glibc here).
 *
 * Citations are file:line into the reference tree.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mantaray_b200.h"
#include "mr_oracle.h"

/* const G: f64 = 9.8;  src/wave_ray_path.rs:23 */
static const double G = 9.8;

/* ------------------------------------------------------------------------- */
/* Rust cast helpers                                                          */
/* ------------------------------------------------------------------------- */

/* `f as usize`: saturating, NaN -> 0. */
static size_t f32_as_usize(float f)
{
    if (!(f > 0.0f)) return 0;                 /* NaN, negatives, zero */
    if (f >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)f;
}
static size_t f64_as_usize(double f)
{
    if (!(f > 0.0)) return 0;
    if (f >= 18446744073709551616.0) return SIZE_MAX;
    return (size_t)f;
}

/* ------------------------------------------------------------------------- */
/* interpolator::bilinear — src/interpolator.rs:39-84 (all f32)               */
/* points = [a, b, c, d], each (x, y, z); returns 0 = Ok, 1 = Err             */
/* ------------------------------------------------------------------------- */
int orc_bilinear(const float pts[4][3], float tx, float ty, float *out)
{
    /* :46-50 target coincident with a point */
    for (int i = 0; i < 4; ++i) {
        if (tx == pts[i][0] && ty == pts[i][1]) {
            *out = pts[i][2];
            return 0;
        }
    }
    const float *a = pts[0], *b = pts[1], *c = pts[2], *d = pts[3];

    /* :59-61 translate b, d and the target with respect to a */
    float bt0 = b[0] - a[0], bt1 = b[1] - a[1];
    float dt0 = d[0] - a[0], dt1 = d[1] - a[1];
    float tt0 = tx - a[0],  tt1 = ty - a[1];

    /* :64-67 */
    float p0 = bt0 * dt1;
    float p1 = dt0 * bt1;
    float det = p0 - p1;
    if (det == 0.0f) return 1;

    /* :69-72 inverse change-of-basis matrix */
    float c00 = dt1 / det;
    float c01 = -(dt0 / det);
    float c10 = -(bt1 / det);
    float c11 = bt0 / det;

    /* :74-75 */
    float m0 = c00 * tt0, m1 = c01 * tt1;
    float x = m0 + m1;
    float m2 = c10 * tt0, m3 = c11 * tt1;
    float y = m2 + m3;

    /* :78-81 */
    float a00 = a[2];
    float a10 = b[2] - a[2];
    float a01 = d[2] - a[2];
    float s0 = c[2] - a[2];
    float s1 = s0 - a10;
    float a11 = s1 - a01;

    /* :83  a00 + a10*x + a01*y + a11*x*y, left to right */
    float q0 = a10 * x;
    float r0 = a00 + q0;
    float q1 = a01 * y;
    float r1 = r0 + q1;
    float q2 = a11 * x;
    float q3 = q2 * y;
    *out = r1 + q3;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* CartesianNetcdf3 — src/bathymetry/cartesian_netcdf3.rs                     */
/* ------------------------------------------------------------------------- */

/* nearest :274-296 (f32 fractional index); 0 = Ok, 1 = Err */
int orc_bathy_nearest(float target, const float *arr, int n, float *index)
{
    if (n <= 0) return 1;                       /* :276-278 */
    if (n == 1) { *index = 0.0f; return 0; }    /* :281-283 */
    float spacing = fabsf(arr[1] - arr[0]);     /* :287 */
    float t = target - arr[0];
    float idx = t / spacing;                    /* :289 */
    if (idx < 0.0f || idx > (float)(n - 1)) return 1;   /* :291-292 */
    *index = idx;
    return 0;
}

/* The edge / on-gridline / interior rule of four_corners, :344-387 (f32). */
static void cell_rule_f32(float index, int n, size_t *i1, size_t *i2)
{
    float low = 0.0f, high = (float)(n - 1);
    if (index == low) {
        *i1 = f32_as_usize(index);
        *i2 = f32_as_usize(index) + 1;
    } else if (index == high) {
        *i1 = f32_as_usize(index) - 1;
        *i2 = f32_as_usize(index);
    } else if (index - truncf(index) == 0.0f) {       /* fract() == 0.0 */
        *i1 = f32_as_usize(roundf(index));
        *i2 = *i1 + 1;
    } else {
        *i1 = f32_as_usize(floorf(index));
        *i2 = f32_as_usize(ceilf(index));
    }
}

/* four_corners :334-390 -> [(x1,y1),(x1,y2),(x2,y2),(x2,y1)] */
int orc_bathy_four_corners(const mr_bathymetry_desc *b, float x, float y, size_t c[4][2])
{
    float xi, yi;
    if (orc_bathy_nearest(x, b->x, b->nx, &xi)) return 1;   /* :315 */
    if (orc_bathy_nearest(y, b->y, b->ny, &yi)) return 1;   /* :316 */
    size_t x1, x2, y1, y2;
    cell_rule_f32(xi, b->nx, &x1, &x2);
    cell_rule_f32(yi, b->ny, &y1, &y2);
    c[0][0] = x1; c[0][1] = y1;
    c[1][0] = x1; c[1][1] = y2;
    c[2][0] = x2; c[2][1] = y2;
    c[3][0] = x2; c[3][1] = y1;
    return 0;
}

/* depth_at_indexes :465-471 */
static int bathy_depth_at(const mr_bathymetry_desc *b, size_t xi, size_t yi, double *v)
{
    size_t index = (size_t)b->nx * yi + xi;
    if (index >= (size_t)b->nx * (size_t)b->ny) return 1;
    *v = b->depth[index];
    return 0;
}

/* interpolate :417-445 */
static int bathy_interpolate(const mr_bathymetry_desc *b, size_t c[4][2], float tx, float ty, float *out)
{
    float pts[4][3];
    for (int i = 0; i < 4; ++i) {
        double d;
        if (c[i][0] >= (size_t)b->nx || c[i][1] >= (size_t)b->ny) return 1; /* Rust would panic on x[..]; unreachable for nx,ny>=2 */
        if (bathy_depth_at(b, c[i][0], c[i][1], &d)) return 1;
        pts[i][0] = b->x[c[i][0]];
        pts[i][1] = b->y[c[i][1]];
        pts[i][2] = (float)d;
    }
    return orc_bilinear(pts, tx, ty, out);
}

/* depth_and_gradient for the gridded file :98-135 */
static int grid_depth_and_gradient(const mr_bathymetry_desc *b, float x, float y,
                                   float *h, float *dhdx, float *dhdy)
{
    if (isnan(x) || isnan(y)) {                 /* :101-103 */
        *h = NAN; *dhdx = NAN; *dhdy = NAN;
        return 0;
    }
    size_t c[4][2];
    if (orc_bathy_four_corners(b, x, y, c)) return 1;       /* :105-108 */
    if (bathy_interpolate(b, c, x, y, h)) return 1;         /* :111 */

    double x_space = (double)b->x[1] - (double)b->x[0];     /* :119 */
    double y_space = (double)b->y[1] - (double)b->y[0];     /* :120 */
    double sw, nw, se;
    if (bathy_depth_at(b, c[0][0], c[0][1], &sw)) return 1;
    if (bathy_depth_at(b, c[1][0], c[1][1], &nw)) return 1;
    if (bathy_depth_at(b, c[3][0], c[3][1], &se)) return 1;
    double xg = (se - sw) / x_space;            /* :126-128 */
    double yg = (nw - sw) / y_space;            /* :130-132 */
    *dhdx = (float)xg;                          /* :134 */
    *dhdy = (float)yg;
    return 0;
}

/* BathymetryData::depth_and_gradient for every kind; 0 = Ok, 1 = Err */
int orc_depth_and_gradient(const mr_bathymetry_desc *b, float x, float y,
                           float *h, float *dhdx, float *dhdy)
{
    switch (b->kind) {
    case MR_BATHY_CONSTANT:                     /* constant_depth.rs:39-45 */
        if (isnan(x) || isnan(y)) { *h = NAN; *dhdx = NAN; *dhdy = NAN; }
        else { *h = b->h0; *dhdx = 0.0f; *dhdy = 0.0f; }
        return 0;
    case MR_BATHY_SLOPE:                        /* constant_slope.rs:67-76 */
        if (isnan(x) || isnan(y)) { *h = NAN; *dhdx = NAN; *dhdy = NAN; }
        else {
            float ax = x - b->x0;
            float px = b->dhdx * ax;
            float s  = b->h0 + px;
            float ay = y - b->y0;
            float py = b->dhdy * ay;
            *h = s + py;
            *dhdx = b->dhdx; *dhdy = b->dhdy;
        }
        return 0;
    case MR_BATHY_GRID:
        return grid_depth_and_gradient(b, x, y, h, dhdx, dhdy);
    case MR_BATHY_ARRAY: {                      /* array_depth.rs:27-35 */
        size_t xi = f32_as_usize(x), yi = f32_as_usize(y);
        size_t len = (size_t)b->nx;             /* both compared with the OUTER length */
        if (xi >= len || yi >= len) { *h = NAN; *dhdx = NAN; *dhdy = NAN; }
        else { *h = b->array[xi * (size_t)b->ny + yi]; *dhdx = 0.0f; *dhdy = 0.0f; }
        return 0;
    }
    default:
        return 1;
    }
}

/* BathymetryData::depth (used by a few reference tests; cartesian_netcdf3.rs:65-77) */
int orc_depth(const mr_bathymetry_desc *b, float x, float y, float *h)
{
    if (b->kind == MR_BATHY_GRID) {
        if (isnan(x) || isnan(y)) { *h = NAN; return 0; }
        size_t c[4][2];
        if (orc_bathy_four_corners(b, x, y, c)) return 1;
        return bathy_interpolate(b, c, x, y, h);
    }
    float gx, gy;
    return orc_depth_and_gradient(b, x, y, h, &gx, &gy);
}

/* ------------------------------------------------------------------------- */
/* CartesianCurrent — src/current/cartesian_current.rs                        */
/* ------------------------------------------------------------------------- */

/* nearest :231-253 (f64 fractional index; NO NaN pre-check) */
int orc_current_nearest(double target, const double *arr, int n, double *index)
{
    if (n <= 0) return 1;
    if (n == 1) { *index = 0.0; return 0; }
    double spacing = fabs(arr[1] - arr[0]);     /* :244 */
    double t = target - arr[0];
    double idx = t / spacing;                   /* :246 */
    if (idx < 0.0 || idx > (double)(n - 1)) return 1;   /* :248-249 */
    *index = idx;
    return 0;
}

static void cell_rule_f64(double index, int n, size_t *i1, size_t *i2)
{
    double low = 0.0, high = (double)(n - 1);   /* :287-290 */
    if (index == low) {
        *i1 = f64_as_usize(index);
        *i2 = f64_as_usize(index) + 1;
    } else if (index == high) {
        *i1 = f64_as_usize(index) - 1;
        *i2 = f64_as_usize(index);
    } else if (index - trunc(index) == 0.0) {
        *i1 = f64_as_usize(round(index));
        *i2 = *i1 + 1;
    } else {
        *i1 = f64_as_usize(floor(index));
        *i2 = f64_as_usize(ceil(index));
    }
}

/* four_corners :283-339 */
int orc_current_four_corners(const mr_current_desc *cu, double x, double y, size_t c[4][2])
{
    double xi, yi;
    if (orc_current_nearest(x, cu->x, cu->nx, &xi)) return 1;
    if (orc_current_nearest(y, cu->y, cu->ny, &yi)) return 1;
    size_t x1, x2, y1, y2;
    cell_rule_f64(xi, cu->nx, &x1, &x2);
    cell_rule_f64(yi, cu->ny, &y1, &y2);
    c[0][0] = x1; c[0][1] = y1;
    c[1][0] = x1; c[1][1] = y2;
    c[2][0] = x2; c[2][1] = y2;
    c[3][0] = x2; c[3][1] = y1;
    return 0;
}

/* val_from_arr :421-427 */
static int current_val(const mr_current_desc *cu, size_t xi, size_t yi, const double *arr, double *v)
{
    size_t index = (size_t)cu->nx * yi + xi;
    if (index >= (size_t)cu->nx * (size_t)cu->ny) return 1;
    *v = arr[index];
    return 0;
}

/* interpolate :365-398 (corners and target cast to f32) */
static int current_interpolate(const mr_current_desc *cu, size_t c[4][2], float tx, float ty,
                               const double *arr, float *out)
{
    float pts[4][3];
    for (int i = 0; i < 4; ++i) {
        double z;
        if (c[i][0] >= (size_t)cu->nx || c[i][1] >= (size_t)cu->ny) return 1;
        if (current_val(cu, c[i][0], c[i][1], arr, &z)) return 1;
        pts[i][0] = (float)cu->x[c[i][0]];
        pts[i][1] = (float)cu->y[c[i][1]];
        pts[i][2] = (float)z;
    }
    return orc_bilinear(pts, tx, ty, out);
}

/* CurrentData::current_and_gradient for every kind; 0 = Ok, 1 = Err.
 * grad = (dudx, dudy, dvdx, dvdy) */
int orc_current_and_gradient(const mr_current_desc *cu, double x, double y,
                             double *u, double *v, double grad[4])
{
    if (cu->kind == MR_CURRENT_CONSTANT) {      /* constant_current.rs:69-77 */
        *u = cu->u0; *v = cu->v0;
        grad[0] = grad[1] = grad[2] = grad[3] = 0.0;
        return 0;
    }
    if (cu->kind != MR_CURRENT_GRID) return 1;

    size_t c[4][2];
    if (orc_current_four_corners(cu, x, y, c)) return 1;    /* :492-495 */
    float uf, vf;
    if (current_interpolate(cu, c, (float)x, (float)y, cu->u, &uf)) return 1;   /* :498-502 */
    if (current_interpolate(cu, c, (float)x, (float)y, cu->v, &vf)) return 1;   /* :503-507 */

    double x_space = cu->x[1] - cu->x[0];       /* :515 */
    double y_space = cu->y[1] - cu->y[0];       /* :516 */
    double usw, unw, use_, vsw, vnw, vse;
    if (current_val(cu, c[3][0], c[3][1], cu->u, &use_)) return 1;
    if (current_val(cu, c[0][0], c[0][1], cu->u, &usw)) return 1;
    if (current_val(cu, c[1][0], c[1][1], cu->u, &unw)) return 1;
    if (current_val(cu, c[3][0], c[3][1], cu->v, &vse)) return 1;
    if (current_val(cu, c[0][0], c[0][1], cu->v, &vsw)) return 1;
    if (current_val(cu, c[1][0], c[1][1], cu->v, &vnw)) return 1;
    grad[0] = (use_ - usw) / x_space;           /* dudx :522-524 */
    grad[1] = (unw - usw) / y_space;            /* dudy :526-528 */
    grad[2] = (vse - vsw) / x_space;            /* dvdx :530-532 */
    grad[3] = (vnw - vsw) / y_space;            /* dvdy :534-536 */
    *u = (double)uf; *v = (double)vf;           /* :539 */
    return 0;
}

/* CurrentData::current for every kind; 0 = Ok, 1 = Err.
 * constant_current.rs:51-53 (the point is ignored); cartesian_current.rs:448-467 */
int orc_current(const mr_current_desc *cu, double x, double y, double *u, double *v)
{
    if (cu->kind == MR_CURRENT_CONSTANT) { *u = cu->u0; *v = cu->v0; return 0; }
    if (cu->kind != MR_CURRENT_GRID) return 1;
    size_t c[4][2];
    if (orc_current_four_corners(cu, x, y, c)) return 1;                        /* :450-453 */
    float uf, vf;
    if (current_interpolate(cu, c, (float)x, (float)y, cu->u, &uf)) return 1;   /* :456-460 */
    if (current_interpolate(cu, c, (float)x, (float)y, cu->v, &vf)) return 1;   /* :461-465 */
    *u = (double)uf; *v = (double)vf;                                           /* :467 */
    return 0;
}

/* depth() at (x as f32, y as f32) (wave_ray_path.rs:122) and current() at (x, y) for `count` points:
 * the depth and current columns of Ray{time,state,depth,current}, datatype.rs:165-194.  Err -> NaN. */
void orc_sample_fields(const mr_bathymetry_desc *b, const mr_current_desc *cu, int64_t count,
                       const double *x, const double *y, float *depth, double *u, double *v)
{
    for (int64_t i = 0; i < count; ++i) {
        if (depth) {
            float h;
            depth[i] = orc_depth(b, (float)x[i], (float)y[i], &h) ? NAN : h;
        }
        if (u || v) {
            double uu, vv;
            if (orc_current(cu, x[i], y[i], &uu, &vv)) uu = vv = NAN;
            if (u) u[i] = uu;
            if (v) v[i] = vv;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* WaveRayPath — src/wave_ray_path.rs                                         */
/* ------------------------------------------------------------------------- */

/* group_velocity :177-188; 0 = Ok, 1 = Err(ArgumentOutOfBounds) */
int orc_group_velocity(double k, double h, double *cg)
{
    if (h <= 0.0) { *cg = NAN; return 0; }      /* :178-180 */
    if (k <= 0.0) return 1;                     /* :181-183 */
    double kh = k * h;
    double c = cosh(kh);
    double t1 = tanh(kh) + kh / (c * c);        /* (k*h).cosh().powi(2) */
    double t2 = sqrt(k * G * tanh(kh));
    *cg = (G / 2.0) * (t1 / t2);                /* :184-186 */
    return 0;
}

/* dkdt_bathy :207-216 */
void orc_dkdt_bathy(double k, double h, double dhdx, double dhdy, double *ox, double *oy)
{
    /* (-0.5) * k * 1.0 / sinh * 1.0 / cosh * sqrt(G*k*tanh) * dhdx, left to right */
    double a = (-0.5) * k;
    a = a * 1.0;
    a = a / sinh(k * h);
    a = a * 1.0;
    a = a / cosh(k * h);
    a = a * sqrt(G * k * tanh(k * h));
    *ox = a * dhdx;
    double b = (-0.5) * k;
    b = b * 1.0;
    b = b / sinh(k * h);
    b = b * 1.0;
    b = b / cosh(k * h);
    b = b * sqrt(G * k * tanh(k * h));
    *oy = b * dhdy;
}

/* odes :118-150; 0 = Ok, 1 = Err */
int orc_odes(const mr_bathymetry_desc *b, const mr_current_desc *cu,
             double x, double y, double kx, double ky, double out[4])
{
    float h32, dhx32, dhy32;
    if (orc_depth_and_gradient(b, (float)x, (float)y, &h32, &dhx32, &dhy32)) return 1;  /* :120-122 */
    double h = (double)h32, dhdx = (double)dhx32, dhdy = (double)dhy32;                 /* :124-126 */

    double u, v, g[4];
    if (orc_current_and_gradient(cu, x, y, &u, &v, g)) return 1;                        /* :129 */

    double k = sqrt(kx * kx + ky * ky);         /* :132 */
    double theta = atan2(ky, kx);               /* :133 */

    double cg;
    if (orc_group_velocity(k, h, &cg)) return 1;    /* :136 */
    double cgx = cg * cos(theta) + u;           /* :137 */
    double cgy = cg * sin(theta) + v;           /* :138 */

    double bx, by;
    orc_dkdt_bathy(k, h, dhdx, dhdy, &bx, &by); /* :144 */

    double dkx = bx - kx * g[0] - ky * g[2];    /* :146  du.dx, dv.dx */
    double dky = by - kx * g[1] - ky * g[3];    /* :147  du.dy, dv.dy */

    out[0] = cgx; out[1] = cgy; out[2] = dkx; out[3] = dky;
    return 0;
}

/* System::system :220-234 — Err becomes four NaNs */
void orc_system(const mr_bathymetry_desc *b, const mr_current_desc *cu, const double s[4], double ds[4])
{
    if (orc_odes(b, cu, s[0], s[1], s[2], s[3], ds)) {
        ds[0] = ds[1] = ds[2] = ds[3] = NAN;
    }
}

/* solout :236-246 */
static int solout(const double y[4], const double dy[4])
{
    return (isnan(dy[0]) && isnan(dy[1]) && isnan(dy[2]) && isnan(dy[3]))
        || (isnan(y[0]) && isnan(y[1]) && isnan(y[2]) && isnan(y[3]));
}

/* ------------------------------------------------------------------------- */
/* ode_solvers 0.4.0 Rk4 (external crate; published algorithm restated)       */
/*   integrate(): push (x0,y0); n = ceil((x_end-x)/h); per step:              */
/*     k0=f(x,y); k1=f(x+h/2, y+k0*(h/2)); k2=f(x+h/2, y+k1*(h/2));           */
/*     k3=f(x+h, y+k2*h); x_new = x+h;                                        */
/*     y_new = y + (k0 + k1*2 + k2*2 + k3)*(h/6);  push; stop if              */
/*     solout(x_new, y_new, k0).                                              */
/*   nalgebra Vector4 ops are elementwise, no FMA.                            */
/* Call site: src/ray.rs:205-212.                                             */
/* ------------------------------------------------------------------------- */
int64_t orc_num_steps(double t0, double t_end, double dt)
{
    if (!(dt > 0.0)) return -1;
    double q = ceil((t_end - t0) / dt);
    if (isnan(q) || !(q < 9.0e15)) return -1;
    if (q < 0.0) return 0;      /* `as usize` saturates a negative quotient: no steps, the initial row only */
    return (int64_t)q;
}

void orc_rk4_step(const mr_bathymetry_desc *b, const mr_current_desc *cu, double dt,
                  const double y[4], double ynew[4], double k0[4])
{
    double half = dt / 2.0;
    double k1[4], k2[4], k3[4], tmp[4];
    orc_system(b, cu, y, k0);
    for (int c = 0; c < 4; ++c) { double p = k0[c] * half; tmp[c] = y[c] + p; }
    orc_system(b, cu, tmp, k1);
    for (int c = 0; c < 4; ++c) { double p = k1[c] * half; tmp[c] = y[c] + p; }
    orc_system(b, cu, tmp, k2);
    for (int c = 0; c < 4; ++c) { double p = k2[c] * dt; tmp[c] = y[c] + p; }
    orc_system(b, cu, tmp, k3);
    double sixth = dt / 6.0;
    for (int c = 0; c < 4; ++c) {
        double a = k1[c] * 2.0;
        double s = k0[c] + a;
        double bb = k2[c] * 2.0;
        s = s + bb;
        s = s + k3[c];
        double inc = s * sixth;
        ynew[c] = y[c] + inc;
    }
}

/* One ray.  Writes row j (step j*stride) of ray `col` at out[j*ld + col]. */
static void trace_one(const mr_bathymetry_desc *b, const mr_current_desc *cu,
                      double x0, double y0, double kx0, double ky0,
                      double dt, int64_t nsteps, int stride,
                      double *ox, double *oy, double *okx, double *oky, int64_t ld, int64_t col,
                      int32_t *rows_out, int32_t *len_out, double *fin, int64_t fin_ld)
{
    double y[4] = { x0, y0, kx0, ky0 };
    double last_ok[4] = { NAN, NAN, NAN, NAN };
    int64_t rows = 1, len = 0;
    int have_nan = isnan(y[0]) || isnan(y[1]) || isnan(y[2]) || isnan(y[3]);
    if (!have_nan) { len = 1; memcpy(last_ok, y, sizeof y); }
    if (ox)  ox [col] = y[0];
    if (oy)  oy [col] = y[1];
    if (okx) okx[col] = y[2];
    if (oky) oky[col] = y[3];

    for (int64_t s = 1; s <= nsteps; ++s) {
        double yn[4], k0[4];
        orc_rk4_step(b, cu, dt, y, yn, k0);
        memcpy(y, yn, sizeof y);
        rows = s + 1;
        if (s % stride == 0) {
            int64_t j = s / stride;
            if (ox)  ox [j * ld + col] = y[0];
            if (oy)  oy [j * ld + col] = y[1];
            if (okx) okx[j * ld + col] = y[2];
            if (oky) oky[j * ld + col] = y[3];
        }
        if (!have_nan) {
            if (isnan(y[0]) || isnan(y[1]) || isnan(y[2]) || isnan(y[3])) have_nan = 1;
            else { len = s + 1; memcpy(last_ok, y, sizeof y); }
        }
        if (solout(y, k0)) break;
    }
    /* rows the ray never reached are NaN (python/mantaray/core.py:115-119) */
    int64_t nrows_out = nsteps / stride + 1;
    for (int64_t j = (rows - 1) / stride + 1; j < nrows_out; ++j) {
        if (ox)  ox [j * ld + col] = NAN;
        if (oy)  oy [j * ld + col] = NAN;
        if (okx) okx[j * ld + col] = NAN;
        if (oky) oky[j * ld + col] = NAN;
    }
    if (rows_out) rows_out[col] = (int32_t)rows;
    if (len_out)  len_out[col]  = (int32_t)len;
    if (fin) for (int c = 0; c < 4; ++c) fin[c * fin_ld + col] = last_ok[c];
}

typedef struct {
    const mr_bathymetry_desc *b; const mr_current_desc *cu;
    const double *x0, *y0, *kx0, *ky0;
    double dt; int64_t nsteps; int stride;
    double *x, *y, *kx, *ky; int64_t n;
    int32_t *rows, *len; double *fin;
    atomic_llong next;                          /* next unclaimed ray */
    int64_t grain;
} job_t;

/* rayon's par_iter steals work; here threads claim `grain` rays at a time so
 * regions where rays stop early do not leave threads idle. */
static void *worker(void *p)
{
    job_t *j = (job_t *)p;
    for (;;) {
        int64_t lo = atomic_fetch_add(&j->next, j->grain);
        if (lo >= j->n) break;
        int64_t hi = lo + j->grain < j->n ? lo + j->grain : j->n;
        for (int64_t i = lo; i < hi; ++i)
            trace_one(j->b, j->cu, j->x0[i], j->y0[i], j->kx0[i], j->ky0[i], j->dt, j->nsteps, j->stride,
                      j->x, j->y, j->kx, j->ky, j->n, i, j->rows, j->len, j->fin, j->n);
    }
    return NULL;
}

/* ManyRays::trace_many src/ray.rs:98-127 (rayon par_iter -> pthreads; order
 * preserved because every ray writes its own column).  Same output contract
 * as mr_trace_many in include/mantaray_b200.h. */
int orc_trace_many(const mr_bathymetry_desc *b, const mr_current_desc *cu, int64_t n,
                   const double *x0, const double *y0, const double *kx0, const double *ky0,
                   double t0, double t_end, double dt, int32_t stride, int32_t nthreads,
                   double *t, double *x, double *y, double *kx, double *ky,
                   int32_t *rows, int32_t *len, double *final_state)
{
    if (stride <= 0) stride = 1;
    int64_t nsteps = orc_num_steps(t0, t_end, dt);
    if (nsteps < 0 || n < 0) return MR_ERR_BAD_ARG;
    if (t) {
        double tt = t0;
        t[0] = tt;
        for (int64_t s = 1; s <= nsteps; ++s) {
            tt = tt + dt;                       /* x_new = x + h, accumulated */
            if (s % stride == 0) t[s / stride] = tt;
        }
    }
    if (n == 0) return MR_OK;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = (int32_t)n;
    job_t job = { b, cu, x0, y0, kx0, ky0, dt, nsteps, stride, x, y, kx, ky, n, rows, len, final_state, 0, 0 };
    atomic_init(&job.next, 0);
    job.grain = n / ((int64_t)nthreads * 16);
    if (job.grain < 1) job.grain = 1;
    if (job.grain > 1024) job.grain = 1024;
    if (nthreads == 1) {
        worker(&job);
        return MR_OK;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return MR_ERR_OOM;
    int started = 0;
    for (int i = 0; i < nthreads; ++i) {
        if (pthread_create(&th[i], NULL, worker, &job) != 0) break;
        started++;
    }
    if (started == 0) worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
    return MR_OK;
}

/* SingleRay::trace_individual as ffi::single_ray returns it (src/ffi.rs:42-47):
 * AoS rows (t,x,y,kx,ky).  Returns number of rows written (<= cap) or <0. */
int64_t orc_single_ray(const mr_bathymetry_desc *b, const mr_current_desc *cu,
                       double x0, double y0, double kx0, double ky0,
                       double t0, double t_end, double dt, double *out, int64_t cap)
{
    int64_t nsteps = orc_num_steps(t0, t_end, dt);
    if (nsteps < 0) return MR_ERR_BAD_ARG;
    double y[4] = { x0, y0, kx0, ky0 };
    double t = t0;
    int64_t rows = 0;
    if (rows < cap) { out[0] = t; memcpy(out + 1, y, sizeof y); }
    rows = 1;
    for (int64_t s = 1; s <= nsteps; ++s) {
        double yn[4], k0[4];
        orc_rk4_step(b, cu, dt, y, yn, k0);
        t = t + dt;
        memcpy(y, yn, sizeof y);
        if (rows < cap) { out[5 * rows] = t; memcpy(out + 5 * rows + 1, y, sizeof y); }
        rows++;
        if (solout(y, k0)) break;
    }
    return rows;
}
