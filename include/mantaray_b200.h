/*
 * mantaray_b200.h — C ABI of the B200-native batch ray tracer.
 *
 * This is the drop-in boundary for ONE path of mines-oceanography/mantaray:
 * the per-ray RK4 integration of the wave ray equations over gridded
 * bathymetry and surface currents.  Citations are file:line into the
 * reference tree.
 *
 *   reference                                   replaced by
 *   ------------------------------------------  ---------------------------
 *   ffi::ray_tracing        src/ffi.rs:51-85     mr_fields_open_netcdf3 + mr_trace_many
 *   ffi::single_ray         src/ffi.rs:25-49     mr_fields_open_netcdf3 + mr_single_ray
 *   ManyRays::trace_many    src/ray.rs:98-127    mr_trace_many
 *   SingleRay::trace_individual src/ray.rs:198-213  mr_single_ray
 *   CartesianNetcdf3::open  src/bathymetry/cartesian_netcdf3.rs:167-256   mr_fields_open_netcdf3
 *   CartesianCurrent::open  src/current/cartesian_current.rs:58-213       mr_fields_open_netcdf3
 *   trait BathymetryData    src/bathymetry/mod.rs:35-41   mr_bathymetry_desc (kind tag)
 *   trait CurrentData       src/current/mod.rs:21-31      mr_current_desc   (kind tag)
 *   (commented extern "C" stub  src/ffi.rs:87-107; cbindgen.toml; Cargo.toml:45-50 `capi`)
 *
 * The reference's live FFI is PyO3; its C ABI is a commented-out stub.  This
 * header is what that stub would have grown into.  Plain C: pointers, sizes,
 * POD structs; no C++/CUDA/torch types.
 *
 * Conventions
 *   - every function returns an int status (MR_OK == 0, errors < 0); the text
 *     of the last error on the calling thread is at mr_last_error().
 *   - numerical conditions are never errors: as in the reference
 *     (src/wave_ray_path.rs:220-234) a failing right-hand side becomes four
 *     NaNs, the NaN row is stored and the ray stops (solout, :236-246).
 *   - the caller owns every output buffer; the library never frees caller
 *     memory.  Field handles are opaque and freed with mr_fields_free.
 *   - there is NO CPU fallback.  Without a CUDA device every compute entry
 *     point fails with MR_ERR_CUDA.
 */
#ifndef MANTARAY_B200_H
#define MANTARAY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MR_ABI_VERSION 1

/* ---- status codes -------------------------------------------------------- */
#define MR_OK            0
#define MR_ERR_IO       -1   /* file cannot be opened / read                        */
#define MR_ERR_BAD_ARG  -2   /* NULL pointer, bad size, bad kind, dt<=0, ...        */
#define MR_ERR_CUDA     -3   /* no device, launch/copy failure                      */
#define MR_ERR_OOM      -4   /* host or device allocation failed                    */
#define MR_ERR_FORMAT   -5   /* not a NetCDF-3 file / variable missing / bad shape  */

/* ---- field descriptors --------------------------------------------------- */

/* Implementors of BathymetryData (src/bathymetry/mod.rs:35-41). */
#define MR_BATHY_CONSTANT 0  /* ConstantDepth  src/bathymetry/constant_depth.rs:26-46 */
#define MR_BATHY_SLOPE    1  /* ConstantSlope  src/bathymetry/constant_slope.rs:52-76 */
#define MR_BATHY_GRID     2  /* CartesianNetcdf3 src/bathymetry/cartesian_netcdf3.rs:35-43 */
#define MR_BATHY_ARRAY    3  /* ArrayDepth     src/bathymetry/array_depth.rs:9-36 (test aid) */

/* Implementors of CurrentData (src/current/mod.rs:21-31). */
#define MR_CURRENT_CONSTANT 0 /* ConstantCurrent  src/current/constant_current.rs:51-77 */
#define MR_CURRENT_GRID     1 /* CartesianCurrent src/current/cartesian_current.rs:18-27 */

/*
 * Bathymetry.  All pointers are HOST pointers; mr_fields_create copies what it
 * needs, so they may be released as soon as it returns.
 *   CONSTANT: depth h0 everywhere.
 *   SLOPE   : h = h0 + dhdx*(x-x0) + dhdy*(y-y0), evaluated in float.
 *   GRID    : x[nx], y[ny] float coordinates (equally spaced, ascending),
 *             depth[ny*nx] double, row-major with x fastest: depth[nx*yi+xi]
 *             (src/bathymetry/cartesian_netcdf3.rs:465-471).  nx,ny >= 2.
 *   ARRAY   : array[nx*ny] float, array[xi*ny + yi] is the reference's
 *             array[xi][yi]; the reference bounds-checks BOTH indices against
 *             the outer length nx (src/bathymetry/array_depth.rs:30), so
 *             ny >= nx is required.
 */
typedef struct mr_bathymetry_desc {
    int32_t kind;
    int32_t nx, ny;
    const float  *x, *y;
    const double *depth;
    const float  *array;
    float h0, x0, y0, dhdx, dhdy;
} mr_bathymetry_desc;

/*
 * Surface current.
 *   CONSTANT: (u0, v0) everywhere, zero gradients.
 *   GRID    : x[nx], y[ny], u[ny*nx], v[ny*nx], all double, u[nx*yi+xi]
 *             (src/current/cartesian_current.rs:421-427).  nx,ny >= 2.
 */
typedef struct mr_current_desc {
    int32_t kind;
    int32_t nx, ny;
    const double *x, *y, *u, *v;
    double u0, v0;
} mr_current_desc;

/* Opaque: the two fields, resident on every device selected at creation. */
typedef struct mr_fields mr_fields;

/* ---- options -------------------------------------------------------------- */

/* Arithmetic of the f64 part of the right-hand side.  The f32 part (position
 * rounding, fractional index, bilinear interpolation) is value-identical to
 * the reference in both modes.
 *   MR_MATH_FAST  : restructured f64 math (one exponential per evaluation,
 *                   direction cosines from kx/k, reciprocal-multiply), FMA on.
 *   MR_MATH_STRICT: the reference's expression tree operation by operation
 *                   (atan2/sin/cos, tanh/sinh/cosh, every divide), FMA off.   */
#define MR_MATH_FAST   0
#define MR_MATH_STRICT 1

typedef struct mr_trace_opts {
    int32_t stride;       /* store every stride-th row (row j = step j*stride); 1 = reference; 0 -> 1 */
    int32_t math;         /* MR_MATH_FAST (default) or MR_MATH_STRICT                                 */
    int32_t chunk_rays;   /* host path: rays per device slab (0 = automatic)                          */
    int32_t flags;        /* MR_OPT_* bits; 0 = defaults                                              */
} mr_trace_opts;

/* Depth-floor map.  MR_MATH_FAST on affine gridded bathymetry skips the depth lookup wherever a per-block
 * lower bound of the depth already proves kh >= 22, where the reference's own formulas no longer depend on
 * h.  Same rows / len; values equal to the path without the map up to the sign of an
 * exact zero.  By default (flags = 0) the library uses the map when at least a quarter of the grid's blocks
 * are deep for a 10 s wave (mr_depth_floor_map's *deep_frac >= 0.25); on shallower grids it only adds a
 * dependent load in front of every depth lookup.
 *   MR_OPT_DEEP_MAP    : use the map whenever the grid has one, whatever its deep share
 *   MR_OPT_NO_DEEP_MAP : never use it (wins over MR_OPT_DEEP_MAP)                                     */
#define MR_OPT_DEEP_MAP    1
#define MR_OPT_NO_DEEP_MAP 2
/* Same-grid shortcut.  When the current is given on the bathymetry's own grid (same shape, same coordinates)
 * MR_MATH_FAST can derive the current's cell from the bathymetry's f32 fractional index wherever that index is
 * further from a grid line than the two indices of the reference (f32 for the bathymetry, f64 for the current) can
 * disagree, and evaluate both lookups separately elsewhere: the same cells, hence the same looked-up values, with
 * one index instead of two.  By default the library uses it where the grids coincide and NEITHER map
 * (depth-floor, uniform-current) is in use: measured there it is never slower (C4, C2 without maps: +0.3 %) and up to
 * 8 % faster (C3 without maps -8 %, C5 -3 %); beside a map it costs (C4 with the depth-floor map +10 %), because the
 * maps already skip the work it would share.  Ignored where the grids differ.
 *   MR_OPT_SAME_GRID    : use the shortcut wherever the two grids coincide, maps or not
 *   MR_OPT_NO_SAME_GRID : never use it (wins)                                                          */
#define MR_OPT_SAME_GRID 4
#define MR_OPT_NO_SAME_GRID 32
/* Uniform-current map.  The API always takes a current file, so "no current" is a grid of zeros; step currents are
 * piecewise constant.  Where a block of 8 x 8 cells of an affine current grid holds one u and one v, the reference's
 * bilinear returns exactly that value and its finite differences are exactly 0, so MR_MATH_FAST reads 8 bytes of a
 * small map there instead of the 64-byte cell record and skips the interpolation: the same values.
 * By default the library uses the map when at least half of the grid's blocks are uniform.
 *   MR_OPT_CURRENT_MAP    : use it whenever the grid has one
 *   MR_OPT_NO_CURRENT_MAP : never use it (wins)                                                       */
#define MR_OPT_CURRENT_MAP    8
#define MR_OPT_NO_CURRENT_MAP 16

/* ---- library ------------------------------------------------------------- */

int         mr_abi_version(void);
/* Number of usable CUDA devices (0 when there is none; never an error). */
int         mr_device_count(void);
/* Last error message of the calling thread ("" if none). Never NULL. */
const char *mr_last_error(void);

/* ---- fields -------------------------------------------------------------- */

/* Upload both fields to every device in device_mask (bit i = CUDA device i;
 * 0 means "device 0").  Grids are replicated per device. */
int  mr_fields_create(const mr_bathymetry_desc *bathy, const mr_current_desc *current,
                      uint32_t device_mask, mr_fields **out);

/* What src/ffi.rs:36-38 / :62-64 does: open two NetCDF-3 files, variables
 * "x","y","depth" and "x","y","u","v".  Same dtype handling as the reference:
 * any of {i8,u8,i16,i32,f32,f64}; bathymetry coordinates become float, depth
 * double; current everything double; the dimension order recorded in the file
 * is ignored and the flat variable is indexed [y][x].  Either path may be NULL
 * to get the reference defaults (ConstantDepth 2000 m, src/bathymetry/
 * constant_depth.rs:9; ConstantCurrent (0,0), src/current/constant_current.rs:10). */
int  mr_fields_open_netcdf3(const char *bathymetry_path, const char *current_path,
                            uint32_t device_mask, mr_fields **out);

void mr_fields_free(mr_fields *f);

/* Devices the handle lives on, as a bit mask. */
uint32_t mr_fields_device_mask(const mr_fields *f);

/* How the last host-buffer call on this handle (mr_trace_many, mr_trace_many_env)
 * was shared out: rays_per_device[g] receives the number of rays the g-th device
 * of the mask (ascending device number) took from the slab queue.  Returns the
 * number of devices in the handle (entries beyond `cap` are not written), or
 * MR_ERR_BAD_ARG. */
int  mr_fields_last_split(mr_fields *f, int64_t *rays_per_device, int32_t cap);

/* Which specialisations of the kernel a trace with these options runs on this handle (opts NULL = defaults): a
 * mask of MR_PLAN_* bits, or a negative MR_ERR_*.  The choice never changes rows / len and keeps every value within
 * the tolerances stated with the MR_OPT_* flags; this is for reports and tests. */
#define MR_PLAN_AFFINE      1   /* coordinates exactly affine in f32: index and corner arithmetic from launch constants */
#define MR_PLAN_DEEP_MAP    2   /* depth-floor map    */
#define MR_PLAN_SAME_GRID   4   /* same-grid shortcut */
#define MR_PLAN_CURRENT_MAP 8   /* uniform-current map */
int  mr_trace_plan(const mr_fields *f, const mr_trace_opts *opts);

/* The host-buffer entry points (mr_trace_many, mr_trace_many_env, mr_single_ray)
 * keep their device work buffers in the handle so that the next call does not
 * allocate again: up to two slabs of <= 16 GB per device.  This gives them back
 * to the device now; mr_fields_free does so too. */
void mr_fields_trim(mr_fields *f);

/* ---- sizes --------------------------------------------------------------- */

/* Number of RK4 steps the stepper takes: ceil((t_end - t0)/dt)
 * (ode_solvers 0.4.0 Rk4::integrate; called from src/ray.rs:207-208).
 * A negative quotient (t_end < t0) is 0 steps: the reference's
 * `((x_end - x) / h).ceil() as usize` saturates it, and the ray is its initial
 * row only.  Returns -1 if dt is not > 0, or the quotient is NaN, infinite or
 * >= 2^31 - 2 (the reference panics or never returns on those). */
int64_t mr_num_steps(double t0, double t_end, double dt);
/* Rows a trajectory buffer must hold: num_steps/stride + 1. */
int64_t mr_num_rows(double t0, double t_end, double dt, int32_t stride);

/* ---- tracing: host buffers ----------------------------------------------- */

/*
 * ManyRays::trace_many (src/ray.rs:98-127) for n rays, initial states
 * (x0[i], y0[i], kx0[i], ky0[i]).  With several devices in the handle the batch
 * is one queue of contiguous slabs of rays and every device takes the next slab
 * when it has a free buffer — the dynamic balance rayon's par_iter gives the
 * reference (src/ray.rs:105-123), so rays that stop early do not leave a device
 * idle.  There is no cross-ray communication; each slab is copied into its
 * column block of the outputs.
 *
 * Outputs (all HOST memory, caller-owned, any of them may be NULL):
 *   t [rows_cap]            time of row j: t0, then accumulated += dt*stride steps
 *   x,y,kx,ky [rows_cap][n] step-major, ray fastest: x[j*n + i].  Rows a ray
 *                           never reached are NaN, as mantaray.core.ray_tracing
 *                           pads them (python/mantaray/core.py:115-119).
 *   rows [n]                rows the reference would have stored for ray i
 *                           (1 + executed steps, including the trailing all-NaN
 *                           row that stops the ray), counted at stride 1.
 *   len  [n]                leading rows without any NaN (the rule of
 *                           RayResult::from, src/ray_result.rs:123-151), stride 1.
 *   final_state [4][n]      x,y,kx,ky of the last NaN-free row.
 * rows_cap = mr_num_rows(t0,t_end,dt,stride).
 */
int  mr_trace_many(mr_fields *f, int64_t n,
                   const double *x0, const double *y0, const double *kx0, const double *ky0,
                   double t0, double t_end, double dt, const mr_trace_opts *opts,
                   double *t, double *x, double *y, double *kx, double *ky,
                   int32_t *rows, int32_t *len, double *final_state);

/*
 * SingleRay::trace_individual (src/ray.rs:198-213) as ffi::single_ray returns
 * it (src/ffi.rs:42-47): rows of (t, x, y, kx, ky), array-of-structs,
 * out[5*j + c].  out holds out_cap rows; *n_rows receives the number of rows
 * the ray produced (including the trailing NaN row).  If out_cap is too small
 * the call fails with MR_ERR_BAD_ARG and *n_rows holds the required size.
 */
int  mr_single_ray(mr_fields *f, double x0, double y0, double kx0, double ky0,
                   double t0, double t_end, double dt, const mr_trace_opts *opts,
                   double *out, int64_t out_cap, int64_t *n_rows);

/* ---- tracing: device-resident -------------------------------------------- */

/*
 * The same integration with every buffer already in device memory of CUDA
 * device `device` (which must be in the handle's mask), enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = the legacy default stream) and NOT
 * synchronised: the call returns as soon as the work is queued.
 *   d_x..d_ky  [rows_cap][ld] with ld >= n the row pitch in elements
 *              (may all be NULL: "len and final state only");
 *   d_rows, d_len [n]; d_final [4][n]  (each may be NULL).
 * *launches, if not NULL, receives the number of kernels enqueued.
 */
int  mr_trace_device(mr_fields *f, int device, void *stream, int64_t n,
                     const double *d_x0, const double *d_y0, const double *d_kx0, const double *d_ky0,
                     double t0, double t_end, double dt, const mr_trace_opts *opts,
                     double *d_x, double *d_y, double *d_kx, double *d_ky, int64_t ld,
                     int32_t *d_rows, int32_t *d_len, double *d_final, int32_t *launches);

/* ---- environment along rays ------------------------------------------------ */

/*
 * The reference sketches a per-ray record Ray{time, state, depth: Vec<f32>,
 * current: Vec<Current<T>>} (src/datatype.rs:165-194) that nothing fills yet.
 * These entry points produce its depth and current columns: the fields
 * evaluated at points, by the reference's own accessors
 *   depth = BathymetryData::depth(&Point<f32>)   src/bathymetry/mod.rs:38
 *           (the point is (x as f32, y as f32), src/wave_ray_path.rs:122)
 *   u, v  = CurrentData::current(&Point<f64>)    src/current/mod.rs:24
 * operation by operation; where the accessor returns Err (outside the grid)
 * the value is NaN.  ConstantCurrent ignores the point, NaN or not
 * (src/current/constant_current.rs:51-53), and so does this.
 */
typedef struct mr_env_planes {
    float  *depth;             /* [rows_cap][n] or NULL */
    double *u, *v;             /* [rows_cap][n] or NULL */
} mr_env_planes;

/* mr_trace_many plus the environment at every stored row (same layout as x:
 * depth[j*n + i]).  Needs the x, y, kx, ky planes.  env == NULL, or all three
 * planes NULL, is mr_trace_many. */
int  mr_trace_many_env(mr_fields *f, int64_t n,
                       const double *x0, const double *y0, const double *kx0, const double *ky0,
                       double t0, double t_end, double dt, const mr_trace_opts *opts,
                       double *t, double *x, double *y, double *kx, double *ky,
                       int32_t *rows, int32_t *len, double *final_state,
                       const mr_env_planes *env);

/* The fields at `count` arbitrary points (HOST arrays; outputs may be NULL).
 * Runs on the first device of the handle. */
int  mr_sample_fields(mr_fields *f, int64_t count, const double *x, const double *y,
                      float *depth, double *u, double *v);

/* Device-resident: points laid out [rows][ld] (n <= ld used per row), e.g. the
 * d_x, d_y planes of mr_trace_device.  Asynchronous on `stream`. */
int  mr_sample_device(mr_fields *f, int device, void *stream, int64_t rows, int64_t n, int64_t ld,
                      const double *d_x, const double *d_y,
                      float *d_depth, double *d_u, double *d_v, int32_t *launches);

/* ---- NetCDF-3 ingest -------------------------------------------------------- */

/*
 * The reader behind mr_fields_open_netcdf3, exposed so the host-side mirrors of
 * CartesianNetcdf3::open / CartesianCurrent::open can be tested and reused.
 * Classic (CDF-1) and 64-bit-offset (CDF-2) files, fixed and record variables.
 * nc_type codes: 1 byte(i8) 2 char(u8) 3 short(i16) 4 int(i32) 5 float 6 double.
 * Works without a CUDA device.
 */
typedef struct mr_nc3 mr_nc3;
#define MR_NC3_MAX_DIMS 8
int  mr_nc3_open(const char *path, mr_nc3 **out);
void mr_nc3_close(mr_nc3 *f);
int  mr_nc3_var_count(const mr_nc3 *f);
/* Name of variable `index` copied into buf (NUL-terminated, truncated to cap). */
int  mr_nc3_var_name(const mr_nc3 *f, int index, char *buf, size_t cap);
/* Type, element count and shape of a variable; MR_ERR_FORMAT if absent. */
int  mr_nc3_var_info(const mr_nc3 *f, const char *name, int32_t *nc_type, int64_t *n_elems,
                     int32_t *ndims, int64_t dims[MR_NC3_MAX_DIMS]);
/* Whole variable, flat, cast the way the reference's dtype switch casts it
 * (`as f32` / `as f64`).  cap = capacity of out in elements. */
int  mr_nc3_read_f32(const mr_nc3 *f, const char *name, float *out, int64_t cap);
int  mr_nc3_read_f64(const mr_nc3 *f, const char *name, double *out, int64_t cap);

/* ---- pinned host memory ---------------------------------------------------- */

/* Page-locked host allocations for trajectory buffers, so the device-to-host
 * gather of mr_trace_many runs at full PCIe rate and overlaps the kernels. */
int  mr_host_alloc(size_t bytes, void **out);
void mr_host_free(void *p);

/* ---- inspection ------------------------------------------------------------ */

/* The depth-floor map MR_OPT_DEEP_MAP consults, as the library builds it for a GRID bathymetry at upload:
 * one float per block of 8 x 8 cells, row-major [*nby][*nbx], holding the square (rounded down) of a lower
 * bound of every depth the reference's f32 bilinear lookup (src/bathymetry/cartesian_netcdf3.rs:98-135) can
 * return inside the block, or 0 where there is no such bound (a NaN, infinite or non-positive node).
 * `out` may be NULL to query the shape; `cap` is its capacity in floats.  *deep_frac (may be NULL): the share of
 * blocks whose bound is at least 550 m.  *affine (may be NULL): 1 if the grid's f32 coordinates are exactly affine,
 * which is when the fast path uses launch-constant cell geometry and the map at all (on other grids the reference's
 * lookup may extrapolate beyond a cell's corners, and the bound does not hold).  Host only: works without a
 * CUDA device. */
int  mr_depth_floor_map(const mr_bathymetry_desc *bathy, float *out, size_t cap,
                        int32_t *nbx, int32_t *nby, float *deep_frac, int32_t *affine);

/* The uniform-current map MR_OPT_CURRENT_MAP consults, as the library builds it for a GRID current at upload: two
 * floats {u, v} per block of 8 x 8 cells, row-major [*nby][*nbx][2] — the value of EVERY node the block's cells touch
 * where they all hold the same finite u and the same finite v (compared as doubles), NaNs elsewhere.  `out` may be
 * NULL to query the shape; `cap` is its capacity in floats.  *uniform_frac (may be NULL): the share of uniform
 * blocks.  *affine (may be NULL): 1 if the grid's coordinates, cast to f32, are exactly affine, which is when the
 * fast path uses the map at all.  Host only: works without a CUDA device. */
int  mr_uniform_current_map(const mr_current_desc *current, float *out, size_t cap,
                            int32_t *nbx, int32_t *nby, float *uniform_frac, int32_t *affine);

#ifdef __cplusplus
}
#endif
#endif /* MANTARAY_B200_H */
