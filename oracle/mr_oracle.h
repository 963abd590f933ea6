/*
 * mr_oracle.h — CPU ORACLE (test infrastructure, not product code).
 * See mr_oracle.c.  Uses the field descriptor structs of
 * include/mantaray_b200.h so tests build one descriptor for both sides.
 */
#ifndef MR_ORACLE_H
#define MR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/mantaray_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* src/interpolator.rs:39-84 */
int  orc_bilinear(const float pts[4][3], float tx, float ty, float *out);

/* src/bathymetry/cartesian_netcdf3.rs:274-296, 334-390 */
int  orc_bathy_nearest(float target, const float *arr, int n, float *index);
int  orc_bathy_four_corners(const mr_bathymetry_desc *b, float x, float y, size_t c[4][2]);
/* BathymetryData (src/bathymetry/mod.rs:35-41), all kinds */
int  orc_depth(const mr_bathymetry_desc *b, float x, float y, float *h);
int  orc_depth_and_gradient(const mr_bathymetry_desc *b, float x, float y, float *h, float *dhdx, float *dhdy);

/* src/current/cartesian_current.rs:231-253, 283-339, 487-542 */
int  orc_current_nearest(double target, const double *arr, int n, double *index);
int  orc_current_four_corners(const mr_current_desc *cu, double x, double y, size_t c[4][2]);
int  orc_current(const mr_current_desc *cu, double x, double y, double *u, double *v);
void orc_sample_fields(const mr_bathymetry_desc *b, const mr_current_desc *cu, int64_t count,
                       const double *x, const double *y, float *depth, double *u, double *v);
int  orc_current_and_gradient(const mr_current_desc *cu, double x, double y, double *u, double *v, double grad[4]);

/* src/wave_ray_path.rs:118-247 */
int  orc_group_velocity(double k, double h, double *cg);
void orc_dkdt_bathy(double k, double h, double dhdx, double dhdy, double *ox, double *oy);
int  orc_odes(const mr_bathymetry_desc *b, const mr_current_desc *cu, double x, double y, double kx, double ky, double out[4]);
void orc_system(const mr_bathymetry_desc *b, const mr_current_desc *cu, const double s[4], double ds[4]);

/* ode_solvers 0.4.0 Rk4 restated; call site src/ray.rs:205-212 */
int64_t orc_num_steps(double t0, double t_end, double dt);
void orc_rk4_step(const mr_bathymetry_desc *b, const mr_current_desc *cu, double dt, const double y[4], double ynew[4], double k0[4]);

/* src/ray.rs:98-127 / src/ffi.rs:51-85 */
int  orc_trace_many(const mr_bathymetry_desc *b, const mr_current_desc *cu, int64_t n,
                    const double *x0, const double *y0, const double *kx0, const double *ky0,
                    double t0, double t_end, double dt, int32_t stride, int32_t nthreads,
                    double *t, double *x, double *y, double *kx, double *ky,
                    int32_t *rows, int32_t *len, double *final_state);
/* src/ray.rs:198-213 / src/ffi.rs:25-49 */
int64_t orc_single_ray(const mr_bathymetry_desc *b, const mr_current_desc *cu,
                       double x0, double y0, double kx0, double ky0,
                       double t0, double t_end, double dt, double *out, int64_t cap);
#ifdef __cplusplus
}
#endif
#endif
