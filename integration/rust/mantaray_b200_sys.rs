//! Rust binding of `include/mantaray_b200.h` and a safe wrapper with the shape of
//! `ManyRays::new(..).trace_many(..)` (src/ray.rs:54-64, 98-103 of mantaray).
//!
//! SOURCE ONLY: the build image of this repository has no cargo/rustc, so this file has never
//! been compiled.  It is what a mantaray maintainer would drop into `src/` (behind the existing
//! `capi` feature of Cargo.toml:49-50) to route `ffi::ray_tracing` through the B200 library.
//! `build.rs` would add `println!("cargo:rustc-link-lib=dylib=mantaray_b200");`.

#![allow(non_camel_case_types, dead_code)]

use std::ffi::{c_char, c_void, CStr};
use std::os::raw::c_int;

pub const MR_OK: c_int = 0;
pub const MR_BATHY_CONSTANT: i32 = 0;
pub const MR_BATHY_SLOPE: i32 = 1;
pub const MR_BATHY_GRID: i32 = 2;
pub const MR_BATHY_ARRAY: i32 = 3;
pub const MR_CURRENT_CONSTANT: i32 = 0;
pub const MR_CURRENT_GRID: i32 = 1;

#[repr(C)]
pub struct mr_bathymetry_desc {
    pub kind: i32,
    pub nx: i32,
    pub ny: i32,
    pub x: *const f32,
    pub y: *const f32,
    pub depth: *const f64,
    pub array: *const f32,
    pub h0: f32,
    pub x0: f32,
    pub y0: f32,
    pub dhdx: f32,
    pub dhdy: f32,
}

#[repr(C)]
pub struct mr_current_desc {
    pub kind: i32,
    pub nx: i32,
    pub ny: i32,
    pub x: *const f64,
    pub y: *const f64,
    pub u: *const f64,
    pub v: *const f64,
    pub u0: f64,
    pub v0: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct mr_trace_opts {
    pub stride: i32,
    pub math: i32,
    pub chunk_rays: i32,
    pub flags: i32, // MR_OPT_* bits, 0 = defaults
}

pub const MR_OPT_DEEP_MAP: i32 = 1;
pub const MR_OPT_NO_DEEP_MAP: i32 = 2;
pub const MR_OPT_SAME_GRID: i32 = 4; // one fractional index for both fields on a shared grid, also beside a map
pub const MR_OPT_NO_SAME_GRID: i32 = 32; // never (the default uses it where the grids coincide and no map is in use)
pub const MR_OPT_CURRENT_MAP: i32 = 8;
pub const MR_OPT_NO_CURRENT_MAP: i32 = 16;

#[repr(C)]
pub struct mr_fields {
    _private: [u8; 0],
}

extern "C" {
    pub fn mr_last_error() -> *const c_char;
    pub fn mr_fields_create(b: *const mr_bathymetry_desc, c: *const mr_current_desc, device_mask: u32,
                            out: *mut *mut mr_fields) -> c_int;
    pub fn mr_fields_open_netcdf3(bathy: *const c_char, current: *const c_char, device_mask: u32,
                                  out: *mut *mut mr_fields) -> c_int;
    pub fn mr_fields_free(f: *mut mr_fields);
    /// rays each device of the handle took from the slab queue in the last host-buffer call; returns the device count
    pub fn mr_fields_last_split(f: *mut mr_fields, rays_per_device: *mut i64, cap: i32) -> i32;
    /// MR_PLAN_* bits (1 affine, 2 depth-floor map, 4 same-grid shortcut, 8 uniform-current map) of a trace with `opts`
    pub fn mr_trace_plan(f: *const mr_fields, opts: *const mr_trace_opts) -> c_int;
    pub fn mr_num_rows(t0: f64, t_end: f64, dt: f64, stride: i32) -> i64;
    pub fn mr_trace_many(f: *mut mr_fields, n: i64,
                         x0: *const f64, y0: *const f64, kx0: *const f64, ky0: *const f64,
                         t0: f64, t_end: f64, dt: f64, opts: *const mr_trace_opts,
                         t: *mut f64, x: *mut f64, y: *mut f64, kx: *mut f64, ky: *mut f64,
                         rows: *mut i32, len: *mut i32, final_state: *mut f64) -> c_int;
}

/// What the field traits lower to.  `BathymetryData` / `CurrentData` implementors
/// (src/bathymetry/mod.rs:35-41, src/current/mod.rs:21-31) gain one method each:
/// `fn descriptor(&self) -> mr_bathymetry_desc` — the four bathymetry kinds and two current
/// kinds map one to one onto the `kind` tags.
pub trait ToBathymetryDesc { fn descriptor(&self) -> mr_bathymetry_desc; }
pub trait ToCurrentDesc { fn descriptor(&self) -> mr_current_desc; }

#[derive(Debug)]
pub struct B200Error(pub c_int, pub String);

fn check(rc: c_int) -> Result<(), B200Error> {
    if rc == MR_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(mr_last_error()) }.to_string_lossy().into_owned();
    Err(B200Error(rc, msg))
}

/// Step-major structure-of-arrays result of one batch (`[rows_cap][n]`, ray fastest).
pub struct RayBatch {
    pub n: usize,
    pub rows_cap: usize,
    pub t: Vec<f64>,
    pub x: Vec<f64>,
    pub y: Vec<f64>,
    pub kx: Vec<f64>,
    pub ky: Vec<f64>,
    /// rows the reference stepper stores for each ray (incl. the trailing NaN row)
    pub rows: Vec<i32>,
}

impl RayBatch {
    /// The reference's `Vec<Option<SolverResult>>` element for ray `i` as `(t, [x,y,kx,ky])` rows —
    /// what `ffi::ray_tracing` re-packs into tuples (src/ffi.rs:73-83).
    pub fn ray(&self, i: usize) -> Vec<(f64, f64, f64, f64, f64)> {
        (0..self.rows[i] as usize)
            .map(|r| { let o = r * self.n + i; (self.t[r], self.x[o], self.y[o], self.kx[o], self.ky[o]) })
            .collect()
    }
}

/// Drop-in for `ManyRays` (src/ray.rs:24-127): same constructor arguments, same `trace_many`.
pub struct ManyRaysB200 {
    fields: *mut mr_fields,
    x0: Vec<f64>, y0: Vec<f64>, kx0: Vec<f64>, ky0: Vec<f64>,
}

impl ManyRaysB200 {
    /// `initial_rays` as (x, y, kx, ky), the order of `From<RayState<f64>> for State` (src/datatype.rs:140-150).
    pub fn new(bathymetry: &dyn ToBathymetryDesc, current: &dyn ToCurrentDesc,
               initial_rays: &[(f64, f64, f64, f64)], device_mask: u32) -> Result<Self, B200Error> {
        let (b, c) = (bathymetry.descriptor(), current.descriptor());
        let mut fields = std::ptr::null_mut();
        check(unsafe { mr_fields_create(&b, &c, device_mask, &mut fields) })?;
        Ok(ManyRaysB200 {
            fields,
            x0: initial_rays.iter().map(|r| r.0).collect(),
            y0: initial_rays.iter().map(|r| r.1).collect(),
            kx0: initial_rays.iter().map(|r| r.2).collect(),
            ky0: initial_rays.iter().map(|r| r.3).collect(),
        })
    }

    pub fn trace_many(&self, start_time: f64, end_time: f64, step_size: f64) -> Result<RayBatch, B200Error> {
        let n = self.x0.len();
        let rows_cap = unsafe { mr_num_rows(start_time, end_time, step_size, 1) };
        if rows_cap < 0 { return Err(B200Error(-2, "bad time arguments".into())); }
        let rows_cap = rows_cap as usize;
        let mut out = RayBatch {
            n, rows_cap,
            t: vec![0.0; rows_cap],
            x: vec![0.0; rows_cap * n], y: vec![0.0; rows_cap * n],
            kx: vec![0.0; rows_cap * n], ky: vec![0.0; rows_cap * n],
            rows: vec![0; n],
        };
        check(unsafe {
            mr_trace_many(self.fields, n as i64, self.x0.as_ptr(), self.y0.as_ptr(), self.kx0.as_ptr(), self.ky0.as_ptr(),
                          start_time, end_time, step_size, std::ptr::null(),
                          out.t.as_mut_ptr(), out.x.as_mut_ptr(), out.y.as_mut_ptr(), out.kx.as_mut_ptr(), out.ky.as_mut_ptr(),
                          out.rows.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut())
        })?;
        Ok(out)
    }
}

impl Drop for ManyRaysB200 {
    fn drop(&mut self) { unsafe { mr_fields_free(self.fields) } }
}

// ---- what src/ffi.rs:51-85 becomes --------------------------------------------------------------
//
// #[pyfunction]
// fn ray_tracing(x0: Vec<f64>, y0: Vec<f64>, kx0: Vec<f64>, ky0: Vec<f64>, duration: f64, step_size: f64,
//                bathymetry_filename: String, current_filename: String)
//     -> PyResult<Vec<Vec<(f64, f64, f64, f64, f64)>>> {
//     let (b, c) = (CString::new(bathymetry_filename)?, CString::new(current_filename)?);
//     let mut fields = std::ptr::null_mut();
//     check(unsafe { mr_fields_open_netcdf3(b.as_ptr(), c.as_ptr(), 0, &mut fields) })   // "x","y","depth" / "x","y","u","v"
//         .expect("could not open bathymetry file");                                        // same panic text as ffi.rs:37
//     ... mr_trace_many(fields, n, ..., 0.0, duration, step_size, ...) ...                 // t0 = 0, ffi.rs:72
//     Ok((0..n).map(|i| batch.ray(i)).collect())
// }
