// mr_capi.cu — the C ABI of include/mantaray_b200.h.
//
// Host side of the batch driver: what ffi.rs (src/ffi.rs:25-85) and
// ManyRays/SingleRay (src/ray.rs:24-214) do around the integration — open the
// two field files, build the ray states, run, hand the rows back — with the
// integration itself being the CUDA kernel of mr_trace_kernel.cuh.
//
// Multi-GPU: rays are independent (src/ray.rs:112-123 maps them independently),
// so a handle replicates the field grids on every selected device and
// mr_trace_many gives each device one contiguous block of rays, driven by its
// own host thread and streams.  No collective is involved; the "gather" is each
// device copying its column block of the step-major [rows][n] host arrays.
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>

#include "mr_internal.hpp"
#include "mr_launch.hpp"
#include "mr_trace_kernel.cuh"

namespace mr {

// ---- errors ---------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char *what)
{
    int code = (e == cudaErrorMemoryAllocation) ? MR_ERR_OOM : MR_ERR_CUDA;
    return fail(code, std::string(what) + ": " + cudaGetErrorString(e));
}
#define MR_CUDA(call)                                                 \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);         \
    } while (0)

// ---- DFMA probe -------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(double *sink, int iters)
{
    // 8 independent chains per thread keep the FP64 pipe full at any occupancy
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999999, c = 1e-12;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s;     // never true; keeps the chains alive
}
cudaError_t launch_dfma_probe(double *sink, int iters, int blocks, cudaStream_t stream)
{
    dfma_probe_kernel<<<blocks, 256, 0, stream>>>(sink, iters);
    return cudaGetLastError();
}

// ---- exhaustive check of fdiv_const ---------------------------------------------
__global__ void fdiv_selftest_kernel(float s, float r, unsigned long long *bad)
{
    unsigned long long local = 0;
    // every non-negative finite float: bit patterns 0 .. 0x7f7fffff
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= 0x7f7fffffull;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float t = __uint_as_float((unsigned)b);
        const float want = __fdiv_rn(t, s), got = fdiv_const(t, s, r);
        // Below |t| = 2^-100 the exact residual t - q*s underflows and the last bits of the
        // quotient may differ; there (s > 1e-30 is enforced) both quotients are in [0, 1): cell 0
        // and in bounds either way, which is all the caller derives from the index.
        const bool tiny = t > 0.0f && t < 7.8886090522101181e-31f;
        const bool same = (__float_as_uint(want) == __float_as_uint(got)) || (isinf(want) && !(got == got)) ||
                          (!(want == want) && !(got == got)) || (tiny && want < 1.0f && got >= 0.0f && got < 1.0f);
        local += same ? 0 : 1;
    }
    if (local) atomicAdd(bad, local);
}

// ---- field handle -----------------------------------------------------------
// Work memory of the host-buffer path (mr_trace_many), kept by the handle between calls: allocating and
// freeing two ~16 GB slabs cost 20-480 ms per call on B200 (MR_DEBUG_TIMING), against 1.26 s of drain.
struct WorkArea {
    void *arena[2] = {nullptr, nullptr};
    size_t bytes[2] = {0, 0};
    cudaStream_t s_comp = nullptr, s_copy = nullptr;
    cudaEvent_t computed[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    size_t held() const { return bytes[0] + bytes[1]; }
};

struct DeviceFields {
    int dev = -1;
    BathyDev b{};
    CurrentDev c{};
    float deep_frac = 0.0f;    // share of the depth-floor map's blocks that are deep water for a 10 s wave
    std::vector<void *> allocs;
    WorkArea work;             // guarded by mr_fields::mu
};

}  // namespace mr

struct mr_fields {
    uint32_t mask = 0;
    std::vector<mr::DeviceFields> devs;
    std::mutex mu;      // one trace at a time per handle and device set
};

struct mr_nc3 {
    mr::Nc3File file;
};

namespace mr {

static int device_count_quiet()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

template <typename T>
static int upload(DeviceFields &d, const T *host, size_t count, const T **out)
{
    void *p = nullptr;
    MR_CUDA(cudaMalloc(&p, count * sizeof(T)));
    d.allocs.push_back(p);
    MR_CUDA(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T *)p;
    return MR_OK;
}

static int validate(const mr_bathymetry_desc *b, const mr_current_desc *c)
{
    if (!b || !c) return fail(MR_ERR_BAD_ARG, "mr_fields_create: NULL descriptor");
    switch (b->kind) {
    case MR_BATHY_CONSTANT: case MR_BATHY_SLOPE: break;
    case MR_BATHY_GRID:
        if (b->nx < 2 || b->ny < 2) return fail(MR_ERR_BAD_ARG, "bathymetry grid needs nx >= 2 and ny >= 2");
        if ((int64_t)(b->nx - 1) * (b->ny - 1) > 0x7fffffffLL) return fail(MR_ERR_BAD_ARG, "bathymetry grid has more than 2^31 cells");
        if (!b->x || !b->y || !b->depth) return fail(MR_ERR_BAD_ARG, "bathymetry grid: NULL x / y / depth");
        {
            float sx = fabsf(b->x[1] - b->x[0]), sy = fabsf(b->y[1] - b->y[0]);
            if (!(sx > 0.0f) || !(sy > 0.0f) || std::isinf(sx) || std::isinf(sy))
                return fail(MR_ERR_BAD_ARG, "bathymetry grid: x[1]-x[0] and y[1]-y[0] must be finite and non-zero");
        }
        break;
    case MR_BATHY_ARRAY:
        if (b->nx < 1 || b->ny < b->nx || !b->array)
            return fail(MR_ERR_BAD_ARG, "bathymetry array needs nx >= 1, ny >= nx and a non-NULL array");
        break;
    default: return fail(MR_ERR_BAD_ARG, "unknown bathymetry kind");
    }
    switch (c->kind) {
    case MR_CURRENT_CONSTANT: break;
    case MR_CURRENT_GRID:
        if (c->nx < 2 || c->ny < 2) return fail(MR_ERR_BAD_ARG, "current grid needs nx >= 2 and ny >= 2");
        if ((int64_t)(c->nx - 1) * (c->ny - 1) > 0x3fffffffLL) return fail(MR_ERR_BAD_ARG, "current grid has more than 2^30 cells");
        if (!c->x || !c->y || !c->u || !c->v) return fail(MR_ERR_BAD_ARG, "current grid: NULL x / y / u / v");
        {
            double sx = fabs(c->x[1] - c->x[0]), sy = fabs(c->y[1] - c->y[0]);
            if (!(sx > 0.0) || !(sy > 0.0) || std::isinf(sx) || std::isinf(sy))
                return fail(MR_ERR_BAD_ARG, "current grid: x[1]-x[0] and y[1]-y[0] must be finite and non-zero");
        }
        break;
    default: return fail(MR_ERR_BAD_ARG, "unknown current kind");
    }
    return MR_OK;
}

// ---- per-cell records of the fast path ----------------------------------------------
// Built once at upload from the f64 node grids with exactly the reference's operations:
// corner values `as f32` (cartesian_netcdf3.rs:426, cartesian_current.rs:378), gradients as
// IEEE f64 quotients (cartesian_netcdf3.rs:126-134 then `as f32`; cartesian_current.rs:522-536).
__global__ void build_bathy_cells(const double *depth, int nx, int ny, double x_space, double y_space, float4 *cell)
{
    const size_t ncell = (size_t)(nx - 1) * (ny - 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (size_t)gridDim.x * blockDim.x) {
        const size_t x1 = i % (size_t)(nx - 1), y1 = i / (size_t)(nx - 1);
        const double *p = depth + (size_t)nx * y1 + x1;
        const double sw = p[0], se = p[1], nw = p[nx], ne = p[nx + 1];
        cell[2 * i] = bilinear_coeffs((float)sw, (float)nw, (float)ne, (float)se);
        // `as f32` (cartesian_netcdf3.rs:134) then `as f64` (wave_ray_path.rs:125-126)
        const double gx = (double)(float)__ddiv_rn(__dsub_rn(se, sw), x_space);
        const double gy = (double)(float)__ddiv_rn(__dsub_rn(nw, sw), y_space);
        *reinterpret_cast<double2 *>(cell + 2 * i + 1) = make_double2(gx, gy);
    }
}

__global__ void build_current_cells(const double *u, const double *v, int nx, int ny, double x_space, double y_space,
                                    float4 *cell)
{
    const size_t ncell = (size_t)(nx - 1) * (ny - 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (size_t)gridDim.x * blockDim.x) {
        const size_t x1 = i % (size_t)(nx - 1), y1 = i / (size_t)(nx - 1);
        const double *pu = u + (size_t)nx * y1 + x1, *pv = v + (size_t)nx * y1 + x1;
        const double usw = pu[0], use_ = pu[1], unw = pu[nx], une = pu[nx + 1];
        const double vsw = pv[0], vse = pv[1], vnw = pv[nx], vne = pv[nx + 1];
        const float4 cu = bilinear_coeffs((float)usw, (float)unw, (float)une, (float)use_);
        const float4 cv = bilinear_coeffs((float)vsw, (float)vnw, (float)vne, (float)vse);
        cell[4 * i] = make_float4(cu.x, cv.x, cu.y, cv.y);          // interleaved for the packed (u, v) bilinear
        cell[4 * i + 1] = make_float4(cu.z, cv.z, cu.w, cv.w);
        double2 *g = reinterpret_cast<double2 *>(cell + 4 * i + 2);
        g[0] = make_double2(__ddiv_rn(__dsub_rn(use_, usw), x_space), __ddiv_rn(__dsub_rn(unw, usw), y_space));
        g[1] = make_double2(__ddiv_rn(__dsub_rn(vse, vsw), x_space), __ddiv_rn(__dsub_rn(vnw, vsw), y_space));
    }
}

// Are the f32 coordinates exactly affine, i.e. does the kernel's own arithmetic
// (xa = fmaf(i, d, c[0]); xb = xa + d) reproduce every c[i] bit for bit, AND is every cell exactly d wide
// in f32 (the reference forms dx = x2 - x1 per cell, interpolator.rs:62-63; the affine path replaces it by
// the launch constant d)?  A non-representable origin passes the first test and fails the second: its
// coordinates round differently from binade to binade (found by tests/test_gpu_fuzz.py).
static bool affine_f32(const float *c, int n, float *d_out)
{
    const float d = c[1] - c[0];
    if (!(d > 0.0f) || std::isinf(d)) return false;
    if (n > (1 << 24)) return false;       // cell numbers are formed as floats: (float)i must be exact
    for (int i = 0; i < n; ++i) {
        const float xa = std::fmaf((float)i, d, c[0]);
        if (xa != c[i]) return false;
        if (i + 1 < n) {
            volatile float xb = xa + d;
            volatile float w = xb - xa;
            if (xb != c[i + 1] || w != d) return false;
        }
    }
    *d_out = d;
    return true;
}

// Depth-floor map of the fast path (FastRay, DMAP): for every block of kDeepBlock x kDeepBlock cells, the square
// (rounded down) of H = zmin - 1e-5 zmax over the f32 depths of all nodes the block's cells touch.  The f32
// bilinear of interpolator.rs:59-83 returns, inside a cell, a value between its corners up to a few roundings of
// terms no larger than ~2 zmax (< 1e-6 zmax), so every depth the lookup can produce in the block is >= H.
// A block with a node that is NaN, infinite or <= 0 gets 0: no bound, the kernel looks the depth up.
// *deep_frac: the share of blocks with H >= 550 m (kh >= 22 for periods up to ~10 s), what an automatic choice
// of MR_OPT_DEEP_MAP would look at.
static std::vector<float> depth_floor_map(const double *depth, int nx, int ny, int *nbx_out, int *nby_out, float *deep_frac)
{
    size_t n_deep = 0;
    const int B = kDeepBlock;
    const int nbx = (nx - 1 + B - 1) / B, nby = (ny - 1 + B - 1) / B;
    std::vector<float> out((size_t)nbx * (size_t)nby, 0.0f);
    for (int by = 0; by < nby; ++by) {
        const int j0 = by * B, j1 = std::min(j0 + B, ny - 1);          // nodes j0..j1 inclusive
        for (int bx = 0; bx < nbx; ++bx) {
            const int i0 = bx * B, i1 = std::min(i0 + B, nx - 1);
            float zmin = INFINITY, zmax = 0.0f;
            bool bad = false;
            for (int j = j0; j <= j1 && !bad; ++j)
                for (int i = i0; i <= i1; ++i) {
                    const float z = (float)depth[(size_t)j * (size_t)nx + (size_t)i];   // `as f32`, cartesian_netcdf3.rs:426
                    if (!(z > 0.0f) || std::isinf(z)) { bad = true; break; }
                    zmin = std::min(zmin, z); zmax = std::max(zmax, z);
                }
            if (bad) continue;
            const double H = (double)zmin - 1e-5 * (double)zmax;
            if (!(H > 0.0)) continue;
            const double sq = H * H;
            float v = sq >= (double)FLT_MAX ? FLT_MAX : (float)sq;
            if ((double)v > sq) v = std::nextafterf(v, 0.0f);
            out[(size_t)by * (size_t)nbx + (size_t)bx] = v;
            n_deep += H >= 550.0;
        }
    }
    *nbx_out = nbx; *nby_out = nby;
    *deep_frac = out.empty() ? 0.0f : (float)n_deep / (float)out.size();
    return out;
}

// change-of-basis coefficients of interpolator.rs:64-72 for a (dx, dy) cell, in f32
static bool basis_coeffs(float dx, float dy, float *c01, float *c10)
{
    volatile float p = dx * dy;
    volatile float det = 0.0f - p;
    if (det == 0.0f || std::isinf(det) || std::isnan(det)) return false;
    volatile float a = dx / det, b = dy / det;
    *c01 = -a;
    *c10 = -b;
    return std::isfinite(*c01) && std::isfinite(*c10);
}

// RN(1/s) for fdiv_const; refused for subnormal / huge s and for the all-ones significand
// that Markstein's theorem excludes
static bool recip_ok(float s, float *r)
{
    uint32_t bits;
    std::memcpy(&bits, &s, 4);
    if (!(s > 1e-30f) || !(s < 1e30f) || (bits & 0x7fffffu) == 0x7fffffu) return false;
    volatile float q = 1.0f / s;
    *r = q;
    return true;
}

template <typename T>
static int device_alloc(DeviceFields &d, size_t count, T **out)
{
    void *p = nullptr;
    MR_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    d.allocs.push_back(p);
    *out = (T *)p;
    return MR_OK;
}

static int upload_fields(DeviceFields &d, const mr_bathymetry_desc *b, const mr_current_desc *c)
{
    MR_CUDA(cudaSetDevice(d.dev));
    BathyDev &B = d.b;
    B.kind = b->kind; B.nx = b->nx; B.ny = b->ny;
    B.h0 = b->h0; B.x0 = b->x0; B.y0 = b->y0; B.dhdx = b->dhdx; B.dhdy = b->dhdy;
    if (b->kind == MR_BATHY_GRID) {
        int rc;
        if ((rc = upload(d, b->x, (size_t)b->nx, &B.x))) return rc;
        if ((rc = upload(d, b->y, (size_t)b->ny, &B.y))) return rc;
        if ((rc = upload(d, b->depth, (size_t)b->nx * b->ny, &B.depth))) return rc;
        B.xf0 = b->x[0]; B.yf0 = b->y[0];
        B.sx = fabsf(b->x[1] - b->x[0]);                       // cartesian_netcdf3.rs:287
        B.sy = fabsf(b->y[1] - b->y[0]);
        B.x_space = (double)b->x[1] - (double)b->x[0];         // :119
        B.y_space = (double)b->y[1] - (double)b->y[0];         // :120
        float4 *cell = nullptr;
        const size_t ncell = (size_t)(b->nx - 1) * (b->ny - 1);
        if (ncell >= ((size_t)1 << 30)) return fail(MR_ERR_BAD_ARG, "bathymetry grid has 2^30 cells or more (32-bit cell addressing)");
        if ((rc = device_alloc(d, 2 * ncell, &cell))) return rc;
        build_bathy_cells<<<(unsigned)std::min<size_t>((ncell + 255) / 256, 148 * 16), 256>>>(B.depth, b->nx, b->ny, B.x_space, B.y_space, cell);
        MR_CUDA(cudaGetLastError());
        B.cell = cell;
        B.nxm1f = (float)(b->nx - 1); B.nym1f = (float)(b->ny - 1);
        B.zero = 0;
        auto pack2 = [](float lo, float hi) {
            uint32_t a, b2;
            std::memcpy(&a, &lo, 4); std::memcpy(&b2, &hi, 4);
            return (unsigned long long)a | ((unsigned long long)b2 << 32);
        };
        B.fastdiv = recip_ok(B.sx, &B.rsx) && recip_ok(B.sy, &B.rsy);
        if (!B.fastdiv) B.rsx = B.rsy = 0.0f;
        B.uniform = B.fastdiv && affine_f32(b->x, b->nx, &B.dxf) && affine_f32(b->y, b->ny, &B.dyf) &&
                    basis_coeffs(B.dxf, B.dyf, &B.c01, &B.c10);
        B.p0 = pack2(B.xf0, B.yf0); B.rs2 = pack2(B.rsx, B.rsy); B.ns2 = pack2(-B.sx, -B.sy);
        B.d2 = pack2(B.dxf, B.dyf); B.c2 = pack2(B.c10, B.c01);
        B.dmap = nullptr; B.dmap_nbx = 0;
        if (B.uniform) {
            int nbx = 0, nby = 0;
            std::vector<float> dm = depth_floor_map(b->depth, b->nx, b->ny, &nbx, &nby, &d.deep_frac);
            if ((rc = upload(d, dm.data(), dm.size(), &B.dmap))) return rc;
            B.dmap_nbx = nbx;
        }
    } else if (b->kind == MR_BATHY_ARRAY) {
        int rc;
        if ((rc = upload(d, b->array, (size_t)b->nx * b->ny, &B.array))) return rc;
    }
    CurrentDev &C = d.c;
    C.kind = c->kind; C.nx = c->nx; C.ny = c->ny; C.u0 = c->u0; C.v0 = c->v0;
    if (c->kind == MR_CURRENT_GRID) {
        int rc;
        if ((rc = upload(d, c->x, (size_t)c->nx, &C.x))) return rc;
        if ((rc = upload(d, c->y, (size_t)c->ny, &C.y))) return rc;
        if ((rc = upload(d, c->u, (size_t)c->nx * c->ny, &C.u))) return rc;
        if ((rc = upload(d, c->v, (size_t)c->nx * c->ny, &C.v))) return rc;
        C.xd0 = c->x[0]; C.yd0 = c->y[0];
        C.sx = fabs(c->x[1] - c->x[0]);                        // cartesian_current.rs:244
        C.sy = fabs(c->y[1] - c->y[0]);
        C.inv_sx = 1.0 / C.sx; C.inv_sy = 1.0 / C.sy;
        C.x_space = c->x[1] - c->x[0];                         // :515
        C.y_space = c->y[1] - c->y[0];                         // :516
        const size_t ncell = (size_t)(c->nx - 1) * (c->ny - 1);
        float4 *ccells = nullptr;
        if (ncell >= ((size_t)1 << 30)) return fail(MR_ERR_BAD_ARG, "current grid has 2^30 cells or more (32-bit cell addressing)");
        if ((rc = device_alloc(d, 4 * ncell, &ccells))) return rc;
        build_current_cells<<<(unsigned)std::min<size_t>((ncell + 255) / 256, 148 * 16), 256>>>(C.u, C.v, c->nx, c->ny, C.x_space, C.y_space, ccells);
        MR_CUDA(cudaGetLastError());
        C.cell = ccells;
        C.nxm1d = (double)(c->nx - 1); C.nym1d = (double)(c->ny - 1);
        std::vector<float> xf((size_t)c->nx), yf((size_t)c->ny);   // `as f32`, cartesian_current.rs:375-376
        for (int i = 0; i < c->nx; ++i) xf[(size_t)i] = (float)c->x[i];
        for (int i = 0; i < c->ny; ++i) yf[(size_t)i] = (float)c->y[i];
        if ((rc = upload(d, xf.data(), xf.size(), &C.xf))) return rc;
        if ((rc = upload(d, yf.data(), yf.size(), &C.yf))) return rc;
        C.xf0 = xf[0]; C.yf0 = yf[0];
        C.uniform = affine_f32(xf.data(), c->nx, &C.dxf) && affine_f32(yf.data(), c->ny, &C.dyf) &&
                    basis_coeffs(C.dxf, C.dyf, &C.c01, &C.c10);
        auto pack2c = [](float lo, float hi) {
            uint32_t a, b2;
            std::memcpy(&a, &lo, 4); std::memcpy(&b2, &hi, 4);
            return (unsigned long long)a | ((unsigned long long)b2 << 32);
        };
        C.p0 = pack2c(C.xf0, C.yf0); C.d2 = pack2c(C.dxf, C.dyf); C.c2 = pack2c(C.c10, C.c01);
    }
    MR_CUDA(cudaDeviceSynchronize());
    return MR_OK;
}

static void release_work(DeviceFields &d, bool everything)
{
    WorkArea &w = d.work;
    if (d.dev < 0) return;
    if (!w.arena[0] && !w.arena[1] && !w.s_comp) return;
    cudaSetDevice(d.dev);
    for (int b = 0; b < 2; ++b) {
        cudaFree(w.arena[b]);
        w.arena[b] = nullptr; w.bytes[b] = 0;
    }
    if (!everything) return;
    for (int b = 0; b < 2; ++b) {
        if (w.computed[b]) cudaEventDestroy(w.computed[b]);
        if (w.drained[b]) cudaEventDestroy(w.drained[b]);
        w.computed[b] = w.drained[b] = nullptr;
    }
    if (w.s_comp) cudaStreamDestroy(w.s_comp);
    if (w.s_copy) cudaStreamDestroy(w.s_copy);
    w.s_comp = w.s_copy = nullptr;
}

static void free_device_fields(DeviceFields &d)
{
    release_work(d, true);
    if (d.dev >= 0 && !d.allocs.empty()) {
        cudaSetDevice(d.dev);
        for (void *p : d.allocs) cudaFree(p);
    }
    d.allocs.clear();
}

static const DeviceFields *find_device(const mr_fields *f, int dev)
{
    for (auto &d : f->devs) if (d.dev == dev) return &d;
    return nullptr;
}

static void normalise_opts(const mr_trace_opts *in, mr_trace_opts &o)
{
    o = mr_trace_opts{1, MR_MATH_FAST, 0, 0};       // stride, math, chunk_rays, flags
    if (in) o = *in;
    if (o.stride <= 0) o.stride = 1;
}

// The depth-floor map.  Automatic choice: a quarter of the blocks deep for a 10 s wave — on
// shallower grids the map only adds a dependent load in front of every depth lookup.  MR_OPT_DEEP_MAP forces
// it on for any grid that has one, MR_OPT_NO_DEEP_MAP off.
static constexpr float kDeepMapAutoShare = 0.25f;
static int want_deep_map(const DeviceFields &d, const mr_trace_opts &o)
{
    if (o.flags & MR_OPT_NO_DEEP_MAP) return 0;
    return (o.flags & MR_OPT_DEEP_MAP) != 0 || d.deep_frac >= kDeepMapAutoShare;
}

static int enqueue_trace(const DeviceFields &d, cudaStream_t stream, int64_t n,
                         const double *x0, const double *y0, const double *kx0, const double *ky0,
                         double dt, int64_t nsteps, const mr_trace_opts &o,
                         double *x, double *y, double *kx, double *ky, int64_t ld,
                         int32_t *rows, int32_t *len, double *fin)
{
    TraceArgs a;
    a.b = d.b; a.c = d.c; a.n = n;
    a.x0 = x0; a.y0 = y0; a.kx0 = kx0; a.ky0 = ky0;
    a.dt = dt; a.nsteps = nsteps; a.stride = o.stride;
    a.x = x; a.y = y; a.kx = kx; a.ky = ky; a.ld = ld;
    a.rows = rows; a.len = len; a.fin = fin;
    a.deep_map = want_deep_map(d, o);
    cudaError_t e;
    if (o.math == MR_MATH_STRICT) e = launch_trace_strict(a, stream);
    else if (o.math == MR_MATH_FAST) e = launch_trace_fast(a, stream);
    else return fail(MR_ERR_BAD_ARG, "mr_trace_opts.math must be MR_MATH_FAST or MR_MATH_STRICT");
    if (e != cudaSuccess) return cuda_fail(e, "trace kernel launch");
    return MR_OK;
}

// ---- host-buffer path: one device's share --------------------------------------
struct HostJob {
    int64_t n_total;
    const double *x0, *y0, *kx0, *ky0;
    double dt; int64_t nsteps; int64_t rows_cap;
    mr_trace_opts o;
    double *x, *y, *kx, *ky;
    int32_t *rows, *len; double *fin;
    mr_env_planes env;           // depth/u/v planes at the stored rows, each may be NULL
};

struct DevBuf {
    double *ic = nullptr;        // [4][chunk] initial conditions
    double *traj = nullptr;      // [4][rows_cap][chunk]
    int32_t *rows = nullptr, *len = nullptr;
    double *fin = nullptr;       // [4][chunk]
    float *depth = nullptr;      // [rows_cap][chunk] environment planes (mr_trace_many_env)
    double *u = nullptr, *v = nullptr;
    cudaEvent_t computed = nullptr, drained = nullptr;
};

// Lays the arrays of one slab buffer out in an arena (256-byte aligned each); with base == nullptr it only
// measures.  Returns the bytes used.
static size_t carve(DevBuf &B, char *base, int64_t chunk, int64_t rows_cap, bool traj, bool fin, const mr_env_planes &env)
{
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return p;
    };
    const size_t c = (size_t)chunk, plane = (size_t)rows_cap * c;
    B.ic = (double *)take(sizeof(double) * 4 * c);
    B.traj = traj ? (double *)take(sizeof(double) * 4 * plane) : nullptr;
    B.rows = (int32_t *)take(sizeof(int32_t) * c);
    B.len = (int32_t *)take(sizeof(int32_t) * c);
    B.fin = fin ? (double *)take(sizeof(double) * 4 * c) : nullptr;
    B.depth = env.depth ? (float *)take(sizeof(float) * plane) : nullptr;
    B.u = env.u ? (double *)take(sizeof(double) * plane) : nullptr;
    B.v = env.v ? (double *)take(sizeof(double) * plane) : nullptr;
    return off;
}

// Traces rays [lo, hi) on device d.  Rays are cut into slabs of `chunk` rays;
// slab k+1 integrates on the compute stream while slab k drains to the host on
// the copy stream (two device buffers).  Buffers, streams and events live in the
// handle's WorkArea and are reused by the next call.
static int trace_block_on_device(DeviceFields &d, const HostJob &j, int64_t lo, int64_t hi, std::string &err)
{
    auto bail = [&](int code, const std::string &m) { err = m; return code; };
#define MR_TRY(call)                                                                      \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            rc = bail(e__ == cudaErrorMemoryAllocation ? MR_ERR_OOM : MR_ERR_CUDA,        \
                      std::string(#call) + ": " + cudaGetErrorString(e__));               \
            goto done;                                                                    \
        }                                                                                 \
    } while (0)

    int rc = MR_OK;
    const int64_t n = hi - lo;
    if (n <= 0) return MR_OK;
    // MR_DEBUG_TIMING=1: where a host-buffer call spends its wall time (stderr)
    static const bool timing = std::getenv("MR_DEBUG_TIMING") != nullptr;
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    auto t_alloc = t_begin, t_enqueued = t_begin, t_synced = t_begin;
    const bool want_traj = j.x || j.y || j.kx || j.ky;
    WorkArea &w = d.work;
    DevBuf buf[2];
    int nbuf = 1;
    int64_t chunk = n;
    {
        cudaError_t e = cudaSetDevice(d.dev);
        if (e != cudaSuccess) return bail(MR_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    }
    {
        size_t free_b = 0, total_b = 0;
        MR_TRY(cudaMemGetInfo(&free_b, &total_b));
        // bytes one ray needs on the device
        const double env_row = (j.env.depth ? 4.0 : 0.0) + (j.env.u ? 8.0 : 0.0) + (j.env.v ? 8.0 : 0.0);
        const double per_ray = 32.0 + (want_traj ? (32.0 + env_row) * (double)j.rows_cap : 0.0) + 8.0 + 32.0;
        double budget = 0.80 * (double)(free_b + w.held());      // what this handle already holds is ours to reuse
        if (j.o.chunk_rays > 0) {
            chunk = std::min<int64_t>(n, j.o.chunk_rays);
        } else if (per_ray * (double)n > budget || (want_traj && per_ray * (double)n > 4e9)) {
            // does not fit, or is big enough that overlapping the drain with the next
            // slab's integration pays: two slabs in flight
            const double cap = std::min(budget / 2.0, 16e9);
            const int64_t wave = 148 * 4 * kBlock;          // rays that fill the machine once
            chunk = (int64_t)(cap / per_ray);
            if (chunk >= wave) chunk = chunk / wave * wave;
            chunk = std::max<int64_t>(std::min(chunk, n), 1);
        }
        nbuf = chunk < n ? 2 : 1;
    }
    if (!w.s_comp) MR_TRY(cudaStreamCreateWithFlags(&w.s_comp, cudaStreamNonBlocking));
    if (!w.s_copy) MR_TRY(cudaStreamCreateWithFlags(&w.s_copy, cudaStreamNonBlocking));
    {
        const size_t need = carve(buf[0], nullptr, chunk, j.rows_cap, want_traj, j.fin != nullptr, j.env);
        // an arena is reused when it is large enough and not wastefully larger; growing frees first so that
        // the peak is the new size, not old + new
        for (int b = 0; b < 2; ++b) {
            const bool wanted = b < nbuf;
            if (w.arena[b] && (!wanted || w.bytes[b] < need || w.bytes[b] / 4 > need)) {
                MR_TRY(cudaFree(w.arena[b]));
                w.arena[b] = nullptr; w.bytes[b] = 0;
            }
        }
        for (int b = 0; b < nbuf; ++b) {
            if (!w.arena[b]) {
                MR_TRY(cudaMalloc(&w.arena[b], need));
                w.bytes[b] = need;
            }
            carve(buf[b], (char *)w.arena[b], chunk, j.rows_cap, want_traj, j.fin != nullptr, j.env);
            if (!w.computed[b]) MR_TRY(cudaEventCreateWithFlags(&w.computed[b], cudaEventDisableTiming));
            if (!w.drained[b]) MR_TRY(cudaEventCreateWithFlags(&w.drained[b], cudaEventDisableTiming));
            buf[b].computed = w.computed[b];
            buf[b].drained = w.drained[b];
        }
    }
    t_alloc = clk::now();
    {
        cudaStream_t s_comp = w.s_comp, s_copy = w.s_copy;
        int k = 0;
        for (int64_t c0 = lo; c0 < hi; c0 += chunk, ++k) {
            DevBuf &B = buf[k % nbuf];
            const int64_t m = std::min(chunk, hi - c0);
            if (k >= nbuf) MR_TRY(cudaStreamWaitEvent(s_comp, B.drained, 0));
            MR_TRY(cudaMemcpyAsync(B.ic,             j.x0  + c0, sizeof(double) * m, cudaMemcpyHostToDevice, s_comp));
            MR_TRY(cudaMemcpyAsync(B.ic + chunk,     j.y0  + c0, sizeof(double) * m, cudaMemcpyHostToDevice, s_comp));
            MR_TRY(cudaMemcpyAsync(B.ic + 2 * chunk, j.kx0 + c0, sizeof(double) * m, cudaMemcpyHostToDevice, s_comp));
            MR_TRY(cudaMemcpyAsync(B.ic + 3 * chunk, j.ky0 + c0, sizeof(double) * m, cudaMemcpyHostToDevice, s_comp));
            const size_t plane = (size_t)j.rows_cap * (size_t)chunk;
            double *tx = want_traj ? B.traj : nullptr;
            {
                TraceArgs a;
                a.b = d.b; a.c = d.c; a.n = m;
                a.x0 = B.ic; a.y0 = B.ic + chunk; a.kx0 = B.ic + 2 * chunk; a.ky0 = B.ic + 3 * chunk;
                a.dt = j.dt; a.nsteps = j.nsteps; a.stride = j.o.stride;
                a.x = tx; a.y = tx ? tx + plane : nullptr; a.kx = tx ? tx + 2 * plane : nullptr; a.ky = tx ? tx + 3 * plane : nullptr;
                a.ld = chunk;
                a.rows = B.rows; a.len = B.len; a.fin = B.fin;
                a.deep_map = want_deep_map(d, j.o);
                // fin is [4][n] with n = m for the kernel (it uses a.n as the pitch)
                cudaError_t e = j.o.math == MR_MATH_STRICT ? launch_trace_strict(a, s_comp) : launch_trace_fast(a, s_comp);
                if (e != cudaSuccess) { rc = bail(MR_ERR_CUDA, std::string("trace kernel launch: ") + cudaGetErrorString(e)); goto done; }
            }
            if (B.depth || B.u || B.v)       // the fields at every stored state, from the planes just written
                MR_TRY(launch_sample(d.b, d.c, j.rows_cap, m, chunk, tx, tx + plane, B.depth, B.u, B.v, s_comp));
            MR_TRY(cudaEventRecord(B.computed, s_comp));
            MR_TRY(cudaStreamWaitEvent(s_copy, B.computed, 0));
            if (j.env.depth)
                MR_TRY(cudaMemcpy2DAsync(j.env.depth + c0, sizeof(float) * (size_t)j.n_total, B.depth, sizeof(float) * (size_t)chunk,
                                         sizeof(float) * (size_t)m, (size_t)j.rows_cap, cudaMemcpyDeviceToHost, s_copy));
            if (j.env.u)
                MR_TRY(cudaMemcpy2DAsync(j.env.u + c0, sizeof(double) * (size_t)j.n_total, B.u, sizeof(double) * (size_t)chunk,
                                         sizeof(double) * (size_t)m, (size_t)j.rows_cap, cudaMemcpyDeviceToHost, s_copy));
            if (j.env.v)
                MR_TRY(cudaMemcpy2DAsync(j.env.v + c0, sizeof(double) * (size_t)j.n_total, B.v, sizeof(double) * (size_t)chunk,
                                         sizeof(double) * (size_t)m, (size_t)j.rows_cap, cudaMemcpyDeviceToHost, s_copy));
            if (want_traj) {
                double *dsts[4] = { j.x, j.y, j.kx, j.ky };
                for (int f = 0; f < 4; ++f) {
                    if (!dsts[f]) continue;
                    MR_TRY(cudaMemcpy2DAsync(dsts[f] + c0, sizeof(double) * (size_t)j.n_total,
                                             B.traj + f * plane, sizeof(double) * (size_t)chunk,
                                             sizeof(double) * (size_t)m, (size_t)j.rows_cap,
                                             cudaMemcpyDeviceToHost, s_copy));
                }
            }
            if (j.rows) MR_TRY(cudaMemcpyAsync(j.rows + c0, B.rows, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s_copy));
            if (j.len)  MR_TRY(cudaMemcpyAsync(j.len + c0,  B.len,  sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s_copy));
            if (j.fin) {
                for (int f = 0; f < 4; ++f)
                    MR_TRY(cudaMemcpyAsync(j.fin + (size_t)f * j.n_total + c0, B.fin + (size_t)f * m,
                                           sizeof(double) * m, cudaMemcpyDeviceToHost, s_copy));
            }
            MR_TRY(cudaEventRecord(B.drained, s_copy));
        }
        t_enqueued = clk::now();
        MR_TRY(cudaStreamSynchronize(s_comp));
        MR_TRY(cudaStreamSynchronize(s_copy));
        t_synced = clk::now();
    }
done:
    if (rc != MR_OK) {
        // leave nothing of a failed call behind: the next one starts from a clean device state
        cudaDeviceSynchronize();
        release_work(d, false);
        (void)cudaGetLastError();
    }
    if (timing) {
        auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "[mantaray_b200] device %d: %lld rays in slabs of %lld (%d buffers): alloc %.1f ms, enqueue %.1f ms, "
                             "drain %.1f ms, tail %.1f ms, %.2f GB held\n", d.dev, (long long)n, (long long)chunk, nbuf,
                     ms(t_begin, t_alloc), ms(t_alloc, t_enqueued), ms(t_enqueued, t_synced), ms(t_synced, clk::now()),
                     (double)w.held() / 1e9);
    }
    return rc;
#undef MR_TRY
}

}  // namespace mr

using namespace mr;

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

int mr_abi_version(void) { return MR_ABI_VERSION; }
int mr_device_count(void) { return device_count_quiet(); }
const char *mr_last_error(void) { return g_err.c_str(); }

int mr_fields_create(const mr_bathymetry_desc *bathy, const mr_current_desc *current,
                     uint32_t device_mask, mr_fields **out)
{
    if (!out) return fail(MR_ERR_BAD_ARG, "mr_fields_create: out is NULL");
    *out = nullptr;
    int rc = validate(bathy, current);
    if (rc) return rc;
    int ndev = device_count_quiet();
    if (ndev <= 0) return fail(MR_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device_mask == 0) device_mask = 1u;
    for (int i = 0; i < 32; ++i)
        if ((device_mask >> i & 1u) && i >= ndev)
            return fail(MR_ERR_BAD_ARG, "device_mask selects device " + std::to_string(i) + " but only " +
                                            std::to_string(ndev) + " device(s) are visible");
    std::unique_ptr<mr_fields> f(new (std::nothrow) mr_fields);
    if (!f) return fail(MR_ERR_OOM, "out of host memory");
    f->mask = device_mask;
    for (int i = 0; i < 32; ++i) {
        if (!(device_mask >> i & 1u)) continue;
        f->devs.emplace_back();
        f->devs.back().dev = i;
        rc = upload_fields(f->devs.back(), bathy, current);
        if (rc) {
            for (auto &d : f->devs) free_device_fields(d);
            return rc;
        }
    }
    *out = f.release();
    return MR_OK;
}

int mr_fields_open_netcdf3(const char *bathymetry_path, const char *current_path,
                           uint32_t device_mask, mr_fields **out)
{
    if (!out) return fail(MR_ERR_BAD_ARG, "mr_fields_open_netcdf3: out is NULL");
    *out = nullptr;
    std::string err;
    std::vector<float> bx, by;
    std::vector<double> depth, cx, cy, cu, cv;
    mr_bathymetry_desc b{};
    mr_current_desc c{};
    if (bathymetry_path) {
        // CartesianNetcdf3::open(path, "x", "y", "depth")  src/ffi.rs:36, :62
        Nc3File f;
        int rc = Nc3File::open(bathymetry_path, f, err);
        if (rc) return fail(rc, "could not open bathymetry file: " + err);
        if ((rc = f.read_f32("x", bx, err)) || (rc = f.read_f32("y", by, err)) || (rc = f.read_f64("depth", depth, err)))
            return fail(rc, "could not open bathymetry file: " + err);
        if (bx.size() > (size_t)INT32_MAX || by.size() > (size_t)INT32_MAX)
            return fail(MR_ERR_FORMAT, "bathymetry coordinates too long");
        if (depth.size() != bx.size() * by.size())
            return fail(MR_ERR_FORMAT, "bathymetry file: depth has " + std::to_string(depth.size()) + " values, expected len(x)*len(y) = " +
                                           std::to_string(bx.size() * by.size()));
        b.kind = MR_BATHY_GRID; b.nx = (int32_t)bx.size(); b.ny = (int32_t)by.size();
        b.x = bx.data(); b.y = by.data(); b.depth = depth.data();
    } else {
        b.kind = MR_BATHY_CONSTANT; b.h0 = 2000.0f;          // DEFAULT_BATHYMETRY constant_depth.rs:9
    }
    if (current_path) {
        // CartesianCurrent::open(path, "x", "y", "u", "v")  src/ffi.rs:38, :64
        Nc3File f;
        int rc = Nc3File::open(current_path, f, err);
        if (rc) return fail(rc, "could not open current file: " + err);
        if ((rc = f.read_f64("x", cx, err)) || (rc = f.read_f64("y", cy, err)) ||
            (rc = f.read_f64("u", cu, err)) || (rc = f.read_f64("v", cv, err)))
            return fail(rc, "could not open current file: " + err);
        if (cx.size() > (size_t)INT32_MAX || cy.size() > (size_t)INT32_MAX)
            return fail(MR_ERR_FORMAT, "current coordinates too long");
        if (cu.size() != cx.size() * cy.size() || cv.size() != cx.size() * cy.size())
            return fail(MR_ERR_FORMAT, "current file: u/v do not have len(x)*len(y) values");
        c.kind = MR_CURRENT_GRID; c.nx = (int32_t)cx.size(); c.ny = (int32_t)cy.size();
        c.x = cx.data(); c.y = cy.data(); c.u = cu.data(); c.v = cv.data();
    } else {
        c.kind = MR_CURRENT_CONSTANT; c.u0 = 0.0; c.v0 = 0.0;  // DEFAULT_CURRENT constant_current.rs:10
    }
    return mr_fields_create(&b, &c, device_mask, out);
}

void mr_fields_free(mr_fields *f)
{
    if (!f) return;
    for (auto &d : f->devs) free_device_fields(d);
    delete f;
}

uint32_t mr_fields_device_mask(const mr_fields *f) { return f ? f->mask : 0; }

void mr_fields_trim(mr_fields *f)
{
    if (!f) return;
    std::lock_guard<std::mutex> guard(f->mu);
    for (auto &d : f->devs) release_work(d, false);
}

int64_t mr_num_steps(double t0, double t_end, double dt)
{
    if (!(dt > 0.0)) return -1;
    double q = std::ceil((t_end - t0) / dt);
    if (!(q >= 0.0) || !(q < 2147483646.0)) return -1;
    return (int64_t)q;
}

int64_t mr_num_rows(double t0, double t_end, double dt, int32_t stride)
{
    int64_t n = mr_num_steps(t0, t_end, dt);
    if (n < 0) return -1;
    if (stride <= 0) stride = 1;
    return n / stride + 1;
}

static void fill_time(double *t, double t0, double dt, int64_t nsteps, int32_t stride)
{
    if (!t) return;
    double tt = t0;
    t[0] = tt;
    for (int64_t s = 1; s <= nsteps; ++s) {
        tt = tt + dt;                  // x_new = x + h, accumulated (ode_solvers Rk4::step)
        if (s % stride == 0) t[s / stride] = tt;
    }
}

int mr_trace_many(mr_fields *f, int64_t n,
                  const double *x0, const double *y0, const double *kx0, const double *ky0,
                  double t0, double t_end, double dt, const mr_trace_opts *opts,
                  double *t, double *x, double *y, double *kx, double *ky,
                  int32_t *rows, int32_t *len, double *final_state)
{
    return mr_trace_many_env(f, n, x0, y0, kx0, ky0, t0, t_end, dt, opts, t, x, y, kx, ky, rows, len, final_state, nullptr);
}

int mr_trace_many_env(mr_fields *f, int64_t n,
                      const double *x0, const double *y0, const double *kx0, const double *ky0,
                      double t0, double t_end, double dt, const mr_trace_opts *opts,
                      double *t, double *x, double *y, double *kx, double *ky,
                      int32_t *rows, int32_t *len, double *final_state, const mr_env_planes *env)
{
    if (!f) return fail(MR_ERR_BAD_ARG, "mr_trace_many: NULL field handle");
    if (n < 0) return fail(MR_ERR_BAD_ARG, "mr_trace_many: n < 0");
    if (n > 0 && (!x0 || !y0 || !kx0 || !ky0)) return fail(MR_ERR_BAD_ARG, "mr_trace_many: NULL initial-condition array");
    mr_trace_opts o;
    normalise_opts(opts, o);
    if (o.math != MR_MATH_FAST && o.math != MR_MATH_STRICT) return fail(MR_ERR_BAD_ARG, "mr_trace_opts.math must be MR_MATH_FAST or MR_MATH_STRICT");
    const int64_t nsteps = mr_num_steps(t0, t_end, dt);
    if (nsteps < 0) return fail(MR_ERR_BAD_ARG, "mr_trace_many: need dt > 0 and 0 <= (t_end - t0)/dt < 2^31 (the reference panics here)");
    const bool any_traj = x || y || kx || ky;
    if (any_traj && !(x && y && kx && ky)) return fail(MR_ERR_BAD_ARG, "mr_trace_many: pass all four of x, y, kx, ky or none");
    const bool any_env = env && (env->depth || env->u || env->v);
    if (any_env && !any_traj) return fail(MR_ERR_BAD_ARG, "mr_trace_many_env: the environment planes need the x, y, kx, ky planes");
    fill_time(t, t0, dt, nsteps, o.stride);
    if (n == 0) return MR_OK;

    HostJob j;
    j.n_total = n; j.x0 = x0; j.y0 = y0; j.kx0 = kx0; j.ky0 = ky0;
    j.dt = dt; j.nsteps = nsteps; j.rows_cap = nsteps / o.stride + 1; j.o = o;
    j.x = x; j.y = y; j.kx = kx; j.ky = ky; j.rows = rows; j.len = len; j.fin = final_state;
    j.env = any_env ? *env : mr_env_planes{nullptr, nullptr, nullptr};

    std::lock_guard<std::mutex> guard(f->mu);
    const int G = (int)f->devs.size();
    // contiguous blocks of rays per device, rounded to whole thread blocks
    int64_t per = (n + G - 1) / G;
    per = (per + kBlock - 1) / kBlock * kBlock;
    std::vector<int> rcs(G, MR_OK);
    std::vector<std::string> errs(G);
    if (G == 1) {
        rcs[0] = trace_block_on_device(f->devs[0], j, 0, n, errs[0]);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) {
            int64_t lo = std::min<int64_t>((int64_t)g * per, n), hi = std::min<int64_t>(lo + per, n);
            th.emplace_back([&, g, lo, hi] { rcs[g] = trace_block_on_device(f->devs[g], j, lo, hi, errs[g]); });
        }
        for (auto &t_ : th) t_.join();
    }
    for (int g = 0; g < G; ++g)
        if (rcs[g] != MR_OK) return fail(rcs[g], "device " + std::to_string(f->devs[g].dev) + ": " + errs[g]);
    return MR_OK;
}

int mr_single_ray(mr_fields *f, double x0, double y0, double kx0, double ky0,
                  double t0, double t_end, double dt, const mr_trace_opts *opts,
                  double *out, int64_t out_cap, int64_t *n_rows)
{
    if (!f || !n_rows) return fail(MR_ERR_BAD_ARG, "mr_single_ray: NULL argument");
    mr_trace_opts o;
    normalise_opts(opts, o);
    if (o.stride != 1) return fail(MR_ERR_BAD_ARG, "mr_single_ray: stride must be 1");
    const int64_t nsteps = mr_num_steps(t0, t_end, dt);
    if (nsteps < 0) return fail(MR_ERR_BAD_ARG, "mr_single_ray: need dt > 0 and 0 <= (t_end - t0)/dt < 2^31 (the reference panics here)");
    const int64_t cap = nsteps + 1;
    std::vector<double> t((size_t)cap), soa((size_t)cap * 4);
    int32_t rows = 0;
    // the first device of the handle: a single ray cannot be sharded
    fill_time(t.data(), t0, dt, nsteps, 1);
    HostJob j{};
    j.n_total = 1; j.x0 = &x0; j.y0 = &y0; j.kx0 = &kx0; j.ky0 = &ky0;
    j.dt = dt; j.nsteps = nsteps; j.rows_cap = cap; j.o = o;
    j.x = soa.data(); j.y = soa.data() + cap; j.kx = soa.data() + 2 * cap; j.ky = soa.data() + 3 * cap;
    j.rows = &rows; j.len = nullptr; j.fin = nullptr;
    j.env = mr_env_planes{nullptr, nullptr, nullptr};
    {
        std::lock_guard<std::mutex> guard(f->mu);
        std::string err;
        int rc = trace_block_on_device(f->devs[0], j, 0, 1, err);
        if (rc) return fail(rc, "device " + std::to_string(f->devs[0].dev) + ": " + err);
    }
    *n_rows = rows;
    if (rows > out_cap || !out) return fail(MR_ERR_BAD_ARG, "mr_single_ray: out holds " + std::to_string(out_cap) +
                                                        " rows, " + std::to_string(rows) + " needed");
    for (int64_t r = 0; r < rows; ++r) {    // (t, x, y, kx, ky) tuples, src/ffi.rs:42-47
        out[5 * r + 0] = t[(size_t)r];
        for (int c = 0; c < 4; ++c) out[5 * r + 1 + c] = soa[(size_t)c * cap + r];
    }
    return MR_OK;
}

int mr_trace_device(mr_fields *f, int device, void *stream, int64_t n,
                    const double *d_x0, const double *d_y0, const double *d_kx0, const double *d_ky0,
                    double t0, double t_end, double dt, const mr_trace_opts *opts,
                    double *d_x, double *d_y, double *d_kx, double *d_ky, int64_t ld,
                    int32_t *d_rows, int32_t *d_len, double *d_final, int32_t *launches)
{
    if (launches) *launches = 0;
    if (!f) return fail(MR_ERR_BAD_ARG, "mr_trace_device: NULL field handle");
    const DeviceFields *d = find_device(f, device);
    if (!d) return fail(MR_ERR_BAD_ARG, "mr_trace_device: device " + std::to_string(device) + " is not in the handle's mask");
    if (n < 0) return fail(MR_ERR_BAD_ARG, "mr_trace_device: n < 0");
    if (n > 0 && (!d_x0 || !d_y0 || !d_kx0 || !d_ky0)) return fail(MR_ERR_BAD_ARG, "mr_trace_device: NULL initial-condition array");
    mr_trace_opts o;
    normalise_opts(opts, o);
    const int64_t nsteps = mr_num_steps(t0, t_end, dt);
    if (nsteps < 0) return fail(MR_ERR_BAD_ARG, "mr_trace_device: need dt > 0 and 0 <= (t_end - t0)/dt < 2^31");
    const bool any_traj = d_x || d_y || d_kx || d_ky;
    if (any_traj && !(d_x && d_y && d_kx && d_ky)) return fail(MR_ERR_BAD_ARG, "mr_trace_device: pass all four of x, y, kx, ky or none");
    if (any_traj && ld < n) return fail(MR_ERR_BAD_ARG, "mr_trace_device: ld < n");
    if (n == 0) return MR_OK;
    int cur = -1;
    MR_CUDA(cudaGetDevice(&cur));
    if (cur != device) MR_CUDA(cudaSetDevice(device));
    int rc = enqueue_trace(*d, (cudaStream_t)stream, n, d_x0, d_y0, d_kx0, d_ky0, dt, nsteps, o,
                           d_x, d_y, d_kx, d_ky, ld, d_rows, d_len, d_final);
    if (cur != device) cudaSetDevice(cur);
    if (rc == MR_OK && launches) *launches = 1;
    return rc;
}

int mr_sample_device(mr_fields *f, int device, void *stream, int64_t rows, int64_t n, int64_t ld,
                     const double *d_x, const double *d_y,
                     float *d_depth, double *d_u, double *d_v, int32_t *launches)
{
    if (launches) *launches = 0;
    if (!f) return fail(MR_ERR_BAD_ARG, "mr_sample_device: NULL field handle");
    const DeviceFields *d = find_device(f, device);
    if (!d) return fail(MR_ERR_BAD_ARG, "mr_sample_device: device " + std::to_string(device) + " is not in the handle's mask");
    if (rows < 0 || n < 0 || ld < n) return fail(MR_ERR_BAD_ARG, "mr_sample_device: need rows >= 0 and 0 <= n <= ld");
    if (rows == 0 || n == 0 || !(d_depth || d_u || d_v)) return MR_OK;
    if (!d_x || !d_y) return fail(MR_ERR_BAD_ARG, "mr_sample_device: NULL point array");
    int cur = -1;
    MR_CUDA(cudaGetDevice(&cur));
    if (cur != device) MR_CUDA(cudaSetDevice(device));
    cudaError_t e = launch_sample(d->b, d->c, rows, n, ld, d_x, d_y, d_depth, d_u, d_v, (cudaStream_t)stream);
    if (cur != device) cudaSetDevice(cur);
    if (e != cudaSuccess) return fail(MR_ERR_CUDA, std::string("sample kernel launch: ") + cudaGetErrorString(e));
    if (launches) *launches = 1;
    return MR_OK;
}

int mr_sample_fields(mr_fields *f, int64_t count, const double *x, const double *y,
                     float *depth, double *u, double *v)
{
    if (!f) return fail(MR_ERR_BAD_ARG, "mr_sample_fields: NULL field handle");
    if (count < 0) return fail(MR_ERR_BAD_ARG, "mr_sample_fields: count < 0");
    if (count == 0 || !(depth || u || v)) return MR_OK;
    if (!x || !y) return fail(MR_ERR_BAD_ARG, "mr_sample_fields: NULL point array");
    std::lock_guard<std::mutex> guard(f->mu);
    const DeviceFields &d = f->devs[0];
    MR_CUDA(cudaSetDevice(d.dev));
    // slabs bound the device footprint (36 B a point) whatever `count` is
    const int64_t slab = std::min<int64_t>(count, (int64_t)1 << 26);
    double *dx = nullptr, *dy = nullptr, *du = nullptr, *dv = nullptr;
    float *dh = nullptr;
    int rc = MR_OK;
    auto release = [&] { cudaFree(dx); cudaFree(dy); cudaFree(du); cudaFree(dv); cudaFree(dh); };
#define MR_TRY(call)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            rc = fail(e__ == cudaErrorMemoryAllocation ? MR_ERR_OOM : MR_ERR_CUDA,                \
                      std::string(#call) + ": " + cudaGetErrorString(e__));                       \
            cudaDeviceSynchronize(); (void)cudaGetLastError();                                    \
            release();                                                                            \
            return rc;                                                                            \
        }                                                                                         \
    } while (0)
    MR_TRY(cudaMalloc(&dx, sizeof(double) * (size_t)slab));
    MR_TRY(cudaMalloc(&dy, sizeof(double) * (size_t)slab));
    if (depth) MR_TRY(cudaMalloc(&dh, sizeof(float) * (size_t)slab));
    if (u) MR_TRY(cudaMalloc(&du, sizeof(double) * (size_t)slab));
    if (v) MR_TRY(cudaMalloc(&dv, sizeof(double) * (size_t)slab));
    for (int64_t c0 = 0; c0 < count; c0 += slab) {
        const int64_t m = std::min(slab, count - c0);
        MR_TRY(cudaMemcpyAsync(dx, x + c0, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, 0));
        MR_TRY(cudaMemcpyAsync(dy, y + c0, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, 0));
        MR_TRY(launch_sample(d.b, d.c, 1, m, m, dx, dy, dh, du, dv, 0));
        if (depth) MR_TRY(cudaMemcpyAsync(depth + c0, dh, sizeof(float) * (size_t)m, cudaMemcpyDeviceToHost, 0));
        if (u) MR_TRY(cudaMemcpyAsync(u + c0, du, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, 0));
        if (v) MR_TRY(cudaMemcpyAsync(v + c0, dv, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, 0));
        MR_TRY(cudaStreamSynchronize(0));
    }
#undef MR_TRY
    release();
    return MR_OK;
}

int mr_measure_fp64_peak(int device, int millis, double *tflops)
{
    if (!tflops) return fail(MR_ERR_BAD_ARG, "mr_measure_fp64_peak: NULL output");
    *tflops = 0.0;
    if (device < 0 || device >= device_count_quiet()) return fail(MR_ERR_CUDA, "mr_measure_fp64_peak: no such CUDA device");
    int cur = -1;
    MR_CUDA(cudaGetDevice(&cur));
    MR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MR_CUDA(cudaGetDeviceProperties(&prop, device));
    double *sink = nullptr;
    MR_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    MR_CUDA(cudaEventCreate(&e0));
    MR_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8;
    int iters = 2000;
    double best = 0.0;
    float ms = 0.f;
    // warm up, then size the loop for ~millis of work
    MR_CUDA(launch_dfma_probe(sink, iters, blocks, 0));
    MR_CUDA(cudaDeviceSynchronize());
    for (int rep = 0; rep < 4; ++rep) {
        MR_CUDA(cudaEventRecord(e0, 0));
        MR_CUDA(launch_dfma_probe(sink, iters, blocks, 0));
        MR_CUDA(cudaEventRecord(e1, 0));
        MR_CUDA(cudaEventSynchronize(e1));
        MR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
        double tf = flops / ((double)ms * 1e-3) / 1e12;
        if (rep > 0) best = std::max(best, tf);
        if (rep == 0 && ms > 0.f && millis > 0) {
            double scale = (double)millis / ms;
            iters = (int)std::min(2e6, std::max(200.0, iters * scale));
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (cur != device) cudaSetDevice(cur);
    *tflops = best;
    return MR_OK;
}

int mr_depth_floor_map(const mr_bathymetry_desc *b, float *out, size_t cap, int32_t *nbx, int32_t *nby, float *deep_frac,
                       int32_t *affine)
{
    if (!b || !nbx || !nby) return fail(MR_ERR_BAD_ARG, "mr_depth_floor_map: NULL argument");
    if (b->kind != MR_BATHY_GRID || b->nx < 2 || b->ny < 2 || !b->depth || !b->x || !b->y)
        return fail(MR_ERR_BAD_ARG, "mr_depth_floor_map: needs a GRID bathymetry with nx, ny >= 2");
    if (affine) {       // the same test upload_fields applies before it builds (and the kernel consults) the map
        float rsx, rsy, dxf, dyf, c01, c10;
        *affine = recip_ok(fabsf(b->x[1] - b->x[0]), &rsx) && recip_ok(fabsf(b->y[1] - b->y[0]), &rsy) &&
                  affine_f32(b->x, b->nx, &dxf) && affine_f32(b->y, b->ny, &dyf) && basis_coeffs(dxf, dyf, &c01, &c10);
    }
    int bx = 0, by = 0;
    float frac = 0.0f;
    const std::vector<float> m = depth_floor_map(b->depth, b->nx, b->ny, &bx, &by, &frac);
    *nbx = bx; *nby = by;
    if (deep_frac) *deep_frac = frac;
    if (out) {
        if (cap < m.size()) return fail(MR_ERR_BAD_ARG, "mr_depth_floor_map: output too small");
        std::memcpy(out, m.data(), m.size() * sizeof(float));
    }
    return MR_OK;
}

int mr_selftest_fdiv(int device, float spacing, uint64_t *mismatches, int32_t *usable)
{
    if (!mismatches || !usable) return fail(MR_ERR_BAD_ARG, "mr_selftest_fdiv: NULL output");
    *mismatches = 0;
    float r = 0.f;
    *usable = recip_ok(spacing, &r) ? 1 : 0;
    if (!*usable) return MR_OK;
    if (device < 0 || device >= device_count_quiet()) return fail(MR_ERR_CUDA, "mr_selftest_fdiv: no such CUDA device");
    int cur = -1;
    MR_CUDA(cudaGetDevice(&cur));
    MR_CUDA(cudaSetDevice(device));
    unsigned long long *bad = nullptr, h = 0;
    MR_CUDA(cudaMalloc(&bad, sizeof(*bad)));
    MR_CUDA(cudaMemset(bad, 0, sizeof(*bad)));
    fdiv_selftest_kernel<<<148 * 32, 256>>>(spacing, r, bad);
    MR_CUDA(cudaGetLastError());
    MR_CUDA(cudaMemcpy(&h, bad, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(bad);
    if (cur != device) cudaSetDevice(cur);
    *mismatches = h;
    return MR_OK;
}

// ---- NetCDF-3 ---------------------------------------------------------------
int mr_nc3_open(const char *path, mr_nc3 **out)
{
    if (!path || !out) return fail(MR_ERR_BAD_ARG, "mr_nc3_open: NULL argument");
    *out = nullptr;
    std::unique_ptr<mr_nc3> h(new (std::nothrow) mr_nc3);
    if (!h) return fail(MR_ERR_OOM, "out of host memory");
    std::string err;
    int rc = Nc3File::open(path, h->file, err);
    if (rc) return fail(rc, err);
    *out = h.release();
    return MR_OK;
}
void mr_nc3_close(mr_nc3 *f) { delete f; }
int mr_nc3_var_count(const mr_nc3 *f) { return f ? (int)f->file.vars.size() : 0; }
int mr_nc3_var_name(const mr_nc3 *f, int index, char *buf, size_t cap)
{
    if (!f || !buf || cap == 0 || index < 0 || index >= (int)f->file.vars.size()) return fail(MR_ERR_BAD_ARG, "mr_nc3_var_name: bad argument");
    std::strncpy(buf, f->file.vars[(size_t)index].name.c_str(), cap - 1);
    buf[cap - 1] = 0;
    return MR_OK;
}
int mr_nc3_var_info(const mr_nc3 *f, const char *name, int32_t *nc_type, int64_t *n_elems, int32_t *ndims, int64_t dims[MR_NC3_MAX_DIMS])
{
    if (!f || !name) return fail(MR_ERR_BAD_ARG, "mr_nc3_var_info: NULL argument");
    const Nc3Var *v = f->file.find(name);
    if (!v) return fail(MR_ERR_FORMAT, "'" + f->file.path + "': no variable named '" + name + "'");
    if (nc_type) *nc_type = v->type;
    if (n_elems) *n_elems = (int64_t)f->file.num_elems(*v);
    if (ndims) *ndims = (int32_t)v->dimids.size();
    if (dims)
        for (size_t k = 0; k < v->dimids.size() && k < MR_NC3_MAX_DIMS; ++k) {
            uint32_t l = f->file.dims[v->dimids[k]].len;
            dims[k] = (k == 0 && v->is_record) ? (int64_t)f->file.numrecs : (int64_t)l;
        }
    return MR_OK;
}
int mr_nc3_read_f32(const mr_nc3 *f, const char *name, float *out, int64_t cap)
{
    if (!f || !name || !out) return fail(MR_ERR_BAD_ARG, "mr_nc3_read_f32: NULL argument");
    std::vector<float> v;
    std::string err;
    int rc = f->file.read_f32(name, v, err);
    if (rc) return fail(rc, err);
    if ((int64_t)v.size() > cap) return fail(MR_ERR_BAD_ARG, "mr_nc3_read_f32: buffer too small");
    std::memcpy(out, v.data(), v.size() * sizeof(float));
    return MR_OK;
}
int mr_nc3_read_f64(const mr_nc3 *f, const char *name, double *out, int64_t cap)
{
    if (!f || !name || !out) return fail(MR_ERR_BAD_ARG, "mr_nc3_read_f64: NULL argument");
    std::vector<double> v;
    std::string err;
    int rc = f->file.read_f64(name, v, err);
    if (rc) return fail(rc, err);
    if ((int64_t)v.size() > cap) return fail(MR_ERR_BAD_ARG, "mr_nc3_read_f64: buffer too small");
    std::memcpy(out, v.data(), v.size() * sizeof(double));
    return MR_OK;
}

// ---- pinned host memory -------------------------------------------------------
int mr_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(MR_ERR_BAD_ARG, "mr_host_alloc: NULL out");
    *out = nullptr;
    if (device_count_quiet() <= 0) return fail(MR_ERR_CUDA, "no CUDA device available");
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { (void)cudaGetLastError(); *out = nullptr; return cuda_fail(e, "cudaHostAlloc"); }
    return MR_OK;
}
void mr_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"
