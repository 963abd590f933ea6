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
#include <atomic>
#include <mutex>
#include <new>
#include <stdexcept>
#include <system_error>
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
// No exception crosses the C ABI ("no unwinding across the ABI").  Every extern "C" body that can
// allocate runs between MR_API_BEGIN and MR_API_END; this maps whatever was thrown to a status code.
static int translate_exception(const char *where) noexcept
{
    try {
        try { throw; }
        catch (const std::bad_alloc &) { return fail(MR_ERR_OOM, std::string(where) + ": out of host memory"); }
        catch (const std::length_error &e) { return fail(MR_ERR_OOM, std::string(where) + ": size too large for host memory (" + e.what() + ")"); }
        catch (const std::system_error &e) { return fail(MR_ERR_OOM, std::string(where) + ": could not start a host thread (" + e.what() + ")"); }
        catch (const std::exception &e) { return fail(MR_ERR_FORMAT, std::string(where) + ": " + e.what()); }
        catch (...) { return fail(MR_ERR_FORMAT, std::string(where) + ": unknown C++ exception"); }
    } catch (...) {
        return MR_ERR_OOM;             // even the message could not be built
    }
}
#define MR_API_BEGIN try {
#define MR_API_END(where) } catch (...) { return mr::translate_exception(where); }

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
    float cuni_frac = 0.0f;    // share of the uniform-current map's blocks that are uniform
    size_t max_pitch = 0;      // cudaDevAttrMaxPitch: the largest pitch cudaMemcpy2D accepts
    int64_t last_taken = 0;    // rays this device took from the slab queue in the last host-buffer call
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
// (rounded down) of a lower bound H of every depth the f32 bilinear of interpolator.rs:59-83 can return for a
// point the lookup assigns to one of the block's cells, over the f32 depths of all nodes those cells touch:
//   H = zmin - 1e-5 zmax - max(nx, ny) 2^-22 (zmax - zmin).
// Inside a cell the bilinear returns a value between its corners up to a few roundings of terms no larger than
// ~2 zmax: the first margin (< 1e-6 zmax needed).  The second covers extrapolation: the cell is floor() of an f32
// fractional index that carries a rounding error of up to i 2^-24 cells, so a point up to that far OUTSIDE cell i
// can be assigned to it (the index rounds up to the integer), its basis coordinate is then negative by that much
// (or exceeds 1), and the bilinear extrapolates beyond the corner values by up to ~2 i 2^-24 (zmax - zmin).
// A block with a node that is NaN, infinite or <= 0 gets 0: no bound, the kernel looks the depth up.
// *deep_frac: the share of blocks with H >= 550 m (kh >= 22 for periods up to ~10 s), what the automatic choice
// of MR_OPT_DEEP_MAP looks at.
static std::vector<float> depth_floor_map(const double *depth, int nx, int ny, int *nbx_out, int *nby_out, float *deep_frac)
{
    size_t n_deep = 0;
    const int B = kDeepBlock;
    const int nbx = (nx - 1 + B - 1) / B, nby = (ny - 1 + B - 1) / B;
    const double extrap = (double)std::max(nx, ny) * 0x1p-22;
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
            const double H = (double)zmin - 1e-5 * (double)zmax - extrap * ((double)zmax - (double)zmin);
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

// Uniform-current map of the fast path (FastRay, CMAP): for every block of kDeepBlock x kDeepBlock cells, {u, v} as
// f32 (`as f32`, cartesian_current.rs:378) if all nodes the block's cells touch hold one finite u and one finite v
// (compared as f64: only then are the f64 finite differences of cartesian_current.rs:522-536 exactly 0), else NaNs.
static std::vector<float> uniform_current_map(const double *u, const double *v, int nx, int ny, int *nbx_out, float *frac)
{
    const int B = kDeepBlock;
    const int nbx = (nx - 1 + B - 1) / B, nby = (ny - 1 + B - 1) / B;
    std::vector<float> out((size_t)nbx * (size_t)nby * 2, NAN);
    size_t n_uni = 0;
    for (int by = 0; by < nby; ++by) {
        const int j0 = by * B, j1 = std::min(j0 + B, ny - 1);
        for (int bx = 0; bx < nbx; ++bx) {
            const int i0 = bx * B, i1 = std::min(i0 + B, nx - 1);
            const double u0 = u[(size_t)j0 * (size_t)nx + (size_t)i0], v0 = v[(size_t)j0 * (size_t)nx + (size_t)i0];
            bool same = std::isfinite(u0) && std::isfinite(v0);
            for (int j = j0; j <= j1 && same; ++j)
                for (int i = i0; i <= i1; ++i)
                    if (u[(size_t)j * (size_t)nx + (size_t)i] != u0 || v[(size_t)j * (size_t)nx + (size_t)i] != v0) { same = false; break; }
            if (!same) continue;
            out[2 * ((size_t)by * (size_t)nbx + (size_t)bx)] = (float)u0;
            out[2 * ((size_t)by * (size_t)nbx + (size_t)bx) + 1] = (float)v0;
            ++n_uni;
        }
    }
    *nbx_out = nbx;
    *frac = (float)n_uni / (float)std::max<size_t>((size_t)nbx * (size_t)nby, 1);
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
    {
        int mp = 0;
        MR_CUDA(cudaDeviceGetAttribute(&mp, cudaDevAttrMaxPitch, d.dev));
        d.max_pitch = (size_t)std::max(mp, 0);
        if (const char *e = std::getenv("MR_DEBUG_MAX_PITCH")) d.max_pitch = (size_t)std::strtoull(e, nullptr, 10);   // tests: force the row-by-row gather
    }
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
        C.cmap = nullptr; C.cmap_nbx = 0;
        if (C.uniform) {
            int nbx = 0;
            std::vector<float> cm = uniform_current_map(c->u, c->v, c->nx, c->ny, &nbx, &d.cuni_frac);
            const float *dev = nullptr;
            if ((rc = upload(d, cm.data(), cm.size(), &dev))) return rc;
            C.cmap = reinterpret_cast<const float2 *>(dev);
            C.cmap_nbx = nbx;
        }
    }
    // Same-grid shortcut (FastRay, SG): the current lives on the bathymetry's grid — same shape, bit-identical f32
    // coordinates and basis, f64 coordinates that are exactly the f32 ones widened — so one f32 fractional index
    // serves both fields wherever it lies further than sg_delta from an integer.  Around the exact index I the
    // f32 one (position rounded to f32, f32 subtraction, correctly rounded f32 quotient) is off by at most
    // 2^-24 (3 n + |x0|/s) (1 + 2^-22), the f64 one by 2^-51 n; the distance itself is formed in f32 with an error
    // below 2^-25.  sg_delta is that bound with a 2 % margin plus 2^-22.
    B.same_grid = 0; B.sg_lim = 0.0f;
    if (b->kind == MR_BATHY_GRID && c->kind == MR_CURRENT_GRID && B.uniform && C.uniform &&
        b->nx == c->nx && b->ny == c->ny && B.p0 == C.p0 && B.d2 == C.d2 && B.c2 == C.c2 &&
        C.xd0 == (double)B.xf0 && C.yd0 == (double)B.yf0 && C.sx == (double)B.sx && C.sy == (double)B.sy &&
        C.x_space > 0.0 && C.y_space > 0.0) {
        const double dx = 0x1p-24 * (3.0 * b->nx + std::fabs((double)B.xf0) / (double)B.sx) * 1.02 + 0x1p-22;
        const double dy = 0x1p-24 * (3.0 * b->ny + std::fabs((double)B.yf0) / (double)B.sy) * 1.02 + 0x1p-22;
        const double delta = std::max(dx, dy);
        if (delta < 0.125) {
            float lim = (float)(0.5 - delta);
            if ((double)lim > 0.5 - delta) lim = std::nextafterf(lim, 0.0f);
            B.same_grid = 1; B.sg_lim = lim;
        }
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

// The uniform-current map: used when at least half of the current grid's blocks are uniform (a zero-current file,
// a piecewise-constant current) — elsewhere it only puts a dependent load in front of every record load.
// MR_OPT_CURRENT_MAP forces it on for any grid that has one, MR_OPT_NO_CURRENT_MAP off.
static constexpr float kCurrentMapAutoShare = 0.5f;
static int want_current_map(const DeviceFields &d, const mr_trace_opts &o)
{
    if (o.flags & MR_OPT_NO_CURRENT_MAP) return 0;
    return (o.flags & MR_OPT_CURRENT_MAP) != 0 || d.cuni_frac >= kCurrentMapAutoShare;
}

// The same-grid shortcut: on request wherever the grids coincide; by itself only where neither map is in use (beside
// a map it costs more than it saves, see include/mantaray_b200.h).
static int want_same_grid(const DeviceFields &d, const mr_trace_opts &o)
{
    if ((o.flags & MR_OPT_NO_SAME_GRID) || !d.b.same_grid) return 0;
    if (o.flags & MR_OPT_SAME_GRID) return 1;
    const bool dmap = want_deep_map(d, o) && d.b.dmap != nullptr;
    const bool cmap = want_current_map(d, o) && d.c.cmap != nullptr;
    return !dmap && !cmap;
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
    a.same_grid = want_same_grid(d, o);
    a.current_map = want_current_map(d, o);
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

// Rays are handed out in slabs from one shared cursor (the whole batch is one queue): a device takes the next
// slab when one of its two buffers is free, so a device whose rays stop early simply takes more slabs.  This is
// the dynamic balance rayon's work-stealing par_iter gives the reference (src/ray.rs:105-123); a static split
// into one contiguous block per device leaves whole period bands of a frequency/direction ensemble (C5) on one
// device.  Slabs are contiguous in input order, so the gather stays a concatenation of column blocks.
struct SlabQueue {
    std::atomic<int64_t> next{0};
    std::atomic<bool> failed{false};
    int64_t n = 0;
    int devices = 1;
};

// D2H copy of `height` rows of `width` bytes into a column block of a wider host plane.  cudaMemcpy2DAsync refuses
// pitches above cudaDevAttrMaxPitch (~2 GiB: more than 2.68e8 rays per row of doubles); beyond it the rows go one
// by one.
static cudaError_t copy_columns_d2h(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                                    size_t max_pitch, cudaStream_t s)
{
    if (dpitch <= max_pitch && spitch <= max_pitch)
        return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, s);
    for (size_t r = 0; r < height; ++r) {
        cudaError_t e = cudaMemcpyAsync((char *)dst + r * dpitch, (const char *)src + r * spitch, width, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// One device's worker: takes slabs of `chunk` rays from the queue until it is empty.  Slab k+1 integrates on the
// compute stream while slab k drains to the host on the copy stream (two device buffers).  Buffers, streams and
// events live in the handle's WorkArea and are reused by the next call.
static int trace_slabs_on_device(DeviceFields &d, const HostJob &j, SlabQueue &q, std::string &err)
{
    auto bail = [&](int code, const std::string &m) { err = m; q.failed.store(true); return code; };
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
    const int64_t n = q.n;
    if (n <= 0) return MR_OK;
    // MR_DEBUG_TIMING=1: where a host-buffer call spends its wall time (stderr)
    static const bool timing = std::getenv("MR_DEBUG_TIMING") != nullptr;
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    auto t_alloc = t_begin, t_enqueued = t_begin, t_synced = t_begin;
    const bool want_traj = j.x || j.y || j.kx || j.ky;
    const bool shared_queue = q.devices > 1;
    WorkArea &w = d.work;
    DevBuf buf[2];
    int nbuf = 1;
    int64_t chunk = n, taken = 0;
    int slabs = 0;
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
        const double budget = 0.80 * (double)(free_b + w.held());      // what this handle already holds is ours to reuse
        const int64_t wave = 148 * 512;                     // slab sizes are rounded to this many rays (four 128-ray groups per SM)
        if (j.o.chunk_rays > 0) {
            chunk = std::min<int64_t>(n, j.o.chunk_rays);
        } else {
            if (per_ray * (double)n > budget || (want_traj && per_ray * (double)n > 4e9)) {
                // does not fit, or is big enough that overlapping the drain with the next
                // slab's integration pays: two slabs in flight
                const double cap = std::min(budget / 2.0, 16e9);
                chunk = (int64_t)(cap / per_ray);
                if (chunk >= wave) chunk = chunk / wave * wave;
            }
            if (shared_queue) {
                // several devices on one queue: about sixteen slabs per device, so that the last slab a device
                // takes is a small share of its work — but a slab fills the machine (28 warps on each of 148 SMs)
                // unless the batch is too small to give every device that much
                const int64_t G = q.devices, full = kMachineRays;
                const int64_t per_dev = (n + G - 1) / G, target = (n + G * 16 - 1) / (G * 16);
                int64_t want = std::max(target, std::min(full, per_dev));
                want = (want + 127) / 128 * 128;
                chunk = std::min(chunk, want);
            }
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
        const size_t max_pitch = d.max_pitch;
        for (int k = 0;; ++k) {
            DevBuf &B = buf[k % nbuf];
            if (shared_queue) {
                // With other devices on the queue a slab is taken only when this device can start on it: the
                // previous slab has been integrated (it may still be draining) and this buffer's previous slab has
                // left it.  That is what balances the devices; the gap it leaves between two kernels is a launch
                // latency against tens of milliseconds of integration.
                if (k >= 1) MR_TRY(cudaEventSynchronize(buf[(k - 1) % nbuf].computed));
                if (k >= nbuf) MR_TRY(cudaEventSynchronize(B.drained));
            } else if (k >= nbuf) {
                // alone on the queue the host runs ahead and the compute stream waits for the buffer
                MR_TRY(cudaStreamWaitEvent(s_comp, B.drained, 0));
            }
            if (q.failed.load()) break;
            const int64_t c0 = q.next.fetch_add(chunk);
            if (c0 >= n) break;
            const int64_t m = std::min(chunk, n - c0);
            taken += m; ++slabs;
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
                a.same_grid = want_same_grid(d, j.o);
                a.current_map = want_current_map(d, j.o);
                // fin is [4][n] with n = m for the kernel (it uses a.n as the pitch)
                cudaError_t e = j.o.math == MR_MATH_STRICT ? launch_trace_strict(a, s_comp) : launch_trace_fast(a, s_comp);
                if (e != cudaSuccess) { rc = bail(MR_ERR_CUDA, std::string("trace kernel launch: ") + cudaGetErrorString(e)); goto done; }
            }
            if (B.depth || B.u || B.v)       // the fields at every stored state, from the planes just written
                MR_TRY(launch_sample(d.b, d.c, j.rows_cap, m, chunk, tx, tx + plane, B.depth, B.u, B.v, s_comp));
            MR_TRY(cudaEventRecord(B.computed, s_comp));
            MR_TRY(cudaStreamWaitEvent(s_copy, B.computed, 0));
            const size_t rows_cap = (size_t)j.rows_cap, nt = (size_t)j.n_total, ch = (size_t)chunk, mm = (size_t)m;
            if (j.env.depth)
                MR_TRY(copy_columns_d2h(j.env.depth + c0, sizeof(float) * nt, B.depth, sizeof(float) * ch, sizeof(float) * mm, rows_cap, max_pitch, s_copy));
            if (j.env.u)
                MR_TRY(copy_columns_d2h(j.env.u + c0, sizeof(double) * nt, B.u, sizeof(double) * ch, sizeof(double) * mm, rows_cap, max_pitch, s_copy));
            if (j.env.v)
                MR_TRY(copy_columns_d2h(j.env.v + c0, sizeof(double) * nt, B.v, sizeof(double) * ch, sizeof(double) * mm, rows_cap, max_pitch, s_copy));
            if (want_traj) {
                double *dsts[4] = { j.x, j.y, j.kx, j.ky };
                for (int f = 0; f < 4; ++f) {
                    if (!dsts[f]) continue;
                    MR_TRY(copy_columns_d2h(dsts[f] + c0, sizeof(double) * nt, B.traj + f * plane, sizeof(double) * ch,
                                            sizeof(double) * mm, rows_cap, max_pitch, s_copy));
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
        std::fprintf(stderr, "[mantaray_b200] device %d: %lld of %lld rays in %d slabs of %lld (%d buffers): alloc %.1f ms, enqueue %.1f ms, "
                             "drain %.1f ms, tail %.1f ms, %.2f GB held\n", d.dev, (long long)taken, (long long)n, slabs, (long long)chunk, nbuf,
                     ms(t_begin, t_alloc), ms(t_alloc, t_enqueued), ms(t_enqueued, t_synced), ms(t_synced, clk::now()),
                     (double)w.held() / 1e9);
    }
    d.last_taken = taken;
    return rc;
#undef MR_TRY
}

// The whole batch on the handle's devices: one worker thread per device on a shared slab queue.
static int run_host_job(mr_fields *f, const HostJob &j, int max_devices)
{
    const int G = std::max(1, std::min((int)f->devs.size(), max_devices));
    SlabQueue q;
    q.n = j.n_total; q.devices = G;
    std::vector<int> rcs((size_t)G, MR_OK);
    std::vector<std::string> errs((size_t)G);
    for (auto &d : f->devs) d.last_taken = 0;
    if (G == 1) {
        rcs[0] = trace_slabs_on_device(f->devs[0], j, q, errs[0]);
    } else {
        std::vector<std::thread> th;
        th.reserve((size_t)G);
        auto worker = [&](int g) {
            try {
                rcs[(size_t)g] = trace_slabs_on_device(f->devs[(size_t)g], j, q, errs[(size_t)g]);
            } catch (...) {                 // nothing may leave a thread: it would be std::terminate
                rcs[(size_t)g] = translate_exception("device worker");
                try { errs[(size_t)g] = g_err; } catch (...) {}
                q.failed.store(true);
            }
        };
        int started = 0;
        try {
            for (int g = 1; g < G; ++g) { th.emplace_back(worker, g); ++started; }
        } catch (...) {
            // a thread could not be created: the devices that did start (and this thread) share the queue
            translate_exception("mr_trace_many");
        }
        (void)started;
        worker(0);                          // the calling thread drives the first device
        for (auto &t_ : th) t_.join();
    }
    for (int g = 0; g < G; ++g)
        if (rcs[(size_t)g] != MR_OK) return fail(rcs[(size_t)g], "device " + std::to_string(f->devs[(size_t)g].dev) + ": " + errs[(size_t)g]);
    if (q.next.load() < j.n_total) return fail(MR_ERR_CUDA, "mr_trace_many: the slab queue was abandoned before the last ray");
    return MR_OK;
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
    MR_API_BEGIN
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
    MR_API_END("mr_fields_create")
}

int mr_fields_open_netcdf3(const char *bathymetry_path, const char *current_path,
                           uint32_t device_mask, mr_fields **out)
{
    MR_API_BEGIN
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
    MR_API_END("mr_fields_open_netcdf3")
}

void mr_fields_free(mr_fields *f)
{
    if (!f) return;
    for (auto &d : f->devs) free_device_fields(d);
    delete f;
}

uint32_t mr_fields_device_mask(const mr_fields *f) { return f ? f->mask : 0; }

int mr_fields_last_split(mr_fields *f, int64_t *rays_per_device, int32_t cap)
{
    if (!f || (cap > 0 && !rays_per_device)) return fail(MR_ERR_BAD_ARG, "mr_fields_last_split: NULL argument");
    std::lock_guard<std::mutex> guard(f->mu);
    const int n = (int)f->devs.size();
    for (int g = 0; g < n && g < cap; ++g) rays_per_device[g] = f->devs[(size_t)g].last_taken;
    return n;
}

int mr_trace_plan(const mr_fields *f, const mr_trace_opts *opts)
{
    MR_API_BEGIN
    if (!f || f->devs.empty()) return fail(MR_ERR_BAD_ARG, "mr_trace_plan: NULL handle");
    mr_trace_opts o = opts ? *opts : mr_trace_opts{1, MR_MATH_FAST, 0, 0};
    if (o.math != MR_MATH_FAST && o.math != MR_MATH_STRICT) return fail(MR_ERR_BAD_ARG, "mr_trace_opts.math must be MR_MATH_FAST or MR_MATH_STRICT");
    const DeviceFields &d = f->devs[0];          // every device holds the same fields
    TraceArgs a{};
    a.b = d.b; a.c = d.c;
    a.deep_map = want_deep_map(d, o); a.same_grid = want_same_grid(d, o); a.current_map = want_current_map(d, o);
    const TracePlan p = plan_of(a, o.math == MR_MATH_FAST);
    return (p.uni ? MR_PLAN_AFFINE : 0) | (p.dmap ? MR_PLAN_DEEP_MAP : 0) | (p.sg ? MR_PLAN_SAME_GRID : 0) |
           (p.cmap ? MR_PLAN_CURRENT_MAP : 0);
    MR_API_END("mr_trace_plan")
}

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
    if (std::isnan(q) || !(q < 2147483646.0)) return -1;
    if (q < 0.0) return 0;             // `((x_end - x) / h).ceil() as usize` saturates a negative quotient to 0 steps
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
    MR_API_BEGIN
    if (!f) return fail(MR_ERR_BAD_ARG, "mr_trace_many: NULL field handle");
    if (n < 0) return fail(MR_ERR_BAD_ARG, "mr_trace_many: n < 0");
    if (n > 0 && (!x0 || !y0 || !kx0 || !ky0)) return fail(MR_ERR_BAD_ARG, "mr_trace_many: NULL initial-condition array");
    mr_trace_opts o;
    normalise_opts(opts, o);
    if (o.math != MR_MATH_FAST && o.math != MR_MATH_STRICT) return fail(MR_ERR_BAD_ARG, "mr_trace_opts.math must be MR_MATH_FAST or MR_MATH_STRICT");
    const int64_t nsteps = mr_num_steps(t0, t_end, dt);
    if (nsteps < 0) return fail(MR_ERR_BAD_ARG, "mr_trace_many: need dt > 0 and a finite (t_end - t0)/dt < 2^31");
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
    return run_host_job(f, j, (int)f->devs.size());
    MR_API_END("mr_trace_many")
}

int mr_single_ray(mr_fields *f, double x0, double y0, double kx0, double ky0,
                  double t0, double t_end, double dt, const mr_trace_opts *opts,
                  double *out, int64_t out_cap, int64_t *n_rows)
{
    MR_API_BEGIN
    if (!f || !n_rows) return fail(MR_ERR_BAD_ARG, "mr_single_ray: NULL argument");
    mr_trace_opts o;
    normalise_opts(opts, o);
    if (o.stride != 1) return fail(MR_ERR_BAD_ARG, "mr_single_ray: stride must be 1");
    const int64_t nsteps = mr_num_steps(t0, t_end, dt);
    if (nsteps < 0) return fail(MR_ERR_BAD_ARG, "mr_single_ray: need dt > 0 and a finite (t_end - t0)/dt < 2^31");
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
        int rc = run_host_job(f, j, 1);
        if (rc) return rc;
    }
    *n_rows = rows;
    if (rows > out_cap || !out) return fail(MR_ERR_BAD_ARG, "mr_single_ray: out holds " + std::to_string(out_cap) +
                                                        " rows, " + std::to_string(rows) + " needed");
    for (int64_t r = 0; r < rows; ++r) {    // (t, x, y, kx, ky) tuples, src/ffi.rs:42-47
        out[5 * r + 0] = t[(size_t)r];
        for (int c = 0; c < 4; ++c) out[5 * r + 1 + c] = soa[(size_t)c * cap + r];
    }
    return MR_OK;
    MR_API_END("mr_single_ray")
}

int mr_trace_device(mr_fields *f, int device, void *stream, int64_t n,
                    const double *d_x0, const double *d_y0, const double *d_kx0, const double *d_ky0,
                    double t0, double t_end, double dt, const mr_trace_opts *opts,
                    double *d_x, double *d_y, double *d_kx, double *d_ky, int64_t ld,
                    int32_t *d_rows, int32_t *d_len, double *d_final, int32_t *launches)
{
    MR_API_BEGIN
    if (launches) *launches = 0;
    if (!f) return fail(MR_ERR_BAD_ARG, "mr_trace_device: NULL field handle");
    const DeviceFields *d = find_device(f, device);
    if (!d) return fail(MR_ERR_BAD_ARG, "mr_trace_device: device " + std::to_string(device) + " is not in the handle's mask");
    if (n < 0) return fail(MR_ERR_BAD_ARG, "mr_trace_device: n < 0");
    if (n > 0 && (!d_x0 || !d_y0 || !d_kx0 || !d_ky0)) return fail(MR_ERR_BAD_ARG, "mr_trace_device: NULL initial-condition array");
    mr_trace_opts o;
    normalise_opts(opts, o);
    const int64_t nsteps = mr_num_steps(t0, t_end, dt);
    if (nsteps < 0) return fail(MR_ERR_BAD_ARG, "mr_trace_device: need dt > 0 and a finite (t_end - t0)/dt < 2^31");
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
    MR_API_END("mr_trace_device")
}

int mr_sample_device(mr_fields *f, int device, void *stream, int64_t rows, int64_t n, int64_t ld,
                     const double *d_x, const double *d_y,
                     float *d_depth, double *d_u, double *d_v, int32_t *launches)
{
    MR_API_BEGIN
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
    MR_API_END("mr_sample_device")
}

int mr_sample_fields(mr_fields *f, int64_t count, const double *x, const double *y,
                     float *depth, double *u, double *v)
{
    MR_API_BEGIN
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
    MR_API_END("mr_sample_fields")
}

int mr_depth_floor_map(const mr_bathymetry_desc *b, float *out, size_t cap, int32_t *nbx, int32_t *nby, float *deep_frac,
                       int32_t *affine)
{
    MR_API_BEGIN
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
    MR_API_END("mr_depth_floor_map")
}

int mr_uniform_current_map(const mr_current_desc *c, float *out, size_t cap, int32_t *nbx, int32_t *nby,
                           float *uniform_frac, int32_t *affine)
{
    MR_API_BEGIN
    if (!c || !nbx || !nby) return fail(MR_ERR_BAD_ARG, "mr_uniform_current_map: NULL argument");
    if (c->kind != MR_CURRENT_GRID || c->nx < 2 || c->ny < 2 || !c->x || !c->y || !c->u || !c->v)
        return fail(MR_ERR_BAD_ARG, "mr_uniform_current_map: needs a GRID current with nx, ny >= 2");
    if (affine) {       // the same test upload_fields applies before it builds (and the kernel consults) the map
        std::vector<float> xf((size_t)c->nx), yf((size_t)c->ny);
        for (int i = 0; i < c->nx; ++i) xf[(size_t)i] = (float)c->x[i];
        for (int i = 0; i < c->ny; ++i) yf[(size_t)i] = (float)c->y[i];
        float dxf, dyf, c01, c10;
        *affine = affine_f32(xf.data(), c->nx, &dxf) && affine_f32(yf.data(), c->ny, &dyf) && basis_coeffs(dxf, dyf, &c01, &c10);
    }
    int bx = 0;
    float frac = 0.0f;
    const std::vector<float> m = uniform_current_map(c->u, c->v, c->nx, c->ny, &bx, &frac);
    *nbx = bx; *nby = bx ? (int32_t)(m.size() / 2 / (size_t)bx) : 0;
    if (uniform_frac) *uniform_frac = frac;
    if (out) {
        if (cap < m.size()) return fail(MR_ERR_BAD_ARG, "mr_uniform_current_map: output too small");
        std::memcpy(out, m.data(), m.size() * sizeof(float));
    }
    return MR_OK;
    MR_API_END("mr_uniform_current_map")
}

// ---- NetCDF-3 ---------------------------------------------------------------
int mr_nc3_open(const char *path, mr_nc3 **out)
{
    MR_API_BEGIN
    if (!path || !out) return fail(MR_ERR_BAD_ARG, "mr_nc3_open: NULL argument");
    *out = nullptr;
    std::unique_ptr<mr_nc3> h(new (std::nothrow) mr_nc3);
    if (!h) return fail(MR_ERR_OOM, "out of host memory");
    std::string err;
    int rc = Nc3File::open(path, h->file, err);
    if (rc) return fail(rc, err);
    *out = h.release();
    return MR_OK;
    MR_API_END("mr_nc3_open")
}
void mr_nc3_close(mr_nc3 *f) { delete f; }
int mr_nc3_var_count(const mr_nc3 *f) { return f ? (int)f->file.vars.size() : 0; }
int mr_nc3_var_name(const mr_nc3 *f, int index, char *buf, size_t cap)
{
    MR_API_BEGIN
    if (!f || !buf || cap == 0 || index < 0 || index >= (int)f->file.vars.size()) return fail(MR_ERR_BAD_ARG, "mr_nc3_var_name: bad argument");
    std::strncpy(buf, f->file.vars[(size_t)index].name.c_str(), cap - 1);
    buf[cap - 1] = 0;
    return MR_OK;
    MR_API_END("mr_nc3_var_name")
}
int mr_nc3_var_info(const mr_nc3 *f, const char *name, int32_t *nc_type, int64_t *n_elems, int32_t *ndims, int64_t dims[MR_NC3_MAX_DIMS])
{
    MR_API_BEGIN
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
    MR_API_END("mr_nc3_var_info")
}
int mr_nc3_read_f32(const mr_nc3 *f, const char *name, float *out, int64_t cap)
{
    MR_API_BEGIN
    if (!f || !name || !out) return fail(MR_ERR_BAD_ARG, "mr_nc3_read_f32: NULL argument");
    std::vector<float> v;
    std::string err;
    int rc = f->file.read_f32(name, v, err);
    if (rc) return fail(rc, err);
    if ((int64_t)v.size() > cap) return fail(MR_ERR_BAD_ARG, "mr_nc3_read_f32: buffer too small");
    std::memcpy(out, v.data(), v.size() * sizeof(float));
    return MR_OK;
    MR_API_END("mr_nc3_read_f32")
}
int mr_nc3_read_f64(const mr_nc3 *f, const char *name, double *out, int64_t cap)
{
    MR_API_BEGIN
    if (!f || !name || !out) return fail(MR_ERR_BAD_ARG, "mr_nc3_read_f64: NULL argument");
    std::vector<double> v;
    std::string err;
    int rc = f->file.read_f64(name, v, err);
    if (rc) return fail(rc, err);
    if ((int64_t)v.size() > cap) return fail(MR_ERR_BAD_ARG, "mr_nc3_read_f64: buffer too small");
    std::memcpy(out, v.data(), v.size() * sizeof(double));
    return MR_OK;
    MR_API_END("mr_nc3_read_f64")
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
