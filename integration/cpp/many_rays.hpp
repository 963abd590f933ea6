// many_rays.hpp — C++ host-side mirror of mantaray's batch driver over the C ABI.
//
// Same names and argument meaning as the Rust types it stands for: ManyRays / SingleRay
// (src/ray.rs:24-214), RayState (src/datatype.rs:117-138), the field structs ConstantDepth,
// ConstantSlope, CartesianNetcdf3, ArrayDepth (src/bathymetry/*.rs), ConstantCurrent, CartesianCurrent
// (src/current/*.rs).  The reference's toolchain (Rust) is not in this image, so the host side above
// the C ABI is written in C++ where the reference is compiled code.  Header-only; link with
// -lmantaray_b200.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mantaray_b200.h"

namespace mantaray {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc)
{
    if (rc != MR_OK) throw Error(rc, mr_last_error());
}

struct RayState { double x, y, kx, ky; };                       // src/datatype.rs:117-138, order of :140-150

// ---- BathymetryData implementors ---------------------------------------------------------------
struct BathymetryData {
    virtual ~BathymetryData() = default;
    virtual mr_bathymetry_desc descriptor() const = 0;
};
struct ConstantDepth : BathymetryData {                          // constant_depth.rs:15-18 (default 1000, :16)
    float h;
    explicit ConstantDepth(float h_ = 1000.0f) : h(h_) {}
    mr_bathymetry_desc descriptor() const override { mr_bathymetry_desc d{}; d.kind = MR_BATHY_CONSTANT; d.h0 = h; return d; }
};
struct ConstantSlope : BathymetryData {                          // constant_slope.rs:28-44 defaults
    float h0 = 50.0f, x0 = 0.0f, y0 = 0.0f, dhdx = -5e-2f, dhdy = 0.0f;
    mr_bathymetry_desc descriptor() const override
    {
        mr_bathymetry_desc d{}; d.kind = MR_BATHY_SLOPE; d.h0 = h0; d.x0 = x0; d.y0 = y0; d.dhdx = dhdx; d.dhdy = dhdy; return d;
    }
};
struct CartesianNetcdf3 : BathymetryData {                       // cartesian_netcdf3.rs:35-43
    std::vector<float> x, y;
    std::vector<double> depth;                                   // [ny*nx], depth[nx*yi+xi]
    mr_bathymetry_desc descriptor() const override
    {
        mr_bathymetry_desc d{}; d.kind = MR_BATHY_GRID; d.nx = (int32_t)x.size(); d.ny = (int32_t)y.size();
        d.x = x.data(); d.y = y.data(); d.depth = depth.data(); return d;
    }
};
struct ArrayDepth : BathymetryData {                             // array_depth.rs:9-11
    int nx = 0, ny = 0;
    std::vector<float> array;                                    // array[xi*ny + yi]
    mr_bathymetry_desc descriptor() const override
    {
        mr_bathymetry_desc d{}; d.kind = MR_BATHY_ARRAY; d.nx = nx; d.ny = ny; d.array = array.data(); return d;
    }
};

// ---- CurrentData implementors ---------------------------------------------------------------------
struct CurrentData {
    virtual ~CurrentData() = default;
    virtual mr_current_desc descriptor() const = 0;
};
struct ConstantCurrent : CurrentData {                           // constant_current.rs:13-17
    double u, v;
    ConstantCurrent(double u_ = 0.0, double v_ = 0.0) : u(u_), v(v_) {}
    mr_current_desc descriptor() const override { mr_current_desc d{}; d.kind = MR_CURRENT_CONSTANT; d.u0 = u; d.v0 = v; return d; }
};
struct CartesianCurrent : CurrentData {                          // cartesian_current.rs:18-27
    std::vector<double> x, y, u, v;
    mr_current_desc descriptor() const override
    {
        mr_current_desc d{}; d.kind = MR_CURRENT_GRID; d.nx = (int32_t)x.size(); d.ny = (int32_t)y.size();
        d.x = x.data(); d.y = y.data(); d.u = u.data(); d.v = v.data(); return d;
    }
};

/// One ray's stored rows: t[r] and (x, y, kx, ky)[r] — `SolverResult::get()`.
struct SolverResult {
    std::vector<double> t;
    std::vector<RayState> states;
};

/// ManyRays (src/ray.rs:24-127).
class ManyRays {
public:
    ManyRays(const BathymetryData &b, const CurrentData &c, std::vector<RayState> initial_rays, uint32_t device_mask = 0)
        : rays_(std::move(initial_rays))
    {
        auto bd = b.descriptor();
        auto cd = c.descriptor();
        check(mr_fields_create(&bd, &cd, device_mask, &fields_));
    }
    ~ManyRays() { mr_fields_free(fields_); }
    ManyRays(const ManyRays &) = delete;
    ManyRays &operator=(const ManyRays &) = delete;

    std::vector<SolverResult> trace_many(double start_time, double end_time, double step_size) const
    {
        const int64_t n = (int64_t)rays_.size();
        const int64_t cap = mr_num_rows(start_time, end_time, step_size, 1);
        if (cap < 0) throw Error(MR_ERR_BAD_ARG, "bad time arguments");
        std::vector<double> x0(n), y0(n), kx0(n), ky0(n), t(cap), X(cap * n), Y(cap * n), KX(cap * n), KY(cap * n);
        std::vector<int32_t> rows(n);
        for (int64_t i = 0; i < n; ++i) { x0[i] = rays_[i].x; y0[i] = rays_[i].y; kx0[i] = rays_[i].kx; ky0[i] = rays_[i].ky; }
        check(mr_trace_many(fields_, n, x0.data(), y0.data(), kx0.data(), ky0.data(), start_time, end_time, step_size, nullptr,
                            t.data(), X.data(), Y.data(), KX.data(), KY.data(), rows.data(), nullptr, nullptr));
        std::vector<SolverResult> out(n);
        for (int64_t i = 0; i < n; ++i) {
            out[i].t.assign(t.begin(), t.begin() + rows[i]);
            for (int32_t r = 0; r < rows[i]; ++r) out[i].states.push_back({X[r * n + i], Y[r * n + i], KX[r * n + i], KY[r * n + i]});
        }
        return out;
    }

private:
    std::vector<RayState> rays_;
    mr_fields *fields_ = nullptr;
};

/// SingleRay (src/ray.rs:130-214).
class SingleRay {
public:
    SingleRay(const BathymetryData &b, const CurrentData &c, RayState initial_ray) : many_(b, c, {initial_ray}) {}
    SolverResult trace_individual(double start_time, double end_time, double step_size) const
    {
        return many_.trace_many(start_time, end_time, step_size)[0];
    }

private:
    ManyRays many_;
};

}  // namespace mantaray
