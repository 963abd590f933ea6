// mr_internal.hpp — host-side declarations shared by the translation units of
// libmantaray_b200.so.  Not part of the public ABI.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mantaray_b200.h"

namespace mr {

// thread-local last-error text (mr_last_error)
void set_error(const std::string &msg);
int  fail(int code, const std::string &msg);

// RN(1/s) for the fast path's exact f32 division by a launch constant (fdiv_const, mr_device.cuh); refused for
// subnormal / huge s and for the all-ones significand that Markstein's theorem excludes.  Shared with the
// self-test in tools/csrc/mr_tools.cu, which must apply the library's own rule.
inline bool recip_ok(float s, float *r)
{
    uint32_t bits;
    std::memcpy(&bits, &s, 4);
    if (!(s > 1e-30f) || !(s < 1e30f) || (bits & 0x7fffffu) == 0x7fffffu) return false;
    volatile float q = 1.0f / s;
    *r = q;
    return true;
}

// ---- NetCDF-3 -------------------------------------------------------------
enum { NC3_BYTE = 1, NC3_CHAR = 2, NC3_SHORT = 3, NC3_INT = 4, NC3_FLOAT = 5, NC3_DOUBLE = 6 };

struct Nc3Dim { std::string name; uint32_t len = 0; };
struct Nc3Var {
    std::string name;
    std::vector<uint32_t> dimids;
    int type = 0;
    uint32_t vsize = 0;
    uint64_t begin = 0;
    bool is_record = false;
    uint64_t elems_per_chunk = 0;      // elements per record (record var) or in total
};
struct Nc3File {
    std::string path;
    uint64_t file_size = 0;
    int version = 0;
    uint32_t numrecs = 0;
    uint64_t recsize = 0;
    std::vector<Nc3Dim> dims;
    std::vector<Nc3Var> vars;

    static int open(const char *path, Nc3File &f, std::string &err);
    const Nc3Var *find(const std::string &name) const;
    uint64_t num_elems(const Nc3Var &v) const;
    int read_raw(const Nc3Var &v, std::vector<uint8_t> &raw, std::string &err) const;
    int read_f32(const std::string &name, std::vector<float> &out, std::string &err) const;
    int read_f64(const std::string &name, std::vector<double> &out, std::string &err) const;
};

}  // namespace mr
