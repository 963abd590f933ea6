// mr_netcdf3.cpp — native NetCDF-3 reader (classic CDF-1 and 64-bit-offset CDF-2).
//
// Replaces what the reference gets from the `netcdf3` crate in
// CartesianNetcdf3::open (src/bathymetry/cartesian_netcdf3.rs:167-256) and
// CartesianCurrent::open (src/current/cartesian_current.rs:58-213):
// FileReader::open + read_var(name) + a dtype switch over
// {I8,U8,I16,I32,F32,F64} that casts the whole variable to f32 or f64.
// As there, the dimension order recorded in the file is NOT consulted: a
// variable is returned as the flat array the file stores.
//
// Host-only code; no CUDA here.
#include "mr_internal.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>

namespace mr {

namespace {

constexpr uint32_t NC_DIMENSION = 0x0A, NC_VARIABLE = 0x0B, NC_ATTRIBUTE = 0x0C;

struct Cursor {
    const std::vector<uint8_t> &buf;
    size_t pos = 0;
    bool ok = true;
    explicit Cursor(const std::vector<uint8_t> &b) : buf(b) {}
    bool need(size_t n)
    {
        if (!ok || pos + n > buf.size()) { ok = false; return false; }
        return true;
    }
    uint32_t u32()
    {
        if (!need(4)) return 0;
        uint32_t v = (uint32_t)buf[pos] << 24 | (uint32_t)buf[pos + 1] << 16 | (uint32_t)buf[pos + 2] << 8 | buf[pos + 3];
        pos += 4;
        return v;
    }
    uint64_t u64()
    {
        uint64_t hi = u32(), lo = u32();
        return hi << 32 | lo;
    }
    std::string name()
    {
        uint32_t len = u32();
        if (!need(len)) return std::string();
        std::string s((const char *)&buf[pos], len);
        pos += (len + 3u) & ~3u;          // padded to 4 bytes
        if (pos > buf.size()) ok = false;
        return s;
    }
    void skip(size_t n)
    {
        if (need(n)) pos += n;
    }
};

// a * b without wrapping: false when the product does not fit in 64 bits
bool mul_fits(uint64_t a, uint64_t b, uint64_t &out)
{
    if (a != 0 && b > UINT64_MAX / a) return false;
    out = a * b;
    return true;
}

size_t type_size(int t)
{
    switch (t) {
    case NC3_BYTE: case NC3_CHAR: return 1;
    case NC3_SHORT: return 2;
    case NC3_INT: case NC3_FLOAT: return 4;
    case NC3_DOUBLE: return 8;
    default: return 0;
    }
}

// attribute list: tag, count, then (name, type, nelems, padded values)
bool skip_att_list(Cursor &c)
{
    uint32_t tag = c.u32(), n = c.u32();
    if (!c.ok) return false;
    if (tag == 0 && n == 0) return true;                 // ABSENT
    if (tag != NC_ATTRIBUTE) return false;
    for (uint32_t i = 0; i < n && c.ok; ++i) {
        c.name();
        uint32_t t = c.u32(), ne = c.u32();
        size_t ts = type_size((int)t);
        if (ts == 0) return false;
        c.skip(((size_t)ne * ts + 3u) & ~(size_t)3u);
    }
    return c.ok;
}

}  // namespace

int Nc3File::open(const char *path, Nc3File &f, std::string &err)
{
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) { err = std::string("cannot open '") + path + "'"; return MR_ERR_IO; }
    std::streamoff size = in.tellg();
    if (size < 0) { err = std::string("cannot stat '") + path + "'"; return MR_ERR_IO; }
    f.path = path;
    f.file_size = (uint64_t)size;

    // The header is small; read up to 4 MiB of it (grown if the var list needs more).
    size_t hdr_cap = (size_t)std::min<uint64_t>(f.file_size, 4u << 20);
    for (;;) {
        std::vector<uint8_t> hdr(hdr_cap);
        in.seekg(0);
        in.read((char *)hdr.data(), (std::streamsize)hdr_cap);
        if (!in) { err = std::string("short read on '") + path + "'"; return MR_ERR_IO; }
        Cursor c(hdr);
        if (hdr_cap < 4 || hdr[0] != 'C' || hdr[1] != 'D' || hdr[2] != 'F') {
            err = std::string("'") + path + "' is not a NetCDF-3 file (bad magic)";
            return MR_ERR_FORMAT;
        }
        f.version = hdr[3];
        if (f.version != 1 && f.version != 2) {
            err = std::string("'") + path + "': unsupported NetCDF version byte " + std::to_string(f.version) +
                  " (classic=1 and 64-bit offset=2 are supported, as in the netcdf3 crate)";
            return MR_ERR_FORMAT;
        }
        c.pos = 4;
        f.numrecs = c.u32();
        f.dims.clear();
        f.vars.clear();
        bool good = true;
        // dim_list
        {
            uint32_t tag = c.u32(), n = c.u32();
            if (!(tag == 0 && n == 0)) {
                if (tag != NC_DIMENSION) good = false;
                for (uint32_t i = 0; good && i < n && c.ok; ++i) {
                    Nc3Dim d;
                    d.name = c.name();
                    d.len = c.u32();
                    f.dims.push_back(d);
                }
            }
        }
        if (good) good = skip_att_list(c);
        // var_list
        if (good) {
            uint32_t tag = c.u32(), n = c.u32();
            if (!(tag == 0 && n == 0)) {
                if (tag != NC_VARIABLE) good = false;
                for (uint32_t i = 0; good && i < n && c.ok; ++i) {
                    Nc3Var v;
                    v.name = c.name();
                    uint32_t nd = c.u32();
                    if (nd > 1024) { good = false; break; }
                    for (uint32_t k = 0; k < nd; ++k) v.dimids.push_back(c.u32());
                    if (!skip_att_list(c)) { good = false; break; }
                    v.type = (int)c.u32();
                    v.vsize = c.u32();
                    v.begin = f.version == 2 ? c.u64() : (uint64_t)c.u32();
                    f.vars.push_back(v);
                }
            }
        }
        if (!c.ok && hdr_cap < f.file_size) {            // header longer than what we read
            hdr_cap = (size_t)std::min<uint64_t>(f.file_size, (uint64_t)hdr_cap * 4);
            continue;
        }
        if (!good || !c.ok) {
            err = std::string("'") + path + "': malformed NetCDF-3 header";
            return MR_ERR_FORMAT;
        }
        break;
    }

    // record bookkeeping
    f.recsize = 0;
    int nrecvars = 0;
    for (auto &v : f.vars) {
        for (uint32_t id : v.dimids)
            if (id >= f.dims.size()) { err = "'" + f.path + "': variable '" + v.name + "' has a bad dimension id"; return MR_ERR_FORMAT; }
        v.is_record = !v.dimids.empty() && f.dims[v.dimids[0]].len == 0;
        size_t ts = type_size(v.type);
        if (ts == 0) { err = "'" + f.path + "': variable '" + v.name + "' has an unknown type"; return MR_ERR_FORMAT; }
        // Element and byte counts come from header fields a corrupt file controls: every product is checked, and a
        // variable larger than the file itself is refused here, before anything is sized by it.
        uint64_t per = 1;
        bool fits = true;
        for (size_t k = v.is_record ? 1 : 0; fits && k < v.dimids.size(); ++k)
            fits = mul_fits(per, f.dims[v.dimids[k]].len, per);
        uint64_t chunk_bytes = 0;
        fits = fits && mul_fits(per, ts, chunk_bytes);
        // (a record variable is bounded below through numrecs * recsize: with no records yet it may begin at the
        // very end of the file)
        if (!fits || v.begin > f.file_size || (!v.is_record && chunk_bytes > f.file_size - v.begin)) {
            err = "'" + f.path + "': variable '" + v.name + "' is larger than the file (corrupt header?)";
            return MR_ERR_FORMAT;
        }
        v.elems_per_chunk = per;
        if (v.is_record) {
            nrecvars++;
            const uint64_t padded = (chunk_bytes + 3u) & ~(uint64_t)3u;
            if (f.recsize + padded < f.recsize) { err = "'" + f.path + "': record size overflows"; return MR_ERR_FORMAT; }
            f.recsize += padded;
        }
    }
    if (nrecvars == 1) {                                  // a lone record variable is not padded
        for (auto &v : f.vars)
            if (v.is_record) f.recsize = v.elems_per_chunk * type_size(v.type);
    }
    if (f.numrecs == 0xFFFFFFFFu) {                       // "streaming" marker: derive from the file size
        uint64_t first = UINT64_MAX;
        for (auto &v : f.vars) if (v.is_record) first = std::min(first, v.begin);
        f.numrecs = (first != UINT64_MAX && f.recsize && first <= f.file_size)
                        ? (uint32_t)std::min<uint64_t>((f.file_size - first) / f.recsize, 0xFFFFFFFEu) : 0;
    }
    // the records a header claims must exist in the file too
    if (nrecvars > 0 && f.numrecs > 0) {
        uint64_t all = 0;
        if (!mul_fits(f.recsize, f.numrecs, all) || all > f.file_size) {
            err = "'" + f.path + "': " + std::to_string(f.numrecs) + " records of " + std::to_string(f.recsize) +
                  " bytes do not fit in the file (corrupt header?)";
            return MR_ERR_FORMAT;
        }
    }
    return MR_OK;
}

const Nc3Var *Nc3File::find(const std::string &name) const
{
    for (auto &v : vars) if (v.name == name) return &v;
    return nullptr;
}

uint64_t Nc3File::num_elems(const Nc3Var &v) const
{
    return v.is_record ? v.elems_per_chunk * numrecs : v.elems_per_chunk;
}

// Reads the whole variable as raw big-endian bytes, records de-interleaved.
int Nc3File::read_raw(const Nc3Var &v, std::vector<uint8_t> &raw, std::string &err) const
{
    size_t ts = type_size(v.type);
    uint64_t n = num_elems(v), bytes = 0;
    // extent against the file BEFORE anything is allocated (open() has already refused headers whose products
    // wrap; this is the per-variable bound, records included)
    if (!mul_fits(n, ts, bytes) || bytes > file_size || v.begin > file_size ||
        (!v.is_record && bytes > file_size - v.begin)) {
        err = "'" + path + "': variable '" + v.name + "' runs past the end of the file";
        return MR_ERR_FORMAT;
    }
    if (v.is_record && numrecs > 0) {
        const uint64_t chunk = v.elems_per_chunk * ts;      // fits: checked in open()
        uint64_t span = 0;
        if (!mul_fits(recsize, (uint64_t)numrecs - 1, span) || span > file_size - v.begin || chunk > file_size - v.begin - span) {
            err = "'" + path + "': record variable '" + v.name + "' runs past the end of the file";
            return MR_ERR_FORMAT;
        }
    }
    raw.resize((size_t)bytes);
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open '" + path + "'"; return MR_ERR_IO; }
    if (!v.is_record) {
        in.seekg((std::streamoff)v.begin);
        in.read((char *)raw.data(), (std::streamsize)raw.size());
    } else {
        uint64_t chunk = v.elems_per_chunk * ts;
        for (uint64_t r = 0; r < numrecs; ++r) {
            uint64_t off = v.begin + r * recsize;
            if (off + chunk > file_size) { err = "'" + path + "': record variable '" + v.name + "' runs past the end of the file"; return MR_ERR_FORMAT; }
            in.seekg((std::streamoff)off);
            in.read((char *)raw.data() + r * chunk, (std::streamsize)chunk);
        }
    }
    if (!in) { err = "short read on '" + path + "' (variable '" + v.name + "')"; return MR_ERR_IO; }
    return MR_OK;
}

namespace {
template <typename T> T be_load(const uint8_t *p);
template <> int8_t  be_load<int8_t >(const uint8_t *p) { return (int8_t)p[0]; }
template <> uint8_t be_load<uint8_t>(const uint8_t *p) { return p[0]; }
template <> int16_t be_load<int16_t>(const uint8_t *p) { return (int16_t)((uint16_t)p[0] << 8 | p[1]); }
template <> int32_t be_load<int32_t>(const uint8_t *p)
{
    return (int32_t)((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]);
}
template <> float be_load<float>(const uint8_t *p)
{
    uint32_t u = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
    float f; std::memcpy(&f, &u, 4); return f;
}
template <> double be_load<double>(const uint8_t *p)
{
    uint64_t u = 0;
    for (int i = 0; i < 8; ++i) u = u << 8 | p[i];
    double d; std::memcpy(&d, &u, 8); return d;
}

// `*x as f32` / `*x as f64` of the reference's dtype switch.  NC_CHAR is what the
// netcdf3 crate calls U8.
template <typename Out>
void cast_all(int type, const std::vector<uint8_t> &raw, uint64_t n, Out *out)
{
    const uint8_t *p = raw.data();
    switch (type) {
    case NC3_BYTE:   for (uint64_t i = 0; i < n; ++i) out[i] = (Out)be_load<int8_t >(p + i); break;
    case NC3_CHAR:   for (uint64_t i = 0; i < n; ++i) out[i] = (Out)be_load<uint8_t>(p + i); break;
    case NC3_SHORT:  for (uint64_t i = 0; i < n; ++i) out[i] = (Out)be_load<int16_t>(p + 2 * i); break;
    case NC3_INT:    for (uint64_t i = 0; i < n; ++i) out[i] = (Out)be_load<int32_t>(p + 4 * i); break;
    case NC3_FLOAT:  for (uint64_t i = 0; i < n; ++i) out[i] = (Out)be_load<float  >(p + 4 * i); break;
    case NC3_DOUBLE: for (uint64_t i = 0; i < n; ++i) out[i] = (Out)be_load<double >(p + 8 * i); break;
    }
}
}  // namespace

int Nc3File::read_f32(const std::string &name, std::vector<float> &out, std::string &err) const
{
    const Nc3Var *v = find(name);
    if (!v) { err = "'" + path + "': no variable named '" + name + "'"; return MR_ERR_FORMAT; }
    std::vector<uint8_t> raw;
    int rc = read_raw(*v, raw, err);
    if (rc) return rc;
    out.resize((size_t)num_elems(*v));
    cast_all<float>(v->type, raw, num_elems(*v), out.data());
    return MR_OK;
}

int Nc3File::read_f64(const std::string &name, std::vector<double> &out, std::string &err) const
{
    const Nc3Var *v = find(name);
    if (!v) { err = "'" + path + "': no variable named '" + name + "'"; return MR_ERR_FORMAT; }
    std::vector<uint8_t> raw;
    int rc = read_raw(*v, raw, err);
    if (rc) return rc;
    out.resize((size_t)num_elems(*v));
    cast_all<double>(v->type, raw, num_elems(*v), out.data());
    return MR_OK;
}

}  // namespace mr
