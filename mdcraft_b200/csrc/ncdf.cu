// ncdf.cu — AMBER NetCDF trajectory ingest.
//
// Replaces the READ side of /root/reference/src/mdcraft/openmm/file.py:22-306 (NetCDFFile.get_num_frames,
// get_num_atoms, get_dimensions, get_times, get_positions), which goes through the netCDF4 library, for the files
// that class writes (file.py:52-54: format NETCDF3_64BIT_OFFSET) and for classic NetCDF-3 files.
// The header (dimensions, variables, record layout) is parsed on the host from a memory map; the coordinates of a
// block of frames — big-endian float32 records interleaved with the other record variables — go to HBM with one
// strided copy and are byte-swapped there into float32 (F, N, 3), the layout mdc_sq_push / mdc_rdf_push take with
// on_device = 1.  HBM / PCIe byte work: 12 B per atom in, 12 B out.
#include "common.cuh"

#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

using namespace mdc;

namespace {

constexpr int NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_FLOAT = 5, NC_DOUBLE = 6;   // NC_INT = 4: the default size
constexpr int NC_DIMENSION = 10, NC_VARIABLE = 11, NC_ATTRIBUTE = 12;

__global__ void ncdf_bswap32(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __byte_perm(in[i], 0, 0x0123);
}

struct Var {
    std::string name;
    std::vector<int> dimids;
    int type = 0;
    int64_t vsize = 0, begin = 0;
    bool record = false;
};

struct Cursor {
    const unsigned char *p, *end;
    bool ok = true;
    uint32_t u32()
    {
        if (p + 4 > end) {
            ok = false;
            return 0;
        }
        const uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        p += 4;
        return v;
    }
    uint64_t u64()
    {
        const uint64_t hi = u32();
        return (hi << 32) | u32();
    }
    std::string name()
    {
        const uint32_t n = u32();
        if (!ok || p + n > end) {
            ok = false;
            return "";
        }
        std::string s((const char *)p, n);
        p += (n + 3) & ~3u;
        return s;
    }
    void skip(size_t n)
    {
        n = (n + 3) & ~(size_t)3;
        if (p + n > end) ok = false;
        else p += n;
    }
};

int type_size(int t) { return t == NC_BYTE || t == NC_CHAR ? 1 : t == NC_SHORT ? 2 : t == NC_DOUBLE ? 8 : 4; }

// att_list: ABSENT | NC_ATTRIBUTE nelems [name nc_type nelems values...]; keeps text attributes
void read_atts(Cursor &c, std::map<std::string, std::string> *text)
{
    const uint32_t tag = c.u32(), n = c.u32();
    if (tag == 0) return;
    if (tag != NC_ATTRIBUTE) {
        c.ok = false;
        return;
    }
    for (uint32_t i = 0; i < n && c.ok; ++i) {
        const std::string nm = c.name();
        const int type = (int)c.u32();
        const uint32_t ne = c.u32();
        if (type == NC_CHAR && text && c.ok && c.p + ne <= c.end) (*text)[nm] = std::string((const char *)c.p, ne);
        c.skip((size_t)ne * type_size(type));
    }
}

double be_f64(const unsigned char *p)
{
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
    double d;
    memcpy(&d, &v, 8);
    return d;
}

float be_f32(const unsigned char *p)
{
    const uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    float f;
    memcpy(&f, &v, 4);
    return f;
}

}  // namespace

struct mdc_ncdf {
    int fd = -1;
    const unsigned char *data = nullptr;
    size_t size = 0;
    int64_t n_frames = 0, n_atoms = 0, recsize = 0;
    Var coords, lengths, angles, time, velocities, forces;
    bool has_box = false, has_time = false, has_velocities = false, has_forces = false;
    std::string conventions;
    DevBuf<uint32_t> stage;
    DevBuf<float> out;
    cudaStream_t stream = nullptr;
    int stream_device = -1;
};

namespace {

void ncdf_free(mdc_ncdf *d)
{
    if (!d) return;
    d->stage.release();
    d->out.release();
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->data) munmap((void *)d->data, d->size);
    if (d->fd >= 0) close(d->fd);
    delete d;
}

int ncdf_fail(mdc_ncdf *d, int code, const std::string &msg)
{
    ncdf_free(d);
    return fail(code, msg);
}

}  // namespace

extern "C" int mdc_ncdf_open(const char *path, mdc_ncdf **out)
{
    if (!path || !out) return fail(MDC_ERR_INVALID, "mdc_ncdf_open: NULL argument");
    *out = nullptr;
    mdc_ncdf *d = new mdc_ncdf();
    d->fd = open(path, O_RDONLY);
    if (d->fd < 0) return ncdf_fail(d, MDC_ERR_INVALID, std::string("mdc_ncdf_open: cannot open '") + path + "'");
    struct stat sb;
    if (fstat(d->fd, &sb) != 0 || sb.st_size < 32) return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: empty or unreadable file");
    d->size = (size_t)sb.st_size;
    void *m = mmap(nullptr, d->size, PROT_READ, MAP_PRIVATE, d->fd, 0);
    if (m == MAP_FAILED) {
        d->data = nullptr;
        return ncdf_fail(d, MDC_ERR_NOMEM, "mdc_ncdf_open: mmap failed");
    }
    d->data = (const unsigned char *)m;
    Cursor c{d->data, d->data + d->size};
    if (memcmp(c.p, "CDF", 3) != 0) return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: not a NetCDF classic file");
    const int version = c.p[3];
    if (version != 1 && version != 2)
        return ncdf_fail(d, MDC_ERR_UNSUPPORTED, "mdc_ncdf_open: only NetCDF-3 classic and 64-bit-offset files are supported");
    c.p += 4;
    const uint32_t numrecs = c.u32();
    // dimensions
    std::vector<std::string> dim_names;
    std::vector<int64_t> dim_sizes;
    {
        const uint32_t tag = c.u32(), n = c.u32();
        if (tag != 0 && tag != NC_DIMENSION) return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: corrupt dimension list");
        for (uint32_t i = 0; i < n && c.ok; ++i) {
            dim_names.push_back(c.name());
            dim_sizes.push_back(c.u32());
        }
    }
    std::map<std::string, std::string> gatts;
    read_atts(c, &gatts);
    if (gatts.count("Conventions")) d->conventions = gatts["Conventions"];
    // variables
    std::vector<Var> vars;
    {
        const uint32_t tag = c.u32(), n = c.u32();
        if (tag != 0 && tag != NC_VARIABLE) return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: corrupt variable list");
        for (uint32_t i = 0; i < n && c.ok; ++i) {
            Var v;
            v.name = c.name();
            const uint32_t nd = c.u32();
            for (uint32_t k = 0; k < nd; ++k) v.dimids.push_back((int)c.u32());
            read_atts(c, nullptr);
            v.type = (int)c.u32();
            v.vsize = c.u32();
            v.begin = version == 2 ? (int64_t)c.u64() : (int64_t)c.u32();
            for (int id : v.dimids)
                if (id < 0 || id >= (int)dim_sizes.size()) c.ok = false;
            v.record = !v.dimids.empty() && c.ok && dim_sizes[v.dimids[0]] == 0;
            vars.push_back(v);
        }
    }
    if (!c.ok) return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: truncated header");
    if (d->conventions.rfind("AMBER", 0) != 0 || d->conventions == "AMBERRESTART")
        return ncdf_fail(d, MDC_ERR_INVALID, "The NetCDF file is not a valid AMBER NetCDF trajectory or does not contain any data.");
    int n_rec_vars = 0;
    for (auto &v : vars)
        if (v.record) {
            d->recsize += v.vsize;
            ++n_rec_vars;
        }
    auto find = [&](const char *nm, Var *dst) {
        for (auto &v : vars)
            if (v.name == nm) {
                *dst = v;
                return true;
            }
        return false;
    };
    if (!find("coordinates", &d->coords) || d->coords.type != NC_FLOAT || d->coords.dimids.size() != 3 || !d->coords.record)
        return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: no float32 'coordinates' (frame, atom, spatial) variable");
    d->n_atoms = dim_sizes[d->coords.dimids[1]];
    if (dim_sizes[d->coords.dimids[2]] != 3 || d->n_atoms <= 0)
        return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: 'coordinates' must be (frame, atom, 3)");
    if (n_rec_vars == 1) d->recsize = d->n_atoms * 12;   // a lone record variable is not padded
    d->has_box = find("cell_lengths", &d->lengths) && find("cell_angles", &d->angles) && d->lengths.type == NC_DOUBLE &&
                 d->angles.type == NC_DOUBLE && d->lengths.record && d->angles.record;
    d->has_time = find("time", &d->time) && d->time.type == NC_FLOAT && d->time.record;
    // optional per-atom vectors with the layout of the coordinates (file.py:217-306: get_velocities, get_forces)
    auto per_atom = [&](const char *nm, Var *dst) {
        return find(nm, dst) && dst->type == NC_FLOAT && dst->record && dst->dimids.size() == 3 &&
               dim_sizes[dst->dimids[1]] == d->n_atoms && dim_sizes[dst->dimids[2]] == 3;
    };
    d->has_velocities = per_atom("velocities", &d->velocities);
    d->has_forces = per_atom("forces", &d->forces);
    d->n_frames = numrecs;
    if (numrecs == 0xffffffffu && d->recsize > 0) d->n_frames = ((int64_t)d->size - d->coords.begin + d->recsize - d->n_atoms * 12) / d->recsize;
    if (d->n_frames > 0 && d->coords.begin + (d->n_frames - 1) * d->recsize + d->n_atoms * 12 > (int64_t)d->size)
        return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: file is shorter than its record count");
    for (Var *v : {&d->velocities, &d->forces}) {
        const bool present = v == &d->velocities ? d->has_velocities : d->has_forces;
        if (present && d->n_frames > 0 && v->begin + (d->n_frames - 1) * d->recsize + d->n_atoms * 12 > (int64_t)d->size)
            return ncdf_fail(d, MDC_ERR_INVALID, "mdc_ncdf_open: file is shorter than its record count");
    }
    *out = d;
    return MDC_OK;
}

extern "C" int mdc_ncdf_close(mdc_ncdf *d)
{
    ncdf_free(d);
    return MDC_OK;
}

extern "C" int mdc_ncdf_info(mdc_ncdf *d, int64_t info[4])
{
    if (!d || !info) return fail(MDC_ERR_INVALID, "mdc_ncdf_info: NULL argument");
    info[0] = d->n_frames;
    info[1] = d->n_atoms;
    info[2] = d->has_box;
    info[3] = d->has_time;
    return MDC_OK;
}

// times f32 -> f64 (n_frames), cell_lengths / cell_angles f64 (n_frames, 3); NULL pointers are skipped
extern "C" int mdc_ncdf_headers(mdc_ncdf *d, double *times, double *cell_lengths, double *cell_angles)
{
    if (!d) return fail(MDC_ERR_INVALID, "mdc_ncdf_headers: NULL argument");
    for (int64_t f = 0; f < d->n_frames; ++f) {
        if (times) {
            if (!d->has_time) return fail(MDC_ERR_INVALID, "mdc_ncdf_headers: the file has no 'time' variable");
            times[f] = (double)be_f32(d->data + d->time.begin + f * d->recsize);
        }
        if (cell_lengths || cell_angles) {
            if (!d->has_box) return fail(MDC_ERR_INVALID, "mdc_ncdf_headers: the file has no cell information");
            for (int k = 0; k < 3; ++k) {
                if (cell_lengths) cell_lengths[f * 3 + k] = be_f64(d->data + d->lengths.begin + f * d->recsize + 8 * k);
                if (cell_angles) cell_angles[f * 3 + k] = be_f64(d->data + d->angles.begin + f * d->recsize + 8 * k);
            }
        }
    }
    return MDC_OK;
}

extern "C" int mdc_ncdf_has(mdc_ncdf *d, int which)
{
    if (!d) return 0;
    return which == MDC_NCDF_COORDINATES ? 1 : which == MDC_NCDF_VELOCITIES ? (int)d->has_velocities
                                             : which == MDC_NCDF_FORCES   ? (int)d->has_forces : 0;
}

extern "C" int mdc_ncdf_read(mdc_ncdf *d, mdc_ctx *ctx, int64_t f0, int n_frames, float *pos_out, int on_device)
{
    return mdc_ncdf_read_var(d, ctx, MDC_NCDF_COORDINATES, f0, n_frames, pos_out, on_device);
}

extern "C" int mdc_ncdf_read_var(mdc_ncdf *d, mdc_ctx *ctx, int which, int64_t f0, int n_frames, float *pos_out,
                                 int on_device)
{
    if (!d || !ctx || !pos_out) return fail(MDC_ERR_INVALID, "mdc_ncdf_read: NULL argument");
    if (!mdc_ncdf_has(d, which)) return fail(MDC_ERR_INVALID, "mdc_ncdf_read_var: the file does not hold that variable");
    const Var &var = which == MDC_NCDF_VELOCITIES ? d->velocities : which == MDC_NCDF_FORCES ? d->forces : d->coords;
    if (f0 < 0 || n_frames <= 0 || f0 + n_frames > d->n_frames) return fail(MDC_ERR_INVALID, "mdc_ncdf_read: frame range out of bounds");
    MDC_CUDA(cudaSetDevice(ctx->device));
    if (!d->stream || d->stream_device != ctx->device) {
        if (d->stream) cudaStreamDestroy(d->stream);
        d->stream = nullptr;
        MDC_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        d->stream_device = ctx->device;
    }
    cudaStream_t st = d->stream;
    const size_t row = (size_t)d->n_atoms * 12;
    const int64_t words = (int64_t)n_frames * d->n_atoms * 3;
    MDC_CHECK(d->stage.alloc((size_t)words));
    float *out_dev = pos_out;
    if (!on_device) {
        MDC_CHECK(d->out.alloc((size_t)words));
        out_dev = d->out.p;
    }
    // one strided copy: the coordinates of consecutive records are recsize bytes apart in the file
    MDC_CUDA(cudaMemcpy2DAsync(d->stage.p, row, d->data + var.begin + f0 * d->recsize, (size_t)d->recsize, row,
                               (size_t)n_frames, cudaMemcpyHostToDevice, st));
    ncdf_bswap32<<<(unsigned)ceil_div(words, 256), 256, 0, st>>>(d->stage.p, words, reinterpret_cast<uint32_t *>(out_dev));
    MDC_CUDA(cudaGetLastError());
    if (!on_device) MDC_CUDA(cudaMemcpyAsync(pos_out, out_dev, (size_t)words * 4, cudaMemcpyDeviceToHost, st));
    MDC_CUDA(cudaStreamSynchronize(st));
    return MDC_OK;
}
