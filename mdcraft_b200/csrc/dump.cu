// dump.cu — LAMMPS text-dump frame ingest: the step before the S(q) / g(r) path.
//
// Replaces, for the dump layouts the hot path consumes (orthogonal box, unscaled coordinates x y z or xu yu zu,
// optional atom ids), /root/reference/src/mdcraft/analysis/reader.py:
//   :138-210  _get_offsets          -> frame index built from a memory map with host threads (newline counting)
//   :353-548  _read_next_timestep   -> the per-atom Python loop becomes one GPU pass over the raw text:
//             the bytes of a block of frames go to HBM once, newlines are ranked with a prefix sum, one thread
//             parses one atom line, and positions land in HBM as float32 (F, N, 3) in atom-id order (:550-560),
//             i.e. exactly the layout mdc_sq_push / mdc_rdf_push take with on_device = 1.
// Number semantics are the reference's: ts.positions[i] = tuple(str) is float32(float(str)), a correctly rounded
// decimal -> double conversion followed by a cast (checked against NumPy).  The device converts with Clinger's
// exact fast path (<= 2^53 mantissa, |exponent| <= 22 (+15)), 17-19 digit mantissas with a checked double-double
// division; anything else (more digits, inf / nan, odd syntax, a quotient too close to a rounding boundary)
// is flagged and re-parsed on the host with strtod, so every value is bit-identical.
//
// Roofline: HBM / PCIe byte work: ~45 B of text per atom in, 12 B out.
#include "common.cuh"

#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

using namespace mdc;

namespace {

constexpr int DT_THREADS = 256;
constexpr int DT_TILE = DT_THREADS * 16;   // bytes per block: one uint4 per thread
constexpr int DT_MAXCOL = 64;
constexpr unsigned DT_FIXCAP = 1u << 20;

struct DumpCols {
    int n_cols;
    int id;        // column of the atom id, -1: none
    int xyz[3];    // columns of the three coordinates, -1: axis absent (stays 0, reader.py:503-505)
};

struct FixEntry {
    unsigned off;    // byte offset of the token in the blob
    unsigned dest;   // index into xyz_tmp (coordinates) or ids (dest | 0x80000000)
};

__device__ __forceinline__ bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }

__constant__ double c_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// mant / 10^k (1 <= k <= 22) for a mantissa of up to 64 bits, correctly rounded, or false if the quotient is too
// close to a rounding boundary to tell (probability ~2^-29; the caller defers to strtod).  The mantissa is split
// exactly into its top 53 bits and the rest; q1 + q2 carries the quotient to ~2^-93 relative, far below the
// margin tested against the midpoint.
__device__ bool div_pow10(unsigned long long mant, int k, double *out)
{
    const double p = c_pow10[k];
    const double mh = (double)(mant & ~0x7ffull), ml = (double)(mant & 0x7ffull);   // both exact
    const double q1 = mh / p;
    const double r = fma(-q1, p, mh);          // exact remainder of the first division
    const double q2 = (r + ml) / p;
    const double s = q1 + q2;
    const double err = (q1 - s) + q2;          // exact rounding error of that sum (|q1| >= |q2|)
    const long long bits = __double_as_longlong(s);
    if ((bits & 0xfffffffffffffll) == 0) return false;   // power of two: the spacing below differs
    const double u = __longlong_as_double((bits & 0x7ff0000000000000ll)) * 2.220446049250313e-16;   // ulp(s)
    if (fabs(fabs(err) - 0.5 * u) < u * 9.313225746154785e-10) return false;
    *out = s;
    return true;
}

// [+-]digits[.digits][(e|E)[+-]digits] in [s, e) -> the correctly rounded double, or false (caller defers to strtod)
__device__ bool parse_double(const unsigned char *s, const unsigned char *e, double *out)
{
    if (s >= e) return false;
    bool neg = false;
    if (*s == '+' || *s == '-') {
        neg = *s == '-';
        ++s;
    }
    unsigned long long mant = 0;
    int nd = 0, sig = 0, e10 = 0;
    bool dropped = false;
    for (; s < e && *s >= '0' && *s <= '9'; ++s, ++nd) {
        if (sig < 19) {
            mant = mant * 10 + (unsigned)(*s - '0');
            if (mant) ++sig;
        } else {
            ++e10;
            dropped = dropped || *s != '0';
        }
    }
    if (s < e && *s == '.') {
        ++s;
        for (; s < e && *s >= '0' && *s <= '9'; ++s, ++nd) {
            if (sig < 19) {
                mant = mant * 10 + (unsigned)(*s - '0');
                if (mant) ++sig;
                --e10;
            } else {
                dropped = dropped || *s != '0';
            }
        }
    }
    if (nd == 0) return false;
    if (s < e && (*s == 'e' || *s == 'E')) {
        ++s;
        bool eneg = false;
        if (s < e && (*s == '+' || *s == '-')) {
            eneg = *s == '-';
            ++s;
        }
        int ev = 0, ned = 0;
        for (; s < e && *s >= '0' && *s <= '9'; ++s, ++ned)
            if (ev < 100000) ev = ev * 10 + (*s - '0');
        if (ned == 0) return false;
        e10 += eneg ? -ev : ev;
    }
    if (s != e || dropped) return false;
    double v;
    if (mant == 0) {
        v = 0.0;
    } else {
        if (mant >> 53) {
            // 17-19 significant digits (%.17g, repr): exact double-double division when the exponent allows
            if (e10 >= 0 || e10 < -22 || !div_pow10(mant, -e10, &v)) return false;
        } else if (e10 >= 0) {
            if (e10 <= 22) {
                v = (double)mant * c_pow10[e10];
            } else if (e10 <= 22 + 15) {
                // mant * 10^(e10 - 22) is still exact if it stays below 2^53
                const double m2 = (double)mant * c_pow10[e10 - 22];
                if (m2 >= 9007199254740992.0) return false;
                v = m2 * c_pow10[22];
            } else {
                return false;
            }
        } else {
            if (e10 < -22) return false;
            v = (double)mant / c_pow10[-e10];
        }
    }
    *out = neg ? -v : v;
    return true;
}

__device__ bool parse_int(const unsigned char *s, const unsigned char *e, long long *out)
{
    if (s >= e) return false;
    bool neg = false;
    if (*s == '+' || *s == '-') {
        neg = *s == '-';
        ++s;
    }
    if (s >= e || e - s > 18) return false;
    long long v = 0;
    for (; s < e; ++s) {
        if (*s < '0' || *s > '9') return false;
        v = v * 10 + (*s - '0');
    }
    *out = neg ? -v : v;
    return true;
}

__device__ __forceinline__ unsigned newline_mask(uint4 w)
{
    unsigned m = 0;
    const unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (((v[k] >> (8 * b)) & 0xffu) == '\n') m |= 1u << (4 * k + b);
    return m;
}

// text is padded with spaces to a multiple of DT_TILE
__global__ void __launch_bounds__(DT_THREADS)
dump_count(const uint4 *__restrict__ text, unsigned *__restrict__ tile_count)
{
    __shared__ unsigned sh[DT_THREADS / 32];
    const unsigned m = newline_mask(text[(size_t)blockIdx.x * DT_THREADS + threadIdx.x]);
    unsigned c = __popc(m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < DT_THREADS / 32; ++i) t += sh[i];
        tile_count[blockIdx.x] = t;
    }
}

// exclusive scan of n counters in place (one block); total -> *total
__global__ void __launch_bounds__(1024)
dump_scan(unsigned *__restrict__ v, unsigned n, unsigned *__restrict__ total)
{
    __shared__ unsigned sh[1024];
    const unsigned per = (n + 1023) / 1024, lo = threadIdx.x * per, hi = min(n, lo + per);
    unsigned s = 0;
    for (unsigned i = lo; i < hi; ++i) s += v[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    unsigned run = sh[threadIdx.x] - s;
    for (unsigned i = lo; i < hi; ++i) {
        const unsigned c = v[i];
        v[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) *total = sh[1023];
}

// line_start[k + 1] = byte after the k-th newline (line_start[0] = 0 is set by the host)
__global__ void __launch_bounds__(DT_THREADS)
dump_starts(const uint4 *__restrict__ text, const unsigned *__restrict__ tile_off, unsigned max_lines,
            unsigned *__restrict__ line_start)
{
    __shared__ unsigned sh[DT_THREADS / 32];
    const size_t chunk = (size_t)blockIdx.x * DT_THREADS + threadIdx.x;
    unsigned m = newline_mask(text[chunk]);
    const unsigned c = __popc(m);
    unsigned inc = c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    unsigned base = tile_off[blockIdx.x];
    for (int i = 0; i < warp; ++i) base += sh[i];
    unsigned rank = base + inc - c;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        if (rank + 1 < max_lines) line_start[rank + 1] = (unsigned)(chunk * 16 + b + 1);
        ++rank;
    }
}

// One thread per line of the blob.  The blob starts at the first ATOM line of its first frame; every later frame
// brings its 9 header lines, so line i is line (i + 9) % L of frame (i + 9) / L with L = 9 + N.
__global__ void __launch_bounds__(DT_THREADS)
dump_parse_atoms(const unsigned char *__restrict__ text, const unsigned *__restrict__ line_start, unsigned n_lines,
                 unsigned text_len, int N, DumpCols cols, long long *__restrict__ ids, float *__restrict__ xyz,
                 FixEntry *__restrict__ fix, unsigned *__restrict__ n_fix, int *__restrict__ status)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    const unsigned L = 9u + (unsigned)N, j = i + 9u, f = j / L, rel = j - f * L;
    if (rel < 9u) return;
    const unsigned a = rel - 9u;
    const unsigned char *s = text + line_start[i];
    const unsigned char *end = text + ((i + 1 < n_lines) ? line_start[i + 1] - 1 : text_len);   // at the newline
    const size_t slot = (size_t)f * N + a;
    float v[3] = {0.f, 0.f, 0.f};
    bool have[3] = {cols.xyz[0] < 0, cols.xyz[1] < 0, cols.xyz[2] < 0}, have_id = cols.id < 0;
    int col = 0;
    while (s < end) {
        while (s < end && is_space(*s)) ++s;
        if (s >= end) break;
        const unsigned char *t = s;
        while (s < end && !is_space(*s)) ++s;
        if (col == cols.id) {
            long long id;
            if (parse_int(t, s, &id)) {
                ids[slot] = id;
            } else {
                const unsigned k = atomicAdd(n_fix, 1u);
                if (k < DT_FIXCAP) fix[k] = {(unsigned)(t - text), (unsigned)slot | 0x80000000u};
            }
            have_id = true;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (col == cols.xyz[d]) {
                double x;
                if (parse_double(t, s, &x)) {
                    v[d] = (float)x;   // float32(float(str)): reader.py:503-505 through NumPy's str -> float32 cast
                } else {
                    const unsigned k = atomicAdd(n_fix, 1u);
                    if (k < DT_FIXCAP) fix[k] = {(unsigned)(t - text), (unsigned)(slot * 3 + d)};
                }
                have[d] = true;
            }
        ++col;
    }
    if (!(have[0] && have[1] && have[2] && have_id)) atomicOr(status, 1);   // short line
    xyz[slot * 3 + 0] = v[0];
    xyz[slot * 3 + 1] = v[1];
    xyz[slot * 3 + 2] = v[2];
}

__global__ void dump_min_id(const long long *__restrict__ ids, int N, int F, long long *__restrict__ min_id)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * F) return;
    atomicMin(&min_id[i / N], ids[i]);
}

// positions[order] with order = argsort(ids) (reader.py:550-552), for ids that are a permutation of min .. min + N - 1
__global__ void dump_scatter(const long long *__restrict__ ids, const float *__restrict__ xyz, int N, int F,
                             const long long *__restrict__ min_id, int *__restrict__ mark, float *__restrict__ out,
                             int *__restrict__ status)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * F) return;
    const int64_t f = i / N;
    const long long r = ids[i] - min_id[f];
    if (r < 0 || r >= N) {
        atomicOr(status, 2);
        return;
    }
    if (atomicExch(&mark[f * N + r], 1) != 0) atomicOr(status, 2);   // duplicate id
    float *o = out + ((size_t)f * N + r) * 3;
    o[0] = xyz[i * 3 + 0];
    o[1] = xyz[i * 3 + 1];
    o[2] = xyz[i * 3 + 2];
}

__global__ void dump_gather(const int *__restrict__ order, const float *__restrict__ xyz, int64_t total,
                            float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t src = order[i];
    out[i * 3 + 0] = xyz[src * 3 + 0];
    out[i * 3 + 1] = xyz[src * 3 + 1];
    out[i * 3 + 2] = xyz[src * 3 + 2];
}

// ---- host side ----------------------------------------------------------------------------------------------

struct Line {
    const char *b, *e;   // [b, e): without the newline
};

bool next_line(const char *&p, const char *end, Line *out)
{
    if (p >= end) return false;
    const char *nl = (const char *)memchr(p, '\n', end - p);
    out->b = p;
    out->e = nl ? nl : end;
    p = nl ? nl + 1 : end;
    return true;
}

std::vector<std::string> split_ws(const Line &l)
{
    std::vector<std::string> t;
    const char *s = l.b;
    while (s < l.e) {
        while (s < l.e && isspace((unsigned char)*s)) ++s;
        if (s >= l.e) break;
        const char *b = s;
        while (s < l.e && !isspace((unsigned char)*s)) ++s;
        t.emplace_back(b, s);
    }
    return t;
}

}  // namespace

struct mdc_dump {
    int fd = -1;
    const char *data = nullptr;
    size_t size = 0;
    int64_t n_atoms = 0, n_frames = 0;
    DumpCols cols{};
    int conv[3] = {-1, -1, -1};            // index into {"u", "su", "", "s"}, -1: axis absent
    std::vector<int64_t> offsets;          // frame starts (+ end of the last complete frame)
    std::vector<int64_t> atoms_off;        // first atom line of every frame
    std::vector<int64_t> steps;
    std::vector<float> box;                // (n_frames, 6)
    std::vector<double> origin;            // (n_frames, 3)
    std::string atoms_header;
    // device scratch, grown on demand
    DevBuf<unsigned char> text;
    DevBuf<unsigned> tile_count, line_start, counters;
    DevBuf<long long> ids, min_id;
    DevBuf<float> xyz, out;
    DevBuf<int> mark, order, status;
    DevBuf<FixEntry> fix;
    int last_fix = 0;
    // the parser's own stream: the text upload and the parse kernels of block k + 1 overlap the analysis kernels of
    // block k, which run on the context's streams
    cudaStream_t stream = nullptr;
    int stream_device = -1;
};

namespace {

void dump_free(mdc_dump *d)
{
    if (!d) return;
    d->text.release();
    d->tile_count.release();
    d->line_start.release();
    d->counters.release();
    d->ids.release();
    d->min_id.release();
    d->xyz.release();
    d->out.release();
    d->mark.release();
    d->order.release();
    d->status.release();
    d->fix.release();
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->data) munmap((void *)d->data, d->size);
    if (d->fd >= 0) close(d->fd);
    delete d;
}

int dump_fail(mdc_dump *d, int code, const std::string &msg)
{
    dump_free(d);
    return fail(code, msg);
}

}  // namespace

extern "C" int mdc_dump_open(const char *path, const char *conventions, int n_threads, mdc_dump **out)
{
    if (!path || !out) return fail(MDC_ERR_INVALID, "mdc_dump_open: NULL argument");
    *out = nullptr;
    mdc_dump *d = new mdc_dump();
    d->fd = open(path, O_RDONLY);
    if (d->fd < 0) return dump_fail(d, MDC_ERR_INVALID, std::string("mdc_dump_open: cannot open '") + path + "'");
    struct stat sb;
    if (fstat(d->fd, &sb) != 0 || sb.st_size == 0) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: empty or unreadable file");
    d->size = (size_t)sb.st_size;
    void *m = mmap(nullptr, d->size, PROT_READ, MAP_PRIVATE, d->fd, 0);
    if (m == MAP_FAILED) {
        d->data = nullptr;
        return dump_fail(d, MDC_ERR_NOMEM, "mdc_dump_open: mmap failed");
    }
    d->data = (const char *)m;
    const char *end = d->data + d->size;

    // first frame header (reader.py:262-269, 382-418)
    const char *p = d->data;
    Line ln[9];
    for (int i = 0; i < 9; ++i)
        if (!next_line(p, end, &ln[i])) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: truncated header");
    if (strncmp(ln[0].b, "ITEM: TIMESTEP", 14) != 0) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: not a LAMMPS dump file");
    d->n_atoms = strtoll(std::string(ln[3].b, ln[3].e).c_str(), nullptr, 10);
    if (d->n_atoms <= 0 || d->n_atoms > 2147483647ll) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: bad number of atoms");
    {
        std::string l8(ln[8].b, ln[8].e);
        while (!l8.empty() && isspace((unsigned char)l8.back())) l8.pop_back();
        if (l8 == "ITEM: DIMENSION") return dump_fail(d, MDC_ERR_UNSUPPORTED, "mdc_dump_open: 'style grid' dumps are not supported");
        if (l8.compare(0, 11, "ITEM: ATOMS") != 0) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: no 'ITEM: ATOMS' line");
    }
    d->atoms_header.assign(ln[8].b, ln[8].e);
    std::vector<std::string> names = split_ws(ln[8]);
    names.erase(names.begin(), names.begin() + 2);
    if ((int)names.size() > DT_MAXCOL) return dump_fail(d, MDC_ERR_UNSUPPORTED, "mdc_dump_open: too many columns");
    auto find = [&](const std::string &n) {
        for (size_t i = 0; i < names.size(); ++i)
            if (names[i] == n) return (int)i;
        return -1;
    };
    d->cols.n_cols = (int)names.size();
    d->cols.id = find("id");
    static const char *CONV[4] = {"u", "su", "", "s"};
    // conventions: NULL = automatic (first of u, su, "", s present per axis, reader.py:438-447); otherwise three
    // comma-separated entries, "-" meaning "ignore this axis"
    std::vector<std::string> want;
    if (conventions) {
        std::string c(conventions), cur;
        for (char ch : c) {
            if (ch == ',') {
                want.push_back(cur);
                cur.clear();
            } else {
                cur.push_back(ch);
            }
        }
        want.push_back(cur);
        if (want.size() != 3) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: conventions must name three axes");
    }
    for (int ax = 0; ax < 3; ++ax) {
        const std::string axn(1, "xyz"[ax]);
        d->cols.xyz[ax] = -1;
        if (conventions) {
            if (want[ax] == "-") continue;
            int ci = -1;
            for (int k = 0; k < 4; ++k)
                if (want[ax] == CONV[k]) ci = k;
            if (ci < 0) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: invalid convention '" + want[ax] + "'");
            const int col = find(axn + CONV[ci]);
            if (col < 0)
                return dump_fail(d, MDC_ERR_INVALID, "No " + axn + "-coordinate information with the specified convention '" +
                                                         want[ax] + "' found in '" + path + "'.");
            d->cols.xyz[ax] = col;
            d->conv[ax] = ci;
        } else {
            for (int k = 0; k < 4; ++k) {
                const int col = find(axn + CONV[k]);
                if (col >= 0) {
                    d->cols.xyz[ax] = col;
                    d->conv[ax] = k;
                    break;
                }
            }
        }
    }

    // frame index: every frame has L = 9 + N lines; rank the newlines with host threads (reader.py:138-210)
    const int64_t L = 9 + d->n_atoms;
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    T = std::max(1, std::min(T, 64));
    if (d->size < ((size_t)1 << 22)) T = 1;
    std::vector<int64_t> cnt(T, 0);
    const size_t per = (d->size + T - 1) / T;
    auto count_part = [&](int t) {
        const char *b = d->data + std::min(d->size, per * t), *e = d->data + std::min(d->size, per * (t + 1));
        int64_t c = 0;
        while (b < e) {
            const char *nl = (const char *)memchr(b, '\n', e - b);
            if (!nl) break;
            ++c;
            b = nl + 1;
        }
        cnt[t] = c;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(count_part, t);
        count_part(0);
        for (auto &x : th) x.join();
    }
    int64_t total_nl = 0;
    std::vector<int64_t> first(T);
    for (int t = 0; t < T; ++t) {
        first[t] = total_nl;
        total_nl += cnt[t];
    }
    int64_t total_lines = total_nl + ((d->size > 0 && d->data[d->size - 1] != '\n') ? 1 : 0);
    d->n_frames = total_lines / L;
    if (d->n_frames == 0) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: no complete frame");
    d->offsets.assign(d->n_frames + 1, -1);
    d->offsets[0] = 0;
    auto locate_part = [&](int t) {
        const char *b = d->data + std::min(d->size, per * t), *e = d->data + std::min(d->size, per * (t + 1));
        int64_t g = first[t];   // newlines before b
        while (b < e) {
            const char *nl = (const char *)memchr(b, '\n', e - b);
            if (!nl) break;
            ++g;
            if (g % L == 0 && g / L <= d->n_frames) d->offsets[g / L] = (nl + 1) - d->data;
            b = nl + 1;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(locate_part, t);
        locate_part(0);
        for (auto &x : th) x.join();
    }
    if (d->offsets[d->n_frames] < 0) d->offsets[d->n_frames] = (int64_t)d->size;   // last line without a newline

    // per-frame headers on the host: step, atom count, box, ATOMS columns
    d->steps.resize(d->n_frames);
    d->box.resize(d->n_frames * 6);
    d->origin.resize(d->n_frames * 3);
    d->atoms_off.resize(d->n_frames);
    for (int64_t f = 0; f < d->n_frames; ++f) {
        const char *q = d->data + d->offsets[f];
        Line h[9];
        for (int i = 0; i < 9; ++i)
            if (!next_line(q, end, &h[i])) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: truncated frame header");
        if (strncmp(h[0].b, "ITEM: TIMESTEP", 14) != 0)
            return dump_fail(d, MDC_ERR_UNSUPPORTED,
                             "mdc_dump_open: frame " + std::to_string(f) + " does not start where a constant atom count puts it");
        d->steps[f] = strtoll(std::string(h[1].b, h[1].e).c_str(), nullptr, 10);
        const int64_t n = strtoll(std::string(h[3].b, h[3].e).c_str(), nullptr, 10);
        if (n != d->n_atoms)
            return dump_fail(d, MDC_ERR_INVALID, "Timestep " + std::to_string(d->steps[f]) + " has " + std::to_string(n) +
                                                     " atoms, but the topology has " + std::to_string(d->n_atoms) + " atoms.");
        const std::string bb(h[4].b, h[4].e);
        if (bb.find("abc origin") != std::string::npos)
            return dump_fail(d, MDC_ERR_UNSUPPORTED, "mdc_dump_open: general triclinic boxes (abc origin) are not supported");
        const bool tri = bb.find("xy xz yz") != std::string::npos;   // restricted triclinic: lo hi tilt per line (reader.py:372-375)
        double lo[3], hi[3], tilt[3] = {0.0, 0.0, 0.0};
        for (int ax = 0; ax < 3; ++ax) {
            const std::string l(h[5 + ax].b, h[5 + ax].e);
            char *e1 = nullptr, *e2 = nullptr, *e3 = nullptr;
            lo[ax] = strtod(l.c_str(), &e1);
            hi[ax] = strtod(e1, &e2);
            if (e1 == l.c_str() || e2 == e1) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: bad box bounds");
            if (tri) {
                tilt[ax] = strtod(e2, &e3);
                if (e3 == e2) return dump_fail(d, MDC_ERR_INVALID, "mdc_dump_open: bad box bounds (tilt factor missing)");
            }
        }
        if (!tri) {
            for (int ax = 0; ax < 3; ++ax) {
                d->box[f * 6 + ax] = (float)(hi[ax] - lo[ax]);   // reader.py:418, stored by MDAnalysis as float32
                d->box[f * 6 + 3 + ax] = 90.f;
                d->origin[f * 3 + ax] = lo[ax];
            }
        } else {
            // reader.py:376-388: the bounds of a tilted box hold its extent, take the tilts back out; box vectors
            // a = (lx, 0, 0), b = (xy, ly, 0), c = (xz, yz, lz); convert_cell_representation (algorithm/topology.py:
            // 743-757): lengths = norms, angles = degrees(arccos(dot / (|u| |v|))), float64, stored as float32
            const double xy = tilt[0], xz = tilt[1], yz = tilt[2];
            lo[0] -= std::min(std::min(0.0, xy), std::min(xz, xy + xz));
            hi[0] -= std::max(std::max(0.0, xy), std::max(xz, xy + xz));
            lo[1] -= std::min(0.0, yz);
            hi[1] -= std::max(0.0, yz);
            const double a[3] = {hi[0] - lo[0], 0.0, 0.0}, b[3] = {xy, hi[1] - lo[1], 0.0}, c[3] = {xz, yz, hi[2] - lo[2]};
            auto norm = [](const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
            auto dot = [](const double *u, const double *v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; };
            const double la = norm(a), lb = norm(b), lc = norm(c), deg = 180.0 / 3.141592653589793;
            d->box[f * 6 + 0] = (float)la;
            d->box[f * 6 + 1] = (float)lb;
            d->box[f * 6 + 2] = (float)lc;
            d->box[f * 6 + 3] = (float)(acos(dot(b, c) / (lb * lc)) * deg);
            d->box[f * 6 + 4] = (float)(acos(dot(a, c) / (la * lc)) * deg);
            d->box[f * 6 + 5] = (float)(acos(dot(a, b) / (la * lb)) * deg);
            for (int ax = 0; ax < 3; ++ax) d->origin[f * 3 + ax] = lo[ax];
        }
        if ((size_t)(h[8].e - h[8].b) != d->atoms_header.size() || memcmp(h[8].b, d->atoms_header.data(), d->atoms_header.size()) != 0)
            return dump_fail(d, MDC_ERR_UNSUPPORTED, "mdc_dump_open: the ATOMS columns change between frames");
        d->atoms_off[f] = q - d->data;
    }
    *out = d;
    return MDC_OK;
}

extern "C" int mdc_dump_close(mdc_dump *d)
{
    dump_free(d);
    return MDC_OK;
}

extern "C" int mdc_dump_info(mdc_dump *d, int64_t info[8])
{
    if (!d || !info) return fail(MDC_ERR_INVALID, "mdc_dump_info: NULL argument");
    info[0] = d->n_frames;
    info[1] = d->n_atoms;
    info[2] = d->cols.id >= 0;
    info[3] = d->conv[0];
    info[4] = d->conv[1];
    info[5] = d->conv[2];
    info[6] = d->cols.n_cols;
    info[7] = d->last_fix;
    return MDC_OK;
}

extern "C" int mdc_dump_headers(mdc_dump *d, int64_t *steps, float *box, double *origin, int64_t *offsets)
{
    if (!d) return fail(MDC_ERR_INVALID, "mdc_dump_headers: NULL argument");
    if (steps) memcpy(steps, d->steps.data(), d->n_frames * sizeof(int64_t));
    if (box) memcpy(box, d->box.data(), d->n_frames * 6 * sizeof(float));
    if (origin) memcpy(origin, d->origin.data(), d->n_frames * 3 * sizeof(double));
    if (offsets) memcpy(offsets, d->offsets.data(), (d->n_frames + 1) * sizeof(int64_t));
    return MDC_OK;
}

extern "C" int64_t mdc_dump_block_bytes(mdc_dump *d, int64_t f0, int n_frames)
{
    if (!d || f0 < 0 || n_frames <= 0 || f0 + n_frames > d->n_frames) return -1;
    return d->offsets[f0 + n_frames] - d->atoms_off[f0];
}

extern "C" int mdc_dump_read(mdc_dump *d, mdc_ctx *ctx, int64_t f0, int n_frames, float *pos_out, int on_device)
{
    if (!d || !ctx || !pos_out) return fail(MDC_ERR_INVALID, "mdc_dump_read: NULL argument");
    if (f0 < 0 || n_frames <= 0 || f0 + n_frames > d->n_frames) return fail(MDC_ERR_INVALID, "mdc_dump_read: frame range out of bounds");
    for (int ax = 0; ax < 3; ++ax)
        if (d->conv[ax] == 1 || d->conv[ax] == 3)
            return fail(MDC_ERR_UNSUPPORTED, "mdc_dump_read: scaled coordinates (xs / xsu) are not supported");
    MDC_CUDA(cudaSetDevice(ctx->device));
    const int N = (int)d->n_atoms, F = n_frames;
    const int64_t b0 = d->atoms_off[f0], b1 = d->offsets[f0 + F];
    const int64_t len = b1 - b0;
    if (len <= 0 || len >= ((int64_t)1 << 31)) return fail(MDC_ERR_INVALID, "mdc_dump_read: block of frames must be below 2 GiB of text");
    const int64_t padded = ceil_div(len, DT_TILE) * DT_TILE;
    const unsigned n_tiles = (unsigned)(padded / DT_TILE);
    const unsigned n_lines = (unsigned)((int64_t)F * (9 + N) - 9);
    if (!d->stream || d->stream_device != ctx->device) {
        if (d->stream) cudaStreamDestroy(d->stream);
        d->stream = nullptr;
        MDC_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        d->stream_device = ctx->device;
    }
    cudaStream_t st = d->stream;
    MDC_CHECK(d->text.alloc((size_t)padded));
    MDC_CHECK(d->tile_count.alloc(n_tiles));
    MDC_CHECK(d->line_start.alloc((size_t)n_lines + 1));
    MDC_CHECK(d->counters.alloc(4));
    MDC_CHECK(d->status.alloc(1));
    MDC_CHECK(d->fix.alloc(DT_FIXCAP));
    MDC_CHECK(d->xyz.alloc((size_t)F * N * 3));
    if (d->cols.id >= 0) {
        MDC_CHECK(d->ids.alloc((size_t)F * N));
        MDC_CHECK(d->min_id.alloc(F));
        MDC_CHECK(d->mark.alloc((size_t)F * N));
    }
    float *out_dev = pos_out;
    if (!on_device) {
        MDC_CHECK(d->out.alloc((size_t)F * N * 3));
        out_dev = d->out.p;
    }
    // raw text to HBM (pad with spaces), then rank the newlines
    MDC_CUDA(cudaMemsetAsync(d->text.p + (padded - DT_TILE), ' ', DT_TILE, st));
    MDC_CUDA(cudaMemcpyAsync(d->text.p, d->data + b0, (size_t)len, cudaMemcpyHostToDevice, st));
    MDC_CUDA(cudaMemsetAsync(d->counters.p, 0, 4 * sizeof(unsigned), st));
    MDC_CUDA(cudaMemsetAsync(d->status.p, 0, sizeof(int), st));
    MDC_CUDA(cudaMemsetAsync(d->line_start.p, 0, sizeof(unsigned), st));
    dump_count<<<n_tiles, DT_THREADS, 0, st>>>(reinterpret_cast<const uint4 *>(d->text.p), d->tile_count.p);
    dump_scan<<<1, 1024, 0, st>>>(d->tile_count.p, n_tiles, d->counters.p + 1);
    dump_starts<<<n_tiles, DT_THREADS, 0, st>>>(reinterpret_cast<const uint4 *>(d->text.p), d->tile_count.p, n_lines + 1,
                                               d->line_start.p);
    float *xyz = d->cols.id >= 0 ? d->xyz.p : out_dev;
    dump_parse_atoms<<<(unsigned)ceil_div(n_lines, DT_THREADS), DT_THREADS, 0, st>>>(
        d->text.p, d->line_start.p, n_lines, (unsigned)len, N, d->cols, d->ids.p, xyz, d->fix.p, d->counters.p, d->status.p);
    MDC_CUDA(cudaGetLastError());
    unsigned h_counters[4];
    int h_status = 0;
    MDC_CUDA(cudaMemcpyAsync(h_counters, d->counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    MDC_CUDA(cudaMemcpyAsync(&h_status, d->status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    MDC_CUDA(cudaStreamSynchronize(st));
    const unsigned total_nl = h_counters[1], expect_nl = n_lines - ((d->data[b1 - 1] == '\n') ? 0u : 1u);
    if (total_nl != expect_nl) return fail(MDC_ERR_INVALID, "mdc_dump_read: line count of the block does not match its frames");
    if (h_status & 1) return fail(MDC_ERR_INVALID, "mdc_dump_read: an atom line has fewer columns than the ATOMS header");
    // values the device's exact fast path could not convert: strtod on the host (bit-identical to Python's float())
    const unsigned n_fix = h_counters[0];
    d->last_fix = (int)n_fix;
    if (n_fix > DT_FIXCAP) return fail(MDC_ERR_UNSUPPORTED, "mdc_dump_read: too many values need the host parser");
    if (n_fix > 0) {
        std::vector<FixEntry> hf(n_fix);
        MDC_CUDA(cudaMemcpy(hf.data(), d->fix.p, n_fix * sizeof(FixEntry), cudaMemcpyDeviceToHost));
        for (const FixEntry &fe : hf) {
            const char *tok = d->data + b0 + fe.off;
            const char *te = tok;
            while (te < d->data + b1 && !isspace((unsigned char)*te)) ++te;
            const std::string s(tok, te);
            char *endp = nullptr;
            if (fe.dest & 0x80000000u) {
                const long long id = strtoll(s.c_str(), &endp, 10);
                if (endp == s.c_str() || *endp) return fail(MDC_ERR_INVALID, "invalid literal for int() with base 10: '" + s + "'");
                MDC_CUDA(cudaMemcpy(d->ids.p + (fe.dest & 0x7fffffffu), &id, sizeof(id), cudaMemcpyHostToDevice));
            } else {
                const double x = strtod(s.c_str(), &endp);
                if (endp == s.c_str() || *endp) return fail(MDC_ERR_INVALID, "could not convert string to float: '" + s + "'");
                const float v = (float)x;
                MDC_CUDA(cudaMemcpy(xyz + fe.dest, &v, sizeof(v), cudaMemcpyHostToDevice));
            }
        }
    }
    if (d->cols.id >= 0) {
        const int64_t total = (int64_t)F * N;
        const unsigned blocks = (unsigned)ceil_div(total, 256);
        std::vector<long long> big(F, 0x7fffffffffffffffll);
        MDC_CUDA(cudaMemcpyAsync(d->min_id.p, big.data(), F * sizeof(long long), cudaMemcpyHostToDevice, st));
        MDC_CUDA(cudaMemsetAsync(d->mark.p, 0, total * sizeof(int), st));
        MDC_CUDA(cudaMemsetAsync(d->status.p, 0, sizeof(int), st));
        dump_min_id<<<blocks, 256, 0, st>>>(d->ids.p, N, F, d->min_id.p);
        dump_scatter<<<blocks, 256, 0, st>>>(d->ids.p, d->xyz.p, N, F, d->min_id.p, d->mark.p, out_dev, d->status.p);
        MDC_CUDA(cudaGetLastError());
        MDC_CUDA(cudaMemcpyAsync(&h_status, d->status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        MDC_CUDA(cudaStreamSynchronize(st));
        if (h_status & 2) {
            // ids are not a contiguous range: argsort on the host, gather on the device
            std::vector<long long> hid(total);
            MDC_CUDA(cudaMemcpy(hid.data(), d->ids.p, total * sizeof(long long), cudaMemcpyDeviceToHost));
            std::vector<int> order(total);
            for (int f = 0; f < F; ++f) {
                int *o = order.data() + (int64_t)f * N;
                const long long *id = hid.data() + (int64_t)f * N;
                for (int i = 0; i < N; ++i) o[i] = i;
                std::stable_sort(o, o + N, [&](int a, int b) { return id[a] < id[b]; });
                for (int i = 0; i < N; ++i) o[i] += f * N;
            }
            MDC_CHECK(d->order.alloc(total));
            MDC_CUDA(cudaMemcpyAsync(d->order.p, order.data(), total * sizeof(int), cudaMemcpyHostToDevice, st));
            dump_gather<<<blocks, 256, 0, st>>>(d->order.p, d->xyz.p, total, out_dev);
            MDC_CUDA(cudaGetLastError());
            MDC_CUDA(cudaStreamSynchronize(st));
        }
    }
    if (!on_device)
        MDC_CUDA(cudaMemcpyAsync(pos_out, out_dev, (size_t)F * N * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    MDC_CUDA(cudaStreamSynchronize(st));   // pos_out is complete on return (the plans consume it on other streams)
    return MDC_OK;
}
