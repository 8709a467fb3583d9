// rdf.cu — g(r): RadialDistributionFunction per-frame arithmetic on sm_100a.
//
// Replaces (paths relative to /root/reference/src/mdcraft/):
//   analysis/structure.py:34-131     radial_histogram   (capped_distance + exclusion + histogram)
//   algorithm/accelerated.py:46-81   numba_histogram    (the bin formula)
// and the third-party distance routine behind them, MDAnalysis.lib.distances.capped_distance
// (brute-force orthorhombic path of calc_distances.h), whose arithmetic the pair loop reproduces
// bit for bit:  dx = (double)(b - a in float32);  s = (double)inv_box_f32 * dx;
//               dx = (double)box_f32 * (s - round(s));  r2 = (dx^2 + dy^2) + dz^2   (no FMA).
//
// Bit-exact binning without a per-pair FP64 sqrt: the reference bin
//     int(((sqrt(r2) - r0) * B) * (1 / (r1 - r0)))          (LLVM's lowering of accelerated.py:72-81)
// is monotone in r2, so the plan precomputes, on the host and with exactly that arithmetic, the
// smallest double r2 that reaches every bin (bisection over the bit patterns of r2).  On the device a
// pair's exact FP64 r2 is compared against those thresholds; FP32 only proposes the starting bin.
//
// Data layout in HBM per frame block: raw f32 (F, N, 3) as pushed -> SoA f32 (F, 3, N) per group;
// counts u64 (B) accumulate over all frames pushed.  Histograms are privatised per warp in shared
// memory (u32) and flushed with 64-bit global atomics.
#include "common.cuh"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <vector>

using namespace mdc;

namespace {

struct FrameBox {
    double box[3];
    double inv[3];
    int pbc[3];
    int pad;
};

// Triclinic frames (any box angle != 90): the lower-triangular box matrix of MDAnalysis' triclinic_vectors, float32
// values widened: a = (ax, 0, 0), b = (bx, by, 0), c = (cx, cy, cz).  on == 0: orthorhombic frame (FrameBox applies).
struct FrameTri {
    double ax, bx, by, cx, cy, cz;
    double w[3];      // perpendicular widths of the cell along a, b, c (distance between opposite faces)
    int on, pad;
};

// ------------------------------------------------------------------------------------------
// host: thresholds (the plan's own statement of numba_histogram's arithmetic)
// ------------------------------------------------------------------------------------------

inline double from_bits(uint64_t u)
{
    double d;
    memcpy(&d, &u, sizeof(d));
    return d;
}

template <typename Pred>
double first_true(Pred pred)
{
    // smallest non-negative double x with pred(x); pred is monotone (false ... false true ... true).
    // Non-negative doubles order like their bit patterns.
    uint64_t lo = 0, hi = 0x7ff0000000000000ull;   // +inf: returned when pred never holds
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (pred(from_bits(mid))) hi = mid;
        else lo = mid + 1;
    }
    return from_bits(lo);
}

struct Thresholds {
    std::vector<double> T;   // B + 1 entries: r2 in [T[k], T[k+1]) -> bin k
    double e_lo, e_hi;       // r2 in [e_lo, e_hi): d == r1 exactly -> last bin
};

Thresholds make_thresholds(int B, double r0, double r1)
{
    // accelerated.py:72-81 under fastmath, as lowered by LLVM: reciprocal hoisted, multiply, fptosi
    volatile double nd = (double)B;
    volatile double invw = 1.0 / (r1 - r0);
    auto bin = [&](double d) -> int64_t {
        volatile double t = (d - r0);
        t = t * nd;
        t = t * invw;
        if (!(t < 9.0e18)) return INT64_MAX;
        return (int64_t)t;
    };
    const double min_cut = r0 - DBL_EPSILON;   // structure.py:112-113
    Thresholds th;
    th.T.resize(B + 1);
    th.e_lo = first_true([&](double r2) { return sqrt(r2) >= r1; });
    th.e_hi = first_true([&](double r2) { return sqrt(r2) > r1; });
    th.T[0] = first_true([&](double r2) { return sqrt(r2) > min_cut; });
    for (int k = 1; k <= B; ++k) {
        double t = first_true([&](double r2) { return bin(sqrt(r2)) >= k; });
        if (t < th.T[k - 1]) t = th.T[k - 1];
        if (t > th.e_lo) t = th.e_lo;
        th.T[k] = t;
    }
    if (th.T[0] > th.e_lo) th.T[0] = th.e_lo;
    // d == r1 goes to the last bin (accelerated.py:75-76).  In the normal case the formula's own last
    // bin ends exactly where d == r1 begins and the two ranges merge.
    if (th.T[B] == th.e_lo) th.T[B] = th.e_hi;
    return th;
}

// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------

// Per-frame constants of the filter kernel (written by rdf_frame_params).
struct FrameFast {
    float ax, ay, az;   // (float)(int32)(u_j - u_i) * a_k  =  separation along k in units of one bin width
    float c0;           // v = |separation| + c0 = bin coordinate - 0.5   (c0 = -r0 * S - 0.5)
    float thr;          // 0.5 - margin: a pair is decided in FP32 iff |v - rint(v)| <= thr
    int kind;           // FK_CUBIC: ax == ay == az; FK_ORTHO; FK_TRI: triclinic cell, the separation is
                        //   (tx ax + ty bx + tz cx, ty ay + tz cy, tz az) with t = differences of the FRACTIONAL coordinates
    int usable;         // 0: this frame must take the exact kernel's arithmetic for every pair
    float bx, cx, cy;   // FK_TRI: off-diagonal entries of the cell matrix, scaled like ax / ay / az (= its diagonal)
};
enum { FK_CUBIC = 0, FK_ORTHO = 1, FK_TRI = 2 };

// MDAnalysis' _triclinic_pbc for one coordinate (calc_distances.h; restated in oracle/oracle.c:orc_triclinic_wrap, same
// operations in the same order, no FMA contraction): into the primary cell along c, then b, then a.
__device__ __forceinline__ void tri_wrap(float &xf, float &yf, float &zf, const FrameTri &t)
{
    const double bi0 = __ddiv_rn(1.0, t.ax), bi4 = __ddiv_rn(1.0, t.by), bi8 = __ddiv_rn(1.0, t.cz);
    const double a_y = __dmul_rn(t.bx, bi4), b_z = __dmul_rn(t.cy, bi8);
    const double a_z = __dmul_rn(__dsub_rn(t.cx, __dmul_rn(t.cy, a_y)), bi8);
    double x = (double)xf, y = (double)yf, z = (double)zf;
    if (z < 0.0 || z >= t.cz) {
        const double s = floor(__dmul_rn(z, bi8));
        z = __dsub_rn(z, __dmul_rn(s, t.cz));
        y = __dsub_rn(y, __dmul_rn(s, t.cy));
        x = __dsub_rn(x, __dmul_rn(s, t.cx));
    }
    double lb = __dmul_rn(z, b_z);
    if (y < lb || y >= __dadd_rn(lb, t.by)) {
        const double s = floor(__dmul_rn(__dsub_rn(y, lb), bi4));
        y = __dsub_rn(y, __dmul_rn(s, t.by));
        x = __dsub_rn(x, __dmul_rn(s, t.bx));
    }
    lb = __dadd_rn(__dmul_rn(y, a_y), __dmul_rn(z, a_z));
    if (x < lb || x >= __dadd_rn(lb, t.ax)) {
        const double s = floor(__dmul_rn(__dsub_rn(x, lb), bi0));
        x = __dsub_rn(x, __dmul_rn(s, t.ax));
    }
    xf = (float)x;
    yf = (float)y;
    zf = (float)z;
}

// raw f32 (F, N, 3) -> soa f32 (F, 3, N); also fix u32 (F, 3, N) = frac(x / L) * 2^32 and, per frame, the largest
// coordinate and the largest NEGATED coordinate (float bits, atomicMax; both clamped at 0): their sum bounds the range
// of the coordinates, hence every |a - b| that the reference subtracts in float32.  extent = {max x, max -x} x (F).
// Triclinic frames: soa holds the coordinates moved into the primary cell (what the reference's distance routine
// works on); fix the fractions of THOSE along the cell vectors a, b, c (the filter kernel's minimum image is the wrap
// of the fractional differences, valid for cutoffs below half the smallest width of the cell: see rdf_frame_params).
__global__ void rdf_prep(const float *__restrict__ raw, int64_t N, int64_t total, const FrameBox *__restrict__ boxes,
                         const FrameTri *__restrict__ tri, int F, float *__restrict__ soa, uint32_t *__restrict__ fix,
                         unsigned *__restrict__ extent)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float mp = 0.f, mn = 0.f;
    int64_t f = 0;
    if (i < total) {
        f = i / N;
        const int64_t j = i - f * N;
        const FrameBox fb = boxes[f];
        float p[3] = {raw[3 * i], raw[3 * i + 1], raw[3 * i + 2]};
        const FrameTri ft = tri[f];
        if (ft.on) tri_wrap(p[0], p[1], p[2], ft);
        // triclinic frames: fractions along a, b, c of the (wrapped) position, x = sa a + sb b + sc c
        double fr[3] = {0.0, 0.0, 0.0};
        if (ft.on) {
            fr[2] = (double)p[2] / ft.cz;
            fr[1] = ((double)p[1] - fr[2] * ft.cy) / ft.by;
            fr[0] = ((double)p[0] - fr[1] * ft.bx - fr[2] * ft.cx) / ft.ax;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float x = p[k];
            soa[(f * 3 + k) * N + j] = x;
            double s = ft.on ? fr[k] : (fb.pbc[k] ? (double)x / fb.box[k] : 0.0);
            s -= floor(s);
            fix[(f * 3 + k) * N + j] = (uint32_t)(unsigned long long)(s * 4294967296.0 + 0.5);
            mp = fmaxf(mp, x);
            mn = fmaxf(mn, -x);
        }
    }
    // one pair of atomics per warp (lanes of a warp may straddle two frames: fall back to per-lane atomics then)
    const int64_t f0 = __shfl_sync(0xffffffffu, f, 0);
    const bool same = __all_sync(0xffffffffu, (i >= total) || f == f0);
    if (same) {
        for (int o = 16; o > 0; o >>= 1) {
            mp = fmaxf(mp, __shfl_xor_sync(0xffffffffu, mp, o));
            mn = fmaxf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (mp > 0.f) atomicMax(&extent[f0], __float_as_uint(mp));
            if (mn > 0.f) atomicMax(&extent[F + f0], __float_as_uint(mn));
        }
    } else if (i < total) {
        if (mp > 0.f) atomicMax(&extent[f], __float_as_uint(mp));
        if (mn > 0.f) atomicMax(&extent[F + f], __float_as_uint(mn));
    }
}

// One thread per frame: turns the box, the plan's range and the coordinate range into the filter's constants.  The
// margin bounds |v_filter - v_reference|:
//   reference vs true separation, per axis: float32 subtraction of two coordinates (<= 2^-24 |a - b|) and the
//     (box * inv_box) factor of its minimum image (inverse box stored as float32: <= 2^-24 |a - b|); |a - b| <= R, the
//     range of the coordinates of both groups (R = L for wrapped positions): 2^-23 R.  A pair within that of half a box
//     may take the other image in the reference; its |separation| then differs by <= 2^-24 L <= 2^-23 R as well;
//   fixed point vs true separation: 2^-32 L per axis (2^-31 L is used);
//   FP32 evaluation (I2F, squares, sum, MUFU.SQRT, FMA): <= 3.5e-7 relative + one ulp of v (4.7e-7 is used);
//   all of it times 1.25.
__global__ void rdf_frame_params(const FrameBox *__restrict__ boxes, const FrameTri *__restrict__ tri,
                                 const unsigned *__restrict__ extent_a, const unsigned *__restrict__ extent_b, int Fa, int F,
                                 int n_bins, double r0, double r1, FrameFast *__restrict__ out)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const FrameBox fb = boxes[f];
    const FrameTri ft = tri[f];
    const double S = (double)n_bins / (r1 - r0);
    const double hi = fmax((double)__uint_as_float(extent_a[f]), (double)__uint_as_float(extent_b[f]));
    const double lo = fmax((double)__uint_as_float(extent_a[Fa + f]), (double)__uint_as_float(extent_b[Fa + f]));
    const double R = hi + lo;   // >= max - min over both groups and all three axes
    FrameFast p;
    p.c0 = (float)(-r0 * S - 0.5);
    p.bx = p.cx = p.cy = 0.f;
    const double vmax = (double)n_bins + 2.0 + fabs(r0 * S);
    double margin, v_far;
    bool ok = true;
    if (ft.on) {
        // Triclinic cell.  The filter's image of a pair is the wrap of its FRACTIONAL differences (|s_k| <= 1/2).  A
        // separation shorter than half the smallest width w of the cell has |s_k| = |d . n_k| / w_k < 1/2 on every
        // axis, so it IS that image; and if that image is longer than the cutoff, so is the true one (or it is at
        // least w / 2): for r1 < w / 2 the filter and the reference (minimum_image_triclinic: the shortest of 27
        // images) see the same image of every pair that can lie in the window.
        p.ax = (float)(ft.ax * S / 4294967296.0);
        p.ay = (float)(ft.by * S / 4294967296.0);
        p.az = (float)(ft.cz * S / 4294967296.0);
        p.bx = (float)(ft.bx * S / 4294967296.0);
        p.cx = (float)(ft.cx * S / 4294967296.0);
        p.cy = (float)(ft.cy * S / 4294967296.0);
        p.kind = FK_TRI;
        const double la = fabs(ft.ax), lb = sqrt(ft.bx * ft.bx + ft.by * ft.by),
                     lc = sqrt(ft.cx * ft.cx + ft.cy * ft.cy + ft.cz * ft.cz), Lsum = la + lb + lc;
        // reference: float32 subtraction of two wrapped coordinates per axis (<= 2^-24 R; the image shifts are exact);
        // filter: fractions to 2^-32 (2^-31 |a| + |b| + |c|), the two conversions of the relative rows (2^-26 + 2^-28),
        // and the matrix product in FP32: six roundings of terms of up to half a cell vector (2^-22 (|a| + |b| + |c|))
        const double e_abs = 1.7320508 * R / 8388608.0 +
                             Lsum * (1.0 / 2147483648.0 + 1.0 / 67108864.0 + 1.0 / 268435456.0 + 1.0 / 4194304.0);
        margin = 1.25 * (S * e_abs + 4.7e-7 * vmax) + 1e-6;
        v_far = 0.5 * Lsum * S + fabs(r0 * S) + 1.0;
        const double wmin = fmin(ft.w[0], fmin(ft.w[1], ft.w[2]));
        ok = r1 * (1.0 + 1e-4) + 1e-5 * Lsum < 0.5 * wmin;
    } else {
        const double Lmax = fmax(fb.box[0], fmax(fb.box[1], fb.box[2]));
        p.ax = (float)(fb.box[0] * S / 4294967296.0);
        p.ay = (float)(fb.box[1] * S / 4294967296.0);
        p.az = (float)(fb.box[2] * S / 4294967296.0);
        // + the relative-coordinate rows of the filter (rows_rel): the row's difference to the box centre is rounded to FP32
        // on its own (<= 2^-26 L: 32-bit integers below 2^31), the particle's too (<= 2^-28 L: boxes below 2^-3 L); the
        // rounding of their FP32 difference is relative (<= 1.3 x 2^-24 of the wrapped difference at the seam) and part of
        // the 4.7e-7 below, like the conversion it replaces
        const double e_abs = 1.7320508 * (R / 8388608.0 + Lmax / 2147483648.0 + Lmax * (1.0 / 67108864.0 + 1.0 / 268435456.0));
        margin = 1.25 * (S * e_abs + 4.7e-7 * vmax) + 1e-6;
        p.kind = (p.ax == p.ay && p.ay == p.az) ? FK_CUBIC : FK_ORTHO;
        v_far = 0.5 * sqrt(fb.box[0] * fb.box[0] + fb.box[1] * fb.box[1] + fb.box[2] * fb.box[2]) * S + fabs(r0 * S) + 1.0;
        ok = fb.pbc[0] && fb.pbc[1] && fb.pbc[2];
    }
    p.thr = (float)fmax(0.5 - margin, -1.0);
    // the magic-number rounding below needs |v| < 2^22; all three axes must be periodic
    p.usable = (ok && v_far < 4.0e6 && margin < 0.5) ? 1 : 0;
    out[f] = p;
}

struct RdfArgs {
    const float *a, *b;        // SoA f32 (F, 3, Na), (F, 3, Nb); b == a for a self-RDF
    const uint32_t *ua, *ub;   // SoA u32 fixed-point fractions, same shapes
    const FrameBox *boxes;     // (F)
    const FrameTri *tri;       // (F): triclinic frames (exact kernel only)
    const FrameFast *fast;     // (F)
    const double *T;           // (B + 1)
    unsigned long long *counts;
    double e_lo, e_hi;
    float r0f, scalef;         // candidate bin = (sqrt(r2) - r0) * B / (r1 - r0) in FP32
    int64_t na, nb;
    int64_t ex0, ex1;          // ex0 <= 0: no exclusion
    int n_bins, n_copies;
    int F, nti, ntj;           // tiles
    int sym;                   // 1: self-RDF over unordered pairs (each counted twice)
    int64_t items_per_frame;
    unsigned long long *ticket;   // next tile pair to hand out (filter kernel: blocks draw work items dynamically)
    // spatially sorted frames (filter kernel, cutoffs well below half the box): particles are in Hilbert order of a
    // cell grid, so a tile of consecutive particles is a compact region and tile pairs farther apart than the
    // cutoff are skipped.  oa / ob: original index of every sorted particle (the exclusion i / ex0 == j / ex1 is
    // defined on those); tbox_a / tbox_b: bounding box of every tile {lo x, hi x, lo y, hi y, lo z, hi z} as
    // 32-bit box fractions, (F, n_tiles, 6).  All NULL: original order, no culling.
    const int *oa, *ob;
    const uint32_t *tbox_a, *tbox_b;
    double cull_r;                // a tile pair is skipped iff its boxes are farther apart than this (minimum image)
    // filter kernel, extended histograms: ext_k > 0 = words per histogram copy, ext_off = index of bin 0 in a copy.
    // Every bin coordinate the filter can produce in these frames, in the window or not, then has a word of its
    // own, so the hot loop increments without a range test (bins outside [0, B) are never read back).
    int ext_k, ext_off;
    // filter kernel: histogram layout in shared memory, in words: bin k of copy c sits at k * hist_ks + c * hist_cs.
    // Copies one after the other (hist_ks = 1, hist_cs = words per copy, 8 copies chosen by lane & 7), or -- extended
    // histograms that fit -- 32 INTERLEAVED copies, one per lane (hist_ks = 32, hist_cs = 1): every lane then owns a
    // bank and the increments of a warp never conflict (tools/atoms_probe.cu: 1.0 instead of 3.1 cycles per ATOMS).
    int hist_ks, hist_cs;
    int qcap;                     // filter kernel: capacity of the queue of undecided rows (words, after the histograms)
    unsigned char *ovf;           // filter kernel: per block, RF_TJ x RF_THREADS bytes for the rows that do not fit the queue (park_row)
    int rel;                      // filter kernel, sorted frames: relative-coordinate rows allowed (MDC_RDF_REL=0 turns them off)
};

// ---- spatial sort (Hilbert order of an nc^3 cell grid, nc = 2^bits <= 32) ----------------------------------------

// Hilbert index of cell (x, y, z) of a (2^bits)^3 grid (Skilling's transpose form).  Consecutive indices are
// face-adjacent cells, so a run of consecutive particles has a tighter bounding box than in Morton order (measured:
// half as many chunk pairs survive the culling at small cutoffs).
__device__ __forceinline__ unsigned hilbert3(unsigned x, unsigned y, unsigned z, int bits)
{
    unsigned X[3] = {x, y, z};
    const unsigned M = 1u << (bits - 1);
    for (unsigned Q = M; Q > 1u; Q >>= 1) {
        const unsigned P = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else {
                const unsigned t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    unsigned t = 0u;
    for (unsigned Q = M; Q > 1u; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1u;
    X[0] ^= t;
    X[1] ^= t;
    X[2] ^= t;
    unsigned k = 0u;
    for (int b = 0; b < bits; ++b)
        k |= (((X[0] >> b) & 1u) << (3 * b + 2)) | (((X[1] >> b) & 1u) << (3 * b + 1)) | (((X[2] >> b) & 1u) << (3 * b));
    return k;
}

__device__ __forceinline__ unsigned cell_key(const uint32_t *fix, int64_t f, int64_t N, int64_t j, int bits)
{
    const int sh = 32 - bits;
    return bits == 0 ? 0u : hilbert3(fix[(f * 3 + 0) * N + j] >> sh, fix[(f * 3 + 1) * N + j] >> sh, fix[(f * 3 + 2) * N + j] >> sh, bits);
}

__global__ void rdf_cell_count(const uint32_t *__restrict__ fix, int64_t N, int64_t total, int bits, int *__restrict__ cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t f = i / N, j = i - f * N;
    atomicAdd(&cnt[(f << (3 * bits)) + cell_key(fix, f, N, j, bits)], 1);
}

// exclusive scan of 8^bits <= 32768 counters per frame (one block of 1024 threads per frame)
__global__ void __launch_bounds__(1024) rdf_cell_scan(int bits, int *__restrict__ cnt)
{
    __shared__ int sh[1024];
    const int n = 1 << (3 * bits);
    int *c = cnt + ((int64_t)blockIdx.x << (3 * bits));
    const int per = (n + 1023) / 1024, lo = threadIdx.x * per, hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += c[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        const int v = c[i];
        c[i] = run;
        run += v;
    }
}

// soa / fix (F, 3, N) -> the same in cell order, plus the original index of every sorted particle
__global__ void rdf_cell_scatter(const float *__restrict__ soa, const uint32_t *__restrict__ fix, int64_t N, int64_t total,
                                 int bits, int *__restrict__ cursor, float *__restrict__ soa_s, uint32_t *__restrict__ fix_s,
                                 int *__restrict__ orig)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t f = i / N, j = i - f * N;
    const int pos = atomicAdd(&cursor[(f << (3 * bits)) + cell_key(fix, f, N, j, bits)], 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        soa_s[(f * 3 + k) * N + pos] = soa[(f * 3 + k) * N + j];
        fix_s[(f * 3 + k) * N + pos] = fix[(f * 3 + k) * N + j];
    }
    orig[f * N + pos] = (int)j;
}

// bounding box of every tile of `tile` consecutive (sorted) particles: one block per (tile, frame)
__global__ void __launch_bounds__(256) rdf_tile_boxes(const uint32_t *__restrict__ fix, int64_t N, int tile, int n_tiles,
                                                      uint32_t *__restrict__ tbox)
{
    __shared__ uint32_t lo[3][256], hi[3][256];
    const int t = blockIdx.x, f = blockIdx.y;
    const int64_t i0 = (int64_t)t * tile, i1 = min(N, i0 + tile);
    uint32_t l[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, h[3] = {0u, 0u, 0u};
    for (int64_t i = i0 + threadIdx.x; i < i1; i += 256)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = fix[((int64_t)f * 3 + k) * N + i];
            l[k] = min(l[k], v);
            h[k] = max(h[k], v);
        }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k][threadIdx.x] = l[k];
        hi[k][threadIdx.x] = h[k];
    }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                lo[k][threadIdx.x] = min(lo[k][threadIdx.x], lo[k][threadIdx.x + o]);
                hi[k][threadIdx.x] = max(hi[k][threadIdx.x], hi[k][threadIdx.x + o]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 3) {
        uint32_t *o = tbox + ((int64_t)f * n_tiles + t) * 6;
        o[2 * threadIdx.x] = lo[threadIdx.x][0];
        o[2 * threadIdx.x + 1] = hi[threadIdx.x][0];
    }
}

// squared minimum-image distance between two tile boxes (fractions of the box per axis, scaled by the box lengths)
// Triclinic frames: the boxes bound the FRACTIONS along a, b, c; two points whose fractions differ by g along k are at
// least g w_k apart (w_k: width of the cell along k), so the bound is the LARGEST of the three, not their root sum.
__device__ __forceinline__ double tile_gap2(const uint32_t *a, const uint32_t *b, const FrameBox &fb, const FrameTri *ftp)
{
    double d2 = 0.0;
    const bool tri = ftp->on != 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double alo = a[2 * k] * (1.0 / 4294967296.0), ahi = a[2 * k + 1] * (1.0 / 4294967296.0);
        const double blo = b[2 * k] * (1.0 / 4294967296.0), bhi = b[2 * k + 1] * (1.0 / 4294967296.0);
        double gap = 1.0;
#pragma unroll
        for (int m = -1; m <= 1; ++m) gap = fmin(gap, fmax(0.0, fmax(blo + m - ahi, alo - (bhi + m))));
        gap *= tri ? ftp->w[k] : fb.box[k];
        d2 = tri ? fmax(d2, gap * gap) : d2 + gap * gap;
    }
    return d2;
}

__device__ __forceinline__ double min_image_sq(float dxf, double inv, double box, int pbc)
{
    double dx = (double)dxf;
    if (pbc) {
        const double s = __dmul_rn(inv, dx);
        // round to nearest integer; ties go to even instead of away from zero, which only flips the
        // sign of a +-box/2 separation and leaves its square unchanged
        const double r = __dsub_rn(__dadd_rn(s, 6755399441055744.0), 6755399441055744.0);
        dx = __dmul_rn(box, __dsub_rn(s, r));
    }
    return __dmul_rn(dx, dx);
}

// The reference's r^2 for one pair, bit for bit (calc_distances.h arithmetic, see file header).
__device__ __forceinline__ double exact_r2(float xi, float yi, float zi, float xj, float yj, float zj,
                                           const FrameBox &fb)
{
    double r2 = min_image_sq(xj - xi, fb.inv[0], fb.box[0], fb.pbc[0]);
    r2 = __dadd_rn(r2, min_image_sq(yj - yi, fb.inv[1], fb.box[1], fb.pbc[1]));
    r2 = __dadd_rn(r2, min_image_sq(zj - zi, fb.inv[2], fb.box[2], fb.pbc[2]));
    return r2;
}

// The reference's r^2 for one pair of a TRICLINIC frame, bit for bit (calc_distances.h: minimum_image_triclinic on the
// float32 difference of coordinates already moved into the primary cell; restated in oracle/oracle.c): the shortest
// of the 27 images dx + ix a + iy b + iz c.  The image shifts are +-(one float32 matrix entry), exact in double.
__device__ __forceinline__ double tri_r2(float xi, float yi, float zi, float xj, float yj, float zj, const FrameTri &t)
{
    const double d0 = (double)(xj - xi), d1 = (double)(yj - yi), d2 = (double)(zj - zi);
    double best = 3.4028234663852886e38;   // FLT_MAX
#pragma unroll
    for (int ix = -1; ix < 2; ++ix) {
        const double rx = __dadd_rn(d0, ix * t.ax);
#pragma unroll
        for (int iy = -1; iy < 2; ++iy) {
            const double ry0 = __dadd_rn(rx, iy * t.bx), ry1 = __dadd_rn(d1, iy * t.by);
#pragma unroll
            for (int iz = -1; iz < 2; ++iz) {
                const double rz0 = __dadd_rn(ry0, iz * t.cx), rz1 = __dadd_rn(ry1, iz * t.cy), rz2 = __dadd_rn(d2, iz * t.cz);
                const double dsq = __dadd_rn(__dadd_rn(__dmul_rn(rz0, rz0), __dmul_rn(rz1, rz1)), __dmul_rn(rz2, rz2));
                best = fmin(best, dsq);
            }
        }
    }
    return best;
}

// Exact bin of an exact r^2 (-1: not counted), by comparison with the precomputed thresholds.
__device__ __forceinline__ int exact_bin(double r2, const RdfArgs &A, const double *sT)
{
    if (!(r2 >= sT[0]) || !(r2 < A.e_hi)) return -1;
    if (r2 < sT[A.n_bins]) {
        const float d = sqrtf((float)r2);
        int k = (int)((d - A.r0f) * A.scalef);
        k = max(0, min(A.n_bins - 1, k));
        while (r2 < sT[k]) --k;
        while (r2 >= sT[k + 1]) ++k;
        return k;
    }
    return (r2 >= A.e_lo) ? A.n_bins - 1 : -1;
}

// ---- exact kernel: FP64 for every pair (frames the filter cannot take: non-periodic axes, huge v) ----

constexpr int RDF_THREADS = 256;
constexpr int RDF_IPT = 2;                       // i-particles per thread
constexpr int RDF_TI = RDF_THREADS * RDF_IPT;    // i-tile
constexpr int RDF_TJ = 512;                      // j-tile staged in shared memory
constexpr int RDF_FLUSH_ITEMS = 4096;            // u32 privatised counters: 4096 * 512 * 512 < 2^31

template <bool CHECK>
__device__ __forceinline__ void tile_pairs(const RdfArgs &A, const float4 *__restrict__ sj, int nj, int64_t j0,
                                           const float (&xi)[RDF_IPT], const float (&yi)[RDF_IPT],
                                           const float (&zi)[RDF_IPT], const int (&bi)[RDF_IPT],
                                           const int64_t (&ig)[RDF_IPT], const FrameBox &fb, const FrameTri &ft,
                                           const double *sT, unsigned *hist, bool diag)
{
    for (int j = 0; j < nj; ++j) {
        const float4 pj = sj[j];
        const int bj = __float_as_int(pj.w);
#pragma unroll
        for (int u = 0; u < RDF_IPT; ++u) {
            const double r2 = ft.on ? tri_r2(xi[u], yi[u], zi[u], pj.x, pj.y, pj.z, ft)
                                    : exact_r2(xi[u], yi[u], zi[u], pj.x, pj.y, pj.z, fb);
            bool ok = bi[u] != bj;
            if (CHECK) ok = ok && (ig[u] < A.na) && (!diag || (j0 + j) > ig[u]);
            if (ok) {
                const int k = exact_bin(r2, A, sT);
                if (k >= 0) atomicAdd(&hist[k], 1u);
            }
        }
    }
}

// decodes a work item into (frame, I, J); false if the item is a duplicate to skip
__device__ __forceinline__ bool decode_item(const RdfArgs &A, int64_t item, int &f, int &I, int &J, bool &diag)
{
    f = (int)(item / A.items_per_frame);
    const int64_t w = item - (int64_t)f * A.items_per_frame;
    diag = false;
    if (A.sym) {
        // unordered tile pairs: J = (I + off) mod nt, off in [0, nt/2]; for even nt the
        // off == nt/2 column would visit each pair twice, keep the lower half
        const int nt = A.nti, half = nt / 2 + 1;
        I = (int)(w / half);
        const int off = (int)(w - (int64_t)I * half);
        J = I + off;
        if (J >= nt) J -= nt;
        diag = off == 0;
        if (((nt & 1) == 0) && (off == nt / 2) && (I >= nt / 2)) return false;
    } else {
        I = (int)(w / A.ntj);
        J = (int)(w - (int64_t)I * A.ntj);
    }
    return true;
}

__device__ __forceinline__ void flush_hist(unsigned *sh, int B, int n_copies, unsigned long long weight,
                                           unsigned long long *counts, bool clear, int stride = 0, int ks = 1)
{
    if (stride == 0) stride = B;   // sh points at bin 0 of copy 0; copies are `stride` words apart, bins `ks` words
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        unsigned long long s = 0;
        for (int c = 0; c < n_copies; ++c) {
            s += sh[c * stride + i * ks];
            if (clear) sh[c * stride + i * ks] = 0u;
        }
        if (s) atomicAdd(&counts[i], s * weight);
    }
}

__global__ void __launch_bounds__(RDF_THREADS)
rdf_pairs_exact(RdfArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sj = reinterpret_cast<float4 *>(smem_raw);
    double *sT = reinterpret_cast<double *>(smem_raw + sizeof(float4) * RDF_TJ);
    unsigned *sh = reinterpret_cast<unsigned *>(sT + (A.n_bins + 1));
    const int B = A.n_bins;
    for (int i = threadIdx.x; i <= B; i += RDF_THREADS) sT[i] = A.T[i];
    for (int i = threadIdx.x; i < B * A.n_copies; i += RDF_THREADS) sh[i] = 0u;
    unsigned *hist = sh + ((threadIdx.x >> 5) & (A.n_copies - 1)) * B;
    const unsigned long long weight = A.sym ? 2ull : 1ull;
    const int64_t total = A.items_per_frame * A.F;
    int since_flush = 0;

    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        int f, I, J;
        bool diag;
        if (!decode_item(A, item, f, I, J, diag)) continue;
        const FrameBox fb = A.boxes[f];
        const FrameTri ft = A.tri[f];
        const float *ax = A.a + ((int64_t)f * 3) * A.na, *ay = ax + A.na, *az = ay + A.na;
        const float *bx = A.b + ((int64_t)f * 3) * A.nb, *by = bx + A.nb, *bz = by + A.nb;

        float xi[RDF_IPT], yi[RDF_IPT], zi[RDF_IPT];
        int bi[RDF_IPT];
        int64_t ig[RDF_IPT];
#pragma unroll
        for (int u = 0; u < RDF_IPT; ++u) {
            ig[u] = (int64_t)I * RDF_TI + threadIdx.x + u * RDF_THREADS;
            const int64_t ii = min(ig[u], A.na - 1);
            xi[u] = ax[ii];
            yi[u] = ay[ii];
            zi[u] = az[ii];
            bi[u] = A.ex0 > 0 ? (int)(ii / A.ex0) : -1;
        }
        const int64_t j0 = (int64_t)J * RDF_TJ;
        const int nj = (int)min((int64_t)RDF_TJ, A.nb - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < nj; j += RDF_THREADS) {
            const int64_t jj = j0 + j;
            const int bj = A.ex0 > 0 ? (int)(jj / A.ex1) : -2;
            sj[j] = make_float4(bx[jj], by[jj], bz[jj], __int_as_float(bj));
        }
        __syncthreads();
        const bool edge = (int64_t)(I + 1) * RDF_TI > A.na;
        if (diag || edge)
            tile_pairs<true>(A, sj, nj, j0, xi, yi, zi, bi, ig, fb, ft, sT, hist, diag);
        else
            tile_pairs<false>(A, sj, nj, j0, xi, yi, zi, bi, ig, fb, ft, sT, hist, false);

        if (++since_flush >= RDF_FLUSH_ITEMS) {
            __syncthreads();
            flush_hist(sh, B, A.n_copies, weight, A.counts, true);
            since_flush = 0;
        }
    }
    __syncthreads();
    flush_hist(sh, B, A.n_copies, weight, A.counts, false);
}

// ---- filter kernel: integer minimum image + FP32 bin proposal, FP64 only for undecided pairs ----
//
// Positions are 32-bit binary fractions of the box, so u_j - u_i wraps to the minimum image for
// free.  The FP32 bin coordinate v differs from the reference's by at most `margin` (FrameFast);
// a pair whose v is further than that from a bin edge is binned at once (shared-memory atomic),
// the rest (~1e-3 of the pairs) are queued and resolved with the exact arithmetic after the tile,
// so the counts are exactly those of the exact kernel.

constexpr int RF_THREADS = 256;
constexpr int RF_IPT = 4;
constexpr int RF_TI = RF_THREADS * RF_IPT;   // 1024
constexpr int RF_TJ = 1024;
constexpr int RF_QCAP = 4096;                // undecided rows per tile pair before falling back to in-line resolution
constexpr int RF_QCAP_L32 = 1024;            // ... when the 32 interleaved histogram copies need the room
constexpr int RF_FLUSH_ITEMS = 1024;         // 1024 * 1024 * 1024 pairs < 2^31
// the filter rounds with the magic number 1.5 * 2^23 = 12582912.0f = 0x4B400000 (bin_or_flag)

struct TileCtx;
struct FastShared {
    uint4 sj[RF_TJ];            // {ux, uy, uz, exclusion block of j}
    // undecided rows: row j | lane << 10 | group << 15 | mask of the lane's four particles << 18 | relative form << 22;
    // RdfArgs::qcap words behind the histograms
    unsigned *queue;
    unsigned qcount;
    unsigned qcap;
    unsigned char *ovf;         // this block's overflow table (RdfArgs::ovf), RF_TJ x RF_THREADS bytes, all zero between tile pairs
    unsigned overflow;          // a row of this tile pair went to the table
    unsigned pad0[3];
    unsigned trash[RF_THREADS]; // per-thread dummy target of the branch-free histogram increment
    unsigned char ctx[512];     // TileCtx of the tile in flight (cold passes read it from here)
    // sorted frames only: bounding boxes {lo x, lo y, lo z, -, hi x, hi y, hi z, -} (32-bit box fractions) of the 128
    // particles i of every warp (followed by the box's centre and half extents) and of every chunk of 32 rows j, and
    // {L_k / 2^32 (k = x, y, z), cull_r^2}
    uint4 ibox[RF_THREADS / 32][8];   // + {centre}, {half extent} of the box, of its first 64 and of its last 64 particles
    uint4 jbox[RF_TJ / 32][2];
    float cull[4];
    int cull_tri;               // triclinic frame: the boxes hold fractions along the cell vectors (gap_norm2)
    int pad3[3];
    // sorted frames: next chunk of 32 rows to hand out for the 128 particles i of every warp ("group").  A warp draws the
    // chunks of its own group from here and, once those are gone, the remaining chunks of the other groups (it then
    // reloads that group's particles): uneven culling no longer leaves warps idle at the end of a tile pair.
    int next_chunk[RF_THREADS / 32];
    int pad2[8 - (RF_THREADS / 32) % 8];
    // sorted frames, relative-coordinate rows (rows_rel): per warp, the 32 rows of the chunk in flight as FP32
    // differences to the centre of the group's box {x, y, z, -}
    float4 rel[RF_THREADS / 32][32];
};

// local index (within the tile) of particle u of this thread: a warp holds 128 consecutive particles, which in a
// sorted frame is a compact region
__device__ __forceinline__ int rf_il(int u) { return (int)(threadIdx.x & ~31u) * RF_IPT + u * 32 + (int)(threadIdx.x & 31u); }
// the same for the particles of group g (the warp's own group is g = threadIdx.x / 32)
__device__ __forceinline__ int rf_il_of(int g, int u) { return g * (32 * RF_IPT) + u * 32 + (int)(threadIdx.x & 31u); }

// sorted frames: is every particle of box a farther than the cutoff from every particle of box b (minimum image)?
// Warp-uniform.  Per axis the gap between two intervals on the 2^32 circle that do not overlap is the smaller of the
// two wrapped differences.
__device__ __forceinline__ float gap_norm2(float gx, float gy, float gz, bool tri)
{
    // orthorhombic: Euclidean; triclinic (gaps along the cell vectors times the cell's widths): the largest one
    return tri ? fmaxf(gx * gx, fmaxf(gy * gy, gz * gz)) : fmaf(gz, gz, fmaf(gy, gy, gx * gx));
}

__device__ __forceinline__ bool boxes_apart(const uint4 *a, const uint4 *b, const float (&c)[4], bool tri)
{
    auto gap = [](uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi) -> float {
        const bool overlap = blo <= ahi && alo <= bhi;
        return overlap ? 0.f : (float)min(blo - ahi, alo - bhi);
    };
    const float gx = gap(a[0].x, a[1].x, b[0].x, b[1].x) * c[0];
    const float gy = gap(a[0].y, a[1].y, b[0].y, b[1].y) * c[1];
    const float gz = gap(a[0].z, a[1].z, b[0].z, b[1].z) * c[2];
    return gap_norm2(gx, gy, gz, tri) > c[3];
}

// sorted frames: which of the 32 rows j of a chunk (one row per lane) have ANY of this warp's 128 particles within the
// cutoff?  Gap between the point and the warp's box per axis: the wrapped difference to the box centre is the
// nearest image, minus the half extent.  Returns the ballot; the tests of 32 rows cost what one would.
__device__ __forceinline__ unsigned rows_in_reach(const FastShared *S, const uint4 *ch, int c, int nj)
{
    const int j = c + (int)(threadIdx.x & 31u);
    const uint4 pj = S->sj[min(j, nj - 1)];
    const uint4 ctr = ch[0], hf = ch[1];   // {centre, half extent}
    auto gap = [](uint32_t u, uint32_t centre, uint32_t half) -> float {
        const unsigned ad = (unsigned)abs((int)(u - centre));
        return (float)max((int)(ad - half), 0);
    };
    const float gx = gap(pj.x, ctr.x, hf.x) * S->cull[0];
    const float gy = gap(pj.y, ctr.y, hf.y) * S->cull[1];
    const float gz = gap(pj.z, ctr.z, hf.z) * S->cull[2];
    const bool in = j < nj && gap_norm2(gx, gy, gz, S->cull_tri != 0) <= S->cull[3];
    return __ballot_sync(0xffffffffu, in);
}

// sorted frames, relative-coordinate rows: every lane takes its row of the chunk relative to the centre of the group's
// box (ch = {centre, half extent}) with the integer wrap, leaves the three differences as FP32 in the warp's scratch
// and reports whether the row lies at the periodic SEAM of the box: if |u_j - centre| + half extent reaches 2^31 on
// some axis, the particles of the box do not all see the same image of j and the differences need the period back
// (rel_wrap).  For every other row  (u_j - centre) - (u_i - centre)  IS the wrapped difference u_j - u_i, as plain
// integers.
__device__ __forceinline__ unsigned rows_rel(FastShared *S, const uint4 *ch, int c, int nj)
{
    const int lane = (int)(threadIdx.x & 31u);
    const uint4 pj = S->sj[min(c + lane, nj - 1)];
    const uint4 ctr = ch[0], hf = ch[1];
    const int dx = (int)(pj.x - ctr.x), dy = (int)(pj.y - ctr.y), dz = (int)(pj.z - ctr.z);
    // half extents are below 2^29 here (rel_ok), so the sums cannot overflow
    const bool seam = (unsigned)abs(dx) + hf.x > 0x7fffffffu || (unsigned)abs(dy) + hf.y > 0x7fffffffu ||
                      (unsigned)abs(dz) + hf.z > 0x7fffffffu;
    __syncwarp();   // the previous chunk's rows have been read by every lane
    S->rel[threadIdx.x >> 5][lane] = make_float4((float)dx, (float)dy, (float)dz, __uint_as_float(pj.w));   // + exclusion block of j
    const unsigned m = __ballot_sync(0xffffffffu, seam);   // also orders the scratch writes before the reads below
    __syncwarp();
    return m;
}

// The filter's FP32 bin coordinate of one pair from its three coordinate differences (2^-32 box fractions; for a
// triclinic cell: differences of the fractions along a, b, c): v = bin coordinate - 0.5 (so that rint(v) is the bin).
template <int KIND>
__device__ __forceinline__ float filter_v_tail(float tx, float ty, float tz, const FrameFast &ff)
{
    float d;
    if (KIND == FK_CUBIC) {
        const float t2 = __fmaf_rn(tz, tz, __fmaf_rn(ty, ty, __fmul_rn(tx, tx)));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(t2));
        return __fmaf_rn(d, ff.ax, ff.c0);
    }
    float ex, ey, ez;
    if (KIND == FK_TRI) {
        ex = __fmaf_rn(tz, ff.cx, __fmaf_rn(ty, ff.bx, __fmul_rn(tx, ff.ax)));
        ey = __fmaf_rn(tz, ff.cy, __fmul_rn(ty, ff.ay));
        ez = __fmul_rn(tz, ff.az);
    } else {
        ex = __fmul_rn(tx, ff.ax);
        ey = __fmul_rn(ty, ff.ay);
        ez = __fmul_rn(tz, ff.az);
    }
    const float t2 = __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex)));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(t2));
    return __fadd_rn(d, ff.c0);
}

template <int KIND>
__device__ __forceinline__ float filter_v(uint32_t ujx, uint32_t ujy, uint32_t ujz, uint32_t uix, uint32_t uiy,
                                          uint32_t uiz, const FrameFast &ff)
{
    return filter_v_tail<KIND>((float)(int)(ujx - uix), (float)(int)(ujy - uiy), (float)(int)(ujz - uiz), ff);
}

// Relative-coordinate form of the filter (sorted frames, rows clear of the periodic seam of the group's box, see
// rows_rel): both particles are taken relative to the centre of the group's box with the integer wrap, converted to
// FP32 ONCE, and the pair's difference is a floating-point subtraction -- packed, two pairs per instruction, in the
// hot loop (filter_v2_rel); this is the same arithmetic in scalar form, operation for operation, for the resolve pass.
// Rows AT the seam get the period back in floating point (rel_wrap: three more packed operations per axis, exact);
// applied to a row clear of the seam it changes nothing, so the resolve pass need not know which form a pair took.
// difference of two FP32 coordinates relative to the box centre, brought back to the nearest image: at most one period
// (2^32) off, and then an exact correction; rows clear of the seam are left as they are (q = 0)
__device__ __forceinline__ float rel_wrap(float e)
{
    const float q = __fadd_rn(__fmaf_rn(e, 2.3283064365386963e-10f, 12582912.f), -12582912.f);   // rint(e / 2^32)
    return __fmaf_rn(q, -4294967296.f, e);
}

template <int KIND>
__device__ __forceinline__ float filter_v_rel(uint32_t ujx, uint32_t ujy, uint32_t ujz, uint32_t uix, uint32_t uiy,
                                              uint32_t uiz, const uint4 &ctr, const FrameFast &ff)
{
    const float tx = rel_wrap(__fadd_rn((float)(int)(ujx - ctr.x), -(float)(int)(uix - ctr.x)));
    const float ty = rel_wrap(__fadd_rn((float)(int)(ujy - ctr.y), -(float)(int)(uiy - ctr.y)));
    const float tz = rel_wrap(__fadd_rn((float)(int)(ujz - ctr.z), -(float)(int)(uiz - ctr.z)));
    return filter_v_tail<KIND>(tx, ty, tz, ff);
}

struct TileCtx {            // everything the cold passes need (kept in shared memory, one per block)
    RdfArgs A;
    FrameFast ff;
    FrameBox fb;
    FrameTri ft;
    int64_t i0, j0;
    int f, diag;
    float thr;
};

// Undecided pairs are PARKED, not handled where they are found: a flagged lane only appends one word to the block's
// queue — {row j (10 bits), lane (5), group (3), mask of its four particles (4)} — and the tile's resolve pass, in
// which every lane works on an entry, does the rest: take the filter's increment back (extended histograms), evaluate
// the reference's FP64 distance, count.  (In round 1 the flagged lane did the take-back and one queue slot per pair
// on the spot: one active lane of 32 for ~60 instructions in 13 % of all rows, a tenth of the kernel's instructions.)

// tri_r2 behind a call: inlined into resolve_entry, its 27 images raise the register needs of every function on the call
// chain, and the hot loops (which call park_row in their cold branch) spill around the call
__device__ __noinline__ double tri_r2_call(float xi, float yi, float zi, float xj, float yj, float zj, const FrameTri *t)
{
    return tri_r2(xi, yi, zi, xj, yj, zj, *t);
}

// cold: the reference's exact FP64 result for a pair the filter could not decide
__device__ __forceinline__ void resolve_pair(const TileCtx &T, int il, int jl, const double *sT, unsigned *hist)
{
    const int64_t ig = T.i0 + il, jg = T.j0 + jl;
    const float *ax = T.A.a + ((int64_t)T.f * 3) * T.A.na, *bx = T.A.b + ((int64_t)T.f * 3) * T.A.nb;
    const double r2 = T.ft.on ? tri_r2_call(ax[ig], ax[T.A.na + ig], ax[2 * T.A.na + ig], bx[jg], bx[T.A.nb + jg],
                                             bx[2 * T.A.nb + jg], &T.ft)
                              : exact_r2(ax[ig], ax[T.A.na + ig], ax[2 * T.A.na + ig], bx[jg], bx[T.A.nb + jg],
                                         bx[2 * T.A.nb + jg], T.fb);
    const int k = exact_bin(r2, T.A, sT);
    if (k >= 0) atomicAdd(&hist[k * T.A.hist_ks], 1u);
}

// cold: one queue entry.  copies = bin 0 of histogram copy 0 (layout: RdfArgs::hist_ks / hist_cs); hist = bin 0 of the
// calling thread's own copy (any copy may receive the exact count; the take-back must hit the copy the flagged lane used).
__device__ __noinline__ void resolve_entry(const FastShared *S, unsigned entry, bool uncount, const double *sT, unsigned *copies,
                                           unsigned *hist)
{
    const TileCtx &T = *reinterpret_cast<const TileCtx *>(S->ctx);
    const int jl = (int)(entry & 1023u), lane = (int)((entry >> 10) & 31u), group = (int)((entry >> 15) & 7u);
    const unsigned mask = (entry >> 18) & 15u;
    const bool rel = (entry >> 22) & 1u;   // the filter's bin came from the relative-coordinate form
    for (int u = 0; u < RF_IPT; ++u) {
        if (!((mask >> u) & 1u)) continue;
        const int il = group * (32 * RF_IPT) + u * 32 + lane;
        if (uncount) {
            // the filter's bin of this pair, recomputed with the filter's own arithmetic (bit-identical: same operations
            // in the same order as filter_v2 in the hot loop)
            const int64_t ig = min(T.i0 + il, T.A.na - 1);
            const uint32_t *ua = T.A.ua + ((int64_t)T.f * 3) * T.A.na;
            const uint4 pj = S->sj[jl];
            float v;
            const uint32_t uix = ua[ig], uiy = ua[T.A.na + ig], uiz = ua[2 * T.A.na + ig];
            if (rel) {
                const uint4 ctr = S->ibox[group][2];
                v = T.ff.kind == FK_CUBIC  ? filter_v_rel<FK_CUBIC>(pj.x, pj.y, pj.z, uix, uiy, uiz, ctr, T.ff)
                    : T.ff.kind == FK_TRI ? filter_v_rel<FK_TRI>(pj.x, pj.y, pj.z, uix, uiy, uiz, ctr, T.ff)
                                          : filter_v_rel<FK_ORTHO>(pj.x, pj.y, pj.z, uix, uiy, uiz, ctr, T.ff);
            } else {
                v = T.ff.kind == FK_CUBIC  ? filter_v<FK_CUBIC>(pj.x, pj.y, pj.z, uix, uiy, uiz, T.ff)
                    : T.ff.kind == FK_TRI ? filter_v<FK_TRI>(pj.x, pj.y, pj.z, uix, uiy, uiz, T.ff)
                                          : filter_v<FK_ORTHO>(pj.x, pj.y, pj.z, uix, uiy, uiz, T.ff);
            }
            const int k = (int)(__float_as_uint(__fadd_rn(v, 12582912.f)) - 0x4B400000u);
            // rint(v) beyond B (or below -1): |v_ref - v| < 0.5 puts the pair outside the window whatever the rounding;
            // it sits in a word that is never read back and needs no FP64 pass (40 % of the flagged pairs at L/2)
            if (k > T.A.n_bins || k < -1) continue;
            atomicSub(&copies[(lane & (T.A.n_copies - 1)) * T.A.hist_cs + k * T.A.hist_ks], 1u);
        }
        resolve_pair(T, il, jl, sT, hist);
    }
}

// cold: this lane's pairs (row jl, particles with a bit in `mask`) are undecided: one word into the queue
// Queue full (lattices whose spacings sit on bin edges, margins of a sizeable fraction of a bin): the row is noted in the
// block's overflow table in global memory instead -- one byte per (row, group, lane) of the tile pair, written only by
// the thread that owns it -- and the resolve pass scans that table when the flag is up.  (The first version resolved
// such rows on the spot; the FP64 code behind that call raised the register needs of park_row's whole call tree, and
// every hot loop, which calls park_row in its cold branch, spilled its particles around the call.)
__device__ __forceinline__ void park_row(FastShared *S, int jl, unsigned mask, int group, bool, unsigned rel = 0u)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned slot = atomicAdd(&S->qcount, 1u);
    if (slot < S->qcap) {
        S->queue[slot] = (unsigned)jl | (lane << 10) | ((unsigned)group << 15) | (mask << 18) | (rel << 22);
        return;
    }
    unsigned char *t = S->ovf + ((unsigned)jl * (RF_THREADS / 32) + (unsigned)group) * 32u + lane;
    *t = (unsigned char)(*t | mask | (rel << 4));
    S->overflow = 1u;
}

// The hot loop after v is known: a pair is binned if v is clear of every bin edge, otherwise reported undecided.
// Written in PTX to keep it branch free (bin_or_flag4 / bin_count4 below):
//   w = v + MAGIC;  k = bits(w) - bits(MAGIC) = rint(v);  undecided = |v - (w - MAGIC)| > thr

#ifndef MDC_RF_UNROLL2
#define MDC_RF_UNROLL2 1   // pairs of rows per loop iteration (relative-coordinate form)
#endif
#ifndef MDC_RF_UNROLL
#define MDC_RF_UNROLL 2   // rows of j per loop iteration
#endif

__device__ __forceinline__ unsigned long long pk2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

// filter_v for two particles (u, u + 1) against the same j: identical operations in identical order, issued as
// packed FP32 (mul / fma .f32x2), so the values are bit-identical to the scalar form
// the frame's constants (FrameFast), each twice in a 64-bit register pair (packed FP32 operands): separate values, not a
// struct -- a struct captured by the row lambdas ends up in local memory
#define RF_K_PARAMS unsigned long long kax, unsigned long long kay, unsigned long long kaz, unsigned long long kc0, \
                    unsigned long long kbx, unsigned long long kcx, unsigned long long kcy
#define RF_K_ARGS kax, kay, kaz, kc0, kbx, kcx, kcy

// the filter's v for two pairs from their packed coordinate differences (in 2^-32 box fractions): filter_v_tail, packed
template <int KIND>
__device__ __forceinline__ unsigned long long filter_v2_tail(unsigned long long tx, unsigned long long ty, unsigned long long tz,
                                                             RF_K_PARAMS)
{
    unsigned long long t2, v;
    float a, b, da, db;
    if (KIND == FK_TRI) {
        asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(tx) : "l"(kax));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(tx) : "l"(ty), "l"(kbx));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(tx) : "l"(tz), "l"(kcx));
        asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(ty) : "l"(kay));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ty) : "l"(tz), "l"(kcy));
        asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(tz) : "l"(kaz));
    } else if (KIND == FK_ORTHO) {
        asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(tx) : "l"(kax));
        asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(ty) : "l"(kay));
        asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(tz) : "l"(kaz));
    }
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(t2) : "l"(tx));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(t2) : "l"(ty));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(t2) : "l"(tz));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(t2));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(da) : "f"(a));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(db) : "f"(b));
    const unsigned long long d = pk2(da, db);
    if (KIND == FK_CUBIC) {
        v = kc0;
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(v) : "l"(d), "l"(kax));
    } else {
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(d), "l"(kc0));
    }
    return v;
}

template <int KIND>
__device__ __forceinline__ unsigned long long filter_v2(uint32_t ujx, uint32_t ujy, uint32_t ujz, uint32_t ax0,
                                                        uint32_t ay0, uint32_t az0, uint32_t ax1, uint32_t ay1,
                                                        uint32_t az1, RF_K_PARAMS)
{
    const unsigned long long tx = pk2((float)(int)(ujx - ax0), (float)(int)(ujx - ax1));
    const unsigned long long ty = pk2((float)(int)(ujy - ay0), (float)(int)(ujy - ay1));
    const unsigned long long tz = pk2((float)(int)(ujz - az0), (float)(int)(ujz - az1));
    return filter_v2_tail<KIND>(tx, ty, tz, RF_K_ARGS);
}

// filter_v_rel for two particles: r = the row's FP32 coordinates relative to the centre of the group's box, n = MINUS the
// two particles' (packed): three packed additions replace six integer subtractions and six conversions.  WRAP: rows at
// the seam of the box (rel_wrap, packed).
template <int KIND, bool WRAP>
__device__ __forceinline__ unsigned long long filter_v2_rel(float rx, float ry, float rz, unsigned long long nx,
                                                            unsigned long long ny, unsigned long long nz, RF_K_PARAMS)
{
    unsigned long long tx, ty, tz;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(tx) : "l"(pk2(rx, rx)), "l"(nx));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ty) : "l"(pk2(ry, ry)), "l"(ny));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(tz) : "l"(pk2(rz, rz)), "l"(nz));
    if (WRAP) {
        const unsigned long long inv = pk2(2.3283064365386963e-10f, 2.3283064365386963e-10f);
        const unsigned long long mg = pk2(12582912.f, 12582912.f), nmg = pk2(-12582912.f, -12582912.f);
        const unsigned long long per = pk2(-4294967296.f, -4294967296.f);
        unsigned long long q;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(tx), "l"(inv), "l"(mg));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(q) : "l"(nmg));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(tx) : "l"(q), "l"(per));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(ty), "l"(inv), "l"(mg));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(q) : "l"(nmg));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ty) : "l"(q), "l"(per));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(tz), "l"(inv), "l"(mg));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(q) : "l"(nmg));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(tz) : "l"(q), "l"(per));
    }
    return filter_v2_tail<KIND>(tx, ty, tz, RF_K_ARGS);
}

// bin_or_flag for the four pairs of one row (two packed v's): the round-to-bin chain as six packed instructions,
// then per pair  k = bits(w), undecided = |df| > thr, hist[(ok && !undecided && bin < B) ? bin : dummy] += 1.
// With r0 == 0 (R0ZERO) v >= -0.5, so bits(w) >= bits(MAGIC) and neither the subtraction of bits(MAGIC) nor a
// lower bound check is needed: the caller folds bits(MAGIC) into the limit and into the histogram base address.
// Returns whether ANY pair is undecided (predicate OR, one SEL) and the four |df| so that the cold path can tell
// which; that replaces four SELs and the adds that combined them.
// GATE (sorted frames, i.e. cutoffs far below the box size): a pair whose k lies beyond the last bin + 1 is out of range
// whatever the rounding, so it is never reported as undecided -- with narrow bins the margin is a sizeable
// fraction of a bin and nearly every row would otherwise hold an "undecided" pair that is nowhere near the window.
template <bool HAS_OK, bool R0ZERO, bool GATE>
__device__ __forceinline__ unsigned bin_or_flag4(unsigned long long v01, unsigned long long v23, unsigned long long magic,
                                                 unsigned long long nmagic, unsigned long long neg1, float thr,
                                                 unsigned blim, unsigned hbase, unsigned trash_addr, unsigned one,
                                                 const unsigned (&ok)[4], float (&adf)[4])
{
    unsigned any;
#define RF_HEAD                                                   \
    "{\n\t"                                                       \
    ".reg .pred p, q, q0, q1, q2, q3, r, g;\n\t"                  \
    ".reg .b64 W01, W23, KF, DF01, DF23;\n\t"                     \
    ".reg .f32 w0, w1, w2, w3;\n\t"                               \
    ".reg .u32 k, kk, ad;\n\t"                                    \
    "add.rn.f32x2 W01, %5, %12;\n\t"                              \
    "add.rn.f32x2 W23, %6, %12;\n\t"                              \
    "add.rn.f32x2 KF, W01, %13;\n\t"                              \
    "fma.rn.f32x2 DF01, KF, %14, %5;\n\t"                         \
    "add.rn.f32x2 KF, W23, %13;\n\t"                              \
    "fma.rn.f32x2 DF23, KF, %14, %6;\n\t"                         \
    "mov.b64 {w0, w1}, W01;\n\t"                                  \
    "mov.b64 {w2, w3}, W23;\n\t"                                  \
    "mov.b64 {%1, %2}, DF01;\n\t"                                 \
    "mov.b64 {%3, %4}, DF23;\n\t"                                 \
    "abs.f32 %1, %1;\n\t"                                         \
    "abs.f32 %2, %2;\n\t"                                         \
    "abs.f32 %3, %3;\n\t"                                         \
    "abs.f32 %4, %4;\n\t"
#define RF_FOOT                                                   \
    "or.pred q, q0, q1;\n\t"                                      \
    "or.pred q, q, q2;\n\t"                                       \
    "or.pred q, q, q3;\n\t"                                       \
    "selp.u32 %0, 1, 0, q;\n\t"                                   \
    "}"
#define RF_SUB "sub.u32 k, k, 0x4B400000;\n\t"
// without the ok flags: p = (k < limit) && !q_u
#define RF_P4(SUBK)                                                                                              \
    "mov.b32 k, w0;\n\t" SUBK "setp.gt.f32 q0, %1, %7;\n\t setp.lt.and.u32 p, k, %8, !q0;\n\t"                  \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"                \
    "mov.b32 k, w1;\n\t" SUBK "setp.gt.f32 q1, %2, %7;\n\t setp.lt.and.u32 p, k, %8, !q1;\n\t"                  \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"                \
    "mov.b32 k, w2;\n\t" SUBK "setp.gt.f32 q2, %3, %7;\n\t setp.lt.and.u32 p, k, %8, !q2;\n\t"                  \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"                \
    "mov.b32 k, w3;\n\t" SUBK "setp.gt.f32 q3, %4, %7;\n\t setp.lt.and.u32 p, k, %8, !q3;\n\t"                  \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"
// with them: q_u = undecided && ok_u, p = (k < limit) && ok_u && !q_u
#define RF_P4_OK(SUBK)                                                                                           \
    "mov.b32 k, w0;\n\t" SUBK "setp.ne.u32 r, %15, 0;\n\t setp.gt.and.f32 q0, %1, %7, r;\n\t"                   \
    "setp.lt.and.u32 p, k, %8, r;\n\t and.pred p, p, !q0;\n\t"                                                  \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"                \
    "mov.b32 k, w1;\n\t" SUBK "setp.ne.u32 r, %16, 0;\n\t setp.gt.and.f32 q1, %2, %7, r;\n\t"                   \
    "setp.lt.and.u32 p, k, %8, r;\n\t and.pred p, p, !q1;\n\t"                                                  \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"                \
    "mov.b32 k, w2;\n\t" SUBK "setp.ne.u32 r, %17, 0;\n\t setp.gt.and.f32 q2, %3, %7, r;\n\t"                   \
    "setp.lt.and.u32 p, k, %8, r;\n\t and.pred p, p, !q2;\n\t"                                                  \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"                \
    "mov.b32 k, w3;\n\t" SUBK "setp.ne.u32 r, %18, 0;\n\t setp.gt.and.f32 q3, %4, %7, r;\n\t"                   \
    "setp.lt.and.u32 p, k, %8, r;\n\t and.pred p, p, !q3;\n\t"                                                  \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"
// gated forms: undecided only if -1 <= bin <= B.  GATEK sets predicate g from k (see RF_GATE_*)
#define RF_GATE_R0ZERO "setp.le.u32 g, k, %8;\n\t"                              /* bits(w) <= B + bits(MAGIC) */
#define RF_GATE_ANY(LIM1) "add.u32 kk, k, 1;\n\t setp.le.u32 g, kk, " LIM1 ";\n\t" /* k + 1 <= B + 1, unsigned */
#define RF_G4(SUBK, GATEK) \
    "mov.b32 k, w0;\n\t" SUBK GATEK "setp.gt.and.f32 q0, %1, %7, g;\n\t setp.lt.and.u32 p, k, %8, !q0;\n\t" \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t" \
    "mov.b32 k, w1;\n\t" SUBK GATEK "setp.gt.and.f32 q1, %2, %7, g;\n\t setp.lt.and.u32 p, k, %8, !q1;\n\t" \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t" \
    "mov.b32 k, w2;\n\t" SUBK GATEK "setp.gt.and.f32 q2, %3, %7, g;\n\t setp.lt.and.u32 p, k, %8, !q2;\n\t" \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t" \
    "mov.b32 k, w3;\n\t" SUBK GATEK "setp.gt.and.f32 q3, %4, %7, g;\n\t setp.lt.and.u32 p, k, %8, !q3;\n\t" \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"
#define RF_G4_OK(SUBK, GATEK) \
    "mov.b32 k, w0;\n\t" SUBK "setp.ne.u32 r, %15, 0;\n\t" GATEK "and.pred g, g, r;\n\t setp.gt.and.f32 q0, %1, %7, g;\n\t" \
    "setp.lt.and.u32 p, k, %8, r;\n\t and.pred p, p, !q0;\n\t" \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t" \
    "mov.b32 k, w1;\n\t" SUBK "setp.ne.u32 r, %16, 0;\n\t" GATEK "and.pred g, g, r;\n\t setp.gt.and.f32 q1, %2, %7, g;\n\t" \
    "setp.lt.and.u32 p, k, %8, r;\n\t and.pred p, p, !q1;\n\t" \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t" \
    "mov.b32 k, w2;\n\t" SUBK "setp.ne.u32 r, %17, 0;\n\t" GATEK "and.pred g, g, r;\n\t setp.gt.and.f32 q2, %3, %7, g;\n\t" \
    "setp.lt.and.u32 p, k, %8, r;\n\t and.pred p, p, !q2;\n\t" \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t" \
    "mov.b32 k, w3;\n\t" SUBK "setp.ne.u32 r, %18, 0;\n\t" GATEK "and.pred g, g, r;\n\t setp.gt.and.f32 q3, %4, %7, g;\n\t" \
    "setp.lt.and.u32 p, k, %8, r;\n\t and.pred p, p, !q3;\n\t" \
    "mad.lo.u32 ad, k, 4, %9;\n\t selp.u32 ad, ad, %10, p;\n\t red.shared.add.u32 [ad], %11;\n\t"
#define RF_OUT "=r"(any), "=f"(adf[0]), "=f"(adf[1]), "=f"(adf[2]), "=f"(adf[3])
#define RF_IN "l"(v01), "l"(v23), "f"(thr), "r"(blim), "r"(hbase), "r"(trash_addr), "r"(one), "l"(magic), "l"(nmagic), "l"(neg1)
    if (GATE) {
        const unsigned lim1 = blim + 1u;
        if (HAS_OK) {
            if (R0ZERO)
                asm volatile(RF_HEAD RF_G4_OK("", RF_GATE_R0ZERO) RF_FOOT : RF_OUT : RF_IN, "r"(ok[0]), "r"(ok[1]), "r"(ok[2]), "r"(ok[3]) : "memory");
            else
                asm volatile(RF_HEAD RF_G4_OK(RF_SUB, RF_GATE_ANY("%19")) RF_FOOT : RF_OUT : RF_IN, "r"(ok[0]), "r"(ok[1]), "r"(ok[2]), "r"(ok[3]), "r"(lim1) : "memory");
        } else {
            if (R0ZERO) asm volatile(RF_HEAD RF_G4("", RF_GATE_R0ZERO) RF_FOOT : RF_OUT : RF_IN : "memory");
            else asm volatile(RF_HEAD RF_G4(RF_SUB, RF_GATE_ANY("%15")) RF_FOOT : RF_OUT : RF_IN, "r"(lim1) : "memory");
        }
    } else if (HAS_OK) {
        if (R0ZERO)
            asm volatile(RF_HEAD RF_P4_OK("") RF_FOOT : RF_OUT : RF_IN, "r"(ok[0]), "r"(ok[1]), "r"(ok[2]), "r"(ok[3]) : "memory");
        else
            asm volatile(RF_HEAD RF_P4_OK(RF_SUB) RF_FOOT : RF_OUT : RF_IN, "r"(ok[0]), "r"(ok[1]), "r"(ok[2]), "r"(ok[3]) : "memory");
    } else {
        if (R0ZERO) asm volatile(RF_HEAD RF_P4("") RF_FOOT : RF_OUT : RF_IN : "memory");
        else asm volatile(RF_HEAD RF_P4(RF_SUB) RF_FOOT : RF_OUT : RF_IN : "memory");
    }
#undef RF_HEAD
#undef RF_FOOT
#undef RF_SUB
#undef RF_P4
#undef RF_P4_OK
#undef RF_G4
#undef RF_G4_OK
#undef RF_GATE_R0ZERO
#undef RF_GATE_ANY
#undef RF_OUT
#undef RF_IN
    return any;
}

// The four pairs of one row with extended histograms (RdfArgs::ext_k): every k the filter can produce has a word, so
// the increment needs no range test and no dummy word: hist[k] += 1 (IMAD + ATOMS per pair) and ONE predicate that
// says whether any pair is undecided (FSETP chain).  The cold path takes the increment of an undecided pair back.
// hbase = shared address of bin 0 of this thread's copy - 4 * bits(MAGIC): bits(w) = bits(MAGIC) + k for any sign of k.
// LEA: the word address as a funnel shift + add (LEA.HI, ALU pipe) instead of a multiply-add (IMAD, the FMA pipe that
// the packed arithmetic of the relative-coordinate rows already fills); the integer rows keep IMAD (their ALU pipe
// carries the conversions).  SH: log2 of the BYTES between consecutive bins of one copy: 2 (copies one after the other)
// or 7 (32 interleaved copies, one per lane = per bank: RdfArgs::hist_ks).
template <bool HAS_OK, bool LEA, int SH>
__device__ __forceinline__ unsigned bin_count4(unsigned long long v01, unsigned long long v23, unsigned long long magic,
                                               unsigned long long nmagic, unsigned long long neg1, float thr,
                                               unsigned hbase, unsigned trash_addr, unsigned one, const unsigned (&ok)[4],
                                               float (&adf)[4])
{
    unsigned any;
#define RX_HEAD                                                   \
    "{\n\t"                                                       \
    ".reg .pred q, r;\n\t"                                        \
    ".reg .b64 W01, W23, KF, DF01, DF23;\n\t"                     \
    ".reg .f32 w0, w1, w2, w3;\n\t"                               \
    ".reg .u32 k, ad;\n\t"                                        \
    "add.rn.f32x2 W01, %5, %11;\n\t"                              \
    "add.rn.f32x2 W23, %6, %11;\n\t"                              \
    "add.rn.f32x2 KF, W01, %12;\n\t"                              \
    "fma.rn.f32x2 DF01, KF, %13, %5;\n\t"                         \
    "add.rn.f32x2 KF, W23, %12;\n\t"                              \
    "fma.rn.f32x2 DF23, KF, %13, %6;\n\t"                         \
    "mov.b64 {w0, w1}, W01;\n\t"                                  \
    "mov.b64 {w2, w3}, W23;\n\t"                                  \
    "mov.b64 {%1, %2}, DF01;\n\t"                                 \
    "mov.b64 {%3, %4}, DF23;\n\t"                                 \
    "abs.f32 %1, %1;\n\t"                                         \
    "abs.f32 %2, %2;\n\t"                                         \
    "abs.f32 %3, %3;\n\t"                                         \
    "abs.f32 %4, %4;\n\t"
#define RX_OUT "=r"(any), "=f"(adf[0]), "=f"(adf[1]), "=f"(adf[2]), "=f"(adf[3])
#define RX_IN "l"(v01), "l"(v23), "f"(thr), "r"(hbase), "r"(trash_addr), "r"(one), "l"(magic), "l"(nmagic), "l"(neg1), "n"(1 << SH), "n"(SH)
#define RX_MAD "mad.lo.u32 ad, k, %14, %8;\n\t"
#define RX_LEA "shf.l.wrap.b32 ad, 0, k, %15;\n\t add.u32 ad, ad, %8;\n\t"
    if (HAS_OK) {
        asm volatile(RX_HEAD
            "mov.b32 k, w0;\n\t setp.ne.u32 r, %16, 0;\n\t" RX_MAD "selp.u32 ad, ad, %9, r;\n\t"
            "red.shared.add.u32 [ad], %10;\n\t setp.gt.and.f32 q, %1, %7, r;\n\t"
            "mov.b32 k, w1;\n\t setp.ne.u32 r, %17, 0;\n\t" RX_MAD "selp.u32 ad, ad, %9, r;\n\t"
            "red.shared.add.u32 [ad], %10;\n\t setp.gt.and.f32 r, %2, %7, r;\n\t or.pred q, q, r;\n\t"
            "mov.b32 k, w2;\n\t setp.ne.u32 r, %18, 0;\n\t" RX_MAD "selp.u32 ad, ad, %9, r;\n\t"
            "red.shared.add.u32 [ad], %10;\n\t setp.gt.and.f32 r, %3, %7, r;\n\t or.pred q, q, r;\n\t"
            "mov.b32 k, w3;\n\t setp.ne.u32 r, %19, 0;\n\t" RX_MAD "selp.u32 ad, ad, %9, r;\n\t"
            "red.shared.add.u32 [ad], %10;\n\t setp.gt.and.f32 r, %4, %7, r;\n\t or.pred q, q, r;\n\t"
            "selp.u32 %0, 1, 0, q;\n\t}"
            : RX_OUT : RX_IN, "r"(ok[0]), "r"(ok[1]), "r"(ok[2]), "r"(ok[3]) : "memory");
    } else if (LEA) {
        asm volatile(RX_HEAD
            "mov.b32 k, w0;\n\t" RX_LEA "red.shared.add.u32 [ad], 1;\n\t"
            "mov.b32 k, w1;\n\t" RX_LEA "red.shared.add.u32 [ad], 1;\n\t"
            "mov.b32 k, w2;\n\t" RX_LEA "red.shared.add.u32 [ad], 1;\n\t"
            "mov.b32 k, w3;\n\t" RX_LEA "red.shared.add.u32 [ad], 1;\n\t"
            "setp.gt.f32 q, %1, %7;\n\t setp.gt.or.f32 q, %2, %7, q;\n\t setp.gt.or.f32 q, %3, %7, q;\n\t"
            "setp.gt.or.f32 q, %4, %7, q;\n\t selp.u32 %0, 1, 0, q;\n\t}"
            : RX_OUT : RX_IN : "memory");
    } else {
        asm volatile(RX_HEAD
            "mov.b32 k, w0;\n\t" RX_MAD "red.shared.add.u32 [ad], 1;\n\t"
            "mov.b32 k, w1;\n\t" RX_MAD "red.shared.add.u32 [ad], 1;\n\t"
            "mov.b32 k, w2;\n\t" RX_MAD "red.shared.add.u32 [ad], 1;\n\t"
            "mov.b32 k, w3;\n\t" RX_MAD "red.shared.add.u32 [ad], 1;\n\t"
            "setp.gt.f32 q, %1, %7;\n\t setp.gt.or.f32 q, %2, %7, q;\n\t setp.gt.or.f32 q, %3, %7, q;\n\t"
            "setp.gt.or.f32 q, %4, %7, q;\n\t selp.u32 %0, 1, 0, q;\n\t}"
            : RX_OUT : RX_IN : "memory");
    }
#undef RX_HEAD
#undef RX_OUT
#undef RX_IN
#undef RX_MAD
#undef RX_LEA
    return any;
}

// bin_count4 for half a row: the two pairs of one packed v (extended histograms, no ok flags)
template <bool LEA, int SH>
__device__ __forceinline__ unsigned bin_count2(unsigned long long v, unsigned long long magic, unsigned long long nmagic,
                                               unsigned long long neg1, float thr, unsigned hbase, float (&adf)[2])
{
    unsigned any;
#define RY_HEAD                                                   \
    "{\n\t"                                                       \
    ".reg .pred q;\n\t"                                           \
    ".reg .b64 W, KF, DF;\n\t"                                    \
    ".reg .f32 w0, w1;\n\t"                                       \
    ".reg .u32 k, ad;\n\t"                                        \
    "add.rn.f32x2 W, %3, %6;\n\t"                                 \
    "add.rn.f32x2 KF, W, %7;\n\t"                                 \
    "fma.rn.f32x2 DF, KF, %8, %3;\n\t"                            \
    "mov.b64 {w0, w1}, W;\n\t"                                    \
    "mov.b64 {%1, %2}, DF;\n\t"                                   \
    "abs.f32 %1, %1;\n\t"                                         \
    "abs.f32 %2, %2;\n\t"
#define RY_FOOT                                                   \
    "setp.gt.f32 q, %1, %4;\n\t setp.gt.or.f32 q, %2, %4, q;\n\t" \
    "selp.u32 %0, 1, 0, q;\n\t}"
#define RY_OPS "=r"(any), "=f"(adf[0]), "=f"(adf[1]) : "l"(v), "f"(thr), "r"(hbase), "l"(magic), "l"(nmagic), "l"(neg1), "n"(1 << SH), "n"(SH) : "memory"
    if (LEA) {
        asm volatile(RY_HEAD
            "mov.b32 k, w0;\n\t shf.l.wrap.b32 ad, 0, k, %10;\n\t add.u32 ad, ad, %5;\n\t red.shared.add.u32 [ad], 1;\n\t"
            "mov.b32 k, w1;\n\t shf.l.wrap.b32 ad, 0, k, %10;\n\t add.u32 ad, ad, %5;\n\t red.shared.add.u32 [ad], 1;\n\t"
            RY_FOOT : RY_OPS);
    } else {
        asm volatile(RY_HEAD
            "mov.b32 k, w0;\n\t mad.lo.u32 ad, k, %9, %5;\n\t red.shared.add.u32 [ad], 1;\n\t"
            "mov.b32 k, w1;\n\t mad.lo.u32 ad, k, %9, %5;\n\t red.shared.add.u32 [ad], 1;\n\t"
            RY_FOOT : RY_OPS);
    }
#undef RY_HEAD
#undef RY_FOOT
#undef RY_OPS
    return any;
}

template <bool CHECK, int KIND, bool EXCL, bool R0ZERO, bool CULL, bool EXT, bool L32>
__device__ __forceinline__ void tile_pairs_fast(const TileCtx *Tp, FastShared *S, int nj, uint32_t (&ux)[RF_IPT],
                                                uint32_t (&uy)[RF_IPT], uint32_t (&uz)[RF_IPT],
                                                int (&bi)[RF_IPT], const FrameFast &ff, float thr, unsigned B,
                                                int64_t na, int64_t i0, int64_t j0, bool diag, const double *sT,
                                                unsigned *hist)
{
    const unsigned hist_addr = (unsigned)__cvta_generic_to_shared(hist);
    const unsigned trash_addr = (unsigned)__cvta_generic_to_shared(&S->trash[threadIdx.x]);
    // an increment the compiler cannot see through (a literal 1 becomes ATOMS.POPC.INC inside a branch)
    const unsigned one = (unsigned)Tp->A.sym | 1u;
    constexpr bool HAS_OK = CHECK || EXCL;
    static_assert(RF_IPT == 4, "the undecided-pair mask is built for four particles per thread");
    const unsigned long long kax = pk2(ff.ax, ff.ax), kay = pk2(ff.ay, ff.ay), kaz = pk2(ff.az, ff.az);
    const unsigned long long kc0 = pk2(ff.c0, ff.c0);
    const unsigned long long kbx = pk2(ff.bx, ff.bx), kcx = pk2(ff.cx, ff.cx), kcy = pk2(ff.cy, ff.cy);
    const unsigned long long magic = pk2(12582912.f, 12582912.f), nmagic = pk2(-12582912.f, -12582912.f);
    const unsigned long long neg1 = pk2(-1.f, -1.f);
    // R0ZERO: bits(w) is used as it is; bits(MAGIC) = 0x4B400000 moves into the limit and the base address
    const unsigned blim = R0ZERO ? B + 0x4B400000u : B;
    // bytes between consecutive bins of this thread's copy: 4, or 128 with the 32 interleaved copies (L32, EXT only)
    constexpr int SH = L32 ? 7 : 2;
    static_assert(!L32 || EXT, "the interleaved copies come with the extended histograms");
    const unsigned hbase = (R0ZERO || EXT) ? hist_addr - (1u << SH) * 0x4B400000u : hist_addr;
    constexpr int RF_UNROLL = MDC_RF_UNROLL;
    constexpr int RF_UNROLL2 = MDC_RF_UNROLL2;
    int group = (int)(threadIdx.x >> 5);   // whose particles ux / uy / uz hold: the warp's own, or a group it helps out
    auto row = [&](int j) {
        const uint4 pj = S->sj[j];
        unsigned ok[RF_IPT];
#pragma unroll
        for (int u = 0; u < RF_IPT; ++u) {
            ok[u] = 1u;
            if (EXCL) ok[u] = bi[u] != (int)pj.w;
            if (CHECK) {
                const int64_t ig = i0 + rf_il_of(group, u);
                ok[u] = ok[u] && (ig < na) && (!diag || (j0 + j) > ig);
            }
        }
        const unsigned long long v01 = filter_v2<KIND>(pj.x, pj.y, pj.z, ux[0], uy[0], uz[0], ux[1], uy[1], uz[1], RF_K_ARGS);
        const unsigned long long v23 = filter_v2<KIND>(pj.x, pj.y, pj.z, ux[2], uy[2], uz[2], ux[3], uy[3], uz[3], RF_K_ARGS);
        float adf[4];
        const unsigned any = EXT ? bin_count4<HAS_OK, false, SH>(v01, v23, magic, nmagic, neg1, thr, hbase, trash_addr, one, ok, adf)
                                 : bin_or_flag4<HAS_OK, R0ZERO, CULL>(v01, v23, magic, nmagic, neg1, thr, blim, hbase,
                                                                      trash_addr, one, ok, adf);
        if (any) {
            unsigned und = 0;   // cold: which of the four (without the gate: a far pair parked here resolves to no bin)
#pragma unroll
            for (int u = 0; u < RF_IPT; ++u)
                if (adf[u] > thr && ok[u]) und |= 1u << u;
            if (und) park_row(S, j, und, group, EXT);
        }
    };
    // half a row: the pairs of particles (2 H, 2 H + 1) only (extended histograms, tiles without ok flags)
    auto half_row = [&](int j, int H) {
        const uint4 pj = S->sj[j];
        const int a = 2 * H, b = 2 * H + 1;
        const unsigned long long v = filter_v2<KIND>(pj.x, pj.y, pj.z, ux[a], uy[a], uz[a], ux[b], uy[b], uz[b], RF_K_ARGS);
        float adf[2];
        if (bin_count2<false, SH>(v, magic, nmagic, neg1, thr, hbase, adf)) {
            unsigned und = 0;
            if (adf[0] > thr) und |= 1u << a;
            if (adf[1] > thr) und |= 1u << b;
            if (und) park_row(S, j, und, group, true);
        }
    };
    // the same in the relative-coordinate form (rows_rel).  The group's particles then live in ux / uy / uz as MINUS their
    // FP32 differences to the centre of the group's box (the same twelve registers: the integer form is not used for
    // such a group), packed (0, 1) and (2, 3); groups whose box is wider than an eighth of the cell keep the integers.
    constexpr bool REL = CULL && EXT;
    bool rel_ok = false;
    // shared address of the warp's scratch row 0, minus 16 c for the chunk in flight: a value the compiler must keep
    // (it otherwise rebuilds the address from the thread index for every row)
    unsigned rel_at = 0;
    auto rel_row = [&](int j) {
        float4 r;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(rel_at + 16u * (unsigned)j));
        return r;
    };
    auto load_rel = [&](const uint4 *ib) {
        if (!REL) return;
        const uint4 ctr = ib[2], hf = ib[3];
        rel_ok = Tp->A.rel && (hf.x | hf.y | hf.z) < (1u << 29);   // warp-uniform
        if (!rel_ok) return;
#pragma unroll
        for (int u = 0; u < RF_IPT; ++u) {
            ux[u] = __float_as_uint(-(float)(int)(ux[u] - ctr.x));
            uy[u] = __float_as_uint(-(float)(int)(uy[u] - ctr.y));
            uz[u] = __float_as_uint(-(float)(int)(uz[u] - ctr.z));
        }
    };
    auto nrel = [&](const uint32_t (&w)[RF_IPT], int H) { return pk2(__uint_as_float(w[2 * H]), __uint_as_float(w[2 * H + 1])); };
    // One row, branch free: the four v's, the increments, whether any pair is undecided (adf: |v - rint(v)| per pair).
    auto vs_rel = [&](const float4 &r, auto wrap, unsigned long long (&v)[2]) {
        constexpr bool W = decltype(wrap)::value;
        v[0] = filter_v2_rel<KIND, W>(r.x, r.y, r.z, nrel(ux, 0), nrel(uy, 0), nrel(uz, 0), RF_K_ARGS);
        v[1] = filter_v2_rel<KIND, W>(r.x, r.y, r.z, nrel(ux, 1), nrel(uy, 1), nrel(uz, 1), RF_K_ARGS);
    };
    auto count_rel = [&](const unsigned long long (&v)[2], float (&adf)[4]) -> unsigned {
        const unsigned ok[RF_IPT] = {1u, 1u, 1u, 1u};
        return bin_count4<false, true, SH>(v[0], v[1], magic, nmagic, neg1, thr, hbase, trash_addr, one, ok, adf);
    };
    auto park_rel = [&](int j, const float (&adf)[4]) {   // cold
        unsigned und = 0;
#pragma unroll
        for (int u = 0; u < RF_IPT; ++u)
            if (adf[u] > thr) und |= 1u << u;
        if (und) park_row(S, j, und, group, true, 1u);
    };
    // one row with ok flags (tiles on the diagonal, ragged tiles, block exclusion): the row's exclusion block rides in
    // the fourth component of its scratch entry
    auto row_rel_ok = [&](int j, auto wrap) {
        const float4 r = rel_row(j);
        unsigned ok[RF_IPT];
#pragma unroll
        for (int u = 0; u < RF_IPT; ++u) {
            ok[u] = 1u;
            if (EXCL) ok[u] = bi[u] != (int)__float_as_uint(r.w);
            if (CHECK) {
                const int64_t ig = i0 + rf_il_of(group, u);
                ok[u] = ok[u] && (ig < na) && (!diag || (j0 + j) > ig);
            }
        }
        unsigned long long v[2];
        vs_rel(r, wrap, v);
        float adf[4];
        if (bin_count4<true, false, SH>(v[0], v[1], magic, nmagic, neg1, thr, hbase, trash_addr, one, ok, adf)) {
            unsigned und = 0;
#pragma unroll
            for (int u = 0; u < RF_IPT; ++u)
                if (adf[u] > thr && ok[u]) und |= 1u << u;
            if (und) park_row(S, j, und, group, true, 1u);
        }
    };
    auto row_rel = [&](int j, int c, auto wrap) {
        float adf[4];
        unsigned long long v[2];
        vs_rel(rel_row(j), wrap, v);
        if (count_rel(v, adf)) park_rel(j, adf);
    };
    // Two rows at a time with ONE branch to the cold path: the branch around the parking code ends a scheduling region,
    // and one row alone is only two dependent chains of packed operations (fixed-latency waits were the top stall).
    // Both rows are loaded and all four v's formed before the first increment (the increments are volatile and keep
    // their order, the arithmetic before them is free to interleave).
    auto rows2_rel = [&](int j0, int j1, auto wrap) {
        float adf0[4], adf1[4];
        unsigned long long v0[2], v1[2];
        const float4 r0 = rel_row(j0), r1 = rel_row(j1);
        vs_rel(r0, wrap, v0);
        vs_rel(r1, wrap, v1);
        const unsigned a0 = count_rel(v0, adf0), a1 = count_rel(v1, adf1);
        if (a0 | a1) {
            if (a0) park_rel(j0, adf0);
            if (a1) park_rel(j1, adf1);
        }
    };
    auto half_row_rel = [&](int j, int c, int H, auto wrap) {
        constexpr bool W = decltype(wrap)::value;
        const float4 r = rel_row(j);
        const unsigned long long v = filter_v2_rel<KIND, W>(r.x, r.y, r.z, nrel(ux, H), nrel(uy, H), nrel(uz, H), RF_K_ARGS);
        float adf[2];
        if (bin_count2<true, SH>(v, magic, nmagic, neg1, thr, hbase, adf)) {
            unsigned und = 0;
            if (adf[0] > thr) und |= 1u << (2 * H);
            if (adf[1] > thr) und |= 1u << (2 * H + 1);
            if (und) park_row(S, j, und, group, true, 1u);
        }
    };
    // the rows of a mask two at a time
    auto walk2_rel = [&](unsigned m, int c, auto wrap) {
        while (m) {
            const int j0 = c + __ffs((int)m) - 1;
            m &= m - 1u;
            if (!m) {
                row_rel(j0, c, wrap);
                break;
            }
            const int j1 = c + __ffs((int)m) - 1;
            m &= m - 1u;
            rows2_rel(j0, j1, wrap);
        }
    };
    using Yes = std::true_type;
    using No = std::false_type;
    auto walk = [&](unsigned m, int c, auto &&body) {
        while (m) {
            const int j = c + __ffs((int)m) - 1;
            m &= m - 1u;
            body(j);
        }
    };
    if (CULL) {
        // sorted frames: a group of 128 particles against chunks of 32 rows; chunks out of reach are skipped.  Chunks are
        // drawn from the group's counter: first the warp's own group, then what is left of the others (work stealing).
        const int own = group;
        for (int turn = 0; turn < RF_THREADS / 32; ++turn) {
            const int g = (own + turn) & (RF_THREADS / 32 - 1);
            if (turn > 0) {
                if (*(volatile int *)&S->next_chunk[g] * 32 >= nj) continue;   // nothing left there (warp-uniform)
                // take over group g: its particles replace this warp's own (not needed again for this tile pair)
                const RdfArgs &A = Tp->A;
                const uint32_t *ax = A.ua + ((int64_t)Tp->f * 3) * A.na, *ay = ax + A.na, *az = ay + A.na;
                const int *oa = A.oa ? A.oa + (int64_t)Tp->f * A.na : nullptr;
                group = g;
#pragma unroll
                for (int u = 0; u < RF_IPT; ++u) {
                    const int64_t ii = min(i0 + rf_il_of(g, u), na - 1);
                    ux[u] = ax[ii];
                    uy[u] = ay[ii];
                    uz[u] = az[ii];
                    bi[u] = A.ex0 > 0 ? (int)((oa ? (int64_t)oa[ii] : ii) / A.ex0) : -1;
                }
            }
            const uint4 *ib = S->ibox[g];
            load_rel(ib);
            for (;;) {
                int c = 0;
                if ((threadIdx.x & 31u) == 0u) c = atomicAdd(&S->next_chunk[g], 1);
                c = __shfl_sync(0xffffffffu, c, 0) * 32;
                if (c >= nj) break;
                if (boxes_apart(ib, S->jbox[c >> 5], S->cull, S->cull_tri != 0)) continue;
                // then row by row (one test per lane): only the rows within reach of the group's box are evaluated
                if (EXT && !HAS_OK) {
                    // ... and half row by half row: the group's first and last 64 particles have boxes of their own
                    const unsigned ma = rows_in_reach(S, ib + 4, c, nj), mb = rows_in_reach(S, ib + 6, c, nj);
                    const unsigned both = ma & mb;
                    if (REL && rel_ok) {
                        // relative-coordinate form: 3 packed additions instead of 6 subtractions + 6 conversions per two
                        // pairs; rows at the seam of the group's box pay 9 more packed operations for the period
                        const unsigned seam = rows_rel(S, ib + 2, c, nj);
                        asm volatile("mov.u32 %0, %1;" : "=r"(rel_at) : "r"((unsigned)__cvta_generic_to_shared(S->rel[threadIdx.x >> 5]) - 16u * (unsigned)c));
                        if (seam == 0u && __popc(both) >= 29) {
                            const int je = min(nj, c + 32);
                            int j = c;
#pragma unroll RF_UNROLL2
                            for (; j + 1 < je; j += 2) rows2_rel(j, j + 1, No());
                            if (j < je) row_rel(j, c, No());
                            continue;
                        }
                        walk2_rel(both & ~seam, c, No());
                        walk(ma & ~mb & ~seam, c, [&](int j) { half_row_rel(j, c, 0, No()); });
                        walk(mb & ~ma & ~seam, c, [&](int j) { half_row_rel(j, c, 1, No()); });
                        if (seam & (ma | mb)) {
                            walk2_rel(both & seam, c, Yes());
                            walk(ma & ~mb & seam, c, [&](int j) { half_row_rel(j, c, 0, Yes()); });
                            walk(mb & ~ma & seam, c, [&](int j) { half_row_rel(j, c, 1, Yes()); });
                        }
                        continue;
                    }
                    if (__popc(both) >= 29) {   // (nearly) all rows in full: the plain loop has less overhead per row
                        const int je = min(nj, c + 32);
#pragma unroll RF_UNROLL
                        for (int j = c; j < je; ++j) row(j);
                        continue;
                    }
                    walk(both, c, [&](int j) { row(j); });
                    walk(ma & ~mb, c, [&](int j) { half_row(j, 0); });
                    walk(mb & ~ma, c, [&](int j) { half_row(j, 1); });
                    continue;
                }
                unsigned m = rows_in_reach(S, ib + 2, c, nj);
                if (REL && rel_ok) {   // diagonal / edge tiles and block exclusion: the relative rows with their ok flags
                    const unsigned seam = rows_rel(S, ib + 2, c, nj);
                    asm volatile("mov.u32 %0, %1;" : "=r"(rel_at) : "r"((unsigned)__cvta_generic_to_shared(S->rel[threadIdx.x >> 5]) - 16u * (unsigned)c));
                    if (seam == 0u && __popc(m) >= 29) {
                        const int je = min(nj, c + 32);
#pragma unroll RF_UNROLL
                        for (int j = c; j < je; ++j) row_rel_ok(j, No());
                        continue;
                    }
                    walk(m & ~seam, c, [&](int j) { row_rel_ok(j, No()); });
                    walk(m & seam, c, [&](int j) { row_rel_ok(j, Yes()); });
                    continue;
                }
                if (__popc(m) >= 29) {   // (nearly) all of them: the plain loop has less overhead per row than the bit walk
                    const int je = min(nj, c + 32);
#pragma unroll RF_UNROLL
                    for (int j = c; j < je; ++j) row(j);
                    continue;
                }
                walk(m, c, [&](int j) { row(j); });
            }
        }
    } else {
#pragma unroll RF_UNROLL
        for (int j = 0; j < nj; ++j) row(j);
    }
}

#ifndef MDC_RF_BLOCKS
#define MDC_RF_BLOCKS 3   // resident blocks per SM the register budget is sized for (80 registers per thread)
#endif
__global__ void __launch_bounds__(RF_THREADS, MDC_RF_BLOCKS)
rdf_pairs_fast(RdfArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastShared *S = reinterpret_cast<FastShared *>(smem_raw);
    double *sT = reinterpret_cast<double *>(smem_raw + ((sizeof(FastShared) + 15) & ~size_t(15)));
    unsigned *sh = reinterpret_cast<unsigned *>(sT + (A.n_bins + 1));
    const int B = A.n_bins;
    for (int i = threadIdx.x; i <= B; i += RF_THREADS) sT[i] = A.T[i];
    const int HS = A.ext_k > 0 ? A.ext_k : B;   // words per histogram copy
    for (int i = threadIdx.x; i < HS * A.n_copies; i += RF_THREADS) sh[i] = 0u;
    if (threadIdx.x == 0) {
        S->qcount = 0u;
        S->qcap = (unsigned)A.qcap;
        S->queue = sh + HS * A.n_copies;
        S->ovf = A.ovf + (size_t)blockIdx.x * (RF_TJ * RF_THREADS);
        S->overflow = 0u;
    }
    // Histogram copies are private to LANE groups, not to warps: bank conflicts and same-address serialisation only
    // arise among the 32 lanes of one ATOMS instruction, and neighbouring lanes (neighbouring particles in a sorted
    // frame, hence similar distances to the same j) are the ones that would collide.
    unsigned *hist = sh + (threadIdx.x & (unsigned)(A.n_copies - 1)) * A.hist_cs + A.ext_off * A.hist_ks;   // bin 0 of this lane's copy
    const bool ext = A.ext_k > 0, l32 = A.hist_ks == 32;
    const unsigned long long weight = A.sym ? 2ull : 1ull;
    const int64_t total = A.items_per_frame * A.F;
    int since_flush = 0;

    // Tile pairs are drawn from a global ticket counter, so that no block idles while another still has a whole
    // tile pair queued (static striding left up to one item of tail: 11 % of the kernel at N = 1e5).
    __shared__ unsigned long long s_ticket;
    for (;;) {
        if (threadIdx.x == 0) s_ticket = atomicAdd(A.ticket, 1ull);
        __syncthreads();
        const int64_t item = (int64_t)s_ticket;
        if (item >= total) break;
        int f, I, J;
        bool diag;
        if (!decode_item(A, item, f, I, J, diag)) {
            __syncthreads();   // every thread has read the ticket before thread 0 draws the next one
            continue;
        }
        const FrameFast ff = A.fast[f];
        const FrameBox fb = A.boxes[f];
        const FrameTri *ftp = A.tri + f;   // read where it is needed: twenty registers otherwise
        if (A.tbox_a) {   // sorted frames: skip tile pairs that lie farther apart than the cutoff (same for all threads)
            const uint32_t *ta = A.tbox_a + ((int64_t)f * A.nti + I) * 6, *tb = A.tbox_b + ((int64_t)f * A.ntj + J) * 6;
            if (tile_gap2(ta, tb, fb, ftp) > A.cull_r * A.cull_r) {
                __syncthreads();
                continue;
            }
        }
        const uint32_t *ax = A.ua + ((int64_t)f * 3) * A.na, *ay = ax + A.na, *az = ay + A.na;
        const uint32_t *bx = A.ub + ((int64_t)f * 3) * A.nb, *by = bx + A.nb, *bz = by + A.nb;
        const int *oa = A.oa ? A.oa + (int64_t)f * A.na : nullptr, *ob = A.ob ? A.ob + (int64_t)f * A.nb : nullptr;
        const int64_t i0 = (int64_t)I * RF_TI, j0 = (int64_t)J * RF_TJ;

        uint32_t ux[RF_IPT], uy[RF_IPT], uz[RF_IPT];
        int bi[RF_IPT];
#pragma unroll
        for (int u = 0; u < RF_IPT; ++u) {
            const int64_t ii = min(i0 + rf_il(u), A.na - 1);
            ux[u] = ax[ii];
            uy[u] = ay[ii];
            uz[u] = az[ii];
            bi[u] = A.ex0 > 0 ? (int)((oa ? (int64_t)oa[ii] : ii) / A.ex0) : -1;
        }
        const int nj = (int)min((int64_t)RF_TJ, A.nb - j0);
        __syncthreads();
        const bool cull = A.tbox_a != nullptr;
        for (int jb = 0; jb < nj; jb += RF_THREADS) {
            const int j = jb + threadIdx.x;
            uint4 pj = make_uint4(0u, 0u, 0u, 0u);
            if (j < nj) {
                const int64_t jj = j0 + j;
                const int bj = A.ex0 > 0 ? (int)((ob ? (int64_t)ob[jj] : jj) / A.ex1) : -2;
                pj = make_uint4(bx[jj], by[jj], bz[jj], (unsigned)bj);
                S->sj[j] = pj;
            }
            if (cull) {   // one warp stages one chunk of 32 rows
                const bool in = j < nj;
                const uint4 lo = make_uint4(__reduce_min_sync(0xffffffffu, in ? pj.x : 0xffffffffu),
                                            __reduce_min_sync(0xffffffffu, in ? pj.y : 0xffffffffu),
                                            __reduce_min_sync(0xffffffffu, in ? pj.z : 0xffffffffu), 0u);
                const uint4 hi = make_uint4(__reduce_max_sync(0xffffffffu, in ? pj.x : 0u),
                                            __reduce_max_sync(0xffffffffu, in ? pj.y : 0u),
                                            __reduce_max_sync(0xffffffffu, in ? pj.z : 0u), 0u);
                if ((threadIdx.x & 31u) == 0u) {
                    S->jbox[j >> 5][0] = lo;
                    S->jbox[j >> 5][1] = hi;
                }
            }
        }
        if (cull) {
            // bounding boxes of the warp's first 64 particles (u = 0, 1), of its last 64 (u = 2, 3) and of all 128
            uint4 lo2[2], hi2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int a = 2 * h, b = 2 * h + 1;
                lo2[h] = make_uint4(__reduce_min_sync(0xffffffffu, min(ux[a], ux[b])), __reduce_min_sync(0xffffffffu, min(uy[a], uy[b])),
                                    __reduce_min_sync(0xffffffffu, min(uz[a], uz[b])), 0u);
                hi2[h] = make_uint4(__reduce_max_sync(0xffffffffu, max(ux[a], ux[b])), __reduce_max_sync(0xffffffffu, max(uy[a], uy[b])),
                                    __reduce_max_sync(0xffffffffu, max(uz[a], uz[b])), 0u);
            }
            if ((threadIdx.x & 31u) == 0u) {
                uint4 *ib = S->ibox[threadIdx.x >> 5];
                const uint4 lo = make_uint4(min(lo2[0].x, lo2[1].x), min(lo2[0].y, lo2[1].y), min(lo2[0].z, lo2[1].z), 0u);
                const uint4 hi = make_uint4(max(hi2[0].x, hi2[1].x), max(hi2[0].y, hi2[1].y), max(hi2[0].z, hi2[1].z), 0u);
                ib[0] = lo;
                ib[1] = hi;
                // centre (rounded down) and half extent (rounded up, + 1): |u - centre| <= half for every u in the box
                auto centre_half = [](const uint4 &l, const uint4 &h, uint4 *out) {
                    out[0] = make_uint4(l.x + ((h.x - l.x) >> 1), l.y + ((h.y - l.y) >> 1), l.z + ((h.z - l.z) >> 1), 0u);
                    out[1] = make_uint4(((h.x - l.x) >> 1) + 1u, ((h.y - l.y) >> 1) + 1u, ((h.z - l.z) >> 1) + 1u, 0u);
                };
                centre_half(lo, hi, ib + 2);
                centre_half(lo2[0], hi2[0], ib + 4);
                centre_half(lo2[1], hi2[1], ib + 6);
            }
            if (threadIdx.x < 3) S->cull[threadIdx.x] = (float)((ftp->on ? ftp->w[threadIdx.x] : fb.box[threadIdx.x]) * (1.0 / 4294967296.0));
            if (threadIdx.x == 4) S->cull_tri = ftp->on;
            if (threadIdx.x == 3) S->cull[3] = (float)(A.cull_r * A.cull_r);
            if (threadIdx.x >= 32 && threadIdx.x < 32 + RF_THREADS / 32) S->next_chunk[threadIdx.x - 32] = 0;
        }
        __syncthreads();
        static_assert(sizeof(TileCtx) <= sizeof(S->ctx), "TileCtx must fit its shared-memory slot");
        TileCtx *Tp = reinterpret_cast<TileCtx *>(S->ctx);
        // frames whose margin is too wide for the filter (usable == 0) send every pair to the FP64 pass
        const float thr = ff.usable ? ff.thr : -1.f;
        if (threadIdx.x == 0) {
            Tp->A = A;
            Tp->ff = ff;
            Tp->fb = fb;
            Tp->ft = *ftp;
            Tp->i0 = i0;
            Tp->j0 = j0;
            Tp->f = f;
            Tp->diag = diag ? 1 : 0;
            Tp->thr = thr;
        }
        __syncthreads();
        const bool check = diag || (i0 + RF_TI > A.na);
        const bool r0zero = ff.c0 == -0.5f;   // r0 == 0: v >= -0.5, the bin index cannot be negative
        // the block exclusion i / ex0 == j / ex1 can only bite where the two tiles' block ranges overlap
        bool excl = false;
        if (A.ex0 > 0) {
            const int64_t ilo = i0 / A.ex0, ihi = min(i0 + RF_TI - 1, A.na - 1) / A.ex0;
            const int64_t jlo = j0 / A.ex1, jhi = (j0 + nj - 1) / A.ex1;
            // sorted frames: a tile holds any original indices; but blocks of one particle within one group exclude
            // the self pairs only, which a symmetric sweep (j > i) never evaluates
            excl = oa != nullptr ? !(A.sym && A.ex0 == 1 && A.ex1 == 1) : (ilo <= jhi && jlo <= ihi);
        }
#define RF_CALL2(C, Q, E, Z, U, X, L) tile_pairs_fast<C, Q, E, Z, U, X, L>(Tp, S, nj, ux, uy, uz, bi, ff, thr, (unsigned)B, A.na, i0, j0, diag, sT, hist)
#define RF_CALL(C, Q, E)                                                                                              \
    do {                                                                                                              \
        if (ext && l32) { if (cull) RF_CALL2(C, Q, E, true, true, true, true); else RF_CALL2(C, Q, E, true, false, true, true); } \
        else if (ext) { if (cull) RF_CALL2(C, Q, E, true, true, true, false); else RF_CALL2(C, Q, E, true, false, true, false); } \
        else if (cull) { if (r0zero) RF_CALL2(C, Q, E, true, true, false, false); else RF_CALL2(C, Q, E, false, true, false, false); } \
        else { if (r0zero) RF_CALL2(C, Q, E, true, false, false, false); else RF_CALL2(C, Q, E, false, false, false, false); }        \
    } while (0)
#define RF_KIND(C, E)                                            \
    do {                                                         \
        if (ff.kind == FK_CUBIC) RF_CALL(C, FK_CUBIC, E);        \
        else if (ff.kind == FK_TRI) RF_CALL(C, FK_TRI, E);       \
        else RF_CALL(C, FK_ORTHO, E);                            \
    } while (0)
        if (check) {
            if (excl) RF_KIND(true, true); else RF_KIND(true, false);
        } else {
            if (excl) RF_KIND(false, true); else RF_KIND(false, false);
        }
#undef RF_KIND
#undef RF_CALL
#undef RF_CALL2
        // resolve the undecided pairs of this tile with the reference's FP64 arithmetic
        __syncthreads();
        const unsigned nq = min(S->qcount, S->qcap);
        for (unsigned e = threadIdx.x; e < nq; e += RF_THREADS)
            resolve_entry(S, S->queue[e], ext, sT, sh + A.ext_off * A.hist_ks, hist);
        if (S->overflow) {   // block-uniform: rows that did not fit the queue, one byte per (row, group, lane)
            unsigned *tab = reinterpret_cast<unsigned *>(S->ovf);
            for (unsigned w = threadIdx.x; w < (unsigned)(RF_TJ * RF_THREADS / 4); w += RF_THREADS) {
                unsigned four = tab[w];
                if (!four) continue;
                tab[w] = 0u;
                for (unsigned b = 0; b < 4u; ++b, four >>= 8) {
                    const unsigned v = four & 0xffu;
                    if (!v) continue;
                    const unsigned idx = 4u * w + b, jl = idx / RF_THREADS, gl = idx % RF_THREADS;   // gl = group * 32 + lane
                    resolve_entry(S, jl | ((gl & 31u) << 10) | ((gl >> 5) << 15) | ((v & 15u) << 18) | ((v >> 4) << 22), ext, sT,
                                  sh + A.ext_off * A.hist_ks, hist);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            S->qcount = 0u;
            S->overflow = 0u;
        }

        if (++since_flush >= RF_FLUSH_ITEMS) {
            flush_hist(sh + A.ext_off * A.hist_ks, B, A.n_copies, weight, A.counts, true, A.hist_cs, A.hist_ks);
            since_flush = 0;
        }
    }
    __syncthreads();
    flush_hist(sh + A.ext_off * A.hist_ks, B, A.n_copies, weight, A.counts, false, A.hist_cs, A.hist_ks);
}

__global__ void rdf_add_const(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }

}  // namespace

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------

struct mdc_rdf_plan {
    mdc_ctx *ctx = nullptr;
    int n_bins = 0, n_copies = 1, n_copies_fast = 1;
    double r0 = 0, r1 = 0;
    int64_t na = 0, nb = 0, ex0 = 0, ex1 = 0;
    bool same = false, sym = false;
    int self_bin = -1;          // bin receiving the i == j pairs of a symmetric self-RDF (-1: none)
    int force_exact = 0;        // testing / diagnostics: MDC_RDF_EXACT=1 in the environment
    int occ_fast = 1, occ_exact = 1;   // resident blocks per SM (persistent grids)
    int blocks_per_sm = 0;             // > 0: the filter kernel launches at most this many blocks per SM
    int ext_cap = 0;                   // > 0: words per histogram copy the filter kernel's shared memory is sized for
    int ext_cap32 = 0;                 // > 0: the same with 32 interleaved copies (RdfArgs::hist_ks = 32)
    Thresholds th;
    DevBuf<double> T;
    DevBuf<unsigned long long> counts, ticket;
    DevBuf<unsigned char> ovf;         // filter kernel: overflow tables, one per block of the persistent grid, kept zero
    DevBuf<float> raw_a[2], raw_b[2], soa_a[2], soa_b[2];
    DevBuf<uint32_t> fix_a[2], fix_b[2];
    DevBuf<unsigned> absmax[2];        // (2 groups, {max x, max -x}, fblk): coordinate extents per frame
    DevBuf<FrameBox> boxes[2];
    DevBuf<FrameTri> tri[2];
    DevBuf<FrameFast> fast[2];
    // spatially sorted copies (cutoffs well below half the box), allocated on first use
    DevBuf<float> sort_soa_a[2], sort_soa_b[2];
    DevBuf<uint32_t> sort_fix_a[2], sort_fix_b[2], tbox_a[2], tbox_b[2];
    DevBuf<int> orig_a[2], orig_b[2], cells[2];
    int sort_bits_a = 0, sort_bits_b = 0;
    int64_t last_culled_hint = 0;
    Streams st;
    int fblk = 0;
    size_t smem = 0, smem_fast = 0;
    int64_t n_frames = 0;
    int last_launches = 0;
    int last_info[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // mdc_rdf_plan_info: the kernel path of the last frame block pushed
};

namespace {

void rdf_free(mdc_rdf_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();
    p->T.release();
    p->counts.release();
    p->ticket.release();
    p->ovf.release();
    for (int i = 0; i < 2; ++i) {
        p->raw_a[i].release();
        p->raw_b[i].release();
        p->soa_a[i].release();
        p->soa_b[i].release();
        p->fix_a[i].release();
        p->fix_b[i].release();
        p->absmax[i].release();
        p->boxes[i].release();
        p->tri[i].release();
        p->fast[i].release();
        p->sort_soa_a[i].release();
        p->sort_soa_b[i].release();
        p->sort_fix_a[i].release();
        p->sort_fix_b[i].release();
        p->tbox_a[i].release();
        p->tbox_b[i].release();
        p->orig_a[i].release();
        p->orig_b[i].release();
        p->cells[i].release();
    }
    p->st.destroy();
    delete p;
}

int rdf_configure(mdc_rdf_plan *p, int n_frames_hint)
{
    if (p->fblk > 0) return MDC_OK;
    const int64_t per_frame = (p->na + (p->same ? 0 : p->nb)) * 3 * (int64_t)sizeof(float);
    int fblk = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_frames_hint, 64), ((int64_t)1 << 30) / per_frame));
    p->fblk = fblk;
    for (int i = 0; i < 2; ++i) {
        MDC_CHECK(p->raw_a[i].alloc((size_t)fblk * p->na * 3));
        MDC_CHECK(p->soa_a[i].alloc((size_t)fblk * p->na * 3));
        MDC_CHECK(p->fix_a[i].alloc((size_t)fblk * p->na * 3));
        if (!p->same) {
            MDC_CHECK(p->raw_b[i].alloc((size_t)fblk * p->nb * 3));
            MDC_CHECK(p->soa_b[i].alloc((size_t)fblk * p->nb * 3));
            MDC_CHECK(p->fix_b[i].alloc((size_t)fblk * p->nb * 3));
        }
        MDC_CHECK(p->absmax[i].alloc((size_t)4 * fblk));
        MDC_CHECK(p->boxes[i].alloc(fblk));
        MDC_CHECK(p->tri[i].alloc(fblk));
        MDC_CHECK(p->fast[i].alloc(fblk));
    }
    return MDC_OK;
}

// MDAnalysis.lib.mdamath.triclinic_vectors(dimensions, dtype=float32), restated (see oracle/oracle.c): the box matrix is
// formed in float64 and rounded to float32; cos(90 deg) is exactly 0.  false: not a valid cell.
bool triclinic_vectors(const float *b, FrameTri &t)
{
    const double lx = b[0], ly = b[1], lz = b[2], alpha = b[3], beta = b[4], gamma = b[5];
    const double deg = 0.017453292519943295;
    if (!(lx > 0 && ly > 0 && lz > 0 && alpha > 0 && beta > 0 && gamma > 0 && alpha < 180 && beta < 180 && gamma < 180))
        return false;
    const double cos_alpha = alpha == 90.0 ? 0.0 : cos(alpha * deg);
    const double cos_beta = beta == 90.0 ? 0.0 : cos(beta * deg);
    double cos_gamma = 0.0, sin_gamma = 1.0;
    if (gamma != 90.0) {
        cos_gamma = cos(gamma * deg);
        sin_gamma = sin(gamma * deg);
    }
    const double bx = ly * cos_gamma, by = ly * sin_gamma;
    const double cx = lz * cos_beta, cy = lz * (cos_alpha - cos_beta * cos_gamma) / sin_gamma;
    const double cz = sqrt(lz * lz - cx * cx - cy * cy);
    if (!(cz > 0.0)) return false;
    t.ax = (double)(float)lx;
    t.bx = (double)(float)bx;
    t.by = (double)(float)by;
    t.cx = (double)(float)cx;
    t.cy = (double)(float)cy;
    t.cz = (double)(float)cz;
    // perpendicular widths: volume / area of the opposite face (a x b = (0, 0, ax by), c x a = (0, cz ax, -cy ax),
    // b x c = (by cz, -bx cz, bx cy - by cx))
    const double vol = t.ax * t.by * t.cz;
    const double bc = sqrt(t.by * t.cz * t.by * t.cz + t.bx * t.cz * t.bx * t.cz + (t.bx * t.cy - t.by * t.cx) * (t.bx * t.cy - t.by * t.cx));
    t.w[0] = vol / bc;
    t.w[1] = vol / (fabs(t.ax) * sqrt(t.cz * t.cz + t.cy * t.cy));
    t.w[2] = fabs(t.cz);
    t.on = 1;
    t.pad = 0;
    return true;
}

int make_boxes(const float *box, int F, std::vector<FrameBox> &out, std::vector<FrameTri> &tri)
{
    out.resize(F);
    tri.resize(F);
    for (int f = 0; f < F; ++f) {
        const float *b = box + 6 * f;
        memset(&tri[f], 0, sizeof(FrameTri));
        if (!(b[3] == 90.f && b[4] == 90.f && b[5] == 90.f)) {
            if (!triclinic_vectors(b, tri[f]))
                return fail(MDC_ERR_INVALID, "mdc_rdf_push: box lengths must be positive and box angles must describe a valid cell");
        }
        for (int k = 0; k < 3; ++k) {
            if (!(b[k] >= 0.f) || !isfinite(b[k])) return fail(MDC_ERR_INVALID, "mdc_rdf_push: invalid box length");
            const float inv = (float)(1.0 / (double)b[k]);   // calc_distances.h: inverse_box stored as float
            out[f].box[k] = (double)b[k];
            out[f].inv[k] = (double)inv;
            out[f].pbc[k] = b[k] > FLT_EPSILON ? 1 : 0;
        }
        out[f].pad = 0;
    }
    return MDC_OK;
}

}  // namespace

extern "C" int mdc_rdf_plan_create(mdc_ctx *ctx, int n_bins, double r0, double r1, int64_t n_a, int64_t n_b,
                                   int same_group, int64_t excl0, int64_t excl1, mdc_rdf_plan **out)
{
    if (!ctx || !out) return fail(MDC_ERR_INVALID, "mdc_rdf_plan_create: NULL argument");
    *out = nullptr;
    if (n_bins < 1) return fail(MDC_ERR_INVALID, "mdc_rdf_plan_create: n_bins must be positive");
    if (!(r1 > r0) || !(r0 >= 0.0) || !isfinite(r1)) return fail(MDC_ERR_INVALID, "mdc_rdf_plan_create: need 0 <= r0 < r1");
    if (n_a < 1 || (!same_group && n_b < 1)) return fail(MDC_ERR_INVALID, "mdc_rdf_plan_create: empty group");
    if (n_a > 2147483647ll || n_b > 2147483647ll) return fail(MDC_ERR_INVALID, "mdc_rdf_plan_create: group too large");
    if (excl0 > 0 && excl1 <= 0) return fail(MDC_ERR_INVALID, "mdc_rdf_plan_create: exclusion needs two positive block sizes");
    MDC_CUDA(cudaSetDevice(ctx->device));

    mdc_rdf_plan *p = new mdc_rdf_plan();
    p->ctx = ctx;
    p->n_bins = n_bins;
    p->r0 = r0;
    p->r1 = r1;
    p->same = same_group != 0;
    p->na = n_a;
    p->nb = p->same ? n_a : n_b;
    p->ex0 = excl0 > 0 ? excl0 : 0;
    p->ex1 = excl0 > 0 ? excl1 : 0;
    // unordered-pair traversal needs a symmetric exclusion rule
    p->sym = p->same && (p->ex0 == 0 || p->ex0 == p->ex1);
    p->th = make_thresholds(n_bins, r0, r1);
    if (p->sym && p->ex0 == 0 && p->th.T[0] == 0.0) {
        // i == j pairs (d = 0) are counted by the reference (structure.py:109-115 keeps d > r0 - eps)
        int k = 0;
        while (k < n_bins && !(0.0 < p->th.T[k + 1])) ++k;
        p->self_bin = (k < n_bins) ? k : ((0.0 >= p->th.e_lo && 0.0 < p->th.e_hi) ? n_bins - 1 : -1);
    }

    // shared memory: j tile + thresholds + privatised histograms (as many warp copies as fit)
    const size_t fixed = sizeof(float4) * RDF_TJ + sizeof(double) * (n_bins + 1);
    const size_t budget = std::min<size_t>(ctx->smem_optin, 200 * 1024);
    int copies = RDF_THREADS / 32;
    while (copies > 1 && fixed + (size_t)copies * n_bins * 4 > budget) copies >>= 1;
    if (fixed + (size_t)copies * n_bins * 4 > budget) {
        delete p;
        return fail(MDC_ERR_UNSUPPORTED, "mdc_rdf_plan_create: n_bins too large for the shared-memory histogram");
    }
    p->n_copies = copies;
    p->smem = fixed + (size_t)copies * n_bins * 4;
    const size_t fixed_fast = ((sizeof(FastShared) + 15) & ~size_t(15)) + sizeof(double) * (n_bins + 1) + sizeof(unsigned) * RF_QCAP;
    int copies_fast = RF_THREADS / 32;
    while (copies_fast > 1 && fixed_fast + (size_t)copies_fast * n_bins * 4 > budget) copies_fast >>= 1;
    p->n_copies_fast = copies_fast;
    p->smem_fast = fixed_fast + (size_t)copies_fast * n_bins * 4;   // > budget: the filter kernel is not used
    // extended histograms (RdfArgs::ext_k): as many words per copy as still leave MDC_RF_BLOCKS blocks per SM -- eight
    // copies one after the other, or (preferred where the bins of a frame block fit) 32 interleaved ones with a
    // shorter queue of undecided rows
    {
        const size_t per_block = ctx->smem_per_sm / MDC_RF_BLOCKS;
        const size_t room = per_block > 1024 + fixed_fast ? per_block - 1024 - fixed_fast : 0;   // 1 KB reserved per block
        int cap = (int)std::min<size_t>(room / ((size_t)copies_fast * 4), 4096);
        if (cap > 36) cap -= (cap - 4) % 32;   // stride = 4 (mod 32) words: equal bins of the 8 copies sit in 8 different banks
        if (copies_fast == RF_THREADS / 32 && cap >= n_bins + 8 && fixed_fast + (size_t)copies_fast * cap * 4 <= budget) {
            p->ext_cap = cap;
            p->smem_fast = fixed_fast + (size_t)copies_fast * cap * 4;
            const size_t room32 = room + sizeof(unsigned) * (RF_QCAP - RF_QCAP_L32);
            const int cap32 = (int)std::min<size_t>(room32 / 128, 4096);
            static const bool l32_off_env = getenv("MDC_RDF_L32") && getenv("MDC_RDF_L32")[0] == '0';
            if (cap32 >= n_bins + 8 && !l32_off_env) {
                p->ext_cap32 = cap32;
                p->smem_fast = std::max(p->smem_fast, fixed_fast - sizeof(unsigned) * (RF_QCAP - RF_QCAP_L32) + (size_t)cap32 * 128);
            }
        }
    }
    const char *env = getenv("MDC_RDF_EXACT");
    p->force_exact = (env && env[0] == '1') || p->smem_fast > budget;

    int rc = p->st.init(ctx);
    if (rc == MDC_OK) rc = p->T.alloc(n_bins + 1);
    if (rc == MDC_OK) rc = p->counts.alloc(n_bins);
    if (rc == MDC_OK) rc = p->ticket.alloc(1);
    if (rc != MDC_OK) {
        rdf_free(p);
        return rc;
    }
    cudaError_t e = cudaMemcpy(p->T.p, p->th.T.data(), (n_bins + 1) * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(p->counts.p, 0, n_bins * sizeof(unsigned long long));
    // The limit is a per-function (per-device) attribute shared by every live plan: it is only ever set to the device
    // maximum, so that a later plan with a smaller histogram cannot lower it under an earlier plan's launches.
    if (e == cudaSuccess) e = raise_smem_limit(rdf_pairs_exact, ctx->smem_optin);
    if (e == cudaSuccess) e = raise_smem_limit(rdf_pairs_fast, ctx->smem_optin);
    if (e == cudaSuccess && !p->force_exact)   // four ~42 KB blocks per SM need more than the default carve-out
        e = cudaFuncSetAttribute(rdf_pairs_fast, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p->occ_exact, rdf_pairs_exact, RDF_THREADS, p->smem);
    if (e == cudaSuccess && !p->force_exact)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p->occ_fast, rdf_pairs_fast, RF_THREADS, p->smem_fast);
    if (e != cudaSuccess) {
        rdf_free(p);
        return fail(MDC_ERR_CUDA, std::string("mdc_rdf_plan_create: ") + cudaGetErrorString(e));
    }
    p->occ_exact = std::max(1, p->occ_exact);
    p->occ_fast = std::max(1, p->occ_fast);
    *out = p;
    return MDC_OK;
}

extern "C" int mdc_rdf_plan_destroy(mdc_rdf_plan *plan)
{
    rdf_free(plan);
    return MDC_OK;
}

extern "C" int mdc_rdf_push(mdc_rdf_plan *p, const float *pos_a, const float *pos_b, const float *box,
                            int n_frames, int on_device)
{
    if (!p || !pos_a || !box) return fail(MDC_ERR_INVALID, "mdc_rdf_push: NULL argument");
    if (!p->same && !pos_b) return fail(MDC_ERR_INVALID, "mdc_rdf_push: pos_b is required for two groups");
    if (n_frames <= 0) return MDC_OK;
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CHECK(rdf_configure(p, n_frames));
    std::vector<FrameBox> hb;
    std::vector<FrameTri> ht;
    MDC_CHECK(make_boxes(box, n_frames, hb, ht));
    cudaStream_t sc = p->st.copy, sk = p->st.compute;
    p->st.n_timed = 0;
    int launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += p->fblk) {
        const int F = std::min(p->fblk, n_frames - f0);
        const int slot = p->st.next;
        p->st.next ^= 1;
        if (p->st.used[slot]) MDC_CUDA(cudaStreamWaitEvent(sc, p->st.consumed[slot], 0));
        const float *ra = pos_a + (size_t)f0 * p->na * 3;
        const float *rb = p->same ? nullptr : pos_b + (size_t)f0 * p->nb * 3;
        if (!on_device) {
            MDC_CUDA(cudaMemcpyAsync(p->raw_a[slot].p, ra, (size_t)F * p->na * 3 * sizeof(float), cudaMemcpyHostToDevice, sc));
            ra = p->raw_a[slot].p;
            if (rb) {
                MDC_CUDA(cudaMemcpyAsync(p->raw_b[slot].p, rb, (size_t)F * p->nb * 3 * sizeof(float), cudaMemcpyHostToDevice, sc));
                rb = p->raw_b[slot].p;
            }
        }
        MDC_CUDA(cudaMemcpyAsync(p->boxes[slot].p, hb.data() + f0, F * sizeof(FrameBox), cudaMemcpyHostToDevice, sc));
        MDC_CUDA(cudaMemcpyAsync(p->tri[slot].p, ht.data() + f0, F * sizeof(FrameTri), cudaMemcpyHostToDevice, sc));
        MDC_CUDA(cudaMemsetAsync(p->absmax[slot].p, 0, 4 * (size_t)p->fblk * sizeof(unsigned), sc));
        rdf_prep<<<(unsigned)ceil_div((int64_t)F * p->na, 256), 256, 0, sc>>>(
            ra, p->na, (int64_t)F * p->na, p->boxes[slot].p, p->tri[slot].p, p->fblk, p->soa_a[slot].p, p->fix_a[slot].p,
            p->absmax[slot].p);
        ++launches;
        if (rb) {
            rdf_prep<<<(unsigned)ceil_div((int64_t)F * p->nb, 256), 256, 0, sc>>>(
                rb, p->nb, (int64_t)F * p->nb, p->boxes[slot].p, p->tri[slot].p, p->fblk, p->soa_b[slot].p, p->fix_b[slot].p,
                p->absmax[slot].p + 2 * p->fblk);
            ++launches;
        }
        rdf_frame_params<<<(unsigned)ceil_div(F, 64), 64, 0, sc>>>(
            p->boxes[slot].p, p->tri[slot].p, p->absmax[slot].p, p->absmax[slot].p + (rb ? 2 * p->fblk : 0), p->fblk, F, p->n_bins, p->r0, p->r1,
            p->fast[slot].p);
        ++launches;
        // the filter kernel needs an orthorhombic cell, three periodic axes and a bin coordinate below 2^22 in every frame
        // (triclinic frames take the exact kernel: minimum_image_triclinic in FP64 for every pair)
        // (triclinic frames: the filter's image is the wrap of the fractional differences, which is the reference's
        // minimum image for cutoffs below half the smallest width of the cell -- rdf_frame_params; wider windows take
        // the exact kernel: minimum_image_triclinic in FP64 for every pair)
        bool fast = !p->force_exact;
        static const bool tri_filter_off = getenv("MDC_RDF_TRI_FILTER") && getenv("MDC_RDF_TRI_FILTER")[0] == '0';
        double max_len = 0.0, v_far_max = 0.0;
        for (int f = 0; f < F && fast; ++f) {
            const FrameBox &fb = hb[f0 + f];
            const FrameTri &ft = ht[f0 + f];
            const double S = (double)p->n_bins / (p->r1 - p->r0);
            double v_far;
            if (ft.on) {
                const double Lsum = fabs(ft.ax) + sqrt(ft.bx * ft.bx + ft.by * ft.by) + sqrt(ft.cx * ft.cx + ft.cy * ft.cy + ft.cz * ft.cz);
                const double wmin = std::min(ft.w[0], std::min(ft.w[1], ft.w[2]));
                v_far = 0.5 * Lsum * S + fabs(p->r0 * S) + 1.0;
                fast = !tri_filter_off && p->r1 * (1.0 + 1e-4) + 1e-5 * Lsum < 0.5 * wmin && v_far < 4.0e6;
                max_len = std::max(max_len, Lsum);
            } else {
                v_far = 0.5 * sqrt(fb.box[0] * fb.box[0] + fb.box[1] * fb.box[1] + fb.box[2] * fb.box[2]) * S + fabs(p->r0 * S) + 1.0;
                fast = fb.pbc[0] && fb.pbc[1] && fb.pbc[2] && v_far < 4.0e6;
                for (int k = 0; k < 3; ++k) max_len = std::max(max_len, fb.box[k]);
            }
            v_far_max = std::max(v_far_max, v_far);
        }
        // Put every frame in Hilbert order of a cell grid, so that runs of consecutive particles are compact regions,
        // and let the pair kernel skip what lies farther apart than the cutoff: whole tile pairs, and inside a tile
        // pair the (warp of 128 i) x (chunk of 32 j) blocks (MDAnalysis switches capped_distance to a grid search
        // for small cutoffs; here the same sweep serves every cutoff: at r_max = L / 2 a few per cent of the blocks
        // drop out, at r_max << L nearly all).  Counts do not depend on the particle order; exclusion uses the
        // original indices.
        const char *sort_env = getenv("MDC_RDF_SORT");   // "0": keep the original order (diagnostics / tests)
        const bool sort_off = sort_env && sort_env[0] == '0';
        const bool sorted = fast && !sort_off && p->na >= 4 * RF_TI && p->nb >= 4 * RF_TJ;
        const int nti_fast = (int)ceil_div(p->na, RF_TI), ntj_fast = (int)ceil_div(p->nb, RF_TJ);
        if (sorted) {
            auto bits_for = [](int64_t n) {
                int b = 0;
                while (b < 5 && ((int64_t)32 << (3 * b)) < n) ++b;   // ~32 particles per cell
                return b;
            };
            p->sort_bits_a = bits_for(p->na);
            p->sort_bits_b = bits_for(p->nb);
            const int max_bits = std::max(p->sort_bits_a, p->sort_bits_b);
            MDC_CHECK(p->cells[slot].alloc((size_t)p->fblk << (3 * max_bits)));
            auto sort_group = [&](const float *soa, const uint32_t *fix, int64_t n, int bits, int n_tiles, DevBuf<float> &soa_s,
                                  DevBuf<uint32_t> &fix_s, DevBuf<int> &orig, DevBuf<uint32_t> &tbox) -> int {
                MDC_CHECK(soa_s.alloc((size_t)p->fblk * n * 3));
                MDC_CHECK(fix_s.alloc((size_t)p->fblk * n * 3));
                MDC_CHECK(orig.alloc((size_t)p->fblk * n));
                MDC_CHECK(tbox.alloc((size_t)p->fblk * n_tiles * 6));
                const int64_t total = (int64_t)F * n;
                MDC_CUDA(cudaMemsetAsync(p->cells[slot].p, 0, ((size_t)F << (3 * bits)) * sizeof(int), sc));
                rdf_cell_count<<<(unsigned)ceil_div(total, 256), 256, 0, sc>>>(fix, n, total, bits, p->cells[slot].p);
                rdf_cell_scan<<<F, 1024, 0, sc>>>(bits, p->cells[slot].p);
                rdf_cell_scatter<<<(unsigned)ceil_div(total, 256), 256, 0, sc>>>(soa, fix, n, total, bits, p->cells[slot].p,
                                                                              soa_s.p, fix_s.p, orig.p);
                rdf_tile_boxes<<<dim3((unsigned)n_tiles, (unsigned)F), 256, 0, sc>>>(fix_s.p, n, RF_TI, n_tiles, tbox.p);
                launches += 4;
                return MDC_OK;
            };
            static_assert(RF_TI == RF_TJ, "one tile size for both groups");
            MDC_CHECK(sort_group(p->soa_a[slot].p, p->fix_a[slot].p, p->na, p->sort_bits_a, nti_fast, p->sort_soa_a[slot],
                                 p->sort_fix_a[slot], p->orig_a[slot], p->tbox_a[slot]));
            if (!p->same)
                MDC_CHECK(sort_group(p->soa_b[slot].p, p->fix_b[slot].p, p->nb, p->sort_bits_b, ntj_fast, p->sort_soa_b[slot],
                                     p->sort_fix_b[slot], p->orig_b[slot], p->tbox_b[slot]));
        }
        MDC_CUDA(cudaGetLastError());
        MDC_CUDA(cudaEventRecord(p->st.staged[slot], sc));
        MDC_CUDA(cudaStreamWaitEvent(sk, p->st.staged[slot], 0));
        RdfArgs A;
        A.a = p->soa_a[slot].p;
        A.b = p->same ? p->soa_a[slot].p : p->soa_b[slot].p;
        A.ua = p->fix_a[slot].p;
        A.ub = p->same ? p->fix_a[slot].p : p->fix_b[slot].p;
        A.ext_k = A.ext_off = 0;
        A.n_copies = fast ? p->n_copies_fast : p->n_copies;
        A.hist_ks = 1;
        A.hist_cs = p->n_bins;
        A.qcap = RF_QCAP;
        {
            static const bool rel_off_env = getenv("MDC_RDF_REL") && getenv("MDC_RDF_REL")[0] == '0';
            A.rel = rel_off_env ? 0 : 1;
        }
        if (fast && p->ext_cap > 0) {
            // bins the filter can produce: k >= -(r0 S) - 1 (d >= 0) and k <= v_far (half the box diagonal), with slack
            const double S = (double)p->n_bins / (p->r1 - p->r0);
            const double k_lo = -(std::ceil(std::fabs(p->r0 * S)) + 2.0), k_hi = std::ceil(v_far_max) + 1.0;
            static const bool ext_off_env = getenv("MDC_RDF_EXT") && getenv("MDC_RDF_EXT")[0] == '0';
            if (k_hi - k_lo + 1.0 <= (double)p->ext_cap32 && !ext_off_env) {
                A.ext_k = p->ext_cap32;   // one copy per lane (= per bank), bins 32 words apart
                A.ext_off = (int)(-k_lo);
                A.n_copies = 32;
                A.hist_ks = 32;
                A.hist_cs = 1;
                A.qcap = RF_QCAP_L32;
            } else if (k_hi - k_lo + 1.0 <= (double)p->ext_cap && !ext_off_env) {
                A.ext_k = p->ext_cap;
                A.ext_off = (int)(-k_lo);
                A.hist_cs = p->ext_cap;
            }
        }
        A.oa = A.ob = nullptr;
        A.tbox_a = A.tbox_b = nullptr;
        A.cull_r = 0.0;
        if (sorted) {
            A.a = p->sort_soa_a[slot].p;
            A.ua = p->sort_fix_a[slot].p;
            A.oa = p->orig_a[slot].p;
            A.tbox_a = p->tbox_a[slot].p;
            A.b = p->same ? A.a : p->sort_soa_b[slot].p;
            A.ub = p->same ? A.ua : p->sort_fix_b[slot].p;
            A.ob = p->same ? A.oa : p->orig_b[slot].p;
            A.tbox_b = p->same ? A.tbox_a : p->tbox_b[slot].p;
            // conservative: far above the difference between the reference's float32-derived distance, the fixed-point
            // boxes and the true separation (all below 1e-6 of the box)
            A.cull_r = p->r1 * (1.0 + 1e-4) + 1e-5 * max_len;
        }
        A.boxes = p->boxes[slot].p;
        A.tri = p->tri[slot].p;
        A.fast = p->fast[slot].p;
        A.T = p->T.p;
        A.counts = p->counts.p;
        A.ticket = p->ticket.p;
        A.e_lo = p->th.e_lo;
        A.e_hi = p->th.e_hi;
        A.r0f = (float)p->r0;
        A.scalef = (float)((double)p->n_bins / (p->r1 - p->r0));
        A.na = p->na;
        A.nb = p->nb;
        A.ex0 = p->ex0;
        A.ex1 = p->ex1;
        A.n_bins = p->n_bins;
        A.F = F;
        A.nti = (int)ceil_div(p->na, fast ? RF_TI : RDF_TI);
        A.ntj = (int)ceil_div(p->nb, fast ? RF_TJ : RDF_TJ);
        A.sym = p->sym ? 1 : 0;
        A.items_per_frame = p->sym ? (int64_t)A.nti * (A.nti / 2 + 1) : (int64_t)A.nti * A.ntj;
        const int64_t total = A.items_per_frame * F;
        A.ovf = nullptr;
        if (fast) {   // (before the timed region: the allocation blocks the host)
            // one table per block of the grid this push launches (small systems: a few blocks, not 113 MB)
            const int64_t max_blocks = std::min<int64_t>(total, (int64_t)p->ctx->sm_count * p->occ_fast);
            const size_t ovf_bytes = (size_t)max_blocks * RF_TJ * RF_THREADS;
            if (p->ovf.n < ovf_bytes) {   // first filtered push (or a larger one): the tables start, and are left, all zero
                MDC_CUDA(cudaStreamSynchronize(sk));   // no kernel of this plan still reads the old tables
                MDC_CHECK(p->ovf.alloc(ovf_bytes));
                MDC_CUDA(cudaMemsetAsync(p->ovf.p, 0, ovf_bytes, sk));
            }
            A.ovf = p->ovf.p;
        }
        {
            bool any_tri = false;
            for (int f = 0; f < F; ++f) any_tri = any_tri || ht[f0 + f].on != 0;
            p->last_info[0] = fast ? 1 : 0;
            p->last_info[1] = sorted ? 1 : 0;
            p->last_info[2] = A.ext_k;
            p->last_info[3] = A.hist_ks;
            p->last_info[4] = (fast && sorted && A.rel) ? 1 : 0;
            p->last_info[5] = any_tri ? 1 : 0;
            p->last_info[6] = fast ? A.qcap : 0;
            p->last_info[7] = F;
        }
        const int ti = p->st.n_timed < 8 ? p->st.n_timed : -1;
        if (ti >= 0) MDC_CUDA(cudaEventRecord(p->st.k0[ti], sk));
        if (fast) {
            // persistent grid: as many blocks as fit, or fewer per SM when the caller wants room left on every SM for
            // the kernels of another stream (mdc_rdf_set_blocks_per_sm)
            const int bps = p->blocks_per_sm > 0 ? std::min(p->blocks_per_sm, p->occ_fast) : p->occ_fast;
            const int grid = (int)std::min<int64_t>(total, (int64_t)p->ctx->sm_count * bps);
            MDC_CUDA(cudaMemsetAsync(p->ticket.p, 0, sizeof(unsigned long long), sk));
            rdf_pairs_fast<<<grid, RF_THREADS, p->smem_fast, sk>>>(A);
        } else {
            const int grid = (int)std::min<int64_t>(total, (int64_t)p->ctx->sm_count * p->occ_exact);
            rdf_pairs_exact<<<grid, RDF_THREADS, p->smem, sk>>>(A);
        }
        ++launches;
        if (ti >= 0) {
            MDC_CUDA(cudaEventRecord(p->st.k1[ti], sk));
            p->st.n_timed = ti + 1;
        }
        if (p->self_bin >= 0) {
            rdf_add_const<<<1, 1, 0, sk>>>(p->counts.p + p->self_bin, (unsigned long long)F * (unsigned long long)p->na);
            ++launches;
        }
        MDC_CUDA(cudaGetLastError());
        MDC_CUDA(cudaEventRecord(p->st.consumed[slot], sk));
        p->st.used[slot] = true;
    }
    MDC_CUDA(cudaStreamSynchronize(sc));   // host buffers (and hb) may be reused on return
    p->n_frames += n_frames;
    p->last_launches = launches;
    return MDC_OK;
}

extern "C" int mdc_rdf_read(mdc_rdf_plan *p, int64_t *counts_out, int64_t *n_frames_out)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_rdf_read: NULL plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->st.compute));
    if (counts_out)
        MDC_CUDA(cudaMemcpy(counts_out, p->counts.p, p->n_bins * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (n_frames_out) *n_frames_out = p->n_frames;
    return MDC_OK;
}

extern "C" int mdc_rdf_set_blocks_per_sm(mdc_rdf_plan *p, int blocks_per_sm)
{
    if (!p || blocks_per_sm < 0) return fail(MDC_ERR_INVALID, "mdc_rdf_set_blocks_per_sm: bad argument");
    p->blocks_per_sm = blocks_per_sm;
    return MDC_OK;
}

extern "C" int mdc_rdf_reset(mdc_rdf_plan *p)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_rdf_reset: NULL plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->st.compute));
    MDC_CUDA(cudaMemset(p->counts.p, 0, p->n_bins * sizeof(unsigned long long)));
    p->n_frames = 0;
    return MDC_OK;
}

extern "C" int mdc_rdf_accum_ptr(mdc_rdf_plan *p, void **dev_ptr, int64_t *n_elems)
{
    if (!p || !dev_ptr || !n_elems) return fail(MDC_ERR_INVALID, "mdc_rdf_accum_ptr: NULL argument");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->st.compute));
    *dev_ptr = p->counts.p;
    *n_elems = p->n_bins;
    return MDC_OK;
}

extern "C" int mdc_rdf_last_push_stats(mdc_rdf_plan *p, float *kernel_ms, int *n_launches)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_rdf_last_push_stats: NULL plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    if (kernel_ms) MDC_CHECK(p->st.kernel_ms(kernel_ms));
    if (n_launches) *n_launches = p->last_launches;
    return MDC_OK;
}

extern "C" int mdc_rdf_plan_info(mdc_rdf_plan *p, int info[8])
{
    if (!p || !info) return fail(MDC_ERR_INVALID, "mdc_rdf_plan_info: NULL argument");
    memcpy(info, p->last_info, sizeof(p->last_info));
    return MDC_OK;
}

extern "C" int mdc_radial_histogram(mdc_ctx *ctx, const float *pos_a, int64_t n_a, const float *pos_b, int64_t n_b,
                                    const float box[6], int n_bins, double r0, double r1, int64_t excl0,
                                    int64_t excl1, int64_t *counts_out)
{
    if (!counts_out) return fail(MDC_ERR_INVALID, "mdc_radial_histogram: NULL argument");
    mdc_rdf_plan *p = nullptr;
    // the functional seam always receives two position arrays (structure.py:34); identical pointers
    // mean a self-RDF
    const int same = (pos_b == nullptr || (pos_b == pos_a && n_a == n_b)) ? 1 : 0;
    MDC_CHECK(mdc_rdf_plan_create(ctx, n_bins, r0, r1, n_a, same ? n_a : n_b, same, excl0, excl1, &p));
    int rc = mdc_rdf_push(p, pos_a, same ? nullptr : pos_b, box, 1, 0);
    if (rc == MDC_OK) rc = mdc_rdf_read(p, counts_out, nullptr);
    mdc_rdf_plan_destroy(p);
    return rc;
}
