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
#include <math.h>
#include <string.h>
#include <vector>

using namespace mdc;

namespace {

constexpr int RDF_THREADS = 256;
constexpr int RDF_IPT = 2;                       // i-particles per thread
constexpr int RDF_TI = RDF_THREADS * RDF_IPT;    // i-tile
constexpr int RDF_TJ = 512;                      // j-tile staged in shared memory
constexpr int RDF_FLUSH_ITEMS = 4096;            // u32 privatised counters: 4096 * 512 * 512 < 2^31

struct FrameBox {
    double box[3];
    double inv[3];
    int pbc[3];
    int pad;
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

// raw f32 (F, N, 3) -> soa f32 (F, 3, N)
__global__ void rdf_prep_soa(const float *__restrict__ raw, int64_t N, int64_t total, float *__restrict__ soa)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t f = i / N, j = i - f * N;
#pragma unroll
    for (int k = 0; k < 3; ++k) soa[(f * 3 + k) * N + j] = raw[3 * i + k];
}

struct RdfArgs {
    const float *a, *b;       // SoA (F, 3, Na), (F, 3, Nb); b == a for a self-RDF
    const FrameBox *boxes;    // (F)
    const double *T;          // (B + 1)
    unsigned long long *counts;
    double e_lo, e_hi;
    float r0f, scalef;        // candidate bin = (sqrt(r2) - r0) * B / (r1 - r0) in FP32
    int64_t na, nb;
    int64_t ex0, ex1;         // ex0 <= 0: no exclusion
    int n_bins, n_copies;
    int F, nti, ntj;          // tiles
    int sym;                  // 1: self-RDF over unordered pairs (each counted twice)
    int64_t items_per_frame;
};

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

template <bool CHECK>
__device__ __forceinline__ void tile_pairs(const RdfArgs &A, const float4 *__restrict__ sj, int nj, int64_t j0,
                                           const float (&xi)[RDF_IPT], const float (&yi)[RDF_IPT],
                                           const float (&zi)[RDF_IPT], const int (&bi)[RDF_IPT],
                                           const int64_t (&ig)[RDF_IPT], const FrameBox &fb, const double *sT,
                                           unsigned *hist, bool diag)
{
    const double T0 = sT[0], TB = sT[A.n_bins];
    for (int j = 0; j < nj; ++j) {
        const float4 pj = sj[j];
        const int bj = __float_as_int(pj.w);
#pragma unroll
        for (int u = 0; u < RDF_IPT; ++u) {
            double r2 = min_image_sq(pj.x - xi[u], fb.inv[0], fb.box[0], fb.pbc[0]);
            r2 = __dadd_rn(r2, min_image_sq(pj.y - yi[u], fb.inv[1], fb.box[1], fb.pbc[1]));
            r2 = __dadd_rn(r2, min_image_sq(pj.z - zi[u], fb.inv[2], fb.box[2], fb.pbc[2]));
            bool ok = (r2 >= T0) && (r2 < A.e_hi) && (bi[u] != bj);
            if (CHECK) ok = ok && (ig[u] < A.na) && (!diag || (j0 + j) > ig[u]);
            if (ok) {
                int k;
                if (r2 < TB) {
                    const float d = sqrtf((float)r2);
                    k = (int)((d - A.r0f) * A.scalef);
                    k = max(0, min(A.n_bins - 1, k));
                    while (r2 < sT[k]) --k;
                    while (r2 >= sT[k + 1]) ++k;
                } else {
                    k = (r2 >= A.e_lo) ? A.n_bins - 1 : -1;
                }
                if (k >= 0) atomicAdd(&hist[k], 1u);
            }
        }
    }
}

__global__ void __launch_bounds__(RDF_THREADS)
rdf_pairs_kernel(RdfArgs A)
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
        const int f = (int)(item / A.items_per_frame);
        const int64_t w = item - (int64_t)f * A.items_per_frame;
        int I, J;
        bool diag = false, skip = false;
        if (A.sym) {
            // unordered tile pairs: J = (I + off) mod nt, off in [0, nt/2]; for even nt the
            // off == nt/2 column would visit each pair twice, keep the lower half
            const int nt = A.nti, half = nt / 2 + 1;
            I = (int)(w / half);
            const int off = (int)(w - (int64_t)I * half);
            J = I + off;
            if (J >= nt) J -= nt;
            diag = off == 0;
            skip = ((nt & 1) == 0) && (off == nt / 2) && (I >= nt / 2);
        } else {
            I = (int)(w / A.ntj);
            J = (int)(w - (int64_t)I * A.ntj);
        }
        if (skip) continue;
        const FrameBox fb = A.boxes[f];
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
            tile_pairs<true>(A, sj, nj, j0, xi, yi, zi, bi, ig, fb, sT, hist, diag);
        else
            tile_pairs<false>(A, sj, nj, j0, xi, yi, zi, bi, ig, fb, sT, hist, false);

        if (++since_flush >= RDF_FLUSH_ITEMS) {
            __syncthreads();
            for (int i = threadIdx.x; i < B; i += RDF_THREADS) {
                unsigned long long s = 0;
                for (int c = 0; c < A.n_copies; ++c) {
                    s += sh[c * B + i];
                    sh[c * B + i] = 0u;
                }
                if (s) atomicAdd(&A.counts[i], s * weight);
            }
            since_flush = 0;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += RDF_THREADS) {
        unsigned long long s = 0;
        for (int c = 0; c < A.n_copies; ++c) s += sh[c * B + i];
        if (s) atomicAdd(&A.counts[i], s * weight);
    }
}

__global__ void rdf_add_const(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }

}  // namespace

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------

struct mdc_rdf_plan {
    mdc_ctx *ctx = nullptr;
    int n_bins = 0, n_copies = 1;
    double r0 = 0, r1 = 0;
    int64_t na = 0, nb = 0, ex0 = 0, ex1 = 0;
    bool same = false, sym = false;
    int self_bin = -1;          // bin receiving the i == j pairs of a symmetric self-RDF (-1: none)
    Thresholds th;
    DevBuf<double> T;
    DevBuf<unsigned long long> counts;
    DevBuf<float> raw_a[2], raw_b[2], soa_a[2], soa_b[2];
    DevBuf<FrameBox> boxes[2];
    Streams st;
    int fblk = 0;
    size_t smem = 0;
    int64_t n_frames = 0;
    int last_launches = 0;
};

namespace {

void rdf_free(mdc_rdf_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();
    p->T.release();
    p->counts.release();
    for (int i = 0; i < 2; ++i) {
        p->raw_a[i].release();
        p->raw_b[i].release();
        p->soa_a[i].release();
        p->soa_b[i].release();
        p->boxes[i].release();
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
        if (!p->same) {
            MDC_CHECK(p->raw_b[i].alloc((size_t)fblk * p->nb * 3));
            MDC_CHECK(p->soa_b[i].alloc((size_t)fblk * p->nb * 3));
        }
        MDC_CHECK(p->boxes[i].alloc(fblk));
    }
    return MDC_OK;
}

int make_boxes(const float *box, int F, std::vector<FrameBox> &out)
{
    out.resize(F);
    for (int f = 0; f < F; ++f) {
        const float *b = box + 6 * f;
        if (!(b[3] == 90.f && b[4] == 90.f && b[5] == 90.f))
            return fail(MDC_ERR_UNSUPPORTED, "mdc_rdf_push: only orthorhombic boxes (90/90/90) are supported");
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

    int rc = p->st.init(ctx);
    if (rc == MDC_OK) rc = p->T.alloc(n_bins + 1);
    if (rc == MDC_OK) rc = p->counts.alloc(n_bins);
    if (rc != MDC_OK) {
        rdf_free(p);
        return rc;
    }
    cudaError_t e = cudaMemcpy(p->T.p, p->th.T.data(), (n_bins + 1) * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(p->counts.p, 0, n_bins * sizeof(unsigned long long));
    if (e == cudaSuccess && p->smem > 48 * 1024)
        e = cudaFuncSetAttribute(rdf_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    if (e != cudaSuccess) {
        rdf_free(p);
        return fail(MDC_ERR_CUDA, std::string("mdc_rdf_plan_create: ") + cudaGetErrorString(e));
    }
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
    MDC_CHECK(make_boxes(box, n_frames, hb));
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
        rdf_prep_soa<<<(unsigned)ceil_div((int64_t)F * p->na, 256), 256, 0, sc>>>(ra, p->na, (int64_t)F * p->na, p->soa_a[slot].p);
        ++launches;
        if (rb) {
            rdf_prep_soa<<<(unsigned)ceil_div((int64_t)F * p->nb, 256), 256, 0, sc>>>(rb, p->nb, (int64_t)F * p->nb, p->soa_b[slot].p);
            ++launches;
        }
        MDC_CUDA(cudaGetLastError());
        MDC_CUDA(cudaEventRecord(p->st.staged[slot], sc));
        MDC_CUDA(cudaStreamWaitEvent(sk, p->st.staged[slot], 0));
        RdfArgs A;
        A.a = p->soa_a[slot].p;
        A.b = p->same ? p->soa_a[slot].p : p->soa_b[slot].p;
        A.boxes = p->boxes[slot].p;
        A.T = p->T.p;
        A.counts = p->counts.p;
        A.e_lo = p->th.e_lo;
        A.e_hi = p->th.e_hi;
        A.r0f = (float)p->r0;
        A.scalef = (float)((double)p->n_bins / (p->r1 - p->r0));
        A.na = p->na;
        A.nb = p->nb;
        A.ex0 = p->ex0;
        A.ex1 = p->ex1;
        A.n_bins = p->n_bins;
        A.n_copies = p->n_copies;
        A.F = F;
        A.nti = (int)ceil_div(p->na, RDF_TI);
        A.ntj = (int)ceil_div(p->nb, RDF_TJ);
        A.sym = p->sym ? 1 : 0;
        A.items_per_frame = p->sym ? (int64_t)A.nti * (A.nti / 2 + 1) : (int64_t)A.nti * A.ntj;
        const int64_t total = A.items_per_frame * F;
        const int grid = (int)std::min<int64_t>(total, (int64_t)p->ctx->sm_count * 8);
        const int ti = p->st.n_timed < 8 ? p->st.n_timed : -1;
        if (ti >= 0) MDC_CUDA(cudaEventRecord(p->st.k0[ti], sk));
        rdf_pairs_kernel<<<grid, RDF_THREADS, p->smem, sk>>>(A);
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
