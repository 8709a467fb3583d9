// sq.cu — S(q): StructureFactor per-frame arithmetic on sm_100a.
//
// Replaces (all paths relative to /root/reference/src/mdcraft/):
//   analysis/structure.py:1822-1887  wavevector grid       -> lat_row_counts / lat_fill (device)
//   analysis/structure.py:1240-1280  delta_fourier_transform_sum
//   algorithm/accelerated.py:84-113  numba_dot             -> sq_lattice_direct / sq_explicit_* kernels
//   analysis/structure.py:1940-1970  _single_frame (exp)   -> sq_pair_accum
//
// Data layout in HBM, per frame block of F frames (one plan, two slots):
//   raw   f32 (F, N, 3)        as pushed (AoS, what AtomGroup.positions yields)
//   fix   u32 (F, 3, N)        SoA, position / L0 as a 32-bit binary fraction of the box   [lattice, FP32 mode]
//   posd  f64 (F, 3, N)        SoA, exact widening of the float32 coordinates             [explicit q / FP64 mode]
//   part  f64x2 (F, n_sets, P, Nq)  per-particle-chunk partial densities (fixed-order reduce => bit-reproducible)
//   ssf   f64 (n_pairs, Nq)    accumulator over all frames pushed
//
// Roofline: one (wavevector, particle) term costs two MUFU ops
// (sin, cos); 16 MUFU lanes per SM per clock => 8 terms/clk/SM.  HBM traffic is
// 12 B per particle per frame and irrelevant.
#include "common.cuh"
#include "sq_math.cuh"
#include "sq_nufft.cuh"

#include <math.h>
#include <string.h>
#include <vector>

using namespace mdc;

namespace {

constexpr int SQ_Q = 8;          // consecutive c (z index) values per thread in the lattice kernel
constexpr int SQ_THREADS = 128;
constexpr int SQ_TILE = 256;     // particles staged in shared memory per inner pass (one exact integer partial sum)
constexpr double TWO_PI = 6.283185307179586;  // == 2 * np.pi


// ------------------------------------------------------------------------------------------
// wavevector grid (structure.py:1837-1841, 1874-1887)
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ double grid_value(int i, double L) { return __ddiv_rn(__dmul_rn(TWO_PI, (double)i), L); }

__device__ __forceinline__ double norm3(double x, double y, double z)
{
    // np.linalg.norm(axis=1): sqrt(((x*x + y*y) + z*z)), no FMA
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// row r = b * n + a  (np.meshgrid 'xy' indexing: flat index b*n^2 + a*n + c, q = (g[a], g[b], g[c]))
__global__ void lat_row_counts(int n, double Lx, double Ly, double Lz, double q_max, int *cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n * n) return;
    const int a = r % n, b = r / n;
    const double qx = grid_value(a, Lx), qy = grid_value(b, Ly);
    int k = 0;
    if (q_max < 0) {
        k = n;
    } else {
        // |q| is non-decreasing in c, so the kept c form a prefix
        for (int c = 0; c < n; ++c) {
            if (norm3(qx, qy, grid_value(c, Lz)) <= q_max) ++k;
            else break;
        }
    }
    cnt[r] = k;
}

// items: {a | b << 16, c0 | count << 16, first wavevector index (low 32), (high 32)}
__global__ void lat_fill(int n, double Lx, double Ly, double Lz, const int *cnt, const int64_t *q_off,
                         const int64_t *item_off, double *q, uint4 *items)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n * n) return;
    const int a = r % n, b = r / n;
    const double qx = grid_value(a, Lx), qy = grid_value(b, Ly);
    const int k = cnt[r];
    const int64_t q0 = q_off[r];
    for (int c = 0; c < k; ++c) {
        q[3 * (q0 + c) + 0] = qx;
        q[3 * (q0 + c) + 1] = qy;
        q[3 * (q0 + c) + 2] = grid_value(c, Lz);
    }
    int64_t it = item_off[r];
    for (int c0 = 0; c0 < k; c0 += SQ_Q, ++it) {
        const int m = min(SQ_Q, k - c0);
        const int64_t qi = q0 + c0;
        items[it] = make_uint4((unsigned)a | ((unsigned)b << 16), (unsigned)c0 | ((unsigned)m << 16),
                               (unsigned)(qi & 0xffffffffll), (unsigned)(qi >> 32));
    }
}

// ------------------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------------------

// raw f32 / f64 (F, N, 3) -> fix u32 (F, 3, N): frac(x / L0) * 2^32, rounded to nearest (wraps mod 2^32)
template <typename T>
__global__ void prep_fix(const T *__restrict__ raw, int64_t N, int64_t total, double iLx, double iLy,
                         double iLz, uint32_t *__restrict__ fix)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t f = i / N, j = i - f * N;
    const T *p = raw + 3 * i;
    const double inv[3] = {iLx, iLy, iLz};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double s = (double)p[k] * inv[k];
        s -= floor(s);
        fix[(f * 3 + k) * N + j] = (uint32_t)(unsigned long long)(s * 4294967296.0 + 0.5);
    }
}

// raw f32 / f64 (F, N, 3) -> posd f64 (F, 3, N)
template <typename T>
__global__ void prep_f64(const T *__restrict__ raw, int64_t N, int64_t total, double *__restrict__ posd)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t f = i / N, j = i - f * N;
#pragma unroll
    for (int k = 0; k < 3; ++k) posd[(f * 3 + k) * N + j] = (double)raw[3 * i + k];
}

// f64 (N, 3) -> f64 (3, N)
__global__ void transpose_f64(const double *__restrict__ rs, int64_t N, double *__restrict__ posd)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) posd[k * N + i] = rs[3 * i + k];
}

// ------------------------------------------------------------------------------------------
// density kernels: part[((f * n_sets + s) * P + p) * nq + q] = sum over the particle chunk
// ------------------------------------------------------------------------------------------

struct Chunk {
    int64_t lo, hi;
};

__device__ __forceinline__ Chunk chunk_range(int2 set, int p, int P)
{
    const int64_t lo = set.x, len = (int64_t)set.y - set.x;
    Chunk c;
    c.lo = lo + len * p / P;
    c.hi = lo + len * (p + 1) / P;
    return c;
}

// DIRECT, reciprocal lattice, FP32 + SFU.  One thread = SQ_Q consecutive c of one (a, b) row;
// particles are broadcast from shared memory.  Phase arithmetic is exact (integers mod 2^32).
__global__ void __launch_bounds__(SQ_THREADS)
sq_lattice_direct(const uint4 *__restrict__ items, int64_t n_items, const uint32_t *__restrict__ fix,
                  int64_t N, const int2 *__restrict__ sets, int n_sets, int P, int64_t nq,
                  double2 *__restrict__ part)
{
    __shared__ uint4 sm[SQ_TILE];
    const int64_t it = (int64_t)blockIdx.x * SQ_THREADS + threadIdx.x;
    const bool valid = it < n_items;
    const uint4 item = items[valid ? it : n_items - 1];
    const uint32_t a = item.x & 0xffffu, b = item.x >> 16, c0 = item.y & 0xffffu;
    const int cnt = (int)(item.y >> 16);
    const int p = blockIdx.y;
    const int f = blockIdx.z / n_sets, s = blockIdx.z - f * n_sets;
    const Chunk ch = chunk_range(sets[s], p, P);
    const uint32_t *fx = fix + ((int64_t)f * 3 + 0) * N;
    const uint32_t *fy = fix + ((int64_t)f * 3 + 1) * N;
    const uint32_t *fz = fix + ((int64_t)f * 3 + 2) * N;

    double dre[SQ_Q], dim[SQ_Q];
#pragma unroll
    for (int k = 0; k < SQ_Q; ++k) dre[k] = dim[k] = 0.0;

    for (int64_t t0 = ch.lo; t0 < ch.hi; t0 += SQ_TILE) {
        const int n = (int)min((int64_t)SQ_TILE, ch.hi - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SQ_THREADS)
            sm[i] = make_uint4(fx[t0 + i], fy[t0 + i], fz[t0 + i], 0u);
        __syncthreads();

        // Partial sums are accumulated EXACTLY: adding 3.0f snaps a term in [-1, 1] onto the 2^-22 grid of
        // the binade [2, 4] (one rounding, +-2^-23, unbiased), after which the float's bit pattern is an
        // integer count of grid steps that an integer add accumulates without error.  A plain FP32 running
        // sum would lose ~eps * |partial sum| per add, which for N = 1e5 exceeds the SFU's own error.
        int re[SQ_Q], im[SQ_Q];
#pragma unroll
        for (int k = 0; k < SQ_Q; ++k) re[k] = im[k] = 0;
#pragma unroll 2
        for (int j = 0; j < n; ++j) {
            const uint4 r = sm[j];
            uint32_t ph = a * r.x + b * r.y + c0 * r.z;
#pragma unroll
            for (int k = 0; k < SQ_Q; ++k) {
                float sn, cs;
                sincos_phase(ph, sn, cs);
                re[k] += __float_as_int(cs + 3.0f) - 0x40400000;
                im[k] += __float_as_int(sn + 3.0f) - 0x40400000;
                ph += r.z;
            }
        }
#pragma unroll
        for (int k = 0; k < SQ_Q; ++k) {
            dre[k] += (double)re[k] * (1.0 / 4194304.0);
            dim[k] += (double)im[k] * (1.0 / 4194304.0);
        }
    }
    if (valid) {
        const int64_t q0 = (int64_t)item.z | ((int64_t)item.w << 32);
        double2 *out = part + ((int64_t)blockIdx.z * P + p) * nq + q0;
#pragma unroll
        for (int k = 0; k < SQ_Q; ++k)
            if (k < cnt) out[k] = make_double2(dre[k], dim[k]);
    }
}

// Arbitrary wavevectors, FP32 + SFU.  The phase is formed in FP64 (3 DFMA-class ops), and the
// magic add  t + 1.5 * 2^20  leaves frac(t) * 2^32 in the low word: the same exact 32-bit phase
// the lattice kernel uses, with no conversion instruction.
constexpr int SQX_TILE = 256;
__global__ void __launch_bounds__(SQ_THREADS)
sq_explicit_f32(const double *__restrict__ q, int64_t q_first, int64_t n_q, const double *__restrict__ posd,
                int64_t N, const int2 *__restrict__ sets, int n_sets, int P, int64_t nq,
                double2 *__restrict__ part)
{
    __shared__ double sx[SQX_TILE], sy[SQX_TILE], sz[SQX_TILE];
    const int64_t iq = (int64_t)blockIdx.x * SQ_THREADS + threadIdx.x;
    const bool valid = iq < n_q;
    const int64_t qi = q_first + (valid ? iq : n_q - 1);
    const double inv2pi = 0.15915494309189535;
    const double qx = q[3 * qi] * inv2pi, qy = q[3 * qi + 1] * inv2pi, qz = q[3 * qi + 2] * inv2pi;
    const int p = blockIdx.y;
    const int f = blockIdx.z / n_sets, s = blockIdx.z - f * n_sets;
    const Chunk ch = chunk_range(sets[s], p, P);
    const double *px = posd + ((int64_t)f * 3 + 0) * N;
    const double *py = posd + ((int64_t)f * 3 + 1) * N;
    const double *pz = posd + ((int64_t)f * 3 + 2) * N;
    double dre = 0.0, dim = 0.0;
    for (int64_t t0 = ch.lo; t0 < ch.hi; t0 += SQX_TILE) {
        const int n = (int)min((int64_t)SQX_TILE, ch.hi - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SQ_THREADS) {
            sx[i] = px[t0 + i];
            sy[i] = py[t0 + i];
            sz[i] = pz[t0 + i];
        }
        __syncthreads();
        int re = 0, im = 0;   // exact integer partial sums on the 2^-22 grid (see sq_lattice_direct)
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const double t = fma(qz, sz[j], fma(qy, sy[j], qx * sx[j]));   // turns
            const uint32_t ph = (uint32_t)__double2loint(t + 1572864.0);    // frac(t) * 2^32
            float sn, cs;
            sincos_phase(ph, sn, cs);
            re += __float_as_int(cs + 3.0f) - 0x40400000;
            im += __float_as_int(sn + 3.0f) - 0x40400000;
        }
        dre += (double)re * (1.0 / 4194304.0);
        dim += (double)im * (1.0 / 4194304.0);
    }
    if (valid) part[((int64_t)blockIdx.z * P + p) * nq + qi] = make_double2(dre, dim);
}

// Arbitrary wavevectors, FP64 throughout: the reference's arithmetic (numba_dot + exp(1j x) in
// complex128), for MDC_PRECISION_FP64.
__global__ void __launch_bounds__(SQ_THREADS)
sq_explicit_f64(const double *__restrict__ q, int64_t q_first, int64_t n_q, const double *__restrict__ posd,
                int64_t N, const int2 *__restrict__ sets, int n_sets, int P, int64_t nq,
                double2 *__restrict__ part)
{
    __shared__ double sx[SQX_TILE], sy[SQX_TILE], sz[SQX_TILE];
    const int64_t iq = (int64_t)blockIdx.x * SQ_THREADS + threadIdx.x;
    const bool valid = iq < n_q;
    const int64_t qi = q_first + (valid ? iq : n_q - 1);
    const double qx = q[3 * qi], qy = q[3 * qi + 1], qz = q[3 * qi + 2];
    const int p = blockIdx.y;
    const int f = blockIdx.z / n_sets, s = blockIdx.z - f * n_sets;
    const Chunk ch = chunk_range(sets[s], p, P);
    const double *px = posd + ((int64_t)f * 3 + 0) * N;
    const double *py = posd + ((int64_t)f * 3 + 1) * N;
    const double *pz = posd + ((int64_t)f * 3 + 2) * N;
    double dre = 0.0, dim = 0.0;
    for (int64_t t0 = ch.lo; t0 < ch.hi; t0 += SQX_TILE) {
        const int n = (int)min((int64_t)SQX_TILE, ch.hi - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SQ_THREADS) {
            sx[i] = px[t0 + i];
            sy[i] = py[t0 + i];
            sz[i] = pz[t0 + i];
        }
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            // accelerated.py:113  a[0]*b[0] + a[1]*b[1] + a[2]*b[2]
            const double x = __dadd_rn(__dadd_rn(__dmul_rn(qx, sx[j]), __dmul_rn(qy, sy[j])), __dmul_rn(qz, sz[j]));
            double sn, cs;
            sincos(x, &sn, &cs);
            dre += cs;
            dim += sn;
        }
    }
    if (valid) part[((int64_t)blockIdx.z * P + p) * nq + qi] = make_double2(dre, dim);
}

// ------------------------------------------------------------------------------------------
// FACTORED, reciprocal lattice, FP32 pipe.  On the lattice
//     exp(i q.r) = X_j[a] * Y_j[b] * Z_j[c],     X_j[a] = exp(2 pi i a s_xj)  etc.,
// so a block that owns a patch of 32 a x 16 b x 8 c wavevectors needs only 56 unit phasors per particle
// for its 4096 terms.  The phasor tables are built once per frame from the exact integer phase with an
// FP64 sincospi (rounded once to FP32, no SFU error), streamed through shared memory, and every term is
// one complex multiply-accumulate = two packed fma.rn.f32x2:
//     acc(re, im) += (zr, zr) * (wr, wi) + (-zi, zi) * (wi, wr),       W = X[a] * Y[b].
// FP32 running sums are folded into FP64 every FC_RUN particles, so the accumulation error stays below
// the rounding of the terms themselves.  Cost: 4 FP32 FMAs per term instead of 2 SFU ops.
// ------------------------------------------------------------------------------------------

constexpr int FC_PA = 32, FC_CQ = 8;                // a block covers one 32-wide a column and 8 c
constexpr int FC_YB = 24;                          // ... and up to 24 consecutive b rows, from which it takes
constexpr int FC_THREADS = 256;                    // up to 256 (a pair, b) work items: thread = 2 adjacent a, 1 b, FC_CQ c
constexpr int FC_T = 32;                           // particles per shared-memory tile
#ifndef MDC_FC_RUN
#define MDC_FC_RUN 8
#endif
constexpr int FC_RUN = MDC_FC_RUN;                 // particles per first-level FP32 run
#ifndef MDC_FC_FOLD
#define MDC_FC_FOLD 8
#endif
constexpr int FC_FOLD = MDC_FC_FOLD;               // first-level runs per second-level FP32 sum before folding into FP64
constexpr size_t FC_SMEM = sizeof(float2) * 2 * FC_T * (FC_PA + 2 * FC_YB + FC_CQ) + sizeof(double2) * 2 * FC_CQ * FC_THREADS;

// tables: X float2 (F*N, nxp), Y float4 {yr, yi, -yi, yr} (F*N, nyp), Z float2 (F*N, nzp)
__global__ void sq_prep_tables(const uint32_t *__restrict__ fix, int64_t N, int64_t total, int n_points, int nxp,
                               int nyp, int nzp, float2 *__restrict__ tx, float4 *__restrict__ ty,
                               float2 *__restrict__ tz)
{
    const int64_t pj = blockIdx.x;   // frame * N + particle
    if (pj >= total) return;
    const int64_t f = pj / N, j = pj - f * N;
    const uint32_t s[3] = {fix[(f * 3 + 0) * N + j], fix[(f * 3 + 1) * N + j], fix[(f * 3 + 2) * N + j]};
    const int n_all = nxp + nyp + nzp;
    for (int e = threadIdx.x; e < n_all; e += blockDim.x) {
        int axis, idx;
        if (e < nxp) { axis = 0; idx = e; }
        else if (e < nxp + nyp) { axis = 1; idx = e - nxp; }
        else { axis = 2; idx = e - nxp - nyp; }
        float re = 0.f, im = 0.f;
        if (idx < n_points) {
            const uint32_t ph = (uint32_t)idx * s[axis];          // exact phase, in 2^-32 turns
            double sn, cs;
            sincospi((double)ph * (2.0 / 4294967296.0), &sn, &cs);
            re = (float)cs;
            im = (float)sn;
        }
        if (axis == 0) tx[pj * nxp + idx] = make_float2(re, im);
        else if (axis == 1) ty[pj * nyp + idx] = make_float4(re, im, -im, re);
        else tz[pj * nzp + idx] = make_float2(re, im);
    }
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a, unsigned long long b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

__global__ void __launch_bounds__(FC_THREADS, 2)
sq_lattice_factored(const uint4 *__restrict__ patches, const unsigned short *__restrict__ entries,
                    const float2 *__restrict__ tabx,
                    const float4 *__restrict__ taby, const float2 *__restrict__ tabz, int nxp, int nyp, int nzp,
                    int64_t N, const int2 *__restrict__ sets, int n_sets, int P, int n_points,
                    const int *__restrict__ row_cnt, const int64_t *__restrict__ row_qoff, int64_t nq,
                    double2 *__restrict__ part)
{
    // shared-memory tiles (double buffered, filled with cp.async): X [FC_T][FC_PA], Y [FC_T][FC_YB],
    // Z [FC_T][FC_CQ] unit phasors (float2 = re, im); and this block's FP64 accumulators
    extern __shared__ __align__(16) unsigned char fc_smem[];
    typedef float2 (*TileX)[FC_T][FC_PA];
    typedef float4 (*TileY)[FC_T][FC_YB];
    typedef float2 (*TileZ)[FC_T][FC_CQ];
    TileX sX = reinterpret_cast<TileX>(fc_smem);
    TileY sY = reinterpret_cast<TileY>(fc_smem + sizeof(float2) * 2 * FC_T * FC_PA);
    TileZ sZ = reinterpret_cast<TileZ>(fc_smem + sizeof(float2) * 2 * FC_T * (FC_PA + 2 * FC_YB));
    double2 *sD = reinterpret_cast<double2 *>(fc_smem + sizeof(float2) * 2 * FC_T * (FC_PA + 2 * FC_YB + FC_CQ));
    const uint4 patch = patches[blockIdx.x];
    // block = {a0 | b0 << 16, c0 | n_items << 16, first entry}; entry = (a pair index in the column) | (b - b0) << 8.
    // Work items are the (a pair, b) rows that still hold kept wavevectors at c0, packed densely, so that blocks
    // on the rim of the q_max sphere are as full as those inside it.
    const int a0 = (int)(patch.x & 0xffffu), b0 = (int)(patch.x >> 16), c0 = (int)(patch.y & 0xffffu);
    const int n_work = (int)(patch.y >> 16);
    const bool active = (int)threadIdx.x < n_work;
    const unsigned entry = entries[patch.z + (active ? threadIdx.x : 0)];
    const int tx = (int)(entry & 0xffu), ty = (int)(entry >> 8);
    const int p = blockIdx.y;
    const int f = blockIdx.z / n_sets, s = blockIdx.z - f * n_sets;
    const Chunk ch = chunk_range(sets[s], p, P);

    // tile loader, 16-byte pieces: X = FC_T * 16 pieces (two per thread), Y = FC_T * 24 (three per thread),
    // Z = FC_T * 4 (threads 0..127).  Particles past the end of the chunk are zero-filled (src-size 0).
    static_assert(FC_THREADS == 256 && FC_T == 32 && FC_PA == 32 && FC_YB == 24 && FC_CQ == 8, "loader layout");
    auto cp16 = [](void *dst, const void *src, bool valid) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
        const int n = valid ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
    };
    auto fetch = [&](int64_t t0, int buf) {
        const int64_t base = (int64_t)f * N;
        {
            const int jj = threadIdx.x >> 4, w = threadIdx.x & 15;
            const int64_t j = t0 + jj, j2 = j + 16;
            cp16(&sX[buf][jj][2 * w], tabx + (base + min(j, ch.hi - 1)) * nxp + a0 + 2 * w, j < ch.hi);
            cp16(&sX[buf][jj + 16][2 * w], tabx + (base + min(j2, ch.hi - 1)) * nxp + a0 + 2 * w, j2 < ch.hi);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int piece = threadIdx.x + r * FC_THREADS, jj = piece / FC_YB, w = piece - jj * FC_YB;
            const int64_t j = t0 + jj;
            cp16(&sY[buf][jj][w], taby + (base + min(j, ch.hi - 1)) * nyp + b0 + w, j < ch.hi);
        }
        if (threadIdx.x < 128) {
            const int jj = threadIdx.x >> 2, w = threadIdx.x & 3;
            const int64_t j = t0 + jj;
            cp16(&sZ[buf][jj][2 * w], tabz + (base + min(j, ch.hi - 1)) * nzp + c0 + 2 * w, j < ch.hi);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Three accumulation levels keep the rounding of the sums below that of the terms:
    //   acc1 (FP32, registers): FC_RUN = 8 particles;  acc2 (FP32, registers): FC_FOLD runs;
    //   sD (FP64, shared memory, [accumulator][thread]): everything else.
#pragma unroll
    for (int k = 0; k < 2 * FC_CQ; ++k) sD[k * FC_THREADS + threadIdx.x] = make_double2(0.0, 0.0);
    unsigned long long acc2[2][FC_CQ];
#pragma unroll
    for (int c = 0; c < FC_CQ; ++c) acc2[0][c] = acc2[1][c] = 0ull;
    auto fold = [&]() {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < FC_CQ; ++c) {
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2[h][c]));
                double2 d = sD[(h * FC_CQ + c) * FC_THREADS + threadIdx.x];
                d.x += (double)lo;
                d.y += (double)hi;
                sD[(h * FC_CQ + c) * FC_THREADS + threadIdx.x] = d;
                acc2[h][c] = 0ull;
            }
    };

    int buf = 0, runs = 0;
    if (ch.lo < ch.hi) fetch(ch.lo, 0);
    for (int64_t t0 = ch.lo; t0 < ch.hi; t0 += FC_T, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                      // tile `buf` landed; everyone is done with `buf ^ 1`
        if (t0 + FC_T < ch.hi) fetch(t0 + FC_T, buf ^ 1);     // next tile's copies fly during this tile's math
        // software pipeline: the phasors of particle jj + 1 are loaded while particle jj is consumed
        float4 x = *reinterpret_cast<const float4 *>(&sX[buf][0][2 * tx]);
        float4 y = sY[buf][0][ty];
        float4 z[FC_CQ / 2];
#pragma unroll
        for (int c = 0; c < FC_CQ / 2; ++c) z[c] = *reinterpret_cast<const float4 *>(&sZ[buf][0][2 * c]);
#pragma unroll 1
        for (int r0 = 0; r0 < FC_T; r0 += FC_RUN) {
            unsigned long long acc[2][FC_CQ];
#pragma unroll
            for (int c = 0; c < FC_CQ; ++c) acc[0][c] = acc[1][c] = 0ull;
#pragma unroll
            for (int jj = r0; jj < r0 + FC_RUN; ++jj) {
                const int jn = (jj + 1) & (FC_T - 1);          // the last particle re-reads row 0 (unused)
                const float4 xn = *reinterpret_cast<const float4 *>(&sX[buf][jn][2 * tx]);
                const float4 yn = sY[buf][jn][ty];
                // With Yp = (yr, yi) and Yq = (-yi, yr) from the table:
                //   W = X * Y       = x.re * Yp + x.im * Yq        V = (-wi, wr) = x.re * Yq - x.im * Yp
                //   acc(re, im)    += zr * W + zi * V              (two packed FMAs per term)
                // Every FFMA2 operand is a scalar (broadcast by the instruction) or a register pair that a
                // packed instruction or a 128-bit load produced: no register shuffling.
                const unsigned long long yp = pack2(y.x, y.y), yq = pack2(y.z, y.w);
                unsigned long long w0, w1, v0, v1;
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(w0) : "l"(pack2(x.x, x.x)), "l"(yp));
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(w1) : "l"(pack2(x.z, x.z)), "l"(yp));
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v0) : "l"(pack2(x.x, x.x)), "l"(yq));
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(v1) : "l"(pack2(x.z, x.z)), "l"(yq));
                fma2(w0, pack2(x.y, x.y), yq);
                fma2(w1, pack2(x.w, x.w), yq);
                fma2(v0, pack2(-x.y, -x.y), yp);
                fma2(v1, pack2(-x.w, -x.w), yp);
#pragma unroll
                for (int c = 0; c < FC_CQ; c += 2) {
                    const float4 zz = z[c / 2];
                    z[c / 2] = *reinterpret_cast<const float4 *>(&sZ[buf][jn][c]);
                    const unsigned long long zr0 = pack2(zz.x, zz.x), zi0 = pack2(zz.y, zz.y);
                    const unsigned long long zr1 = pack2(zz.z, zz.z), zi1 = pack2(zz.w, zz.w);
                    fma2(acc[0][c], zr0, w0);
                    fma2(acc[1][c], zr0, w1);
                    fma2(acc[0][c + 1], zr1, w0);
                    fma2(acc[1][c + 1], zr1, w1);
                    fma2(acc[0][c], zi0, v0);
                    fma2(acc[1][c], zi0, v1);
                    fma2(acc[0][c + 1], zi1, v0);
                    fma2(acc[1][c + 1], zi1, v1);
                }
                x = xn;
                y = yn;
            }
#pragma unroll
            for (int c = 0; c < FC_CQ; ++c) {
                asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc2[0][c]) : "l"(acc[0][c]));
                asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc2[1][c]) : "l"(acc[1][c]));
            }
            if (++runs == FC_FOLD) {
                fold();
                runs = 0;
            }
        }
    }
    fold();
    double dre[2][FC_CQ], dim[2][FC_CQ];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < FC_CQ; ++c) {
            const double2 d = sD[(h * FC_CQ + c) * FC_THREADS + threadIdx.x];
            dre[h][c] = d.x;
            dim[h][c] = d.y;
        }
    const int b = b0 + ty;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int a = a0 + 2 * tx + h;
        if (active && a < n_points && b < n_points) {
            const int row = b * n_points + a;
            const int cnt = row_cnt[row];
            double2 *out = part + ((int64_t)blockIdx.z * P + p) * nq + row_qoff[row];
#pragma unroll
            for (int c = 0; c < FC_CQ; ++c)
                if (c0 + c < cnt) out[c0 + c] = make_double2(dre[h][c], dim[h][c]);
        }
    }
}

// rho[f, s, q] = sum_p part[f, s, p, q] (fixed order) - N_s * dc ;  ssf[pair, q] += products over frames.
// dc = (mean cos error, mean sin error) of the SFU pipeline over all phases: the zeroth Fourier
// coefficient of its error function is the only part a sum over N particles multiplies by N, so it is
// measured once per plan (probe_sincos_kernel) and removed analytically.  Zero in FP64 mode.
__global__ void sq_pair_accum(double2 *__restrict__ part, int F, int n_sets, int P, int64_t nq,
                              const int2 *__restrict__ sets, const int2 *__restrict__ pairs, int n_pairs,
                              double dc_cos, double dc_sin, int64_t dc_first, int64_t p1_below, int sum_pairs,
                              double *__restrict__ ssf)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    if (q < dc_first) dc_cos = dc_sin = 0.0;   // wavevectors evaluated without the SFU
    const int Pq = q < p1_below ? 1 : P;       // ... and as ONE partial (NUFFT): only chunk 0 is written
    double total = 0.0;
    for (int fs = 0; fs < F * n_sets; ++fs) {
        double2 *base = part + (int64_t)fs * P * nq + q;
        double re = 0.0, im = 0.0;
        for (int p = 0; p < Pq; ++p) {
            const double2 v = base[(int64_t)p * nq];
            re += v.x;
            im += v.y;
        }
        const int2 set = sets[fs % n_sets];
        const double n = (double)(set.y - set.x);
        base[0] = make_double2(re - n * dc_cos, im - n * dc_sin);
    }
    for (int pr = 0; pr < n_pairs; ++pr) {
        const int2 jk = pairs[pr];
        double acc = 0.0;
        for (int f = 0; f < F; ++f) {
            const double2 rj = part[((int64_t)f * n_sets + jk.x) * P * nq + q];
            if (jk.x == jk.y) {
                acc += rj.x * rj.x + rj.y * rj.y;   // (rho * conj(rho)).real
            } else {
                const double2 rk = part[((int64_t)f * n_sets + jk.y) * P * nq + q];
                acc += 2.0 * (rj.x * rk.x + rj.y * rk.y);   // 2 * (rho_j * conj(rho_k)).real
            }
        }
        if (sum_pairs) total += acc;
        else ssf[(int64_t)pr * nq + q] += acc;
    }
    if (sum_pairs) ssf[q] += total;
}

// |q|-shell averages (structure.py:2000-2009): out[row, s] = mean over the wavevectors m of shell s of ssf[row, m] / denom.
// The members of a shell are a window [lo, hi) of `order` (wavevector indices sorted by |q|); windows may overlap,
// as np.isclose's do.  One warp per (row, shell), lane-strided partial sums folded in a fixed order: deterministic.
__global__ void sq_shell_mean(const double *__restrict__ ssf, int64_t nq, const int64_t *__restrict__ order,
                              const int64_t *__restrict__ lo, const int64_t *__restrict__ hi, int64_t n_shells,
                              int n_rows, double denom, double *__restrict__ out)
{
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_shells * n_rows) return;
    const int64_t row = warp / n_shells, s = warp - row * n_shells;
    const int64_t a = lo[s], b = hi[s];
    double sum = 0.0;
    for (int64_t m = a + lane; m < b; m += 32) sum += ssf[row * nq + order[m]] / denom;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if (lane == 0) out[row * n_shells + s] = sum / (double)(b - a);   // empty shell: 0 / 0 = NaN, as np.mean([])
}

// ------------------------------------------------------------------------------------------
// intermediate scattering function (structure.py:2198-2835): a ring of the last n_lags densities and,
// for the self part, of the last n_lags position sets
// ------------------------------------------------------------------------------------------

// rho[(f * n_sets + s) * nq + q] = sum_p part[f, s, p, q] - N_s * dc   (same reduction as sq_pair_accum)
__global__ void sq_reduce_rho(const double2 *__restrict__ part, int F, int n_sets, int P, int64_t nq,
                              const int2 *__restrict__ sets, double dc_cos, double dc_sin, int64_t dc_first,
                              int64_t p1_below, double2 *__restrict__ rho)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    if (q < dc_first) dc_cos = dc_sin = 0.0;
    const int Pq = q < p1_below ? 1 : P;
    for (int fs = 0; fs < F * n_sets; ++fs) {
        const double2 *base = part + (int64_t)fs * P * nq + q;
        double re = 0.0, im = 0.0;
        for (int p = 0; p < Pq; ++p) {
            const double2 v = base[(int64_t)p * nq];
            re += v.x;
            im += v.y;
        }
        const int2 set = sets[fs % n_sets];
        const double n = (double)(set.y - set.x);
        rho[(int64_t)fs * nq + q] = make_double2(re - n * dc_cos, im - n * dc_sin);
    }
}

// cisf[lag, pair, q] += Re(rho_j(t - lag) conj rho_k(t)) (+ the j <-> k term for cross pairs), structure.py:2694-2731
__global__ void isf_coherent(const double2 *__restrict__ ring, int n_lags, int n_sets, int64_t nq, int cur,
                             int n_valid, const int2 *__restrict__ pairs, int n_pairs, double *__restrict__ cisf)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const double2 *now = ring + (int64_t)cur * n_sets * nq + q;
    for (int lag = 0; lag < n_valid; ++lag) {
        int s = cur - lag;
        if (s < 0) s += n_lags;
        const double2 *then = ring + (int64_t)s * n_sets * nq + q;
        for (int pr = 0; pr < n_pairs; ++pr) {
            const int2 jk = pairs[pr];
            const double2 aj = then[(int64_t)jk.x * nq], bk = now[(int64_t)jk.y * nq];
            double v = aj.x * bk.x + aj.y * bk.y;
            if (jk.x != jk.y) {
                const double2 ak = then[(int64_t)jk.y * nq], bj = now[(int64_t)jk.x * nq];
                v += ak.x * bj.x + ak.y * bj.y;
            }
            cisf[((int64_t)lag * n_pairs + pr) * nq + q] += v;
        }
    }
}

// displacement "positions" for L consecutive lags: r(t) - r(t - lag).  On the lattice the difference of the
// 32-bit box fractions (mod 2^32) is exact and wrapping is irrelevant because q is commensurate with the box.
__global__ void isf_disp_fix(const uint32_t *__restrict__ ring, int n_lags, int64_t N3, int cur, int lag0, int L,
                             uint32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N3 * L) return;
    const int l = (int)(i / N3);
    const int64_t j = i - (int64_t)l * N3;
    int s = cur - (lag0 + l);
    if (s < 0) s += n_lags;
    out[i] = ring[(int64_t)cur * N3 + j] - ring[(int64_t)s * N3 + j];
}

__global__ void isf_disp_f64(const double *__restrict__ ring, int n_lags, int64_t N3, int cur, int lag0, int L,
                             double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N3 * L) return;
    const int l = (int)(i / N3);
    const int64_t j = i - (int64_t)l * N3;
    int s = cur - (lag0 + l);
    if (s < 0) s += n_lags;
    out[i] = ring[(int64_t)cur * N3 + j] - ring[(int64_t)s * N3 + j];
}

// iisf[lag0 + l, slot(s), q] += Re sum_j exp(i q . (r_j(t) - r_j(t - lag)))   (structure.py:2699-2725)
__global__ void isf_incoherent_accum(const double2 *__restrict__ part, int L, int n_sets, int P, int64_t nq,
                                     const int2 *__restrict__ sets, const int *__restrict__ inc_slot, int n_inc,
                                     int lag0, double dc_cos, int64_t dc_first, int64_t p1_below,
                                     double *__restrict__ iisf)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    if (q < dc_first) dc_cos = 0.0;
    const int Pq = q < p1_below ? 1 : P;
    for (int l = 0; l < L; ++l)
        for (int s = 0; s < n_sets; ++s) {
            const int slot = inc_slot[s];
            if (slot < 0) continue;
            const double2 *base = part + ((int64_t)l * n_sets + s) * P * nq + q;
            double re = 0.0;
            for (int p = 0; p < Pq; ++p) re += base[(int64_t)p * nq].x;
            const int2 set = sets[s];
            iisf[((int64_t)(lag0 + l) * n_inc + slot) * nq + q] += re - (double)(set.y - set.x) * dc_cos;
        }
}

// ------------------------------------------------------------------------------------------
// accuracy probe of sincos_phase over all 2^23 distinct phases
// ------------------------------------------------------------------------------------------
__global__ void probe_sincos_kernel(int variant, double *acc /* [sum ec, sum es, sum e^2, max|e| bits] */)
{
    __shared__ double sh[4][256];
    double sc = 0, ss = 0, s2 = 0, mx = 0;
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < (1u << 23); m += gridDim.x * blockDim.x) {
        float sn, cs;
        if (variant == 0) {
            sincos_phase(m << 9, sn, cs);
        } else {
            // single-constant variant, for the record: x = t * float(2 pi)
            const float t = __uint_as_float(0x3f800000u | m);
            const float x = t * 6.2831854820251465f;
            asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(x));
            asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(x));
        }
        double rs, rc;
        sincospi((double)m * (2.0 / 8388608.0), &rs, &rc);
        const double ec = (double)cs - rc, es = (double)sn - rs;
        sc += ec;
        ss += es;
        s2 += ec * ec + es * es;
        mx = fmax(mx, fmax(fabs(ec), fabs(es)));
    }
    sh[0][threadIdx.x] = sc;
    sh[1][threadIdx.x] = ss;
    sh[2][threadIdx.x] = s2;
    sh[3][threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
            sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
            sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(&acc[0], sh[0][0]);
        atomicAdd(&acc[1], sh[1][0]);
        atomicAdd(&acc[2], sh[2][0]);
        // max of non-negative doubles == max of their bit patterns
        atomicMax((unsigned long long *)&acc[3], (unsigned long long)__double_as_longlong(sh[3][0]));
    }
}

int run_probe_sincos(int variant, double out[4])
{
    DevBuf<double> acc;
    MDC_CHECK(acc.alloc(4));
    MDC_CUDA(cudaMemset(acc.p, 0, 4 * sizeof(double)));
    probe_sincos_kernel<<<1024, 256>>>(variant, acc.p);
    MDC_CUDA(cudaGetLastError());
    double h[4];
    MDC_CUDA(cudaMemcpy(h, acc.p, sizeof(h), cudaMemcpyDeviceToHost));
    acc.release();
    const double n = 8388608.0;
    out[0] = h[0] / n;
    out[1] = h[1] / n;
    out[2] = sqrt(h[2] / (2 * n));
    out[3] = h[3];
    return MDC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------

struct mdc_sq_plan {
    mdc_ctx *ctx = nullptr;
    int n_points = 0;
    double L0[3] = {0, 0, 0};
    int64_t n_lat = 0, n_exp = 0, nq = 0, n_items = 0;
    int64_t N = 0;
    int n_sets = 0, n_pairs = 0;
    int precision = 0, algo = 0;
    bool lattice_fast = false;   // lattice part runs through an FP32 lattice kernel (direct or factored)
    bool factored = false;       // ... the factored one
    bool nufft_on = false;       // lattice part runs through the gridding NUFFT (sq_nufft.cu), either precision
    Nufft nufft;
    double dc[2] = {0, 0};
    DevBuf<double> q, ssf, posd[2];
    DevBuf<uint4> items, patches;
    DevBuf<unsigned short> entries;
    DevBuf<int> row_cnt;
    DevBuf<int64_t> row_qoff;
    DevBuf<float2> tabx[2], tabz[2];
    DevBuf<float4> taby[2];

    int64_t n_patches = 0;
    int nxp = 0, nyp = 0, nzp = 0;
    // intermediate scattering function (mdc_isf_*)
    bool isf = false, incoherent = false;
    int n_lags = 0, n_inc = 0;
    DevBuf<double2> rho_ring;                  // (n_lags, n_sets, nq)
    DevBuf<uint32_t> fix_ring, dfix;           // (n_lags, 3, N), (fblk, 3, N)
    DevBuf<double> posd_ring, dposd;           // same shapes, explicit wavevectors / FP64 mode
    DevBuf<float2> dtabx, dtabz;
    DevBuf<float4> dtaby;
    DevBuf<double> cisf, iisf;                 // (n_lags, n_pairs, nq), (n_lags, n_inc, nq)
    DevBuf<int> inc_slot;                      // set -> row of iisf (-1: none)
    std::vector<int2> pair_sets;
    int64_t isf_in_push = 0;                   // frames of the current push already processed
    int nufft_rc = MDC_OK;                     // status of the last Nufft::density call
    bool sum_pairs = false;                    // all pairs accumulate into row 0 (single-chain structure factor)
    int n_rows = 0;                            // rows of ssf: n_pairs, or 1
    DevBuf<double2> part;
    DevBuf<int64_t> shell_order, shell_lo, shell_hi;   // |q|-shell windows (mdc_sq_set_shells)
    DevBuf<double> shell_out;
    int64_t n_shells = 0;
    DevBuf<float> raw[2];
    DevBuf<double> raw64[2];                           // staging of mdc_sq_push_f64 (allocated on first use)
    DevBuf<uint32_t> fix[2];
    DevBuf<int2> d_sets, d_pairs;
    std::vector<int2> sets;
    Streams st;
    int fblk = 0, P = 1;
    int64_t n_frames = 0;
    float last_ms = 0.f;
    int last_launches = 0;
};

namespace {

void plan_free(mdc_sq_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();
    p->q.release();
    p->ssf.release();
    p->items.release();
    p->patches.release();
    p->entries.release();
    p->rho_ring.release();
    p->fix_ring.release();
    p->dfix.release();
    p->posd_ring.release();
    p->dposd.release();
    p->dtabx.release();
    p->dtaby.release();
    p->dtabz.release();
    p->cisf.release();
    p->iisf.release();
    p->inc_slot.release();
    p->shell_order.release();
    p->shell_lo.release();
    p->shell_hi.release();
    p->shell_out.release();
    p->nufft.release();
    p->row_cnt.release();
    p->row_qoff.release();
    p->part.release();
    p->d_sets.release();
    p->d_pairs.release();
    for (int i = 0; i < 2; ++i) {
        p->posd[i].release();
        p->raw[i].release();
        p->raw64[i].release();
        p->fix[i].release();
        p->tabx[i].release();
        p->taby[i].release();
        p->tabz[i].release();
    }
    p->st.destroy();
    delete p;
}

}  // namespace

static int sq_plan_create_impl(mdc_ctx *ctx, int n_points, const double L0[3], double q_max,
                               const double *q_explicit, int64_t n_explicit, int n_groups,
                               const int64_t *group_sizes, int n_pairs, const int *pairs, int precision, int algo,
                               bool sum_pairs, mdc_sq_plan **out)
{
    if (!ctx || !out) return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: NULL argument");
    *out = nullptr;
    if (n_points < 0 || n_points > 65535) return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: n_points out of range");
    if (n_points > 0 && (!L0 || !(L0[0] > 0) || !(L0[1] > 0) || !(L0[2] > 0)))
        return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: lattice plans need positive L0");
    if (n_explicit < 0 || (n_explicit > 0 && !q_explicit))
        return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: bad explicit wavevectors");
    if (n_points == 0 && n_explicit == 0) return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: no wavevectors");
    if (n_groups < 1 || !group_sizes) return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: need at least one group");
    if (n_pairs < 1 || !pairs) return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: need at least one pair");
    if (precision != MDC_PRECISION_FP32 && precision != MDC_PRECISION_FP64)
        return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: unknown precision");
    if (algo != MDC_SQ_ALGO_AUTO && algo != MDC_SQ_ALGO_DIRECT && algo != MDC_SQ_ALGO_FACTORED &&
        algo != MDC_SQ_ALGO_NUFFT)
        return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: unknown algorithm");
    MDC_CUDA(cudaSetDevice(ctx->device));

    mdc_sq_plan *p = new mdc_sq_plan();
    p->ctx = ctx;
    p->n_points = n_points;
    p->precision = precision;
    p->algo = algo;
    p->n_exp = n_explicit;
    if (n_points > 0) memcpy(p->L0, L0, sizeof(p->L0));

    // density sets: one per group referenced by a pair, plus "all particles" for (-1, -1)
    std::vector<int64_t> off(n_groups + 1, 0);
    for (int g = 0; g < n_groups; ++g) {
        if (group_sizes[g] < 0) {
            delete p;
            return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: negative group size");
        }
        off[g + 1] = off[g] + group_sizes[g];
    }
    p->N = off[n_groups];
    if (p->N <= 0 || p->N > 2147483647ll) {
        delete p;
        return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: total particle count must be in [1, 2^31)");
    }
    std::vector<int> set_of(n_groups + 1, -1);
    std::vector<int2> hp(n_pairs);
    for (int i = 0; i < n_pairs; ++i) {
        int jk[2] = {pairs[2 * i], pairs[2 * i + 1]};
        for (int t = 0; t < 2; ++t) {
            int g = jk[t] < 0 ? n_groups : jk[t];
            if (g > n_groups) {
                delete p;
                return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: pair index out of range");
            }
            if (set_of[g] < 0) {
                set_of[g] = (int)p->sets.size();
                p->sets.push_back(g == n_groups ? make_int2(0, (int)p->N) : make_int2((int)off[g], (int)off[g + 1]));
            }
            jk[t] = set_of[g];
        }
        hp[i] = make_int2(jk[0], jk[1]);
    }
    p->n_sets = (int)p->sets.size();
    p->n_pairs = n_pairs;
    p->pair_sets = hp;
    p->sum_pairs = sum_pairs;
    p->n_rows = sum_pairs ? 1 : n_pairs;
    if (p->n_sets > 65535) {
        delete p;
        return fail(MDC_ERR_UNSUPPORTED, "mdc_sq_plan_create: more than 65535 distinct groups");
    }

    int rc = p->st.init(ctx);
    if (rc != MDC_OK) {
        plan_free(p);
        return rc;
    }

#define PLAN_TRY(expr)              \
    do {                            \
        int _r = (expr);            \
        if (_r != MDC_OK) {         \
            plan_free(p);           \
            return _r;              \
        }                           \
    } while (0)
#define PLAN_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            plan_free(p);                                                                    \
            return fail(MDC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
        }                                                                                    \
    } while (0)

    // grid build, pass 1: kept wavevectors per (a, b) row on the device; prefix sums over the
    // n^2 rows on the host (nothing of size Nq crosses the bus)
    const int rows = n_points * n_points;
    DevBuf<int> &cnt = p->row_cnt;
    std::vector<int64_t> hq(rows), hi(rows);
    std::vector<int> h(rows);
    if (n_points > 0) {
        PLAN_TRY(cnt.alloc(rows));
        lat_row_counts<<<(rows + 255) / 256, 256>>>(n_points, p->L0[0], p->L0[1], p->L0[2], q_max, cnt.p);
        PLAN_CUDA(cudaGetLastError());
        PLAN_CUDA(cudaMemcpy(h.data(), cnt.p, rows * sizeof(int), cudaMemcpyDeviceToHost));
        for (int r = 0; r < rows; ++r) {
            hq[r] = p->n_lat;
            hi[r] = p->n_items;
            p->n_lat += h[r];
            p->n_items += (h[r] + SQ_Q - 1) / SQ_Q;
        }
    }
    p->nq = p->n_lat + p->n_exp;
    if (p->nq == 0) {
        plan_free(p);
        return fail(MDC_ERR_INVALID, "mdc_sq_plan_create: q_max removes every wavevector");
    }
    int64_t max_set = 1, sum_set = 0;
    for (auto &s : p->sets) {
        max_set = std::max<int64_t>(max_set, s.y - s.x);
        sum_set += s.y - s.x;
    }
    if (algo == MDC_SQ_ALGO_NUFFT && p->n_lat > 0 && !Nufft::supported(n_points)) {
        plan_free(p);
        return fail(MDC_ERR_UNSUPPORTED, "mdc_sq_plan_create: the NUFFT path supports 2 <= n_points <= 682");
    }
    // requested relative l2 error of the NUFFT: 1e-9 in FP32 mode (an order below the FP32 kernels), 1e-13 in FP64
    // mode.  MDC_NUFFT_EPS in the environment overrides the FP32-mode value (1e-12 .. 1e-4), e.g. 1e-7: kernel width
    // 10 instead of 12, 42 % fewer reductions, still inside the 1e-5 contract.
    double nufft_eps = precision == MDC_PRECISION_FP64 ? 1e-13 : 1e-9;
    if (precision != MDC_PRECISION_FP64) {
        const char *env = getenv("MDC_NUFFT_EPS");
        const double v = env ? atof(env) : 0.0;
        if (v >= 1e-12 && v <= 1e-4) nufft_eps = v;
    }
    if (p->n_lat > 0 && Nufft::supported(n_points)) {
        if (algo == MDC_SQ_ALGO_NUFFT) {
            p->nufft_on = true;
        } else if (algo == MDC_SQ_ALGO_AUTO) {
            // cost model (rates measured on B200): the O(Nq N) kernel of this precision (factored
            // FP32: 4.4e12 terms/s; FP64 reference arithmetic: 3.4e11 terms/s) vs fine-grid traffic + spreading
            const bool f64 = precision == MDC_PRECISION_FP64;
            NufftGeom ng;
            const double t_terms = (double)p->n_lat * (double)sum_set / (f64 ? 3.4e11 : 4.4e12);
            const double t_nufft =
                Nufft::geometry(n_points, nufft_eps, &ng) ? Nufft::cost(ng, p->n_sets, sum_set) : 1e30;
            p->nufft_on = t_nufft < t_terms;
        }
    }
    p->lattice_fast = (precision == MDC_PRECISION_FP32) && p->n_lat > 0 && !p->nufft_on;
    p->factored = p->lattice_fast && algo != MDC_SQ_ALGO_DIRECT;
    if (p->nufft_on) {
        const int rn = p->nufft.init(ctx, n_points, nufft_eps, max_set);
        if (rn != MDC_OK) {
            plan_free(p);
            return rn;
        }
    }

    PLAN_TRY(p->q.alloc((size_t)p->nq * 3));
    PLAN_TRY(p->ssf.alloc((size_t)p->nq * p->n_rows));
    PLAN_CUDA(cudaMemset(p->ssf.p, 0, (size_t)p->nq * p->n_rows * sizeof(double)));
    PLAN_TRY(p->d_sets.alloc(p->n_sets));
    PLAN_TRY(p->d_pairs.alloc(n_pairs));
    PLAN_CUDA(cudaMemcpy(p->d_sets.p, p->sets.data(), p->n_sets * sizeof(int2), cudaMemcpyHostToDevice));
    PLAN_CUDA(cudaMemcpy(p->d_pairs.p, hp.data(), n_pairs * sizeof(int2), cudaMemcpyHostToDevice));

    if (p->n_lat > 0) {
        // pass 2: wavevectors (reference FP64 operation order) and the lattice kernel's work items
        DevBuf<int64_t> &qo = p->row_qoff;
        DevBuf<int64_t> io;
        PLAN_TRY(qo.alloc(rows));
        PLAN_TRY(io.alloc(rows));
        PLAN_CUDA(cudaMemcpy(qo.p, hq.data(), rows * sizeof(int64_t), cudaMemcpyHostToDevice));
        PLAN_CUDA(cudaMemcpy(io.p, hi.data(), rows * sizeof(int64_t), cudaMemcpyHostToDevice));
        PLAN_TRY(p->items.alloc((size_t)p->n_items));
        lat_fill<<<(rows + 255) / 256, 256>>>(n_points, p->L0[0], p->L0[1], p->L0[2], cnt.p, qo.p, io.p, p->q.p,
                                              p->items.p);
        PLAN_CUDA(cudaGetLastError());
        PLAN_CUDA(cudaDeviceSynchronize());
        io.release();
        if (p->nufft_on) {
            p->items.release();   // the NUFFT scatters through row_cnt / row_qoff
        } else if (!p->factored) {
            cnt.release();
            qo.release();
        } else {
            // work blocks: for every c chunk and every 32-wide a column, runs of at most FC_YB consecutive b rows
            // whose (a pair, b) items with kept wavevectors at c0 are packed, up to FC_THREADS per block
            p->items.release();
            const int n = n_points;
            std::vector<uint4> hp;
            std::vector<unsigned short> he;
            for (int c0 = 0; c0 < n; c0 += FC_CQ)
                for (int a0 = 0; a0 < n; a0 += FC_PA) {
                    int b = 0;
                    while (b < n) {
                        const int b_start = b;
                        const size_t off = he.size();
                        int count = 0;
                        while (b < n && b - b_start < FC_YB) {
                            int m = 0;
                            for (int ap = 0; ap < FC_PA / 2; ++ap) {
                                const int a = a0 + 2 * ap;
                                if (a < n && h[b * n + a] > c0) ++m;
                            }
                            if (count + m > FC_THREADS) break;
                            for (int ap = 0; ap < FC_PA / 2; ++ap) {
                                const int a = a0 + 2 * ap;
                                if (a < n && h[b * n + a] > c0) he.push_back((unsigned short)(ap | ((b - b_start) << 8)));
                            }
                            count += m;
                            ++b;
                            if (m == 0 && count == 0) break;   // nothing kept in this row: start the next block after it
                        }
                        if (count > 0)
                            hp.push_back(make_uint4((unsigned)a0 | ((unsigned)b_start << 16),
                                                    (unsigned)c0 | ((unsigned)count << 16), (unsigned)off, 0u));
                        else
                            he.resize(off);
                    }
                }
            if (he.empty()) he.push_back(0);
            p->n_patches = (int64_t)hp.size();
            PLAN_TRY(p->patches.alloc(hp.size()));
            PLAN_CUDA(cudaMemcpy(p->patches.p, hp.data(), hp.size() * sizeof(uint4), cudaMemcpyHostToDevice));
            PLAN_TRY(p->entries.alloc(he.size()));
            PLAN_CUDA(cudaMemcpy(p->entries.p, he.data(), he.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
            PLAN_CUDA(cudaFuncSetAttribute(sq_lattice_factored, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_SMEM));
            p->nxp = (int)ceil_div(n, FC_PA) * FC_PA;
            p->nyp = (int)ceil_div(n, 8) * 8 + FC_YB;   // a block's Y slice may run past the last row: zero phasors
            p->nzp = (int)ceil_div(n, FC_CQ) * FC_CQ;
        }
    }
    if (p->n_exp > 0)
        PLAN_CUDA(cudaMemcpy(p->q.p + 3 * p->n_lat, q_explicit, (size_t)p->n_exp * 3 * sizeof(double),
                             cudaMemcpyHostToDevice));

    if (precision == MDC_PRECISION_FP32) {
        double pr[4];
        PLAN_TRY(run_probe_sincos(0, pr));
        p->dc[0] = pr[0];
        p->dc[1] = pr[1];
    }
#undef PLAN_TRY
#undef PLAN_CUDA
    *out = p;
    return MDC_OK;
}

extern "C" int mdc_sq_plan_create(mdc_ctx *ctx, int n_points, const double L0[3], double q_max,
                                  const double *q_explicit, int64_t n_explicit, int n_groups,
                                  const int64_t *group_sizes, int n_pairs, const int *pairs, int precision,
                                  int algo, mdc_sq_plan **out)
{
    return sq_plan_create_impl(ctx, n_points, L0, q_max, q_explicit, n_explicit, n_groups, group_sizes, n_pairs, pairs,
                               precision, algo, false, out);
}

// Single-chain structure factor (polymer.py:1054-1445): sum over chains of |rho_chain|^2.  Every chain is a group
// of n_monomers consecutive particles; the n_chains self pairs accumulate into ONE row.
extern "C" int mdc_scsf_plan_create(mdc_ctx *ctx, int n_points, const double L0[3], double q_max,
                                    const double *q_explicit, int64_t n_explicit, int64_t n_chains,
                                    int64_t n_monomers, int precision, int algo, mdc_sq_plan **out)
{
    if (n_chains < 1 || n_monomers < 1 || n_chains > 65535)
        return fail(MDC_ERR_INVALID, "mdc_scsf_plan_create: need 1 <= n_chains <= 65535 and n_monomers >= 1");
    std::vector<int64_t> sizes((size_t)n_chains, n_monomers);
    std::vector<int> pairs((size_t)2 * n_chains);
    for (int64_t c = 0; c < n_chains; ++c) pairs[2 * c] = pairs[2 * c + 1] = (int)c;
    return sq_plan_create_impl(ctx, n_points, L0, q_max, q_explicit, n_explicit, (int)n_chains, sizes.data(),
                               (int)n_chains, pairs.data(), precision, algo, true, out);
}

extern "C" int mdc_sq_plan_destroy(mdc_sq_plan *plan)
{
    plan_free(plan);
    return MDC_OK;
}

extern "C" int64_t mdc_sq_num_wavevectors(mdc_sq_plan *plan) { return plan ? plan->nq : -1; }

extern "C" int mdc_sq_plan_info(mdc_sq_plan *p, int info[8])
{
    if (!p || !info) return fail(MDC_ERR_INVALID, "mdc_sq_plan_info: NULL argument");
    const int lattice_algo = p->n_lat == 0 ? 0
                             : p->nufft_on ? MDC_SQ_ALGO_NUFFT
                             : p->factored ? MDC_SQ_ALGO_FACTORED
                             : p->lattice_fast ? MDC_SQ_ALGO_DIRECT : 0;
    info[0] = lattice_algo;
    info[1] = p->nufft_on ? p->nufft.g.nf : 0;
    info[2] = p->nufft_on ? p->nufft.g.w : 0;
    info[3] = p->n_points;
    info[4] = p->n_sets;
    info[5] = p->P;
    info[6] = p->fblk;
    info[7] = p->nufft_on ? p->nufft.G : 0;
    return MDC_OK;
}

extern "C" int mdc_sq_wavevectors(mdc_sq_plan *plan, double *q_out)
{
    if (!plan || !q_out) return fail(MDC_ERR_INVALID, "mdc_sq_wavevectors: NULL argument");
    MDC_CUDA(cudaSetDevice(plan->ctx->device));
    MDC_CUDA(cudaMemcpy(q_out, plan->q.p, (size_t)plan->nq * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return MDC_OK;
}

namespace {

// frames per block and particle chunks per set, sized so that the grid fills the GPU several times
// over and the partial-density buffer stays within budget
int configure(mdc_sq_plan *p, int n_frames_hint)
{
    if (p->fblk > 0) return MDC_OK;
    const int64_t budget = (int64_t)1 << 31;   // bytes for `part`
    const int64_t per_frame_min = (int64_t)p->n_sets * p->nq * sizeof(double2);
    if (per_frame_min > budget * 8)
        return fail(MDC_ERR_NOMEM, "mdc_sq_push: wavevector grid too large for the partial-density buffer");
    const int64_t blocks_q = (p->factored ? p->n_patches : ceil_div(p->lattice_fast ? p->n_items : 0, SQ_THREADS)) +
                             ceil_div((p->lattice_fast || p->nufft_on) ? p->n_exp : p->nq, SQ_THREADS);
    int fblk = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_frames_hint, 32), budget / per_frame_min));
    if (p->factored) {
        const int64_t tab_frame = p->N * ((int64_t)p->nxp * 8 + (int64_t)p->nyp * 16 + (int64_t)p->nzp * 8);
        fblk = (int)std::max<int64_t>(1, std::min<int64_t>(fblk, ((int64_t)3 << 30) / tab_frame));
    }
    fblk = (int)std::max<int64_t>(1, std::min<int64_t>(fblk, 65535 / p->n_sets));   // gridDim.z = frames x sets
    const int64_t want = (int64_t)p->ctx->sm_count * 32;
    int64_t min_len = p->N;
    for (auto &s : p->sets) min_len = std::min<int64_t>(min_len, std::max(1, s.y - s.x));
    int64_t P = blocks_q > 0 ? ceil_div(want, blocks_q * fblk * p->n_sets) : 1;   // 0: NUFFT only, one partial
    P = std::min<int64_t>(P, std::max<int64_t>(1, min_len / (2 * SQ_TILE)));
    P = std::min<int64_t>(P, std::max<int64_t>(1, budget / (per_frame_min * fblk)));
    P = std::max<int64_t>(1, std::min<int64_t>(P, 65535));
    p->fblk = fblk;
    p->P = (int)P;
    MDC_CHECK(p->part.alloc((size_t)fblk * p->n_sets * P * p->nq));
    for (int i = 0; i < 2; ++i) {
        MDC_CHECK(p->raw[i].alloc((size_t)fblk * p->N * 3));
        if (p->lattice_fast) MDC_CHECK(p->fix[i].alloc((size_t)fblk * p->N * 3));
        if (p->factored) {
            MDC_CHECK(p->tabx[i].alloc((size_t)fblk * p->N * p->nxp));
            MDC_CHECK(p->taby[i].alloc((size_t)fblk * p->N * p->nyp));
            MDC_CHECK(p->tabz[i].alloc((size_t)fblk * p->N * p->nzp));
        }
        if (!p->lattice_fast || p->n_exp > 0) MDC_CHECK(p->posd[i].alloc((size_t)fblk * p->N * 3));
    }
    return MDC_OK;
}

struct DensityIn {
    const uint32_t *fix;
    const float2 *tabx;
    const float4 *taby;
    const float2 *tabz;
    const double *posd;
};

// density kernels of F frames (or pseudo-frames) into p->part, on the compute stream
void launch_density(mdc_sq_plan *p, const DensityIn &in, int F, int *launches)
{
    cudaStream_t sk = p->st.compute;
    const unsigned gz = (unsigned)(F * p->n_sets);
    if (p->nufft_on) {
        const double iL[3] = {1.0 / p->L0[0], 1.0 / p->L0[1], 1.0 / p->L0[2]};
        p->nufft_rc = p->nufft.density(sk, in.posd, iL, p->N, p->d_sets.p, p->n_sets, F, p->row_cnt.p, p->row_qoff.p,
                                       p->nq, p->P, p->part.p, launches);
        if (p->n_exp > 0) {
            dim3 ge((unsigned)ceil_div(p->n_exp, SQ_THREADS), (unsigned)p->P, gz);
            if (p->precision == MDC_PRECISION_FP64)
                sq_explicit_f64<<<ge, SQ_THREADS, 0, sk>>>(p->q.p, p->n_lat, p->n_exp, in.posd, p->N, p->d_sets.p,
                                                          p->n_sets, p->P, p->nq, p->part.p);
            else
                sq_explicit_f32<<<ge, SQ_THREADS, 0, sk>>>(p->q.p, p->n_lat, p->n_exp, in.posd, p->N, p->d_sets.p,
                                                          p->n_sets, p->P, p->nq, p->part.p);
            ++*launches;
        }
    } else if (p->lattice_fast) {
        if (p->factored) {
            dim3 g((unsigned)p->n_patches, (unsigned)p->P, gz);
            sq_lattice_factored<<<g, FC_THREADS, FC_SMEM, sk>>>(p->patches.p, p->entries.p, in.tabx, in.taby, in.tabz,
                                                               p->nxp, p->nyp, p->nzp, p->N, p->d_sets.p, p->n_sets,
                                                               p->P, p->n_points, p->row_cnt.p, p->row_qoff.p, p->nq,
                                                               p->part.p);
        } else {
            dim3 g((unsigned)ceil_div(p->n_items, SQ_THREADS), (unsigned)p->P, gz);
            sq_lattice_direct<<<g, SQ_THREADS, 0, sk>>>(p->items.p, p->n_items, in.fix, p->N, p->d_sets.p, p->n_sets,
                                                       p->P, p->nq, p->part.p);
        }
        ++*launches;
        if (p->n_exp > 0) {
            dim3 ge((unsigned)ceil_div(p->n_exp, SQ_THREADS), (unsigned)p->P, gz);
            sq_explicit_f32<<<ge, SQ_THREADS, 0, sk>>>(p->q.p, p->n_lat, p->n_exp, in.posd, p->N, p->d_sets.p,
                                                      p->n_sets, p->P, p->nq, p->part.p);
            ++*launches;
        }
    } else {
        dim3 g((unsigned)ceil_div(p->nq, SQ_THREADS), (unsigned)p->P, gz);
        if (p->precision == MDC_PRECISION_FP64)
            sq_explicit_f64<<<g, SQ_THREADS, 0, sk>>>(p->q.p, 0, p->nq, in.posd, p->N, p->d_sets.p, p->n_sets, p->P,
                                                      p->nq, p->part.p);
        else
            sq_explicit_f32<<<g, SQ_THREADS, 0, sk>>>(p->q.p, 0, p->nq, in.posd, p->N, p->d_sets.p, p->n_sets, p->P,
                                                      p->nq, p->part.p);
        ++*launches;
    }
}

// the ISF tail of one frame whose partial densities sit in p->part (F == 1): structure.py:2668-2731
int isf_frame(mdc_sq_plan *p, int slot, int *launches)
{
    cudaStream_t sk = p->st.compute;
    const int64_t t = p->n_frames + p->isf_in_push;   // index of this frame in the run
    const int cur = (int)(t % p->n_lags);
    const int n_valid = (int)std::min<int64_t>(p->n_lags, t + 1);
    const unsigned qb = (unsigned)ceil_div(p->nq, 256);
    const int64_t dc_first = (p->factored || p->nufft_on) ? p->n_lat : 0;
    const int64_t p1_below = p->nufft_on ? p->n_lat : 0;
    sq_reduce_rho<<<qb, 256, 0, sk>>>(p->part.p, 1, p->n_sets, p->P, p->nq, p->d_sets.p, p->dc[0], p->dc[1], dc_first,
                                     p1_below, p->rho_ring.p + (size_t)cur * p->n_sets * p->nq);
    isf_coherent<<<qb, 256, 0, sk>>>(p->rho_ring.p, p->n_lags, p->n_sets, p->nq, cur, n_valid, p->d_pairs.p, p->n_pairs,
                                    p->cisf.p);
    *launches += 2;
    if (p->incoherent) {
        const int64_t N3 = 3 * p->N;
        const bool need_posd = !p->lattice_fast || p->n_exp > 0;
        if (p->lattice_fast)
            MDC_CUDA(cudaMemcpyAsync(p->fix_ring.p + (size_t)cur * N3, p->fix[slot].p, N3 * sizeof(uint32_t),
                                     cudaMemcpyDeviceToDevice, sk));
        if (need_posd)
            MDC_CUDA(cudaMemcpyAsync(p->posd_ring.p + (size_t)cur * N3, p->posd[slot].p, N3 * sizeof(double),
                                     cudaMemcpyDeviceToDevice, sk));
        for (int lag0 = 0; lag0 < n_valid; lag0 += p->fblk) {
            const int L = std::min(p->fblk, n_valid - lag0);
            const unsigned db = (unsigned)ceil_div(N3 * L, 256);
            DensityIn in = {nullptr, nullptr, nullptr, nullptr, nullptr};
            if (p->lattice_fast) {
                isf_disp_fix<<<db, 256, 0, sk>>>(p->fix_ring.p, p->n_lags, N3, cur, lag0, L, p->dfix.p);
                ++*launches;
                in.fix = p->dfix.p;
                if (p->factored) {
                    sq_prep_tables<<<(unsigned)(L * p->N), 128, 0, sk>>>(p->dfix.p, p->N, (int64_t)L * p->N, p->n_points,
                                                                        p->nxp, p->nyp, p->nzp, p->dtabx.p, p->dtaby.p,
                                                                        p->dtabz.p);
                    ++*launches;
                    in.tabx = p->dtabx.p;
                    in.taby = p->dtaby.p;
                    in.tabz = p->dtabz.p;
                }
            }
            if (need_posd) {
                isf_disp_f64<<<db, 256, 0, sk>>>(p->posd_ring.p, p->n_lags, N3, cur, lag0, L, p->dposd.p);
                ++*launches;
                in.posd = p->dposd.p;
            }
            launch_density(p, in, L, launches);
            MDC_CHECK(p->nufft_rc);
            isf_incoherent_accum<<<qb, 256, 0, sk>>>(p->part.p, L, p->n_sets, p->P, p->nq, p->d_sets.p, p->inc_slot.p,
                                                    p->n_inc, lag0, p->dc[0], dc_first, p1_below, p->iisf.p);
            ++*launches;
        }
    }
    MDC_CUDA(cudaGetLastError());
    ++p->isf_in_push;
    return MDC_OK;
}

template <typename T>
int launch_block(mdc_sq_plan *p, const T *raw, int F, int slot, int *launches)
{
    cudaStream_t sc = p->st.copy, sk = p->st.compute;
    const int64_t total = (int64_t)F * p->N;
    const int pb = (int)ceil_div(total, 256);
    if (p->lattice_fast) {
        prep_fix<T><<<pb, 256, 0, sc>>>(raw, p->N, total, 1.0 / p->L0[0], 1.0 / p->L0[1], 1.0 / p->L0[2], p->fix[slot].p);
        ++*launches;
    }
    if (p->factored) {
        sq_prep_tables<<<(unsigned)total, 128, 0, sc>>>(p->fix[slot].p, p->N, total, p->n_points, p->nxp, p->nyp,
                                                       p->nzp, p->tabx[slot].p, p->taby[slot].p, p->tabz[slot].p);
        ++*launches;
    }
    if (!p->lattice_fast || p->n_exp > 0) {
        prep_f64<T><<<pb, 256, 0, sc>>>(raw, p->N, total, p->posd[slot].p);
        ++*launches;
    }
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaEventRecord(p->st.staged[slot], sc));
    MDC_CUDA(cudaStreamWaitEvent(sk, p->st.staged[slot], 0));
    const int ti = p->st.n_timed < 8 ? p->st.n_timed : -1;
    if (ti >= 0) MDC_CUDA(cudaEventRecord(p->st.k0[ti], sk));
    DensityIn in;
    in.fix = p->fix[slot].p;
    in.tabx = p->tabx[slot].p;
    in.taby = p->taby[slot].p;
    in.tabz = p->tabz[slot].p;
    in.posd = p->posd[slot].p;
    launch_density(p, in, F, launches);
    MDC_CHECK(p->nufft_rc);
    if (ti >= 0) {
        MDC_CUDA(cudaEventRecord(p->st.k1[ti], sk));
        p->st.n_timed = ti + 1;
    }
    if (p->isf) {
        MDC_CHECK(isf_frame(p, slot, launches));   // F == 1
    } else {
        sq_pair_accum<<<(unsigned)ceil_div(p->nq, 256), 256, 0, sk>>>(p->part.p, F, p->n_sets, p->P, p->nq,
                                                                     p->d_sets.p, p->d_pairs.p, p->n_pairs, p->dc[0],
                                                                     p->dc[1], (p->factored || p->nufft_on) ? p->n_lat : 0,
                                                                     p->nufft_on ? p->n_lat : 0, p->sum_pairs ? 1 : 0, p->ssf.p);
        ++*launches;
    }
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaEventRecord(p->st.consumed[slot], sk));
    p->st.used[slot] = true;
    return MDC_OK;
}

}  // namespace

namespace {

template <typename T>
int sq_push_impl(mdc_sq_plan *p, const T *pos, int n_frames, int on_device, DevBuf<T> (&stage)[2])
{
    if (!p || !pos) return fail(MDC_ERR_INVALID, "mdc_sq_push: NULL argument");
    if (n_frames <= 0) return MDC_OK;
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CHECK(configure(p, p->isf ? std::min(p->n_lags, 8) : n_frames));
    p->st.n_timed = 0;
    p->isf_in_push = 0;
    int launches = 0;
    const int step = p->isf ? 1 : p->fblk;   // ISF frames are consumed one at a time, in time order
    for (int f0 = 0; f0 < n_frames; f0 += step) {
        const int F = std::min(step, n_frames - f0);
        const int slot = p->st.next;
        p->st.next ^= 1;
        // the slot's layout buffers (and `part`, shared by both slots through stream order on
        // `compute`) may still be read by the kernels of two blocks ago
        if (p->st.used[slot]) MDC_CUDA(cudaStreamWaitEvent(p->st.copy, p->st.consumed[slot], 0));
        const T *src = pos + (size_t)f0 * p->N * 3;
        const T *raw = src;
        if (!on_device) {
            MDC_CHECK(stage[slot].alloc((size_t)p->fblk * p->N * 3));
            MDC_CUDA(cudaMemcpyAsync(stage[slot].p, src, (size_t)F * p->N * 3 * sizeof(T), cudaMemcpyHostToDevice,
                                     p->st.copy));
            raw = stage[slot].p;
        }
        MDC_CHECK(launch_block<T>(p, raw, F, slot, &launches));
    }
    // the caller may reuse `pos` on return: wait for the copies / layout kernels (not the analysis kernels)
    MDC_CUDA(cudaStreamSynchronize(p->st.copy));
    p->n_frames += n_frames;
    p->last_launches = launches;
    return MDC_OK;
}

}  // namespace

extern "C" int mdc_sq_push(mdc_sq_plan *p, const float *pos, int n_frames, int on_device)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_sq_push: NULL argument");
    return sq_push_impl<float>(p, pos, n_frames, on_device, p->raw);
}

// float64 positions (centres of mass of residues: structure.py:1943-1944 keeps them in a float64 buffer)
extern "C" int mdc_sq_push_f64(mdc_sq_plan *p, const double *pos, int n_frames, int on_device)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_sq_push_f64: NULL argument");
    return sq_push_impl<double>(p, pos, n_frames, on_device, p->raw64);
}

extern "C" int mdc_sq_read(mdc_sq_plan *p, double *ssf_out, int64_t *n_frames_out)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_sq_read: NULL plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->st.compute));
    if (ssf_out)
        MDC_CUDA(cudaMemcpy(ssf_out, p->ssf.p, (size_t)p->nq * p->n_rows * sizeof(double), cudaMemcpyDeviceToHost));
    if (n_frames_out) *n_frames_out = p->n_frames;
    return MDC_OK;
}

extern "C" int mdc_sq_set_shells(mdc_sq_plan *p, const int64_t *order, int64_t n_order, const int64_t *lo,
                                 const int64_t *hi, int64_t n_shells)
{
    if (!p || !order || !lo || !hi || n_order <= 0 || n_shells <= 0)
        return fail(MDC_ERR_INVALID, "mdc_sq_set_shells: bad argument");
    for (int64_t i = 0; i < n_order; ++i)
        if (order[i] < 0 || order[i] >= p->nq) return fail(MDC_ERR_INVALID, "mdc_sq_set_shells: wavevector index out of range");
    for (int64_t s = 0; s < n_shells; ++s)
        if (lo[s] < 0 || hi[s] < lo[s] || hi[s] > n_order) return fail(MDC_ERR_INVALID, "mdc_sq_set_shells: bad shell window");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CHECK(p->shell_order.alloc((size_t)n_order));
    MDC_CHECK(p->shell_lo.alloc((size_t)n_shells));
    MDC_CHECK(p->shell_hi.alloc((size_t)n_shells));
    MDC_CHECK(p->shell_out.alloc((size_t)n_shells * p->n_rows));
    MDC_CUDA(cudaMemcpy(p->shell_order.p, order, n_order * sizeof(int64_t), cudaMemcpyHostToDevice));
    MDC_CUDA(cudaMemcpy(p->shell_lo.p, lo, n_shells * sizeof(int64_t), cudaMemcpyHostToDevice));
    MDC_CUDA(cudaMemcpy(p->shell_hi.p, hi, n_shells * sizeof(int64_t), cudaMemcpyHostToDevice));
    p->n_shells = n_shells;
    return MDC_OK;
}

extern "C" int mdc_sq_read_shells(mdc_sq_plan *p, double denom, double *out, int64_t *n_frames_out)
{
    if (!p || !out) return fail(MDC_ERR_INVALID, "mdc_sq_read_shells: NULL argument");
    if (p->n_shells <= 0) return fail(MDC_ERR_INVALID, "mdc_sq_read_shells: call mdc_sq_set_shells first");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t sk = p->st.compute;
    const int64_t warps = p->n_shells * p->n_rows;
    sq_shell_mean<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, sk>>>(p->ssf.p, p->nq, p->shell_order.p, p->shell_lo.p,
                                                                      p->shell_hi.p, p->n_shells, p->n_rows, denom,
                                                                      p->shell_out.p);
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaMemcpyAsync(out, p->shell_out.p, (size_t)warps * sizeof(double), cudaMemcpyDeviceToHost, sk));
    MDC_CUDA(cudaStreamSynchronize(sk));
    if (n_frames_out) *n_frames_out = p->n_frames;
    return MDC_OK;
}

extern "C" int mdc_sq_reset(mdc_sq_plan *p)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_sq_reset: NULL plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->st.compute));
    MDC_CUDA(cudaMemset(p->ssf.p, 0, (size_t)p->nq * p->n_rows * sizeof(double)));
    if (p->isf) {
        MDC_CUDA(cudaMemset(p->cisf.p, 0, (size_t)p->n_lags * p->n_pairs * p->nq * sizeof(double)));
        if (p->incoherent) MDC_CUDA(cudaMemset(p->iisf.p, 0, (size_t)p->n_lags * p->n_inc * p->nq * sizeof(double)));
    }
    p->n_frames = 0;
    return MDC_OK;
}

extern "C" int mdc_sq_accum_ptr(mdc_sq_plan *p, void **dev_ptr, int64_t *n_elems)
{
    if (!p || !dev_ptr || !n_elems) return fail(MDC_ERR_INVALID, "mdc_sq_accum_ptr: NULL argument");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->st.compute));
    *dev_ptr = p->ssf.p;
    *n_elems = p->nq * p->n_rows;
    return MDC_OK;
}

extern "C" int mdc_sq_set_frames(mdc_sq_plan *p, int64_t n_frames)
{
    if (!p || n_frames < 0) return fail(MDC_ERR_INVALID, "mdc_sq_set_frames: bad argument");
    p->n_frames = n_frames;
    return MDC_OK;
}

extern "C" int mdc_sq_last_push_stats(mdc_sq_plan *p, float *kernel_ms, int *n_launches)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_sq_last_push_stats: NULL plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    if (kernel_ms) MDC_CHECK(p->st.kernel_ms(kernel_ms));
    if (n_launches) *n_launches = p->last_launches;
    return MDC_OK;
}

// ---- intermediate scattering function ------------------------------------------------------

extern "C" int mdc_isf_plan_create(mdc_ctx *ctx, int n_points, const double L0[3], double q_max,
                                   const double *q_explicit, int64_t n_explicit, int n_groups,
                                   const int64_t *group_sizes, int n_pairs, const int *pairs, int n_lags,
                                   int incoherent, int precision, int algo, mdc_sq_plan **out)
{
    if (n_lags < 1) return fail(MDC_ERR_INVALID, "mdc_isf_plan_create: n_lags must be positive");
    MDC_CHECK(mdc_sq_plan_create(ctx, n_points, L0, q_max, q_explicit, n_explicit, n_groups, group_sizes, n_pairs,
                                 pairs, precision, algo, out));
    mdc_sq_plan *p = *out;
    *out = nullptr;
    p->isf = true;
    p->n_lags = n_lags;
    p->incoherent = incoherent != 0;
    // rows of iisf: one per group (mode None: one); a set contributes only if a pair (j, j) references it
    const bool total = pairs[0] < 0;
    p->n_inc = total ? 1 : n_groups;
    std::vector<int> slot(p->n_sets, -1);
    for (int i = 0; i < n_pairs; ++i)
        if (pairs[2 * i] == pairs[2 * i + 1]) slot[p->pair_sets[i].x] = total ? 0 : pairs[2 * i];
    int rc = configure(p, std::min(n_lags, 8));
    const size_t N3 = (size_t)3 * p->N;
    if (rc == MDC_OK) rc = p->rho_ring.alloc((size_t)n_lags * p->n_sets * p->nq);
    if (rc == MDC_OK) rc = p->cisf.alloc((size_t)n_lags * n_pairs * p->nq);
    if (rc == MDC_OK) rc = p->inc_slot.alloc(p->n_sets);
    if (rc == MDC_OK && p->incoherent) {
        rc = p->iisf.alloc((size_t)n_lags * p->n_inc * p->nq);
        const bool need_posd = !p->lattice_fast || p->n_exp > 0;
        if (rc == MDC_OK && p->lattice_fast) rc = p->fix_ring.alloc((size_t)n_lags * N3);
        if (rc == MDC_OK && p->lattice_fast) rc = p->dfix.alloc((size_t)p->fblk * N3);
        if (rc == MDC_OK && need_posd) rc = p->posd_ring.alloc((size_t)n_lags * N3);
        if (rc == MDC_OK && need_posd) rc = p->dposd.alloc((size_t)p->fblk * N3);
        if (rc == MDC_OK && p->factored) rc = p->dtabx.alloc((size_t)p->fblk * p->N * p->nxp);
        if (rc == MDC_OK && p->factored) rc = p->dtaby.alloc((size_t)p->fblk * p->N * p->nyp);
        if (rc == MDC_OK && p->factored) rc = p->dtabz.alloc((size_t)p->fblk * p->N * p->nzp);
    }
    cudaError_t e = cudaSuccess;
    if (rc == MDC_OK) e = cudaMemcpy(p->inc_slot.p, slot.data(), slot.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (rc == MDC_OK && e == cudaSuccess) e = cudaMemset(p->cisf.p, 0, (size_t)n_lags * n_pairs * p->nq * sizeof(double));
    if (rc == MDC_OK && e == cudaSuccess && p->incoherent)
        e = cudaMemset(p->iisf.p, 0, (size_t)n_lags * p->n_inc * p->nq * sizeof(double));
    if (rc == MDC_OK && e != cudaSuccess) rc = fail(MDC_ERR_CUDA, std::string("mdc_isf_plan_create: ") + cudaGetErrorString(e));
    if (rc != MDC_OK) {
        plan_free(p);
        return rc;
    }
    *out = p;
    return MDC_OK;
}

extern "C" int mdc_isf_push(mdc_sq_plan *p, const float *pos, int n_frames, int on_device)
{
    if (!p || !p->isf) return fail(MDC_ERR_INVALID, "mdc_isf_push: not an ISF plan");
    return mdc_sq_push(p, pos, n_frames, on_device);
}

extern "C" int mdc_isf_read(mdc_sq_plan *p, double *cisf_out, double *iisf_out, int64_t *n_frames_out)
{
    if (!p || !p->isf) return fail(MDC_ERR_INVALID, "mdc_isf_read: not an ISF plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->st.compute));
    if (cisf_out)
        MDC_CUDA(cudaMemcpy(cisf_out, p->cisf.p, (size_t)p->n_lags * p->n_pairs * p->nq * sizeof(double),
                            cudaMemcpyDeviceToHost));
    if (iisf_out) {
        if (!p->incoherent) return fail(MDC_ERR_INVALID, "mdc_isf_read: plan was created without the incoherent part");
        MDC_CUDA(cudaMemcpy(iisf_out, p->iisf.p, (size_t)p->n_lags * p->n_inc * p->nq * sizeof(double),
                            cudaMemcpyDeviceToHost));
    }
    if (n_frames_out) *n_frames_out = p->n_frames;
    return MDC_OK;
}

extern "C" int mdc_delta_fourier_transform_sum(mdc_ctx *ctx, const double *qs, int64_t nq, const double *rs,
                                               int64_t n, int precision, double *rho_out)
{
    if (!ctx || !qs || !rs || !rho_out) return fail(MDC_ERR_INVALID, "mdc_delta_fourier_transform_sum: NULL argument");
    if (nq <= 0 || n <= 0 || n > 2147483647ll) return fail(MDC_ERR_INVALID, "mdc_delta_fourier_transform_sum: bad sizes");
    MDC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<double> dq, drs, dpos;
    DevBuf<double2> part;
    DevBuf<int2> dset;
    int rc = MDC_OK;
    const int64_t blocks_q = ceil_div(nq, SQ_THREADS);
    int64_t P = std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)ctx->sm_count * 16, blocks_q), n / SQX_TILE));
    P = std::min<int64_t>(P, 1024);
    std::vector<double2> h;
    double dc[2] = {0, 0};
    do {
        if ((rc = dq.alloc((size_t)nq * 3)) || (rc = drs.alloc((size_t)n * 3)) || (rc = dpos.alloc((size_t)n * 3)) ||
            (rc = part.alloc((size_t)P * nq)) || (rc = dset.alloc(1)))
            break;
        const int2 set = make_int2(0, (int)n);
        cudaMemcpy(dq.p, qs, (size_t)nq * 3 * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(drs.p, rs, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(dset.p, &set, sizeof(set), cudaMemcpyHostToDevice);
        transpose_f64<<<(unsigned)ceil_div(n, 256), 256>>>(drs.p, n, dpos.p);
        dim3 g((unsigned)blocks_q, (unsigned)P, 1);
        if (precision == MDC_PRECISION_FP64) {
            sq_explicit_f64<<<g, SQ_THREADS>>>(dq.p, 0, nq, dpos.p, n, dset.p, 1, (int)P, nq, part.p);
        } else {
            double pr[4];
            if ((rc = run_probe_sincos(0, pr))) break;
            dc[0] = pr[0];
            dc[1] = pr[1];
            sq_explicit_f32<<<g, SQ_THREADS>>>(dq.p, 0, nq, dpos.p, n, dset.p, 1, (int)P, nq, part.p);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            rc = fail(MDC_ERR_CUDA, std::string("mdc_delta_fourier_transform_sum: ") + cudaGetErrorString(e));
            break;
        }
        h.resize((size_t)P * nq);
        e = cudaMemcpy(h.data(), part.p, h.size() * sizeof(double2), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            rc = fail(MDC_ERR_CUDA, std::string("mdc_delta_fourier_transform_sum: ") + cudaGetErrorString(e));
            break;
        }
        for (int64_t i = 0; i < nq; ++i) {
            double re = 0, im = 0;
            for (int64_t pp = 0; pp < P; ++pp) {
                re += h[pp * nq + i].x;
                im += h[pp * nq + i].y;
            }
            rho_out[2 * i] = re - (double)n * dc[0];
            rho_out[2 * i + 1] = im - (double)n * dc[1];
        }
    } while (0);
    dq.release();
    drs.release();
    dpos.release();
    part.release();
    dset.release();
    return rc;
}

extern "C" int mdc_probe_sincos(mdc_ctx *ctx, int variant, double out[4])
{
    if (!ctx || !out) return fail(MDC_ERR_INVALID, "mdc_probe_sincos: NULL argument");
    MDC_CUDA(cudaSetDevice(ctx->device));
    return run_probe_sincos(variant, out);
}
