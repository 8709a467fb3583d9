// sq_nufft.cu — S(q) on the reciprocal lattice as a type-1 non-uniform FFT (sm_100a).
//
// Replaces, for lattice wavevectors, the same reference code as sq.cu:
//   /root/reference/src/mdcraft/analysis/structure.py:1240-1280  delta_fourier_transform_sum
//   /root/reference/src/mdcraft/algorithm/accelerated.py:84-113  numba_dot
// rho(a, b, c) = sum_j exp(2 pi i (a sx_j + b sy_j + c sz_j)),  s = r / L0 in [0, 1),  a, b, c in [0, n).
//
// The direct and factored kernels spend Nq * N terms per frame (2.1e11 at N = 1e5 on the q_max = 20/sigma
// grid).  Because the wavevectors are ALL the integer modes of a box, the same sums are the Fourier
// coefficients of the particle density, and a gridding NUFFT evaluates them in O(N w^3 + nf^3 log nf):
//   1. centre the modes: k' = k - k0, strength c_j = exp(2 pi i k0 (sx + sy + sz)),
//      so a fine grid of nf >= 1.5 n points per axis suffices instead of 1.5 (2 n - 1);
//   2. spread c_j onto the fine grid with the separable "exponential of semicircle" kernel
//      phi(z) = exp(beta (sqrt(1 - z^2) - 1)), z = (l - s nf) / (w / 2), |z| <= 1  (Barnett et al., FINUFFT):
//      fine grids of 128 points or more BY GATHER (nufft_records / nufft_tile_fill / nufft_spread_tiles: a block owns a
//      tile of the grid, accumulates in registers and writes every cell once), smaller ones by scatter (nufft_spread:
//      one warp per particle, FP64 reductions into an L2-resident grid);
//   3. pruned FFTs, one axis at a time, in FP64, in shared memory: nf inputs -> the n wanted modes, so the
//      intermediates shrink (nf^3 -> nf^2 n -> nf n^2 -> kept wavevectors);
//   4. divide by the kernel's Fourier transform (per-axis table) while scattering into the plan's wavevector order.
// The relative l2 error is set by (nf / n, w): 1e-9 requested in FP32 mode (the factored FP32 kernel reaches
// 8e-8), 1e-13 in FP64 mode; tools/nufft_model.py holds the NumPy model these parameters were derived with.
//
// Rooflines: the FFT passes are HBM bound (every intermediate written and read once: 16 B * (nf^3 + 2 nf^2 n +
// 2 nf n^2 + Nq) per (frame, set)); the gather kernel is FP64-issue / latency bound (4 DMUL + 32 DFMA per particle and
// patch of 4 x 8 x 16 cells) and writes 16 B * nf^3 once; the scatter kernel is bound by the L2 reduction rate
// (N w^3 complex RED.E.ADD.F64).
#include "sq_nufft.cuh"

#include <math.h>
#include <algorithm>
#include <vector>

namespace mdc {

namespace {

constexpr int NU_TC = 8;         // lines per block in the strided FFT passes (8 x 16 B = one 128-byte segment)
#ifndef MDC_NU_THREADS
#define MDC_NU_THREADS 128
#endif
// Threads per block of the spreading and FFT kernels (one line per warp and turn).  128 threads x <= 128 registers
// fit the quarter of an SM that the g(r) pair kernel leaves free when it runs next to these kernels on another
// stream (mdc_rdf_set_blocks_per_sm); alone, 256 threads are 3 % faster.
constexpr int NU_THREADS = MDC_NU_THREADS;
#ifndef MDC_NU_FFT_BLOCKS
#define MDC_NU_FFT_BLOCKS 4   // resident FFT blocks per SM the register budget is sized for (4: <= 128 registers)
#endif

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// shared-memory index of element p of a line: one 16-byte slot of padding per 8 keeps the strided accesses of
// the later FFT stages (stride nf / 8, nf / 64, ...) spread over all banks
__device__ __forceinline__ int padded(int p) { return p + (p >> 3); }

// y_r = sum_q v_q w^(r q), w = exp(+2 pi i / R), in place, natural order
template <int R>
__device__ __forceinline__ void dft(double2 *v);

template <>
__device__ __forceinline__ void dft<2>(double2 *v)
{
    const double2 a = v[0], b = v[1];
    v[0] = make_double2(a.x + b.x, a.y + b.y);
    v[1] = make_double2(a.x - b.x, a.y - b.y);
}

template <>
__device__ __forceinline__ void dft<3>(double2 *v)
{
    const double hs = 0.8660254037844386;   // sin(2 pi / 3)
    const double2 a = v[0], b = v[1], c = v[2];
    const double2 s = make_double2(b.x + c.x, b.y + c.y), d = make_double2(b.x - c.x, b.y - c.y);
    const double2 h = make_double2(a.x - 0.5 * s.x, a.y - 0.5 * s.y);
    v[0] = make_double2(a.x + s.x, a.y + s.y);
    v[1] = make_double2(h.x - hs * d.y, h.y + hs * d.x);   // a + w3 b + w3^2 c
    v[2] = make_double2(h.x + hs * d.y, h.y - hs * d.x);
}

template <>
__device__ __forceinline__ void dft<4>(double2 *v)
{
    const double2 t0 = make_double2(v[0].x + v[2].x, v[0].y + v[2].y), t1 = make_double2(v[0].x - v[2].x, v[0].y - v[2].y);
    const double2 t2 = make_double2(v[1].x + v[3].x, v[1].y + v[3].y), t3 = make_double2(v[1].x - v[3].x, v[1].y - v[3].y);
    v[0] = make_double2(t0.x + t2.x, t0.y + t2.y);
    v[2] = make_double2(t0.x - t2.x, t0.y - t2.y);
    v[1] = make_double2(t1.x - t3.y, t1.y + t3.x);   // t1 + i t3
    v[3] = make_double2(t1.x + t3.y, t1.y - t3.x);   // t1 - i t3
}

template <>
__device__ __forceinline__ void dft<8>(double2 *v)
{
    double2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    dft<4>(e);
    dft<4>(o);
    const double r = 0.7071067811865476;
    // o_r * w8^r:  w8 = (1 + i) / sqrt(2), w8^2 = i, w8^3 = (-1 + i) / sqrt(2)
    const double2 o1 = make_double2(r * (o[1].x - o[1].y), r * (o[1].x + o[1].y));
    const double2 o2 = make_double2(-o[2].y, o[2].x);
    const double2 o3 = make_double2(-r * (o[3].x + o[3].y), r * (o[3].x - o[3].y));
    v[0] = make_double2(e[0].x + o[0].x, e[0].y + o[0].y);
    v[4] = make_double2(e[0].x - o[0].x, e[0].y - o[0].y);
    v[1] = make_double2(e[1].x + o1.x, e[1].y + o1.y);
    v[5] = make_double2(e[1].x - o1.x, e[1].y - o1.y);
    v[2] = make_double2(e[2].x + o2.x, e[2].y + o2.y);
    v[6] = make_double2(e[2].x - o2.x, e[2].y - o2.y);
    v[3] = make_double2(e[3].x + o3.x, e[3].y + o3.y);
    v[7] = make_double2(e[3].x - o3.x, e[3].y - o3.y);
}

// One decimation-in-frequency stage of radix R on the blocks of length L = R << lgsub of one line:
// sub-block r of every block receives the sequence whose length-(L / R) transform is X[R j + r].
// src != nullptr: the inputs come straight from global memory (first stage of a contiguous line) instead of from
// the shared-memory line.  Only W^i is read from the twiddle table; its powers are formed by multiplication
// (depth 3, ~2 ulp), which takes 6 of the 7 table reads of a radix-8 butterfly off the shared-memory pipe.
#ifndef MDC_FFT_POW_Z
#define MDC_FFT_POW_Z 1
#endif
#ifndef MDC_FFT_POW_S
#define MDC_FFT_POW_S 0
#endif
template <int R, bool POW>
__device__ __forceinline__ void fft_stage(double2 *buf, const double2 *tw, int nf, int lgsub, int lane,
                                          const double2 *__restrict__ src = nullptr)
{
    const int sub = 1 << lgsub, L = R << lgsub, tstep = nf / L;
    for (int t = lane; t < nf / R; t += 32) {
        const int blk = t >> lgsub, i = t & (sub - 1);
        const int base = blk * L + i;
        double2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = src ? src[base + q * sub] : buf[padded(base + q * sub)];
        dft<R>(v);
        if (i != 0) {
            if (POW) {
                double2 w[R];
                w[1] = tw[i * tstep];
                if (R > 2) w[2] = cmul(w[1], w[1]);
                if (R > 3) w[3] = cmul(w[2], w[1]);
                if (R > 4) {
                    w[4] = cmul(w[2], w[2]);
                    w[5] = cmul(w[4], w[1]);
                    w[6] = cmul(w[4], w[2]);
                    w[7] = cmul(w[4], w[3]);
                }
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw[r * i * tstep]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) buf[padded(base + r * sub)] = v[r];
    }
    __syncwarp();
}

// One warp transforms one line of nf complex doubles in shared memory (element p at buf[padded(p)]), in place:
//   X[k] = sum_l x[l] exp(+2 pi i k l / nf);   X[k] ends up at element fft_pos(k) (mixed-radix digit reversal,
// undone for free by the pruned read-out).  Radices: 3 (if nf = 3 * 2^a), then 8s, then 4 or 2.
// tw[t] = exp(+2 pi i t / nf).  src != nullptr: the line is still in global memory; the first stage loads it.
template <bool POW>
__device__ __forceinline__ void fft_line(double2 *buf, const double2 *tw, const NufftGeom &g, int lane,
                                         const double2 *__restrict__ src = nullptr)
{
    int lg = g.log2m;   // log2 of the current block length (after the radix-3 stage)
    if (g.r3) {
        fft_stage<3, POW>(buf, tw, g.nf, lg, lane, src);
        src = nullptr;
    }
    for (; lg >= 3; lg -= 3) {
        fft_stage<8, POW>(buf, tw, g.nf, lg - 3, lane, src);
        src = nullptr;
    }
    if (lg == 2) fft_stage<4, POW>(buf, tw, g.nf, 0, lane);
    else if (lg == 1) fft_stage<2, POW>(buf, tw, g.nf, 0, lane);
}

// element index, after fft_line, of X[k]
__device__ __forceinline__ int fft_pos(int k, const NufftGeom &g)
{
    int p = 0, lg = g.log2m;
    if (g.r3) {
        const int d = k % 3;
        k /= 3;
        p = d * g.m;
    }
    for (; lg >= 3; lg -= 3) {
        p += (k & 7) << (lg - 3);
        k >>= 3;
    }
    if (lg == 2) p += k & 3;
    else if (lg == 1) p += k & 1;
    return p;
}

__global__ void nufft_twiddles(int nf, double2 *tw)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nf) return;
    double s, c;
    sincospi(2.0 * (double)t / (double)nf, &s, &c);
    tw[t] = make_double2(c, s);
}

// ---- particle order: counting sort of every set by fine-grid x plane -----------------------------

// fine-grid coordinate of a position: frac(x / L) * nf in [0, nf)
__device__ __forceinline__ double fine_coord(double x, double iL, int nf)
{
    double s = x * iL;
    s -= floor(s);
    s *= (double)nf;
    return s < (double)nf ? s : 0.0;   // frac() of a tiny negative number rounds to 1
}

__global__ void nufft_bin_count(const double *__restrict__ posd, int64_t N, const int2 *__restrict__ sets, int n_sets,
                                int e0, int nf, double iLx, int *__restrict__ cnt)
{
    const int e = e0 + blockIdx.y, f = e / n_sets, s = e - f * n_sets;
    const int2 set = sets[s];
    const int64_t j = (int64_t)set.x + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= set.y) return;
    const int bin = (int)fine_coord(posd[((int64_t)f * 3 + 0) * N + j], iLx, nf);
    atomicAdd(&cnt[blockIdx.y * nf + bin], 1);
}

// exclusive scan of nf <= 1024 counters per element (one block of 1024 threads per element)
__global__ void nufft_bin_scan(int nf, int *__restrict__ cnt)
{
    __shared__ int sh[1024];
    int *c = cnt + blockIdx.x * nf;
    const int t = threadIdx.x;
    const int v = t < nf ? c[t] : 0;
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int add = t >= o ? sh[t - o] : 0;
        __syncthreads();
        sh[t] += add;
        __syncthreads();
    }
    if (t < nf) c[t] = sh[t] - v;
}

__global__ void nufft_bin_scatter(const double *__restrict__ posd, int64_t N, const int2 *__restrict__ sets,
                                  int n_sets, int e0, int nf, double iLx, int *__restrict__ cursor,
                                  int *__restrict__ order, int64_t maxlen)
{
    const int e = e0 + blockIdx.y, f = e / n_sets, s = e - f * n_sets;
    const int2 set = sets[s];
    const int64_t j = (int64_t)set.x + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= set.y) return;
    const int bin = (int)fine_coord(posd[((int64_t)f * 3 + 0) * N + j], iLx, nf);
    const int pos = atomicAdd(&cursor[blockIdx.y * nf + bin], 1);
    order[(int64_t)blockIdx.y * maxlen + pos] = (int)j;
}

// ---- spreading -------------------------------------------------------------------------------------

// One warp per particle.  Lanes cover one z row of the kernel's support as interleaved (re, im) doubles, so a
// warp-wide RED touches 2 W consecutive doubles; the W x W rows (x, y offsets) are walked in a fully unrolled
// loop whose body is one DMUL, one address add and the RED (W <= 8: two rows per instruction, split by x).
// posd f64 (F, 3, N): the float32 coordinates widened exactly; iL = 1 / L0.
template <int W>
__global__ void __launch_bounds__(NU_THREADS)
nufft_spread(const double *__restrict__ posd, int64_t N, const int2 *__restrict__ sets, int n_sets, int e0,
             const int *__restrict__ order, int64_t maxlen, NufftGeom g, double iLx, double iLy, double iLz,
             double *__restrict__ grid)
{
    __shared__ double sv[NU_THREADS / 32][3][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = e0 + blockIdx.y, f = e / n_sets, s = e - f * n_sets;
    const int2 set = sets[s];
    const int len = set.y - set.x;
    const int nf = g.nf;
    const double *px = posd + ((int64_t)f * 3 + 0) * N;
    const double *py = posd + ((int64_t)f * 3 + 1) * N;
    const double *pz = posd + ((int64_t)f * 3 + 2) * N;
    double *G = grid + (size_t)blockIdx.y * nf * nf * nf * 2;
    const double hw = 0.5 * (double)W, ihw = 2.0 / (double)W, inf = 1.0 / (double)nf;

    constexpr int LPR = (W <= 8) ? 16 : 32;      // lanes per z row
    constexpr int RPI = 32 / LPR;                // rows per warp-wide instruction
    const int rsel = lane / LPR, within = lane - rsel * LPR, dz = within >> 1, ri = within & 1;
    const bool act = dz < W;
    const int t = lane & 15;
    const size_t plane_bytes = (size_t)nf * nf * 16;           // one x plane of the grid
    char *Gb = reinterpret_cast<char *>(G);

    for (int idx = blockIdx.x * (NU_THREADS / 32) + warp; idx < len; idx += gridDim.x * (NU_THREADS / 32)) {
        const int64_t j = order ? (int64_t)order[(int64_t)blockIdx.y * maxlen + idx] : (int64_t)set.x + idx;
        const double gx = fine_coord(px[j], iLx, nf), gy = fine_coord(py[j], iLy, nf), gz = fine_coord(pz[j], iLz, nf);
        // kernel values: lanes 0-15 x, 16-31 y; then lanes 0-15 z
        int l0a, l0z;
        {
            const double x = (lane < 16) ? gx : gy;
            const double c = ceil(x - hw);
            l0a = (int)c;
            const double z = (c + (double)t - x) * ihw;
            const double a = fmax(1.0 - z * z, 0.0);
            sv[warp][lane >> 4][t] = t < W ? exp(g.beta * (sqrt(a) - 1.0)) : 0.0;
        }
        {
            const double c = ceil(gz - hw);
            l0z = (int)c;
            const double z = (c + (double)t - gz) * ihw;
            const double a = fmax(1.0 - z * z, 0.0);
            if (lane < 16) sv[warp][2][t] = t < W ? exp(g.beta * (sqrt(a) - 1.0)) : 0.0;
        }
        const int l0x = __shfl_sync(0xffffffffu, l0a, 0), l0y = __shfl_sync(0xffffffffu, l0a, 16);
        // strength exp(2 pi i k0 (sx + sy + sz)) of the centred modes
        double sn, cs;
        sincospi(2.0 * (double)g.k0 * ((gx + gy + gz) * inf), &sn, &cs);
        __syncwarp();
        if (act) {
            const double zc = sv[warp][2][dz] * (ri ? sn : cs);
            int Z = l0z + dz;
            Z = Z < 0 ? Z + nf : (Z >= nf ? Z - nf : Z);
            double vy[W];
            unsigned yo[W];
#pragma unroll
            for (int iy = 0; iy < W; ++iy) {
                int Y = l0y + iy;
                Y = Y < 0 ? Y + nf : (Y >= nf ? Y - nf : Y);
                vy[iy] = sv[warp][1][iy];
                yo[iy] = (((unsigned)Y * (unsigned)nf + (unsigned)Z) * 2u + (unsigned)ri) * 8u;   // bytes, < 2^32
            }
#pragma unroll 2
            for (int ix = rsel; ix < W; ix += RPI) {
                int X = l0x + ix;
                X = X < 0 ? X + nf : (X >= nf ? X - nf : X);
                char *gp = Gb + (size_t)X * plane_bytes;
                const double vxz = zc * sv[warp][0][ix];
#pragma unroll
                for (int iy = 0; iy < W; ++iy) {
                    // grid[gp + yo[iy]] += v: the address in ONE instruction (IMAD.WIDE.U32 with a 64-bit addend),
                    // then the fire-and-forget reduction
                    asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %0, 1, %1;\n\tred.global.add.f64 [a], %2;\n\t}"
                                 ::"r"(yo[iy]), "l"(gp), "d"(vxz * vy[iy])
                                 : "memory");
                }
            }
        }
        __syncwarp();
    }
}


// ---- spreading by gather: one block per 16^3 tile of the fine grid --------------------------------------
//
// The scatter kernel above issues w^3 complex FP64 reductions per particle into L2 (3,456 REDs at w = 12) and runs at
// the L2 atomic rate.  Here the roles are swapped: a block OWNS a tile of 16 x 16 x 16 fine-grid cells, a thread owns
// tile of 8 x 16 x 16 fine-grid cells, each of its four warps a patch of 4 x 8 x 16 cells.  The lanes of a warp are the
// 32 doubles of one z line of the patch (16 cells x (re, im)), and every lane keeps the 4 x 8 (x, y) positions of its
// (z, re / im) in registers.  The block walks over the particles whose support touches the tile (those of the 27
// surrounding tile bins, pre-filtered); per particle and warp the update is an outer product
//     acc[i][j] += (vx[i] * wz_lane) * vy[j]            (4 DMUL + 32 DFMA, vx / vy warp-uniform broadcast loads)
// with vx, vy, wz zero-padded in shared memory, so the body needs no range test whatever the offset of the support.
// Every cell is then written to HBM exactly once, in 256-byte lines: no atomics, no clearing pass, no
// read-modify-write traffic.  Per particle the kernel values are computed once by nufft_records (3 w FP64
// exponentials) and staged in shared memory in chunks of 32.

constexpr int NT_T = 16;                 // bin edge = tile extent in y and z (cells); also the pencil length per thread
constexpr int NT_TX = 8;                 // tile extent in x: half a bin, so that a block is 128 threads x < 128 registers and
                                         // fits the quarter of an SM that the g(r) pair kernel leaves free (see NU_THREADS)
constexpr int NT_PX = 4, NT_PY = 8;      // patch of one warp: 4 x 8 (x, y) positions x 16 z; 2 x 2 patches per tile
constexpr int NT_THREADS = 128;          // 4 warps
constexpr int NT_MAX_TILES = 12;         // tiles a support of <= 16 cells can reach: 3 x 2 x 2

// Tiles (8 x 16 x 16 cells, index (tx2 * nb + ty) * nb + tz) that the support [l0, l0 + W) of a particle reaches, per
// axis: first tile and count (periodic).  At most 3 x 2 x 2 for W <= 16.
template <int W>
__device__ __forceinline__ void tiles_reached(const int4 &l0, int nf, int (&t0)[3], int (&nt)[3])
{
    const int ext[3] = {NT_TX, NT_T, NT_T};
    const int lo[3] = {l0.x, l0.y, l0.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = lo[k] + nf, e = a + W - 1;          // shifted by nf: non-negative (l0 >= -W / 2)
        t0[k] = a / ext[k];
        nt[k] = e / ext[k] - t0[k] + 1;
    }
}

// exclusive scan of n counters per element (one block of 128 threads per element, any n).  128 threads, like every
// kernel of this chain: a larger block would not fit next to the persistent g(r) pair kernel and would wait for it.
__global__ void __launch_bounds__(128) nufft_scan_any(int n, int *__restrict__ cnt)
{
    __shared__ int sh[128];
    int *c = cnt + (size_t)blockIdx.x * n;
    const int per = (n + 127) / 128, lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += c[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 128; o <<= 1) {
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

// Kernel values of every particle of an element (slot = index within its set): rec[slot] = {vx[W], vy[W],
// (vz[W] * strength) as (re, im)[W]}, cell[slot] = first cell of the support per axis (may be negative: periodic); and
// one count per tile the support reaches (tcount: the CSR row lengths of the tile -> particles lists).
// Half a warp per particle (lane t of the half evaluates offset t of the support: x, then y, then z), and two particles
// per half in flight (independent chains of FP64 exponentials: the kernel is latency bound, next to the pair kernel
// it has four warps per SM): 16 particles per block of 128 threads.
constexpr int NR_PER_BLOCK = 16;
template <int W>
__global__ void __launch_bounds__(128)
nufft_records(const double *__restrict__ posd, int64_t N, const int2 *__restrict__ sets, int n_sets, int e0,
              int64_t maxlen, NufftGeom g, int nb, double iLx, double iLy, double iLz, double *__restrict__ rec,
              int4 *__restrict__ cell, int *__restrict__ tcount)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, t = lane & 15;
    const int e = e0 + blockIdx.y, f = e / n_sets, s = e - f * n_sets;
    const int2 set = sets[s];
    const int len = set.y - set.x;
    const int nf = g.nf;
    const double hw = 0.5 * (double)W, ihw = 2.0 / (double)W, inf = 1.0 / (double)nf;
    const double iL[3] = {iLx, iLy, iLz};
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        const int idx = blockIdx.x * NR_PER_BLOCK + warp * 4 + rep * 2 + half;
        if (idx >= len) continue;
        const int64_t slot = (int64_t)blockIdx.y * maxlen + idx;
        const int64_t j = (int64_t)set.x + idx;
        double *R = rec + slot * (4 * W);
        double gk[3], v[3];
        int l0[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            gk[k] = fine_coord(posd[((int64_t)f * 3 + k) * N + j], iL[k], nf);
            const double c = ceil(gk[k] - hw);
            l0[k] = (int)c;
            const double z = (c + (double)t - gk[k]) * ihw;
            const double a = fmax(1.0 - z * z, 0.0);
            v[k] = exp(g.beta * (sqrt(a) - 1.0));
        }
        if (t < W) {
            // strength exp(2 pi i k0 (sx + sy + sz)) of the centred modes
            double sn, cs;
            sincospi(2.0 * (double)g.k0 * ((gk[0] + gk[1] + gk[2]) * inf), &sn, &cs);
            R[t] = v[0];
            R[W + t] = v[1];
            R[2 * W + 2 * t] = v[2] * cs;
            R[2 * W + 2 * t + 1] = v[2] * sn;
        }
        const int4 c4 = make_int4(l0[0], l0[1], l0[2], 0);
        if (t == 0) cell[slot] = c4;
        // one lane per reached tile (at most 12)
        int t0[3], nt[3];
        tiles_reached<W>(c4, nf, t0, nt);
        if (t < nt[0] * nt[1] * nt[2]) {
            const int ix = t / (nt[1] * nt[2]), iy = (t / nt[2]) % nt[1], iz = t % nt[2];
            const int tile = (((t0[0] + ix) % (2 * nb)) * nb + (t0[1] + iy) % nb) * nb + (t0[2] + iz) % nb;
            atomicAdd(&tcount[(size_t)blockIdx.y * 2 * nb * nb * nb + tile], 1);
        }
    }
}

// CSR fill: every particle appends its slot to the list of every tile its support reaches; afterwards cursor[tile] = END
// of the tile's list (cursor starts as the exclusive scan of the counts).  One thread per particle.
template <int W>
__global__ void nufft_tile_fill(const int2 *__restrict__ sets, int n_sets, int e0, int64_t maxlen, int nf, int nb,
                                const int4 *__restrict__ cell, int *__restrict__ cursor, int *__restrict__ tlist,
                                int64_t list_pitch)
{
    const int e = e0 + blockIdx.y, f = e / n_sets, s = e - f * n_sets;
    const int2 set = sets[s];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= set.y - set.x) return;
    const int4 l0 = cell[(int64_t)blockIdx.y * maxlen + idx];
    int t0[3], nt[3];
    tiles_reached<W>(l0, nf, t0, nt);
    int *cur = cursor + (size_t)blockIdx.y * 2 * nb * nb * nb;
    int *out = tlist + (size_t)blockIdx.y * list_pitch;
    for (int ix = 0; ix < nt[0]; ++ix)
        for (int iy = 0; iy < nt[1]; ++iy)
            for (int iz = 0; iz < nt[2]; ++iz) {
                const int tile = (((t0[0] + ix) % (2 * nb)) * nb + (t0[1] + iy) % nb) * nb + (t0[2] + iz) % nb;
                out[atomicAdd(&cur[tile], 1)] = idx;
            }
}

// d = a - b reduced to (-span, nf - span]: the offset of cell a from support start b on the periodic grid
__device__ __forceinline__ int wrap_off(int d, int nf, int span)
{
    d += (d <= -span) ? nf : 0;
    d -= (d > nf - span) ? nf : 0;
    return d;
}

template <int W>
__global__ void __launch_bounds__(NT_THREADS, 4)
nufft_spread_tiles(const double *__restrict__ rec, const int4 *__restrict__ cell, const int *__restrict__ tend,
                   const int *__restrict__ tlist, int64_t list_pitch, int64_t maxlen, int nf, int nb,
                   double2 *__restrict__ grid)
{
    // per-particle kernel values; vx / vy zero-padded by one patch extent on either side, so that a warp reads the
    // NT_PX / NT_PY entries its patch needs at ANY offset of the support without a range test.  Two stages: the
    // records of chunk c + 1 arrive (cp.async) while chunk c is being consumed.
    constexpr int VXP = W + 2 * NT_PX, VYP = W + 2 * NT_PY;
    constexpr int NT_CHUNK = 32;                         // particles per stage
    __shared__ __align__(16) double s_vx[2][NT_CHUNK][VXP], s_vy[2][NT_CHUNK][VYP], s_wz[2][NT_CHUNK][2 * W];   // wz: (re, im)
    __shared__ int4 s_cell[2][NT_CHUNK];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tb = blockIdx.x;                          // tile: (tx2 * nb + ty) * nb + tz, tx2 = half-bin index along x
    const int tz = tb % nb, ty = (tb / nb) % nb, tx2 = tb / (nb * nb);
    const int X0 = tx2 * NT_TX, Y0 = ty * NT_T, Z0 = tz * NT_T;
    // a warp owns a patch of NT_PX x NT_PY x 16 cells; lane = (z, re / im): the 32 lanes cover one z line of the patch
    // as 32 consecutive doubles, and every lane accumulates the NT_PX x NT_PY (x, y) positions of its (z, re / im)
    const int XW = X0 + (warp >> 1) * NT_PX, YW = Y0 + (warp & 1) * NT_PY;
    const size_t ebase = (size_t)blockIdx.y * maxlen;
    // the particles whose support reaches this tile: tlist[lo .. hi) (built by nufft_records / nufft_tile_fill)
    const int *te = tend + (size_t)blockIdx.y * 2 * nb * nb * nb;
    const int lo = tb == 0 ? 0 : te[tb - 1], n = te[tb] - lo;
    const int *list = tlist + (size_t)blockIdx.y * list_pitch + lo;

    double acc[NT_PX][NT_PY];
#pragma unroll
    for (int i = 0; i < NT_PX; ++i)
#pragma unroll
        for (int j = 0; j < NT_PY; ++j) acc[i][j] = 0.0;

    // stage the records of list[c0 .. c0 + nc) into buffer `buf`: four threads per particle, a quarter of the record
    // (W doubles) each: vx | vy | first half of wz | second half of wz; asynchronous (LDGSTS), committed as one group
    auto stage = [&](int buf, int c0, int nc) {
        const int p = tid >> 2, part = tid & 3;
        if (p < nc) {
            const int slot = list[c0 + p];
            const double *src = rec + (ebase + slot) * (4 * W) + part * W;
            double *dst = part == 0 ? &s_vx[buf][p][NT_PX] : part == 1 ? &s_vy[buf][p][NT_PY] : &s_wz[buf][p][(part - 2) * W];
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll
            for (int o = 0; o < W; ++o)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 8u * o), "l"(src + o) : "memory");
            if (part == 0)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&s_cell[buf][p])),
                             "l"(cell + ebase + slot)
                             : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (n > 0) stage(0, 0, min(NT_CHUNK, n));
    // the zero pads of vx / vy (never written by the staging) while the first records are in flight
    for (int i = tid; i < 2 * NT_CHUNK * 2 * NT_PX; i += NT_THREADS) {
        const int r = i / (2 * NT_PX), o = i - r * (2 * NT_PX);
        (&s_vx[0][0][0])[r * VXP + (o < NT_PX ? o : W + o)] = 0.0;
    }
    for (int i = tid; i < 2 * NT_CHUNK * 2 * NT_PY; i += NT_THREADS) {
        const int r = i / (2 * NT_PY), o = i - r * (2 * NT_PY);
        (&s_vy[0][0][0])[r * VYP + (o < NT_PY ? o : W + o)] = 0.0;
    }
    int buf = 0;
    for (int c0 = 0; c0 < n; c0 += NT_CHUNK, buf ^= 1) {
        const int nc = min(NT_CHUNK, n - c0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();   // chunk c0 has landed; every warp is done with the other buffer (chunk c0 - 32)
        if (c0 + NT_CHUNK < n) stage(buf ^ 1, c0 + NT_CHUNK, min(NT_CHUNK, n - c0 - NT_CHUNK));
        // which particles of the chunk reach this warp's patch (one particle per lane), and at which offsets
        int sx = 0, sy = 0, sz = 0;
        bool take = false;
        if (lane < nc) {
            const int4 l0 = s_cell[buf][lane];
            sx = wrap_off(XW - l0.x, nf, NT_PX);    // offset of the patch's first cell in the support, per axis
            sy = wrap_off(YW - l0.y, nf, NT_PY);
            sz = wrap_off(Z0 - l0.z, nf, NT_T);
            take = sx < W && sy < W && sz < W;
        }
        unsigned m = __ballot_sync(0xffffffffu, take);
        while (m) {
            const int p = __ffs((int)m) - 1;
            m &= m - 1u;
            const int px = __shfl_sync(0xffffffffu, sx, p), py = __shfl_sync(0xffffffffu, sy, p);
            const int pz = __shfl_sync(0xffffffffu, sz, p) + (lane >> 1);            // this lane's offset in the z support
            // vz(z) * strength (re or im) of this lane; zero outside the support
            const double wzc = (unsigned)pz < (unsigned)W ? s_wz[buf][p][2 * pz + (lane & 1)] : 0.0;
            const double *vxp = &s_vx[buf][p][NT_PX + px], *vyp = &s_vy[buf][p][NT_PY + py];
            double vy[NT_PY];
#pragma unroll
            for (int j = 0; j < NT_PY; ++j) vy[j] = vyp[j];
            // (loading the next particle's factors while these FMAs issue, with vy fetched just in time, was tried: 10 %
            // slower alone and 20 % slower next to the pair kernel)
#pragma unroll
            for (int i = 0; i < NT_PX; ++i) {
                const double a = vxp[i] * wzc;
#pragma unroll
                for (int j = 0; j < NT_PY; ++j) acc[i][j] = fma(a, vy[j], acc[i][j]);
            }
        }
    }

    // every cell of the tile is written once; a warp store covers one z line of the patch (256 contiguous bytes)
    double *out = reinterpret_cast<double *>(grid) + ((((size_t)blockIdx.y * nf + XW) * nf + YW) * nf + Z0) * 2 + lane;
#pragma unroll
    for (int i = 0; i < NT_PX; ++i)
#pragma unroll
        for (int j = 0; j < NT_PY; ++j) out[((size_t)i * nf + j) * nf * 2] = acc[i][j];
}

typedef void (*records_fn)(const double *, int64_t, const int2 *, int, int, int64_t, NufftGeom, int, double, double, double,
                           double *, int4 *, int *);
typedef void (*fill_fn)(const int2 *, int, int, int64_t, int, int, const int4 *, int *, int *, int64_t);
typedef void (*tiles_fn)(const double *, const int4 *, const int *, const int *, int64_t, int64_t, int, int, double2 *);
struct TileKernels {
    records_fn records;
    fill_fn fill;
    tiles_fn tiles;
};
template <int W>
TileKernels tile_kernels_of()
{
    return TileKernels{nufft_records<W>, nufft_tile_fill<W>, nufft_spread_tiles<W>};
}
TileKernels tile_kernels(int w)
{
    switch (w) {
    case 4: return tile_kernels_of<4>();
    case 5: return tile_kernels_of<5>();
    case 6: return tile_kernels_of<6>();
    case 7: return tile_kernels_of<7>();
    case 8: return tile_kernels_of<8>();
    case 9: return tile_kernels_of<9>();
    case 10: return tile_kernels_of<10>();
    case 11: return tile_kernels_of<11>();
    case 12: return tile_kernels_of<12>();
    case 13: return tile_kernels_of<13>();
    case 14: return tile_kernels_of<14>();
    case 15: return tile_kernels_of<15>();
    default: return tile_kernels_of<16>();
    }
}

typedef void (*spread_fn)(const double *, int64_t, const int2 *, int, int, const int *, int64_t, NufftGeom, double,
                          double, double, double *);
spread_fn spread_kernel(int w)
{
    switch (w) {
    case 4: return nufft_spread<4>;
    case 5: return nufft_spread<5>;
    case 6: return nufft_spread<6>;
    case 7: return nufft_spread<7>;
    case 8: return nufft_spread<8>;
    case 9: return nufft_spread<9>;
    case 10: return nufft_spread<10>;
    case 11: return nufft_spread<11>;
    case 12: return nufft_spread<12>;
    case 13: return nufft_spread<13>;
    case 14: return nufft_spread<14>;
    case 15: return nufft_spread<15>;
    default: return nufft_spread<16>;
    }
}

// ---- pruned FFT passes --------------------------------------------------------------------------------

// z axis: lines are contiguous.  grid (G, nf, nf, nf) -> out (G, nf, nf, ncp), mode c at column c.
__global__ void __launch_bounds__(NU_THREADS, MDC_NU_FFT_BLOCKS)
nufft_fft_z(const double2 *__restrict__ grid, double2 *__restrict__ out, const double2 *__restrict__ tw_g, NufftGeom g)
{
    extern __shared__ __align__(16) unsigned char nu_smem[];
    double2 *tw = reinterpret_cast<double2 *>(nu_smem);
    const int nf = g.nf;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2 *line = tw + nf + warp * (nf + (nf >> 3));
    // where the pruned read-out finds mode c of a transformed line (digit reversal + padding): once per block
    int *ptab = reinterpret_cast<int *>(tw + nf + (NU_THREADS / 32) * (nf + (nf >> 3)));
    for (int i = threadIdx.x; i < nf; i += NU_THREADS) tw[i] = tw_g[i];
    for (int c = threadIdx.x; c < g.n; c += NU_THREADS) {
        int k = c - g.k0;
        if (k < 0) k += nf;
        ptab[c] = padded(fft_pos(k, g));
    }
    __syncthreads();
    const int64_t lines = (int64_t)nf * nf;
    for (int64_t ln = (int64_t)blockIdx.x * (NU_THREADS / 32) + warp; ln < lines; ln += (int64_t)gridDim.x * (NU_THREADS / 32)) {
        const double2 *src = grid + ((size_t)blockIdx.y * lines + ln) * nf;
        fft_line<MDC_FFT_POW_Z != 0>(line, tw, g, lane, src);   // the first radix stage reads the line from global memory
        double2 *dst = out + ((size_t)blockIdx.y * lines + ln) * g.ncp;
        for (int c = lane; c < g.ncp; c += 32) {
            dst[c] = c < g.n ? line[ptab[c]] : make_double2(0.0, 0.0);
        }
        __syncwarp();
    }
}

// y and x axes: lines are strided.  in[(o * nf + l) * S + i] -> a block takes NU_TC consecutive i of one o,
// transposes them into shared memory (odd pitch: conflict free), one warp per line.
//   !FINAL:  out[(o * n + k) * S + i]                      (y pass: O = nf, S = ncp)
//   FINAL:   i = b * ncp + c, k = a;  part[row_qoff[b n + a] + c] = X * dec[a] dec[b] dec[c]   (x pass: O = 1, S = n ncp)
template <bool FINAL>
__global__ void __launch_bounds__(NU_THREADS, MDC_NU_FFT_BLOCKS)
nufft_fft_s(const double2 *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ tw_g, NufftGeom g,
            int O, int64_t S, const double *__restrict__ dec, const int *__restrict__ row_cnt,
            const int64_t *__restrict__ row_qoff, int64_t out_pitch)
{
    extern __shared__ __align__(16) unsigned char nu_smem[];
    double2 *tw = reinterpret_cast<double2 *>(nu_smem);
    const int nf = g.nf, n = g.n, pitch = nf + (nf >> 3) + 1;   // odd: the transposing stores are conflict free
    double2 *tile = tw + nf;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tiles = (S + NU_TC - 1) / NU_TC;
    const int o = (int)(blockIdx.x / tiles);
    const int64_t i0 = (blockIdx.x - (int64_t)o * tiles) * NU_TC;
    if (FINAL) {
        // a = 0 holds the most kept c of a (b, *) row: skip tiles beyond the q_max sphere
        bool keep = false;
        if (threadIdx.x < NU_TC) {
            const int64_t gi = i0 + threadIdx.x;
            if (gi < S) {
                const int b = (int)(gi / g.ncp), c = (int)(gi - (int64_t)b * g.ncp);
                keep = c < row_cnt[b * n];
            }
        }
        if (!__syncthreads_or(keep)) return;
    }
    for (int i = threadIdx.x; i < nf; i += NU_THREADS) tw[i] = tw_g[i];
    int *ptab = reinterpret_cast<int *>(tile + NU_TC * pitch);   // position of mode k in a transformed line
    for (int k = threadIdx.x; k < n; k += NU_THREADS) {
        int kk = k - g.k0;
        if (kk < 0) kk += nf;
        ptab[k] = padded(fft_pos(kk, g));
    }
    const double2 *src = in + ((size_t)blockIdx.y * O + o) * nf * S;
    for (int e = threadIdx.x; e < nf * NU_TC; e += NU_THREADS) {
        const int l = e / NU_TC, ii = e - l * NU_TC;
        const int64_t gi = i0 + ii;
        tile[ii * pitch + padded(l)] = gi < S ? src[(size_t)l * S + gi] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    for (int ln = warp; ln < NU_TC; ln += NU_THREADS / 32) fft_line<MDC_FFT_POW_S != 0>(tile + ln * pitch, tw, g, lane);
    __syncthreads();
    for (int e = threadIdx.x; e < n * NU_TC; e += NU_THREADS) {
        const int k = e / NU_TC, ii = e - k * NU_TC;
        const int64_t gi = i0 + ii;
        if (gi >= S) continue;
        const double2 v = tile[ii * pitch + ptab[k]];
        if (!FINAL) {
            out[(((size_t)blockIdx.y * O + o) * n + k) * S + gi] = v;
        } else {
            const int b = (int)(gi / g.ncp), c = (int)(gi - (int64_t)b * g.ncp);
            const int row = b * n + k;
            if (c < row_cnt[row]) {
                const double d = dec[k] * dec[b] * dec[c];
                out[(size_t)blockIdx.y * out_pitch + row_qoff[row] + c] = make_double2(v.x * d, v.y * d);
            }
        }
    }
}

// Gauss-Legendre nodes / weights on [-1, 1]
void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &wt)
{
    x.resize(n);
    wt.resize(n);
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (fabs(dz) < 1e-16) break;
        }
        x[i] = -z;
        x[n - 1 - i] = z;
        wt[i] = wt[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

// phihat(xi) = int_{-1}^{1} exp(beta (sqrt(1 - z^2) - 1)) exp(i xi z) dz, with z = sin(theta) (analytic integrand)
double phihat(double xi, double beta, const std::vector<double> &gx, const std::vector<double> &gw)
{
    double acc = 0.0;
    for (size_t i = 0; i < gx.size(); ++i) {
        const double th = 0.25 * M_PI * (gx[i] + 1.0);   // [0, pi / 2]
        acc += gw[i] * exp(beta * (cos(th) - 1.0)) * cos(xi * sin(th)) * cos(th);
    }
    return 2.0 * acc * 0.25 * M_PI;
}

}  // namespace

bool Nufft::supported(int n_points) { return n_points >= 2 && 1.5 * n_points <= 1024.0; }

bool Nufft::geometry(int n_points, double eps, NufftGeom *out)
{
    if (!supported(n_points)) return false;
    // fine grid: the smallest 2^a or 3 * 2^a with nf >= 2 n while the grid is small, nf >= 1.5 n otherwise
    std::vector<int> cand;
    for (int a = 5; a <= 10; ++a) cand.push_back(1 << a);
    for (int a = 4; a <= 8; ++a) cand.push_back(3 << a);
    std::sort(cand.begin(), cand.end());
    auto at_least = [&](double v) {
        for (int c : cand)
            if ((double)c >= v) return c;
        return 0;
    };
    int nf = at_least(2.0 * n_points);
    if (nf == 0 || nf > 160) nf = at_least(1.5 * n_points);
    if (nf == 0) return false;
    NufftGeom g{};
    g.n = n_points;
    g.ncp = (n_points + 1) & ~1;
    g.k0 = n_points / 2;
    g.nf = nf;
    g.r3 = (nf % 3 == 0) ? 1 : 0;
    g.m = g.r3 ? nf / 3 : nf;
    g.log2m = 0;
    while ((1 << g.log2m) < g.m) ++g.log2m;
    const double sigma = (double)nf / n_points;
    const int w = (int)ceil((log(1.0 / eps) + log(5.0)) / (M_PI * sqrt(1.0 - 1.0 / sigma)));
    g.w = std::max(4, std::min(16, w));
    g.beta = 0.975 * M_PI * g.w * (1.0 - 1.0 / (2.0 * sigma));
    *out = g;
    return true;
}

// seconds per frame for n_sets sets holding sum_len particles in total (rates measured on B200)
double Nufft::cost(const NufftGeom &g, int n_sets, int64_t sum_len)
{
    const double nf = g.nf, n = g.n;
    const double bytes = 16.0 * (2.0 * nf * nf * nf + 2.0 * nf * nf * n + 2.0 * nf * n * n);
    return n_sets * (bytes / 3.0e12 + 60e-6) + (double)sum_len * g.w * g.w * g.w / 2.5e11;
}

int Nufft::init(const Ctx *ctx, int n_points, double eps_, int64_t max_set_len)
{
    if (!geometry(n_points, eps_, &g)) return fail(MDC_ERR_UNSUPPORTED, "nufft: n_points out of range");
    eps = eps_;
    maxlen = max_set_len;
    const int nf = g.nf;

    std::vector<double> gx, gw, hdec(n_points);
    gauss_legendre(96, gx, gw);
    for (int k = 0; k < n_points; ++k) {
        const double xi = M_PI * (double)(k - g.k0) * g.w / nf;
        hdec[k] = 1.0 / (0.5 * g.w * phihat(xi, g.beta, gx, gw));
    }
    MDC_CHECK(dec.alloc(n_points));
    MDC_CUDA(cudaMemcpy(dec.p, hdec.data(), n_points * sizeof(double), cudaMemcpyHostToDevice));
    MDC_CHECK(tw.alloc(nf));
    nufft_twiddles<<<(nf + 255) / 256, 256>>>(nf, tw.p);
    MDC_CUDA(cudaGetLastError());

    const size_t ptab_bytes = ((size_t)n_points * sizeof(int) + 15) & ~(size_t)15;   // read-out position table
    smem_z = sizeof(double2) * ((size_t)nf + (size_t)(NU_THREADS / 32) * (nf + nf / 8)) + ptab_bytes;
    smem_s = sizeof(double2) * ((size_t)nf + (size_t)NU_TC * (nf + nf / 8 + 1)) + ptab_bytes;
    if (smem_z > ctx->smem_optin || smem_s > ctx->smem_optin)
        return fail(MDC_ERR_UNSUPPORTED, "nufft: fine grid line does not fit shared memory");
    // per-function attribute shared by every live plan (any fine grid): only ever the device maximum
    MDC_CUDA(raise_smem_limit(nufft_fft_z, ctx->smem_optin));
    MDC_CUDA(raise_smem_limit(nufft_fft_s<false>, ctx->smem_optin));
    MDC_CUDA(raise_smem_limit(nufft_fft_s<true>, ctx->smem_optin));
    const char *env = getenv("MDC_NUFFT_SORT");
    sort = !(env && env[0] == '0');
    // spreading by gather (nufft_spread_tiles) where the fine grid holds enough 16^3 tiles to fill the GPU; small grids
    // keep the scatter kernel (their spreading cost is small, and a support may wrap onto its own tile there).
    // MDC_NUFFT_SPREAD=atomic / tiles in the environment forces one (tiles needs nf >= 64).
    nb = nf / NT_T;
    tiles = nf % NT_T == 0 && nf >= 128;
    const char *sp = getenv("MDC_NUFFT_SPREAD");
    if (sp && sp[0] == 'a') tiles = false;
    if (sp && sp[0] == 't') tiles = nf % NT_T == 0 && nf >= 64;
    return MDC_OK;
}

int Nufft::reserve(int batch_hint)
{
    if (G > 0) return MDC_OK;
    const size_t nf = g.nf;
    size_t per = sizeof(double2) * (nf * nf * nf + nf * nf * g.ncp);
    if (tiles) per += (size_t)maxlen * (4 * g.w * sizeof(double) + sizeof(int4) + NT_MAX_TILES * sizeof(int));   // records, lists
    const size_t budget = (size_t)2 << 30;
    int G_ = (int)std::max<size_t>(1, std::min<size_t>(budget / per, 32));
    G_ = std::max(1, std::min(G_, batch_hint));
    MDC_CHECK(grid.alloc((size_t)G_ * nf * nf * nf));
    MDC_CHECK(buf1.alloc((size_t)G_ * nf * nf * g.ncp));
    if (tiles) {
        MDC_CHECK(bins.alloc((size_t)G_ * 2 * nb * nb * nb));               // per tile: count, then end of its list
        MDC_CHECK(order.alloc((size_t)G_ * maxlen * NT_MAX_TILES));         // the lists (CSR)
        MDC_CHECK(rec.alloc((size_t)G_ * maxlen * 4 * g.w));
        MDC_CHECK(cell.alloc((size_t)G_ * maxlen));
    } else if (sort) {
        MDC_CHECK(bins.alloc((size_t)G_ * nf));
        MDC_CHECK(order.alloc((size_t)G_ * maxlen));
    }
    G = G_;
    return MDC_OK;
}

void Nufft::release()
{
    grid.release();
    buf1.release();
    tw.release();
    dec.release();
    bins.release();
    order.release();
    rec.release();
    cell.release();
    G = 0;
}

// MDC_NUFFT_PROFILE=1 in the environment: device time of every stage of the chain (CUDA events on the plan's stream, so
// also when the chain shares the device with kernels of another stream), printed when the process exits.  Diagnostics.
namespace {
struct StageProfile {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;
    double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    StageProfile()
    {
        const char *e = getenv("MDC_NUFFT_PROFILE");
        on = e && e[0] == '1';
    }
    void mark(cudaStream_t st, int t)
    {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        tag.push_back(t);
        if (t == 0 && ev.size() > 150) flush();   // at a batch boundary (waits for the events: diagnostics only)
    }
    void flush()
    {
        for (size_t i = 1; i < ev.size(); ++i) {
            if (tag[i] == 0) continue;   // 0 = start of a batch
            float m = 0.f;
            cudaEventSynchronize(ev[i]);
            cudaEventElapsedTime(&m, ev[i - 1], ev[i]);
            ms[tag[i]] += m;
        }
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
        tag.clear();
    }
    ~StageProfile()
    {
        if (!on) return;
        flush();
        fprintf(stderr, "MDC_NUFFT_PROFILE ms: sort+records %.3f spread %.3f fft_z %.3f fft_y %.3f fft_x %.3f\n", ms[1], ms[2], ms[3],
                ms[4], ms[5]);
    }
};
StageProfile g_prof;
}  // namespace

int Nufft::density(cudaStream_t st, const double *posd, const double iL[3], int64_t N, const int2 *d_sets, int n_sets, int F,
                   const int *row_cnt, const int64_t *row_qoff, int64_t nq, int P, double2 *part, int *launches)
{
    const int total = F * n_sets;
    MDC_CHECK(reserve(total));
    const size_t nf = g.nf;
    const int n = g.n;
    const unsigned pblocks = (unsigned)ceil_div(maxlen, 256);
    for (int e0 = 0; e0 < total; e0 += G) {
        const unsigned B = (unsigned)std::min(G, total - e0);
        g_prof.mark(st, 0);
        if (tiles) {
            // kernel values once per particle (+ one count per tile its support reaches), CSR lists tile -> particles,
            // then one block per tile gathers: the grid is written exactly once (no clearing pass, no atomics)
            const size_t ntiles = (size_t)2 * nb * nb * nb;
            const int64_t pitch = maxlen * NT_MAX_TILES;
            const TileKernels tk = tile_kernels(g.w);
            MDC_CUDA(cudaMemsetAsync(bins.p, 0, (size_t)B * ntiles * sizeof(int), st));
            tk.records<<<dim3((unsigned)ceil_div(maxlen, NR_PER_BLOCK), B), 128, 0, st>>>(posd, N, d_sets, n_sets, e0, maxlen, g, nb, iL[0],
                                                                               iL[1], iL[2], rec.p, cell.p, bins.p);
            nufft_scan_any<<<B, 128, 0, st>>>((int)ntiles, bins.p);
            tk.fill<<<dim3((unsigned)ceil_div(maxlen, 128), B), 128, 0, st>>>(d_sets, n_sets, e0, maxlen, g.nf, nb, cell.p, bins.p,
                                                                              order.p, pitch);
            g_prof.mark(st, 1);
            tk.tiles<<<dim3((unsigned)ntiles, B), NT_THREADS, 0, st>>>(rec.p, cell.p, bins.p, order.p, pitch, maxlen, g.nf, nb,
                                                                       grid.p);
            *launches += 4;
        } else {
        MDC_CUDA(cudaMemsetAsync(grid.p, 0, (size_t)B * nf * nf * nf * sizeof(double2), st));
        const int *ord = nullptr;
        if (sort) {
            MDC_CUDA(cudaMemsetAsync(bins.p, 0, (size_t)B * nf * sizeof(int), st));
            nufft_bin_count<<<dim3(pblocks, B), 256, 0, st>>>(posd, N, d_sets, n_sets, e0, g.nf, iL[0], bins.p);
            nufft_bin_scan<<<B, 1024, 0, st>>>(g.nf, bins.p);
            nufft_bin_scatter<<<dim3(pblocks, B), 256, 0, st>>>(posd, N, d_sets, n_sets, e0, g.nf, iL[0], bins.p, order.p,
                                                                maxlen);
            *launches += 3;
            ord = order.p;
        }
        const unsigned sblocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(maxlen, NU_THREADS / 32), 148 * 8192 / NU_THREADS));
        spread_kernel(g.w)<<<dim3(sblocks, B), NU_THREADS, 0, st>>>(posd, N, d_sets, n_sets, e0, ord, maxlen, g, iL[0], iL[1],
                                                             iL[2], reinterpret_cast<double *>(grid.p));
        ++*launches;
        }
        g_prof.mark(st, 2);
        const unsigned zblocks = (unsigned)std::min<int64_t>(ceil_div((int64_t)nf * nf, NU_THREADS / 32), 148 * 4096 / NU_THREADS);
        nufft_fft_z<<<dim3(zblocks, B), NU_THREADS, smem_z, st>>>(grid.p, buf1.p, tw.p, g);
        g_prof.mark(st, 3);
        // y pass: buf1 (nf, nf, ncp) -> grid reused as (nf, n, ncp)
        const int64_t ytiles = ceil_div(g.ncp, NU_TC);
        nufft_fft_s<false><<<dim3((unsigned)(nf * ytiles), B), NU_THREADS, smem_s, st>>>(
            buf1.p, grid.p, tw.p, g, (int)nf, (int64_t)g.ncp, nullptr, nullptr, nullptr, 0);
        g_prof.mark(st, 4);
        // x pass + deconvolution + scatter into the plan's wavevector order
        const int64_t S = (int64_t)n * g.ncp, xtiles = ceil_div(S, NU_TC);
        nufft_fft_s<true><<<dim3((unsigned)xtiles, B), NU_THREADS, smem_s, st>>>(
            grid.p, part + (size_t)e0 * P * nq, tw.p, g, 1, S, dec.p, row_cnt, row_qoff, (int64_t)P * nq);
        g_prof.mark(st, 5);
        *launches += 3;
    }
    MDC_CUDA(cudaGetLastError());
    return MDC_OK;
}

}  // namespace mdc
