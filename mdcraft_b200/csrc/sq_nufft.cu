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
//      phi(z) = exp(beta (sqrt(1 - z^2) - 1)), z = (l - s nf) / (w / 2), |z| <= 1  (Barnett et al., FINUFFT);
//   3. pruned FFTs, one axis at a time, in FP64, in shared memory: nf inputs -> the n wanted modes, so the
//      intermediates shrink (nf^3 -> nf^2 n -> nf n^2 -> kept wavevectors);
//   4. divide by the kernel's Fourier transform (per-axis table) while scattering into the plan's wavevector order.
// The relative l2 error is set by (nf / n, w): 1e-9 requested in FP32 mode (the factored FP32 kernel reaches
// 8e-8), 1e-13 in FP64 mode; tools/nufft_model.py holds the NumPy model these parameters were derived with.
//
// Roofline: HBM.  Algorithmic bytes per (frame, set) = the fine grid written and read once plus the pruned
// intermediates: 16 B * (2 nf^3 + 2 nf^2 n + 2 nf n^2 + Nq); the spreading kernel adds N w^3 FP64 reductions
// (RED.E.ADD.F64) that land in L2 because particles are visited in x-plane order.
#include "sq_nufft.cuh"

#include <math.h>
#include <algorithm>
#include <vector>

namespace mdc {

namespace {

constexpr int NU_TC = 8;         // lines per block in the strided FFT passes (8 x 16 B = one 128-byte segment)
constexpr int NU_THREADS = 256;  // 8 warps: one line per warp

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// position, after fft_line, of X[k]
__device__ __forceinline__ int fft_pos(int k, const NufftGeom &g)
{
    int r = 0, j = k;
    if (g.r3) {
        j = k / 3;
        r = k - 3 * j;
    }
    const int br = g.log2m ? (int)(__brev((unsigned)j) >> (32 - g.log2m)) : 0;
    return r * g.m + br;
}

// One warp transforms one line of nf complex doubles in shared memory, in place:
//   X[k] = sum_l x[l] exp(+2 pi i k l / nf),   X[k] ends up at buf[fft_pos(k)].
// Decimation in frequency: an optional radix-3 stage (nf = 3 m) followed by log2(m) radix-2 stages on
// contiguous sub-blocks; the bit-reversed order is undone for free by the pruned read-out.
// tw[t] = exp(+2 pi i t / nf).
__device__ __forceinline__ void fft_line(double2 *buf, const double2 *tw, const NufftGeom &g, int lane)
{
    const int nf = g.nf, m = g.m;
    if (g.r3) {
        // X[3 j + r] = FFT_m(Y_r)[j],  Y_r[t] = W^(r t) sum_q w3^(r q) x[t + q m],  w3 = exp(2 pi i / 3)
        const double hs = 0.8660254037844386;   // sin(2 pi / 3)
        for (int t = lane; t < m; t += 32) {
            const double2 a = buf[t], b = buf[t + m], c = buf[t + 2 * m];
            const double2 s = make_double2(b.x + c.x, b.y + c.y);
            const double2 d = make_double2(b.x - c.x, b.y - c.y);
            const double2 y0 = make_double2(a.x + s.x, a.y + s.y);
            const double2 h = make_double2(a.x - 0.5 * s.x, a.y - 0.5 * s.y);
            // w3 b + w3^2 c = -s/2 + i hs d ;  w3^2 b + w3^4 c = -s/2 - i hs d
            const double2 y1 = make_double2(h.x - hs * d.y, h.y + hs * d.x);
            const double2 y2 = make_double2(h.x + hs * d.y, h.y - hs * d.x);
            buf[t] = y0;
            buf[t + m] = cmul(y1, tw[t]);
            buf[t + 2 * m] = cmul(y2, tw[2 * t]);   // 2 t < 2 m < nf
        }
        __syncwarp();
    }
    int tstep = nf / m;   // nf / L for the current stage
    for (int lg = g.log2m; lg >= 1; --lg, tstep <<= 1) {
        const int half = 1 << (lg - 1);
        for (int t = lane; t < (nf >> 1); t += 32) {
            const int blk = t >> (lg - 1), i = t & (half - 1);
            const int base = (blk << lg) + i;
            const double2 a = buf[base], b = buf[base + half];
            buf[base] = make_double2(a.x + b.x, a.y + b.y);
            const double2 d = make_double2(a.x - b.x, a.y - b.y);
            buf[base + half] = (i == 0) ? d : cmul(d, tw[i * tstep]);
        }
        __syncwarp();
    }
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
// warp-wide RED touches 2 w consecutive doubles; rows (x, y offsets) are walked in a loop.
// posd f64 (F, 3, N): the float32 coordinates widened exactly; iL = 1 / L0.
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
    const int nf = g.nf, w = g.w;
    const double *px = posd + ((int64_t)f * 3 + 0) * N;
    const double *py = posd + ((int64_t)f * 3 + 1) * N;
    const double *pz = posd + ((int64_t)f * 3 + 2) * N;
    double *G = grid + (size_t)blockIdx.y * nf * nf * nf * 2;
    const double hw = 0.5 * (double)w, ihw = 2.0 / (double)w, inf = 1.0 / (double)nf;

    const int lpr = (w <= 8) ? 16 : 32;          // lanes per z row
    const int rpi = 32 / lpr;                    // rows per warp-wide instruction
    const int rsel = lane / lpr, within = lane - rsel * lpr, dz = within >> 1, ri = within & 1;
    const bool act = dz < w;
    const int t = lane & 15;

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
            sv[warp][lane >> 4][t] = t < w ? exp(g.beta * (sqrt(a) - 1.0)) : 0.0;
        }
        {
            const double c = ceil(gz - hw);
            l0z = (int)c;
            const double z = (c + (double)t - gz) * ihw;
            const double a = fmax(1.0 - z * z, 0.0);
            if (lane < 16) sv[warp][2][t] = t < w ? exp(g.beta * (sqrt(a) - 1.0)) : 0.0;
        }
        const int l0x = __shfl_sync(0xffffffffu, l0a, 0), l0y = __shfl_sync(0xffffffffu, l0a, 16);
        // strength exp(2 pi i k0 (sx + sy + sz)) of the centred modes
        double sn, cs;
        sincospi(2.0 * (double)g.k0 * ((gx + gy + gz) * inf), &sn, &cs);
        __syncwarp();
        const double zc = act ? sv[warp][2][dz] * (ri ? sn : cs) : 0.0;
        int Z = l0z + dz;
        Z = Z < 0 ? Z + nf : (Z >= nf ? Z - nf : Z);
        const size_t zoff = (size_t)Z * 2 + ri;
        int ix = 0, iy = rsel;
        for (int r = rsel; r < w * w; r += rpi) {
            int X = l0x + ix, Y = l0y + iy;
            X = X < 0 ? X + nf : (X >= nf ? X - nf : X);
            Y = Y < 0 ? Y + nf : (Y >= nf ? Y - nf : Y);
            const double v = zc * (sv[warp][0][ix] * sv[warp][1][iy]);
            if (act) atomicAdd(G + ((size_t)X * nf + Y) * nf * 2 + zoff, v);
            iy += rpi;
            if (iy >= w) {
                iy -= w;
                ++ix;
            }
        }
        __syncwarp();
    }
}

// ---- pruned FFT passes --------------------------------------------------------------------------------

// z axis: lines are contiguous.  grid (G, nf, nf, nf) -> out (G, nf, nf, ncp), mode c at column c.
__global__ void __launch_bounds__(NU_THREADS)
nufft_fft_z(const double2 *__restrict__ grid, double2 *__restrict__ out, const double2 *__restrict__ tw_g, NufftGeom g)
{
    extern __shared__ __align__(16) unsigned char nu_smem[];
    double2 *tw = reinterpret_cast<double2 *>(nu_smem);
    const int nf = g.nf;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2 *line = tw + nf + warp * nf;
    for (int i = threadIdx.x; i < nf; i += NU_THREADS) tw[i] = tw_g[i];
    __syncthreads();
    const int64_t lines = (int64_t)nf * nf;
    for (int64_t ln = (int64_t)blockIdx.x * (NU_THREADS / 32) + warp; ln < lines; ln += (int64_t)gridDim.x * (NU_THREADS / 32)) {
        const double2 *src = grid + ((size_t)blockIdx.y * lines + ln) * nf;
        for (int i = lane; i < nf; i += 32) line[i] = src[i];
        __syncwarp();
        fft_line(line, tw, g, lane);
        double2 *dst = out + ((size_t)blockIdx.y * lines + ln) * g.ncp;
        for (int c = lane; c < g.ncp; c += 32) {
            double2 v = make_double2(0.0, 0.0);
            if (c < g.n) {
                int k = c - g.k0;
                if (k < 0) k += nf;
                v = line[fft_pos(k, g)];
            }
            dst[c] = v;
        }
        __syncwarp();
    }
}

// y and x axes: lines are strided.  in[(o * nf + l) * S + i] -> a block takes NU_TC consecutive i of one o,
// transposes them into shared memory (pitch nf + 1: conflict free), one warp per line.
//   !FINAL:  out[(o * n + k) * S + i]                      (y pass: O = nf, S = ncp)
//   FINAL:   i = b * ncp + c, k = a;  part[row_qoff[b n + a] + c] = X * dec[a] dec[b] dec[c]   (x pass: O = 1, S = n ncp)
template <bool FINAL>
__global__ void __launch_bounds__(NU_THREADS)
nufft_fft_s(const double2 *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ tw_g, NufftGeom g,
            int O, int64_t S, const double *__restrict__ dec, const int *__restrict__ row_cnt,
            const int64_t *__restrict__ row_qoff, int64_t out_pitch)
{
    extern __shared__ __align__(16) unsigned char nu_smem[];
    double2 *tw = reinterpret_cast<double2 *>(nu_smem);
    const int nf = g.nf, n = g.n, pitch = nf + 1;
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
    const double2 *src = in + ((size_t)blockIdx.y * O + o) * nf * S;
    for (int e = threadIdx.x; e < nf * NU_TC; e += NU_THREADS) {
        const int l = e / NU_TC, ii = e - l * NU_TC;
        const int64_t gi = i0 + ii;
        tile[ii * pitch + l] = gi < S ? src[(size_t)l * S + gi] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    fft_line(tile + warp * pitch, tw, g, lane);
    __syncthreads();
    for (int e = threadIdx.x; e < n * NU_TC; e += NU_THREADS) {
        const int k = e / NU_TC, ii = e - k * NU_TC;
        const int64_t gi = i0 + ii;
        if (gi >= S) continue;
        int kk = k - g.k0;
        if (kk < 0) kk += nf;
        const double2 v = tile[ii * pitch + fft_pos(kk, g)];
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

int Nufft::init(const Ctx *ctx, int n_points, double eps_, int64_t max_set_len)
{
    if (!supported(n_points)) return fail(MDC_ERR_UNSUPPORTED, "nufft: n_points out of range");
    eps = eps_;
    maxlen = max_set_len;
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
    if (nf == 0) return fail(MDC_ERR_UNSUPPORTED, "nufft: n_points out of range");
    g.n = n_points;
    g.ncp = (n_points + 1) & ~1;
    g.k0 = n_points / 2;
    g.nf = nf;
    g.r3 = (nf % 3 == 0) ? 1 : 0;
    g.m = g.r3 ? nf / 3 : nf;
    g.log2m = 0;
    while ((1 << g.log2m) < g.m) ++g.log2m;
    const double sigma = (double)nf / n_points;
    int w = (int)ceil((log(1.0 / eps) + log(5.0)) / (M_PI * sqrt(1.0 - 1.0 / sigma)));
    g.w = std::max(4, std::min(16, w));
    g.beta = 0.975 * M_PI * g.w * (1.0 - 1.0 / (2.0 * sigma));

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

    smem_z = sizeof(double2) * ((size_t)nf + (size_t)(NU_THREADS / 32) * nf);
    smem_s = sizeof(double2) * ((size_t)nf + (size_t)NU_TC * (nf + 1));
    if (smem_z > ctx->smem_optin || smem_s > ctx->smem_optin)
        return fail(MDC_ERR_UNSUPPORTED, "nufft: fine grid line does not fit shared memory");
    MDC_CUDA(cudaFuncSetAttribute(nufft_fft_z, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_z));
    MDC_CUDA(cudaFuncSetAttribute(nufft_fft_s<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    MDC_CUDA(cudaFuncSetAttribute(nufft_fft_s<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    const char *env = getenv("MDC_NUFFT_SORT");
    sort = !(env && env[0] == '0');
    return MDC_OK;
}

int Nufft::reserve(int batch_hint)
{
    if (G > 0) return MDC_OK;
    const size_t nf = g.nf;
    const size_t per = sizeof(double2) * (nf * nf * nf + nf * nf * g.ncp);
    const size_t budget = (size_t)2 << 30;
    int G_ = (int)std::max<size_t>(1, std::min<size_t>(budget / per, 32));
    G_ = std::max(1, std::min(G_, batch_hint));
    MDC_CHECK(grid.alloc((size_t)G_ * nf * nf * nf));
    MDC_CHECK(buf1.alloc((size_t)G_ * nf * nf * g.ncp));
    if (sort) {
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
    G = 0;
}

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
        const unsigned sblocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(maxlen, NU_THREADS / 32), 148 * 32));
        nufft_spread<<<dim3(sblocks, B), NU_THREADS, 0, st>>>(posd, N, d_sets, n_sets, e0, ord, maxlen, g, iL[0], iL[1],
                                                             iL[2], reinterpret_cast<double *>(grid.p));
        const unsigned zblocks = (unsigned)std::min<int64_t>(ceil_div((int64_t)nf * nf, NU_THREADS / 32), 148 * 16);
        nufft_fft_z<<<dim3(zblocks, B), NU_THREADS, smem_z, st>>>(grid.p, buf1.p, tw.p, g);
        // y pass: buf1 (nf, nf, ncp) -> grid reused as (nf, n, ncp)
        const int64_t ytiles = ceil_div(g.ncp, NU_TC);
        nufft_fft_s<false><<<dim3((unsigned)(nf * ytiles), B), NU_THREADS, smem_s, st>>>(
            buf1.p, grid.p, tw.p, g, (int)nf, (int64_t)g.ncp, nullptr, nullptr, nullptr, 0);
        // x pass + deconvolution + scatter into the plan's wavevector order
        const int64_t S = (int64_t)n * g.ncp, xtiles = ceil_div(S, NU_TC);
        nufft_fft_s<true><<<dim3((unsigned)xtiles, B), NU_THREADS, smem_s, st>>>(
            grid.p, part + (size_t)e0 * P * nq, tw.p, g, 1, S, dec.p, row_cnt, row_qoff, (int64_t)P * nq);
        *launches += 4;
    }
    MDC_CUDA(cudaGetLastError());
    return MDC_OK;
}

}  // namespace mdc
