// sq_nufft.cuh — type-1 non-uniform FFT for reciprocal-lattice S(q) plans (see sq_nufft.cu).
#pragma once
#include "common.cuh"

namespace mdc {

struct NufftGeom {
    int n;        // lattice points per axis (modes k = 0 .. n-1)
    int ncp;      // n rounded up to even: row pitch of the pruned intermediates (keeps rows 32-byte aligned)
    int k0;       // centring shift: mode k is computed as k' = k - k0 in [-k0, n-1-k0]
    int nf;       // fine-grid points per axis, 2^a or 3 * 2^a
    int m, log2m; // power-of-two factor of nf
    int r3;       // nf == 3 * m
    int w;        // kernel width in fine-grid points (<= 16)
    double beta;  // exp(beta (sqrt(1 - z^2) - 1)) shape parameter
};

struct Nufft {
    NufftGeom g{};
    double eps = 0;
    int G = 0;                        // (frame, set) elements processed per batch
    int64_t maxlen = 0;               // longest particle set
    bool sort = true;                 // order every set by fine-grid x plane before spreading (L2 locality)
    bool tiles = false;               // spread by GATHER: one block per 16^3 tile of the fine grid, no atomics (nf >= 128)
    int nb = 0;                       // tiles per axis (nf / 16)
    DevBuf<double> rec;               // (G, maxlen, 4 w) per-particle kernel values {vx(w), vy(w), wz(w) complex}
    DevBuf<int4> cell;                // (G, maxlen) first fine-grid cell of every particle's support {l0x, l0y, l0z, -}
    size_t smem_z = 0, smem_s = 0;
    DevBuf<double2> grid, buf1, tw;   // (G, nf, nf, nf) [reused as (G, nf, n, ncp)], (G, nf, nf, ncp), (nf)
    DevBuf<double> dec;               // (n) deconvolution factors 1 / ((w / 2) phihat(pi k' w / nf))
    DevBuf<int> bins, order;          // (G, nf), (G, maxlen)

    // chooses nf / w for `n_points` modes per axis and the requested relative accuracy
    static bool supported(int n_points);
    static bool geometry(int n_points, double eps, NufftGeom *out);
    static double cost(const NufftGeom &g, int n_sets, int64_t sum_len);   // seconds per frame (model)
    int init(const Ctx *ctx, int n_points, double eps, int64_t max_set_len);
    int reserve(int batch_hint);      // allocates the work buffers (first push)
    void release();

    // posd f64 (F, 3, N): float32 coordinates widened exactly (no fixed-point quantisation); iL = 1 / L0.
    // rho(q) of F x n_sets (frame, set) elements on the lattice part of the plan's wavevector list:
    //   part[((f * n_sets + s) * P + 0) * nq + row_qoff[b * n + a] + c]   for c < row_cnt[b * n + a]
    int density(cudaStream_t st, const double *posd, const double iL[3], int64_t N, const int2 *d_sets, int n_sets, int F,
                const int *row_cnt, const int64_t *row_qoff, int64_t nq, int P, double2 *part, int *launches);
};

}  // namespace mdc
