/*
 * oracle.c — CPU restatement of MDCraft's S(q) / g(r) per-frame arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mdcraft_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.  It exists to check the CUDA path,
 * never to stand in for it.
 *
 * Parity status:
 *   S(q)  — pinned: checked against the reference's own Numba kernels
 *           (structure.py:1240-1280) run in the build container; golden
 *           outputs are committed under tests/golden/ (make_golden.py).
 *   g(r)  — binning (accelerated.py:46-81) pinned the same way.  The pair
 *           distances come from MDAnalysis.lib.distances.capped_distance,
 *           a third-party C routine that is NOT under /root/reference and is
 *           not installed (requirements.txt:3 "mdanalysis >= 2.2.0", no lock
 *           file).  Its brute-force orthorhombic arithmetic is restated below
 *           from the published algorithm (calc_distances.h:
 *           _calc_distance_array_ortho + minimum_image; for triclinic cells
 *           _calc_distance_array_triclinic = _triclinic_pbc +
 *           minimum_image_triclinic, and mdamath.triclinic_vectors for the
 *           box matrix).  That part is "parity unpinned" against a live
 *           MDAnalysis.
 *
 * Build: see oracle/Makefile (-O2 -fopenmp -ffp-contract=off, no fast-math:
 * FMA contraction would change r^2 in the last bit).
 *
 * All reference citations are relative to /root/reference/src/mdcraft/.
 */

#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* S(q)                                                                     */
/* ------------------------------------------------------------------------ */

/*
 * analysis/structure.py:1240-1280  delta_fourier_transform_sum
 *   F[i] = sum_j exp(1j * dot(q_i, r_j)),  prange over i, sequential complex128
 *   accumulation over j; dot = a0*b0 + a1*b1 + a2*b2 (algorithm/accelerated.py:113).
 * out is interleaved (re, im).
 */
void orc_delta_fourier_transform_sum(const double *qs, int64_t nq,
                                     const double *rs, int64_t n,
                                     double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        const double q0 = qs[3 * i], q1 = qs[3 * i + 1], q2 = qs[3 * i + 2];
        double re = 0.0, im = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            const double x = q0 * rs[3 * j] + q1 * rs[3 * j + 1] + q2 * rs[3 * j + 2];
            double s, c;
            sincos(x, &s, &c);
            re += c;
            im += s;
        }
        out[2 * i] = re;
        out[2 * i + 1] = im;
    }
}

/*
 * analysis/structure.py:1940-1970  StructureFactor._single_frame, form="exp".
 *   positions : (N,3) float64, groups concatenated in order (structure.py:1943)
 *   offsets   : n_groups+1 prefix offsets (the reference's _slices)
 *   pairs     : n_pairs x 2 group indices; (-1,-1) means mode=None (all atoms)
 *   ssf       : (n_pairs, nq) accumulated in place
 * The reference recomputes rho_j per pair; the values are identical, so each
 * group's rho is evaluated once here.
 */
void orc_sq_single_frame(const double *qs, int64_t nq, const double *positions,
                         const int64_t *offsets, int n_groups,
                         const int *pairs, int n_pairs, double *ssf)
{
    double *rho = (double *)malloc(sizeof(double) * 2 * (size_t)nq * (size_t)(n_groups + 1));
    int total_needed = 0;
    for (int p = 0; p < n_pairs; ++p)
        if (pairs[2 * p] < 0) total_needed = 1;
    for (int g = 0; g < n_groups; ++g)
        orc_delta_fourier_transform_sum(qs, nq, positions + 3 * offsets[g],
                                        offsets[g + 1] - offsets[g],
                                        rho + 2 * (size_t)nq * g);
    if (total_needed)
        orc_delta_fourier_transform_sum(qs, nq, positions, offsets[n_groups],
                                        rho + 2 * (size_t)nq * n_groups);
    for (int p = 0; p < n_pairs; ++p) {
        const int j = pairs[2 * p], k = pairs[2 * p + 1];
        const double *rj = rho + 2 * (size_t)nq * (j < 0 ? n_groups : j);
        const double *rk = rho + 2 * (size_t)nq * (k < 0 ? n_groups : k);
        double *out = ssf + (size_t)p * nq;
        if (j == k) {
            for (int64_t i = 0; i < nq; ++i)   /* (rho * conj(rho)).real */
                out[i] += rj[2 * i] * rj[2 * i] + rj[2 * i + 1] * rj[2 * i + 1];
        } else {
            for (int64_t i = 0; i < nq; ++i)   /* 2 * (rho_j * conj(rho_k)).real */
                out[i] += 2.0 * (rj[2 * i] * rk[2 * i] + rj[2 * i + 1] * rk[2 * i + 1]);
        }
    }
    free(rho);
}

/* ------------------------------------------------------------------------ */
/* g(r)                                                                     */
/* ------------------------------------------------------------------------ */

/*
 * algorithm/accelerated.py:13-43  numba_histogram_bin_edges, as Numba 0.65 /
 * LLVM lowers  min + i * ((max - min) / n)  under fastmath=True on the build
 * container's x86-64 (FMA) host, probed with inspect_llvm and confirmed
 * bit-for-bit on 7 ranges:
 *     edges[i] = fma((max - min) * i, 1.0 / n, min)        edges[n] = max
 * (the division is turned into a hoisted reciprocal, the product is
 * re-associated, and the final multiply-add is fused).
 */
void orc_histogram_bin_edges(double min_, double max_, int n_bins, double *edges)
{
    const double w = max_ - min_, inv = 1.0 / (double)n_bins;
    for (int i = 0; i <= n_bins; ++i) edges[i] = fma(w * (double)i, inv, min_);
    edges[n_bins] = max_;
}

/*
 * algorithm/accelerated.py:46-81  numba_histogram, as Numba 0.65 / LLVM lowers
 * it under fastmath=True in the build container (SURVEY.md §7 hard part 1,
 * probed with inspect_llvm):  int(n*(x-min)/(max-min))  becomes
 *     fptosi( ((x - min) * (double)n) * (1.0 / (max - min)) )
 * i.e. reciprocal-multiply with the reciprocal hoisted out of the loop, and
 * truncation toward zero.  x == max goes to the last bin.
 *
 * CAVEAT (measured, tests/golden/make_golden.py records it): that is the
 * arithmetic of the optimised LLVM IR and of the scalar remainder loop.  In
 * the 4x-unrolled main loop the x86 backend re-associates the two `fast`
 * multiplies differently in three of the four slots, so for a value within
 * one ulp of a bin edge the reference's OWN result depends on the value's
 * position in the array (np.full(4, v) lands 1 + 3 in two bins).  Such values
 * have no well-defined reference bin; the oracle (and the CUDA path) define
 * the bin by the IR formula above.  Everywhere else the two agree exactly.
 * Returns the bin, or -1 if the value is not counted.
 */
static inline int orc_bin_of(double x, double min_, double max_, int n_bins,
                             double n_as_double, double inv_width)
{
    if (x == max_) return n_bins - 1;
    const double t = ((x - min_) * n_as_double) * inv_width;
    /* fptosi is undefined out of range; the callers only pass finite,
       bounded distances. */
    const int64_t b = (int64_t)t;
    return (b >= 0 && b < n_bins) ? (int)b : -1;
}

void orc_histogram(const double *x, int64_t n, int n_bins, double min_,
                   double max_, int64_t *hist)
{
    const double nd = (double)n_bins, inv = 1.0 / (max_ - min_);
    for (int64_t i = 0; i < n; ++i) {
        const int b = orc_bin_of(x[i], min_, max_, n_bins, nd, inv);
        if (b >= 0) hist[b] += 1;
    }
}

/*
 * MDAnalysis.lib.distances (third party; restated from the published source,
 * lib/include/calc_distances.h):
 *
 *   static void minimum_image(double *x, float *box, float *inverse_box) {
 *       for i in 0..2: if (box[i] > FLT_EPSILON) {
 *           s = inverse_box[i] * x[i];  x[i] = box[i] * (s - cround(s)); } }
 *
 *   _calc_distance_array_ortho:  inverse_box[i] = 1.0 / box[i]   (stored float)
 *       dx[k] = conf[j][k] - ref[i][k]   (float subtract, widened to double)
 *       minimum_image(dx, box, inverse_box)
 *       rsq = dx0*dx0 + dx1*dx1 + dx2*dx2 ;  d = sqrt(rsq)
 *
 * Coordinates and box are float32 (MDAnalysis coerces them).  round() is C
 * round (half away from zero).
 */
static inline double orc_pair_r2(const float *a, const float *b, const float box[3],
                                 const float inv[3])
{
    double r2 = 0.0;
    for (int k = 0; k < 3; ++k) {
        double dx = (double)(b[k] - a[k]);
        if (box[k] > FLT_EPSILON) {
            const double s = (double)inv[k] * dx;
            dx = (double)box[k] * (s - round(s));
        }
        const double sq = dx * dx;
        r2 = (k == 0) ? sq : r2 + sq;
    }
    return r2;
}

/*
 * Triclinic cells (any box angle != 90), restated from the published MDAnalysis 2.x source:
 *
 *   lib/mdamath.py triclinic_vectors(dimensions, dtype=float32):  lower-triangular box matrix, formed in
 *       float64 and rounded to float32:  a = (lx, 0, 0);  b = (ly cos g, ly sin g, 0);
 *       c = (lz cos b, lz (cos a - cos b cos g) / sin g, sqrt(lz^2 - c_x^2 - c_y^2)); cos(90 deg) is exactly 0.
 *   lib/include/calc_distances.h _calc_distance_array_triclinic:
 *       _triclinic_pbc(ref), _triclinic_pbc(conf): every coordinate is moved into the primary cell (c axis first,
 *         then b, then a; shifts by whole box vectors, arithmetic in double, result stored as float);
 *       dx = conf[j] - ref[i]   (float subtract, widened to double);
 *       minimum_image_triclinic(dx, box): the shortest of the 27 images dx + ix a + iy b + iz c, ix, iy, iz in
 *         {-1, 0, 1} (rsq = rz0^2 + rz1^2 + rz2^2, left to right, first minimum kept);
 *       d = sqrt(rsq).
 * m = {a_x, b_x, b_y, c_x, c_y, c_z} as float32 values.
 */
int orc_triclinic_vectors(const float *box6, float m[6])
{
    const double lx = box6[0], ly = box6[1], lz = box6[2];
    const double alpha = box6[3], beta = box6[4], gamma = box6[5];
    const double deg = 0.017453292519943295;   /* numpy.deg2rad: x * (pi / 180) */
    memset(m, 0, 6 * sizeof(float));
    if (!(lx > 0 && ly > 0 && lz > 0 && alpha > 0 && beta > 0 && gamma > 0 && alpha < 180 && beta < 180 && gamma < 180))
        return -1;
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
    if (!(cz > 0.0)) return -1;
    m[0] = (float)lx;
    m[1] = (float)bx;
    m[2] = (float)by;
    m[3] = (float)cx;
    m[4] = (float)cy;
    m[5] = (float)cz;
    return 0;
}

/* _triclinic_pbc for one coordinate, in place */
void orc_triclinic_wrap(float *p, const float m[6])
{
    const double ax = m[0], bx = m[1], by = m[2], cx = m[3], cy = m[4], cz = m[5];
    const double bi0 = 1.0 / ax, bi4 = 1.0 / by, bi8 = 1.0 / cz;
    const double a_y = bx * bi4, b_z = cy * bi8;
    const double a_z = (cx - cy * a_y) * bi8;
    double x = p[0], y = p[1], z = p[2];
    if (z < 0.0 || z >= cz) {
        const double s = floor(z * bi8);
        z -= s * cz;
        y -= s * cy;
        x -= s * cx;
    }
    double lb = z * b_z;
    if (y < lb || y >= lb + by) {
        const double s = floor((y - lb) * bi4);
        y -= s * by;
        x -= s * bx;
    }
    lb = y * a_y + z * a_z;
    if (x < lb || x >= lb + ax) {
        const double s = floor((x - lb) * bi0);
        x -= s * ax;
    }
    p[0] = (float)x;
    p[1] = (float)y;
    p[2] = (float)z;
}

/* minimum_image_triclinic + rsq for one pair of WRAPPED coordinates */
static inline double orc_pair_r2_triclinic(const float *a, const float *b, const float m[6])
{
    const double d0 = (double)(b[0] - a[0]), d1 = (double)(b[1] - a[1]), d2 = (double)(b[2] - a[2]);
    double best = (double)FLT_MAX;
    for (int ix = -1; ix < 2; ++ix) {
        const double rx = d0 + (double)(m[0] * (float)ix);
        for (int iy = -1; iy < 2; ++iy) {
            const double ry0 = rx + (double)(m[1] * (float)iy);
            const double ry1 = d1 + (double)(m[2] * (float)iy);
            for (int iz = -1; iz < 2; ++iz) {
                const double rz0 = ry0 + (double)(m[3] * (float)iz);
                const double rz1 = ry1 + (double)(m[4] * (float)iz);
                const double rz2 = d2 + (double)(m[5] * (float)iz);
                const double dsq = rz0 * rz0 + rz1 * rz1 + rz2 * rz2;
                if (dsq < best) best = dsq;
            }
        }
    }
    return best;
}

/* exported so tests can pin single pairs (orthorhombic, or triclinic when an angle of box[3..5] is not 90) */
double orc_pair_distance(const float *a, const float *b, const float *box)
{
    float inv[3];
    for (int k = 0; k < 3; ++k) inv[k] = (float)(1.0 / (double)box[k]);
    return sqrt(orc_pair_r2(a, b, box, inv));
}

double orc_pair_distance_triclinic(const float *a, const float *b, const float *box6)
{
    float m[6], pa[3] = {a[0], a[1], a[2]}, pb[3] = {b[0], b[1], b[2]};
    if (orc_triclinic_vectors(box6, m) != 0) return NAN;
    orc_triclinic_wrap(pa, m);
    orc_triclinic_wrap(pb, m);
    return sqrt(orc_pair_r2_triclinic(pa, pb, m));
}

/*
 * analysis/structure.py:34-131  radial_histogram, streamed (the pair list and
 * the distance array of capped_distance are never materialised):
 *   keep ordered pairs (i in a, j in b) with  min_cutoff < d <= max_cutoff,
 *     min_cutoff = range_[0] - DBL_EPSILON            (structure.py:109-115)
 *   drop pairs with i / ex0 == j / ex1                 (structure.py:119-122)
 *   bin with numba_histogram over edges[0]=range_[0], edges[-1]=range_[1].
 * ex0 <= 0 means "no exclusion".  hist is accumulated in place.
 * triclinic != 0: box holds six values and the triclinic routine above is used.
 */
static void orc_radial_histogram_impl(const float *pos_a, int64_t na, const float *pos_b,
                                      int64_t nb, const float *box, int triclinic, int n_bins, double r0,
                                      double r1, int64_t ex0, int64_t ex1, int64_t *hist)
{
    float inv[3], bx[3], m[6];
    for (int k = 0; k < 3; ++k) {
        bx[k] = box[k];
        inv[k] = (float)(1.0 / (double)box[k]);
    }
    float *wa = NULL, *wb = NULL;
    if (triclinic) {
        if (orc_triclinic_vectors(box, m) != 0) return;
        wa = (float *)malloc(sizeof(float) * 3 * (size_t)na);
        wb = (float *)malloc(sizeof(float) * 3 * (size_t)nb);
        memcpy(wa, pos_a, sizeof(float) * 3 * (size_t)na);
        memcpy(wb, pos_b, sizeof(float) * 3 * (size_t)nb);
        for (int64_t i = 0; i < na; ++i) orc_triclinic_wrap(wa + 3 * i, m);
        for (int64_t j = 0; j < nb; ++j) orc_triclinic_wrap(wb + 3 * j, m);
        pos_a = wa;
        pos_b = wb;
    }
    const double min_cut = r0 - DBL_EPSILON;
    const double nd = (double)n_bins, invw = 1.0 / (r1 - r0);
    const double r2_hi = r1 * r1 * (1.0 + 1e-12);   /* cheap prefilter only */

#pragma omp parallel
    {
        int64_t *local = (int64_t *)calloc((size_t)n_bins, sizeof(int64_t));
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < na; ++i) {
            const float *a = pos_a + 3 * i;
            const int64_t bi = ex0 > 0 ? i / ex0 : -1;
            for (int64_t j = 0; j < nb; ++j) {
                if (ex0 > 0 && bi == j / ex1) continue;
                const double r2 = triclinic ? orc_pair_r2_triclinic(a, pos_b + 3 * j, m)
                                            : orc_pair_r2(a, pos_b + 3 * j, bx, inv);
                if (r2 > r2_hi) continue;
                const double d = sqrt(r2);
                if (!(d <= r1 && d > min_cut)) continue;
                const int b = orc_bin_of(d, r0, r1, n_bins, nd, invw);
                if (b >= 0) local[b] += 1;
            }
        }
#pragma omp critical
        for (int b = 0; b < n_bins; ++b) hist[b] += local[b];
        free(local);
    }
    free(wa);
    free(wb);
}

void orc_radial_histogram(const float *pos_a, int64_t na, const float *pos_b,
                          int64_t nb, const float *box, int n_bins, double r0,
                          double r1, int64_t ex0, int64_t ex1, int64_t *hist)
{
    orc_radial_histogram_impl(pos_a, na, pos_b, nb, box, 0, n_bins, r0, r1, ex0, ex1, hist);
}

void orc_radial_histogram_triclinic(const float *pos_a, int64_t na, const float *pos_b,
                                    int64_t nb, const float *box6, int n_bins, double r0,
                                    double r1, int64_t ex0, int64_t ex1, int64_t *hist)
{
    orc_radial_histogram_impl(pos_a, na, pos_b, nb, box6, 1, n_bins, r0, r1, ex0, ex1, hist);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
