/*
 * mdcraft_b200.h — C ABI of libmdcraft_b200.so
 *
 * B200-native (sm_100a) replacement for the per-frame arithmetic of
 * MDCraft's `mdcraft.analysis.structure` S(q) and g(r) path.  The two pure
 * functions below the reference's analysis classes are the cut:
 *
 *   delta_fourier_transform_sum(qs f64[Nq,3], rs f64[N,3]) -> c128[Nq]
 *       /root/reference/src/mdcraft/analysis/structure.py:1240-1280
 *   radial_histogram(pos_a, pos_b, n_bins, range_, dimensions, exclusion) -> intp[B]
 *       /root/reference/src/mdcraft/analysis/structure.py:34-131
 *
 * plus the per-frame accumulation the classes wrap around them
 * (StructureFactor._single_frame structure.py:1940-1991,
 *  RadialDistributionFunction._single_frame structure.py:868-920), which is
 * exposed as "plans" so that frame blocks can be pushed without a host round
 * trip per frame.
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes; no torch / C++ types.
 *   - every entry returns int: 0 = MDC_OK, negative = error; mdc_last_error()
 *     returns a thread-local message.  No C++ exception crosses the ABI.
 *   - "pos" arrays are float32 AoS (F, N, 3), exactly what MDAnalysis'
 *     AtomGroup.positions yields per frame; `on_device` != 0 means the pointer
 *     is a CUDA device pointer on the plan's device, else host memory (pinned
 *     memory makes the copy asynchronous).
 *   - a context owns two CUDA streams (copies + layout kernels / analysis
 *     kernels) shared by its plans; a context and its plans are single-threaded.
 *     Buffers passed to *_push may be reused once the call returns (the library
 *     stages them before returning); the analysis kernels keep running
 *     asynchronously until *_read / mdc_synchronize.
 *   - the library owns all device memory for the lifetime of the plan; the
 *     *_accum_ptr calls expose the on-device accumulators so a caller can hand
 *     them to NCCL (frame-sharded multi-GPU runs all-reduce them once).
 */
#ifndef MDCRAFT_B200_H
#define MDCRAFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDC_OK                 0
#define MDC_ERR_INVALID       -1   /* bad argument */
#define MDC_ERR_CUDA          -2   /* CUDA runtime error (message has details) */
#define MDC_ERR_NOMEM         -3
#define MDC_ERR_UNSUPPORTED   -4   /* e.g. a dump layout the GPU parser does not take */

#define MDC_PRECISION_FP32     0   /* phases exact (fixed point / FP64), sin/cos on the SFU, FP32 partial sums */
#define MDC_PRECISION_FP64     1   /* everything FP64 (reference arithmetic) */

/* S(q) kernel selection for reciprocal-lattice plans in FP32 precision. */
#define MDC_SQ_ALGO_AUTO       0   /* FACTORED or NUFFT for the lattice part, DIRECT arithmetic for explicit wavevectors */
#define MDC_SQ_ALGO_DIRECT     1   /* one fused sin/cos per (q, particle) term on the SFU (2 MUFU ops per term) */
#define MDC_SQ_ALGO_FACTORED   3   /* per-axis unit-phasor tables + one complex FMA per term (4 FP32 FMAs per term) */
#define MDC_SQ_ALGO_NUFFT      4   /* gridding type-1 NUFFT in FP64: O(N w^3 + nf^3 log nf) instead of O(Nq N);
                                      relative l2 error 1e-9 (FP32 precision) / 1e-13 (FP64 precision); also valid
                                      with MDC_PRECISION_FP64.  AUTO picks it when its cost model says it is faster.
                                      MDC_NUFFT_EPS in the environment overrides the FP32-precision request
                                      (1e-12 .. 1e-4; 1e-7 is 23 % faster and stays inside 1e-5). */

typedef struct mdc_ctx      mdc_ctx;
typedef struct mdc_sq_plan  mdc_sq_plan;
typedef struct mdc_rdf_plan mdc_rdf_plan;
typedef struct mdc_dump     mdc_dump;
typedef struct mdc_profile_plan mdc_profile_plan;
typedef struct mdc_ncdf     mdc_ncdf;

/* ---- library / context ------------------------------------------------- */

int          mdc_version(void);
const char  *mdc_last_error(void);

/* Binds a context to CUDA device `device` (creates nothing on other devices). */
int  mdc_create(int device, mdc_ctx **out);
/*
 * The same with a stream priority: plans of a high-priority context get their thread blocks placed first whenever
 * blocks of another context's kernels retire.  Meant for the short, memory-bound S(q) kernels running next to the
 * long, issue-bound g(r) pair kernel of another context on the same device.
 */
int  mdc_create_with_priority(int device, int high_priority, mdc_ctx **out);
int  mdc_destroy(mdc_ctx *ctx);
int  mdc_device_info(mdc_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                     int *sm_clock_khz, int64_t *global_mem_bytes);
/* Blocks until all work queued by plans of this context has finished. */
int  mdc_synchronize(mdc_ctx *ctx);

/* ---- S(q): StructureFactor ---------------------------------------------- */

/*
 * Creates an S(q) plan.
 *
 * Wavevectors = [reciprocal-lattice part] ++ [explicit part], in that order
 * (the order the reference stacks them, structure.py:1837-1872):
 *   lattice part (n_points > 0): q = 2*pi*(a/L0[0], b/L0[1], c/L0[2]),
 *     a,b,c in [0, n_points), flat order b*n^2 + a*n + c (np.meshgrid 'xy'),
 *     filtered by |q| <= q_max when q_max >= 0 (structure.py:1883-1887).  The
 *     grid is generated on the device with the reference's FP64 operation
 *     order; nothing of size Nq is uploaded.
 *   explicit part (n_explicit > 0): q_explicit f64 (n_explicit, 3), used as is
 *     (user `wavevectors=` or the off-lattice `n_surfaces` points); the caller
 *     applies q_max to these.
 *
 * Groups: positions pushed later hold the groups concatenated in order;
 *   group g has group_sizes[g] particles (the reference's _slices,
 *   structure.py:1889-1899).
 * Pairs: n_pairs x 2 group indices (j, k); (-1, -1) means "all particles as
 *   one group" (mode=None).  j == k accumulates |rho_j|^2, j != k accumulates
 *   2 Re(rho_j conj(rho_k))  (structure.py:1953-1970).
 */
int  mdc_sq_plan_create(mdc_ctx *ctx,
                        int n_points, const double L0[3], double q_max,
                        const double *q_explicit, int64_t n_explicit,
                        int n_groups, const int64_t *group_sizes,
                        int n_pairs, const int *pairs,
                        int precision, int algo,
                        mdc_sq_plan **out);
int  mdc_sq_plan_destroy(mdc_sq_plan *plan);

/* Number of wavevectors kept (lattice after q_max + explicit). */
int64_t mdc_sq_num_wavevectors(mdc_sq_plan *plan);
/*
 * What the plan decided: info[0] = kernel family of the lattice part (MDC_SQ_ALGO_DIRECT / FACTORED / NUFFT, 0 =
 * FP64 reference arithmetic or no lattice part), info[1] = NUFFT fine-grid points per axis, info[2] = NUFFT kernel
 * width, info[3] = n_points, info[4] = density sets per frame, info[5] = particle chunks per set, info[6] = frames
 * per block, info[7] = (frame, set) elements per NUFFT batch.  info[5..7] are 0 before the first push.
 */
int  mdc_sq_plan_info(mdc_sq_plan *plan, int info[8]);
/* Copies the device-built wavevectors, f64 (Nq, 3), to host memory. */
int  mdc_sq_wavevectors(mdc_sq_plan *plan, double *q_out);

/*
 * Accumulates n_frames frames:  ssf[p, :] += per-frame pair products.
 * pos: f32 (n_frames, N, 3) with N = sum(group_sizes).  Asynchronous with
 * respect to the device; synchronous with respect to the host buffer.
 */
int  mdc_sq_push(mdc_sq_plan *plan, const float *pos, int n_frames, int on_device);
/*
 * The same with f64 positions (n_frames, N, 3).  structure.py:1943-1944 gathers the groups into a float64
 * buffer: atom positions are float32 values widened, but centres of mass (groupings="residues",
 * algorithm/molecule.py:16-441) are genuine float64 and q . r is evaluated on them in FP64.
 */
int  mdc_sq_push_f64(mdc_sq_plan *plan, const double *pos, int n_frames, int on_device);

/* Un-normalised sums f64 (n_pairs, Nq) and the number of frames pushed. */
int  mdc_sq_read(mdc_sq_plan *plan, double *ssf_out, int64_t *n_frames_out);
/*
 * |q|-shell averages on the device (StructureFactor._conclude with unique=True, structure.py:2000-2009: for every
 * unique wavenumber the mean of the columns np.isclose selects).  mdc_sq_set_shells: `order` = wavevector indices
 * sorted by |q| (n_order of them), shell s owns the window order[lo[s] : hi[s]] (windows may overlap, as
 * np.isclose's do).  mdc_sq_read_shells: out f64 (n_pairs, n_shells) = mean over the window of sums / denom
 * (denom = n_frames * N), so that only N_unique values per pair cross the bus instead of Nq.
 */
int  mdc_sq_set_shells(mdc_sq_plan *plan, const int64_t *order, int64_t n_order,
                       const int64_t *lo, const int64_t *hi, int64_t n_shells);
int  mdc_sq_read_shells(mdc_sq_plan *plan, double denom, double *out, int64_t *n_frames_out);
int  mdc_sq_reset(mdc_sq_plan *plan);
/* Device pointer / element count of the f64 accumulator (for NCCL). */
int  mdc_sq_accum_ptr(mdc_sq_plan *plan, void **dev_ptr, int64_t *n_elems);
/* Overrides the frame counter (after an external all-reduce). */
int  mdc_sq_set_frames(mdc_sq_plan *plan, int64_t n_frames);

/*
 * One-shot functional seam: rho[i] = sum_j exp(i q_i . r_j)
 * (structure.py:1240-1280).  qs f64 (nq,3), rs f64 (n,3) host arrays,
 * rho_out f64 (nq, 2) interleaved (re, im) == complex128.
 */
int  mdc_delta_fourier_transform_sum(mdc_ctx *ctx, const double *qs, int64_t nq,
                                     const double *rs, int64_t n,
                                     int precision, double *rho_out);

/* ---- single-chain structure factor ---------------------------------------- */

/*
 * polymer.py:1054-1445: positions pushed later hold n_chains
 * chains of n_monomers consecutive particles; the plan accumulates
 *   scsf[q] += sum over chains of |rho_chain(q)|^2
 * into ONE row (mdc_sq_read returns f64 (1, Nq)).  Push / read / wavevector queries /
 * accumulator access are the mdc_sq_* calls.
 */
int  mdc_scsf_plan_create(mdc_ctx *ctx,
                          int n_points, const double L0[3], double q_max,
                          const double *q_explicit, int64_t n_explicit,
                          int64_t n_chains, int64_t n_monomers,
                          int precision, int algo,
                          mdc_sq_plan **out);

/* ---- F(q, t): IntermediateScatteringFunction ------------------------------- */

/*
 * Coherent and (optionally) incoherent intermediate scattering functions,
 * structure.py:2198-2835.  Same wavevector / group / pair
 * arguments as mdc_sq_plan_create; the handle IS an mdc_sq_plan (wavevector
 * queries work on it).  Frames must be pushed in time order:
 *   cisf[lag, p, q] += Re(rho_j(t - lag) conj rho_k(t)) (+ j <-> k for cross pairs)
 *   iisf[lag, g, q] += Re sum_i exp(i q . (r_i(t) - r_i(t - lag)))   for pairs (g, g)
 * over a ring of the last n_lags frames kept on the device.  iisf has one row per
 * group (one row when pairs = {(-1, -1)}).  Frames are not independent, so this
 * plan does not shard across GPUs.
 */
int  mdc_isf_plan_create(mdc_ctx *ctx,
                         int n_points, const double L0[3], double q_max,
                         const double *q_explicit, int64_t n_explicit,
                         int n_groups, const int64_t *group_sizes,
                         int n_pairs, const int *pairs,
                         int n_lags, int incoherent,
                         int precision, int algo,
                         mdc_sq_plan **out);
int  mdc_isf_push(mdc_sq_plan *plan, const float *pos, int n_frames, int on_device);
/* Un-normalised sums: cisf f64 (n_lags, n_pairs, Nq); iisf f64 (n_lags, n_rows, Nq) or NULL. */
int  mdc_isf_read(mdc_sq_plan *plan, double *cisf_out, double *iisf_out, int64_t *n_frames_out);

/* ---- g(r): RadialDistributionFunction ------------------------------------ */

/*
 * Creates a g(r) plan for ordered pairs (i in a, j in b).
 *   n_bins, r0, r1 : histogram of distances with r0 - DBL_EPSILON < d <= r1,
 *                    bins as numba_histogram (algorithm/accelerated.py:46-81).
 *   same_group     : b is a (group_b=None in the reference); pos_b is ignored.
 *   excl0, excl1   : drop pairs with i / excl0 == j / excl1 (structure.py:119);
 *                    excl0 <= 0 means no exclusion.
 * Counts are exact integers, independent of launch geometry.
 */
int  mdc_rdf_plan_create(mdc_ctx *ctx, int n_bins, double r0, double r1,
                         int64_t n_a, int64_t n_b, int same_group,
                         int64_t excl0, int64_t excl1,
                         mdc_rdf_plan **out);
int  mdc_rdf_plan_destroy(mdc_rdf_plan *plan);

/*
 * Accumulates n_frames frames.  pos_a f32 (n_frames, n_a, 3); pos_b f32
 * (n_frames, n_b, 3) or NULL when same_group; box f32 (n_frames, 6)
 * [lx, ly, lz, alpha, beta, gamma] (host memory always).  The reference hands all six values to capped_distance
 * (structure.py:109-115).  Frames go through the FP32 filter kernel, which resolves the pairs it cannot decide with the
 * reference's FP64 arithmetic: orthorhombic frames (90/90/90) always; frames with any other angle (MDAnalysis' triclinic
 * routine: triclinic_vectors, _triclinic_pbc, minimum_image_triclinic -- coordinates moved into the primary cell, then the
 * shortest of the 27 nearest images) when r1 is below half the smallest width of the cell, where the wrap of the
 * fractional differences is that image; wider windows, non-periodic axes and MDC_RDF_EXACT=1 take the FP64 kernel for
 * every pair (~20-40x slower per pair).  MDC_ERR_INVALID for angles that do not describe a cell.
 * Groups of 4096 particles or more are put in Hilbert order of a cell grid on the device first, and the pair kernel
 * skips the blocks of pairs farther apart than r1 (what capped_distance's grid search does for the reference,
 * structure.py:109-115); counts do not depend on it.  Environment, for diagnostics: MDC_RDF_SORT=0 keeps the original
 * order, MDC_RDF_EXT=0 the range-tested histogram update, MDC_RDF_EXACT=1 the all-FP64 kernel, MDC_RDF_REL=0 the integer
 * form of the coordinate differences, MDC_RDF_L32=0 eight histogram copies instead of one per bank, MDC_RDF_TRI_FILTER=0
 * the FP64 kernel for every triclinic frame.
 */
int  mdc_rdf_push(mdc_rdf_plan *plan, const float *pos_a, const float *pos_b,
                  const float *box, int n_frames, int on_device);

int  mdc_rdf_read(mdc_rdf_plan *plan, int64_t *counts_out, int64_t *n_frames_out);
/*
 * The pair kernel is a persistent grid that fills every SM (3 blocks of 80 registers x 256 threads per SM).
 * blocks_per_sm > 0 caps it, leaving register file / shared memory on every SM for the kernels of ANOTHER context
 * queued at the same time, which then run side by side with it.  0 = all that fit.
 */
int  mdc_rdf_set_blocks_per_sm(mdc_rdf_plan *plan, int blocks_per_sm);
int  mdc_rdf_reset(mdc_rdf_plan *plan);
int  mdc_rdf_accum_ptr(mdc_rdf_plan *plan, void **dev_ptr, int64_t *n_elems);

/*
 * One-shot functional seam (structure.py:34-131): counts_out i64 (n_bins).
 * exclusion as above.  Host arrays.
 */
int  mdc_radial_histogram(mdc_ctx *ctx, const float *pos_a, int64_t n_a,
                          const float *pos_b, int64_t n_b, const float box[6],
                          int n_bins, double r0, double r1,
                          int64_t excl0, int64_t excl1, int64_t *counts_out);

/* ---- g(r)-derived radial transforms ---------------- */

#define MDC_TRANSFORM_FOURIER_3D  0   /* (4 pi / q) int f(r) r sin(q r) dr      structure.py:176-215 */
#define MDC_TRANSFORM_HANKEL_2D   1   /* 2 pi int f(r) r J0(q r) dr             structure.py:134-173 */

/*
 * out[k] = prefactor(q_k) * sum_i weights[i] f[i] r[i] K(q_k r[i])  in FP64, with K = sin (3-D) or J0 (2-D) and
 * the q = 0 limits of the reference.  `weights` are the quadrature weights of the samples r (the reference uses
 * scipy.integrate.simpson, which is linear: weights = simpson(identity, x = r)).  Host arrays; r, f, weights have
 * n_r elements, q and out n_q.
 */
int  mdc_radial_transform(mdc_ctx *ctx, int kind, const double *r, const double *f,
                          const double *weights, int64_t n_r,
                          const double *q, int64_t n_q, double *out);

/* ---- frame ingest: LAMMPS text dumps ---------------- */

/*
 * Replaces /root/reference/src/mdcraft/analysis/reader.py:24-572 (LAMMPSDumpTrajectoryReader) for orthogonal boxes
 * and unscaled coordinates (x y z / xu yu zu), with or without an id column.
 *
 * mdc_dump_open memory-maps the file, reads the first header, builds the frame index with host threads
 * (reader.py:138-210; every frame must hold the same number of atoms, as the reference requires at read time) and
 * parses every frame's header (time step, box).  `conventions`: NULL = automatic (reader.py:438-447), else
 * "cx,cy,cz" with c in {"u", "su", "", "s", "-"} ("-" ignores the axis).  n_threads <= 0: all host threads.
 *
 * mdc_dump_read parses frames [f0, f0 + n_frames) ON THE GPU: the raw text goes to HBM once, one thread converts
 * one atom line, and positions come out as float32 (n_frames, n_atoms, 3) sorted by atom id (reader.py:550-552),
 * each value equal to float32(float(token)) bit for bit.  pos_out is a device pointer (on_device != 0; ready to be
 * handed to mdc_sq_push / mdc_rdf_push) or host memory.  The block must be below 2 GiB of text
 * (mdc_dump_block_bytes).
 *
 * info: [n_frames, n_atoms, has_id, convention index of x, y, z in {"u","su","","s"} (-1: axis absent), n_columns,
 *        values of the last read that needed the host parser].
 * headers: steps i64 (n_frames), box f32 (n_frames, 6), origin f64 (n_frames, 3) = (xlo, ylo, zlo), offsets i64
 *        (n_frames + 1) byte offsets of the frames; any pointer may be NULL.
 */
int      mdc_dump_open(const char *path, const char *conventions, int n_threads, mdc_dump **out);
int      mdc_dump_close(mdc_dump *dump);
int      mdc_dump_info(mdc_dump *dump, int64_t info[8]);
int      mdc_dump_headers(mdc_dump *dump, int64_t *steps, float *box, double *origin, int64_t *offsets);
int64_t  mdc_dump_block_bytes(mdc_dump *dump, int64_t f0, int n_frames);
int      mdc_dump_read(mdc_dump *dump, mdc_ctx *ctx, int64_t f0, int n_frames, float *pos_out, int on_device);

/* ---- frame ingest: AMBER NetCDF trajectories -------- */

/*
 * Replaces the read side of /root/reference/src/mdcraft/openmm/file.py:22-306 (NetCDFFile.get_num_frames,
 * get_num_atoms, get_dimensions, get_times, get_positions, get_velocities, get_forces), which goes through netCDF4, for NetCDF-3 classic and
 * 64-bit-offset files (what that class writes, file.py:52-54) following the AMBER trajectory convention.
 * mdc_ncdf_open memory-maps the file and parses the header on the host.  mdc_ncdf_read sends the big-endian
 * coordinates of frames [f0, f0 + n_frames) to HBM with one strided copy and byte-swaps them there into float32
 * (n_frames, n_atoms, 3); pos_out is a device pointer (on_device != 0) or host memory.
 * info: [n_frames, n_atoms, has cell_lengths / cell_angles, has time].
 * headers: times (n_frames) as f64, cell_lengths and cell_angles f64 (n_frames, 3); any pointer may be NULL.
 */
int  mdc_ncdf_open(const char *path, mdc_ncdf **out);
int  mdc_ncdf_close(mdc_ncdf *file);
int  mdc_ncdf_info(mdc_ncdf *file, int64_t info[4]);
int  mdc_ncdf_headers(mdc_ncdf *file, double *times, double *cell_lengths, double *cell_angles);
int  mdc_ncdf_read(mdc_ncdf *file, mdc_ctx *ctx, int64_t f0, int n_frames, float *pos_out, int on_device);
/*
 * The other per-atom record variables of the convention, same layout and same path as the coordinates
 * (file.py:217-306, get_velocities / get_forces): which = MDC_NCDF_COORDINATES, _VELOCITIES (angstrom / ps as stored)
 * or _FORCES (kcal / mol / angstrom).  mdc_ncdf_has returns 1 if the file holds the variable, else 0.
 */
#define MDC_NCDF_COORDINATES 0
#define MDC_NCDF_VELOCITIES  1
#define MDC_NCDF_FORCES      2
int  mdc_ncdf_has(mdc_ncdf *file, int which);
int  mdc_ncdf_read_var(mdc_ncdf *file, mdc_ctx *ctx, int which, int64_t f0, int n_frames, float *out, int on_device);

/* ---- 1-D density profiles ------------------------- */

/*
 * Replaces the per-frame arithmetic of DensityProfile._single_frame without recentring
 * (/root/reference/src/mdcraft/analysis/profile.py:1128-1181): positions are wrapped into the box
 * (algorithm/topology.py:620-624) and binned per axis and group with numba_histogram
 * (algorithm/accelerated.py:46-81), in the reference's FP64 operation order: counts are exact.
 *   group_sizes : positions pushed later hold the groups concatenated in order.
 *   axes        : n_axes entries in {0, 1, 2}; n_bins one entry per axis; dims = the three box lengths.
 * Count table layout: [axis][group][bin] (axis k starts at n_groups * sum of n_bins[0..k-1]).
 * mdc_profile_push: pos f32 (n_frames, N, 3); per_frame_out NULL or host i64 (n_frames, total) receiving every
 * frame's own counts (the reference's average=False).  mdc_profile_read: summed counts i64 (total).
 */
int      mdc_profile_plan_create(mdc_ctx *ctx, int n_groups, const int64_t *group_sizes,
                                 int n_axes, const int *axes, const int *n_bins, const double dims[3],
                                 mdc_profile_plan **out);
int      mdc_profile_plan_destroy(mdc_profile_plan *plan);
int64_t  mdc_profile_num_counts(mdc_profile_plan *plan);
int      mdc_profile_push(mdc_profile_plan *plan, const float *pos, int n_frames, int on_device,
                          int64_t *per_frame_out);
/* The same for float64 positions (n_frames, N, 3): what DensityProfile holds after `recenter` (profile.py:1137-1165). */
int      mdc_profile_push_f64(mdc_profile_plan *plan, const double *pos, int n_frames, int on_device,
                              int64_t *per_frame_out);
int      mdc_profile_read(mdc_profile_plan *plan, int64_t *counts_out, int64_t *n_frames_out);
int      mdc_profile_reset(mdc_profile_plan *plan);
int      mdc_profile_accum_ptr(mdc_profile_plan *plan, void **dev_ptr, int64_t *n_elems);

/* ---- timing (CUDA events on the context's compute stream) ---------------- */

/*
 * All plans of a context queue their kernels on the context's compute stream.
 * mdc_mark(ctx, 0 / 1) records a start / stop event there; mdc_elapsed_ms waits
 * for the stop event and returns the device time between the two.
 */
int  mdc_mark(mdc_ctx *ctx, int which);
int  mdc_elapsed_ms(mdc_ctx *ctx, float *ms);

/* Milliseconds the device spent in the DOMINANT kernels of the last *_push call
 * (the density kernels for S(q), the pair kernel for g(r); first 8 frame blocks of
 * the push), and how many kernels of any kind that call launched. */
int  mdc_sq_last_push_stats(mdc_sq_plan *plan, float *kernel_ms, int *n_launches);
int  mdc_rdf_last_push_stats(mdc_rdf_plan *plan, float *kernel_ms, int *n_launches);

/* Which kernel path the LAST frame block of the last mdc_rdf_push took (diagnostics, tests, the verbose run log of the
 * classes -- base.py:378-384 logs what the process farm does): info = {filter kernel (1) or FP64 kernel for every pair
 * (0), frames in Hilbert order with culling, words per extended-histogram copy (0: range-tested update), words between
 * consecutive bins of a copy (32: one copy per bank), relative-coordinate rows, a triclinic frame in the block,
 * capacity of the undecided-row queue, frames in the block}. */
int  mdc_rdf_plan_info(mdc_rdf_plan *plan, int info[8]);

/* ---- on-box probes (recorded next to MEASURED_PEAKS.json by bench.py) ---- */

/*
 * Issue-rate microbenchmarks.  which: 0 = FFMA, 1 = MUFU.SIN+MUFU.COS pairs,
 * 2 = DFMA, 3 = F2F.F64.F32, 4 = shared-memory atomics (spread addresses),
 * 5 = IMAD, 6 = FFMA2 (packed, 2 FMAs each), 7 = ALU (LOP3 + IADD3), 8 = I2F,
 * 9 / 10 = histogram-like shared atomics with / without the atomic,
 * 11 = FFMA2 with scalar x pair operands, 12 / 13 = FFMA2 in the operand pattern of
 * the factored S(q) loop (kernel order / grouped by pair operand).  Returns
 * operations per second over the whole GPU.
 */
int  mdc_probe_rate(mdc_ctx *ctx, int which, double *ops_per_s);

/*
 * Accuracy of the SFU sincos pipeline used by the DIRECT S(q) kernel over all
 * 2^23 phases: out[0..1] = mean error of (cos, sin) (the DC bias a coherent sum
 * over N particles multiplies by N), out[2] = rms error, out[3] = max |error|.
 * variant selects the kernel's phase-to-argument mapping.
 */
int  mdc_probe_sincos(mdc_ctx *ctx, int variant, double out[4]);

#ifdef __cplusplus
}
#endif
#endif /* MDCRAFT_B200_H */
