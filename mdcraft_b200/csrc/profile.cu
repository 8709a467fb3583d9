// profile.cu — 1-D density profiles (the g(r) histogram's 1-D cousin).
//
// Replaces the per-frame arithmetic of /root/reference/src/mdcraft/analysis/profile.py:1128-1181
// (DensityProfile._single_frame without recentring):
//   positions (float32, widened to float64)  ->  wrap into the box (algorithm/topology.py:620-624)
//   -> numba_histogram per axis and group (algorithm/accelerated.py:46-81)
// with the reference's FP64 operation order, so the counts are bit-exact:
//   wrap:  if (x < 0 || x > L)  x -= floor(x / L) * L            (divide, floor, multiply, subtract; no FMA)
//   bin:   x == L ? n - 1 : (int)(((x - 0) * n) * (1 / (L - 0))), counted iff 0 <= bin < n
// One thread per particle and frame; counts go to a per-frame u32 table with global atomics (3 per particle: HBM /
// L2 byte work, 12 B in per particle), then into the u64 accumulator.
#include "common.cuh"

#include <math.h>
#include <string.h>
#include <vector>

using namespace mdc;

namespace {

struct ProfileGeom {
    int n_groups, n_axes;
    int axis[3], n_bins[3], offset[3];   // offset of axis k in the count table; layout [axis][group][bin]
    double L[3], inv[3], nd[3];
    int total;
};

template <typename T>   // float: positions as the trajectory holds them; double: recentred on the host
__global__ void profile_hist(const T *__restrict__ pos, int64_t N, int F, const int *__restrict__ group_end,
                             ProfileGeom g, unsigned *__restrict__ frame_counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * F) return;
    const int64_t f = i / N, j = i - f * N;
    int grp = 0;
    while (grp < g.n_groups - 1 && j >= group_end[grp]) ++grp;
    const T *p = pos + 3 * i;
    unsigned *out = frame_counts + f * g.total;
    for (int k = 0; k < g.n_axes; ++k) {
        double x = (double)p[g.axis[k]];
        const double L = g.L[k];
        if (x < 0.0 || x > L) x = __dsub_rn(x, __dmul_rn(floor(__ddiv_rn(x, L)), L));
        long long b;
        if (x == L) {
            b = g.n_bins[k] - 1;
        } else {
            const double t = __dmul_rn(__dmul_rn(x, g.nd[k]), g.inv[k]);
            // fptosi truncates toward zero; NaN / out-of-range values are simply not counted
            b = (t > -9.0e18 && t < 9.0e18) ? (long long)t : -1;
        }
        if (b >= 0 && b < g.n_bins[k]) atomicAdd(&out[g.offset[k] + grp * g.n_bins[k] + (int)b], 1u);
    }
}

__global__ void profile_accum(const unsigned *__restrict__ frame_counts, int F, int total,
                              unsigned long long *__restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    unsigned long long s = 0;
    for (int f = 0; f < F; ++f) s += frame_counts[(size_t)f * total + i];
    counts[i] += s;
}

}  // namespace

struct mdc_profile_plan {
    mdc_ctx *ctx = nullptr;
    ProfileGeom g{};
    int64_t N = 0, n_frames = 0;
    int fblk = 0;
    DevBuf<int> group_end;
    DevBuf<float> raw;
    DevBuf<double> raw64;
    DevBuf<unsigned> frame_counts;
    DevBuf<unsigned long long> counts;
};

extern "C" int mdc_profile_plan_create(mdc_ctx *ctx, int n_groups, const int64_t *group_sizes, int n_axes,
                                       const int *axes, const int *n_bins, const double dims[3], mdc_profile_plan **out)
{
    if (!ctx || !group_sizes || !axes || !n_bins || !dims || !out) return fail(MDC_ERR_INVALID, "mdc_profile_plan_create: NULL argument");
    *out = nullptr;
    if (n_groups < 1 || n_axes < 1 || n_axes > 3) return fail(MDC_ERR_INVALID, "mdc_profile_plan_create: need >= 1 group and 1-3 axes");
    MDC_CUDA(cudaSetDevice(ctx->device));
    mdc_profile_plan *p = new mdc_profile_plan();
    p->ctx = ctx;
    p->g.n_groups = n_groups;
    p->g.n_axes = n_axes;
    std::vector<int> ends(n_groups);
    int64_t N = 0;
    for (int i = 0; i < n_groups; ++i) {
        if (group_sizes[i] < 0) {
            delete p;
            return fail(MDC_ERR_INVALID, "mdc_profile_plan_create: negative group size");
        }
        N += group_sizes[i];
        ends[i] = (int)N;
    }
    if (N <= 0 || N > 2147483647ll) {
        delete p;
        return fail(MDC_ERR_INVALID, "mdc_profile_plan_create: total particle count must be in [1, 2^31)");
    }
    p->N = N;
    int64_t total = 0;
    for (int k = 0; k < n_axes; ++k) {
        if (axes[k] < 0 || axes[k] > 2 || n_bins[k] < 1 || !(dims[axes[k]] > 0)) {
            delete p;
            return fail(MDC_ERR_INVALID, "mdc_profile_plan_create: bad axis, bin count or box length");
        }
        p->g.axis[k] = axes[k];
        p->g.n_bins[k] = n_bins[k];
        p->g.offset[k] = (int)total;
        p->g.L[k] = dims[axes[k]];
        p->g.inv[k] = 1.0 / (dims[axes[k]] - 0.0);   // accelerated.py:72-81 as compiled: hoisted reciprocal
        p->g.nd[k] = (double)n_bins[k];
        total += (int64_t)n_groups * n_bins[k];
    }
    if (total > (1 << 28)) {
        delete p;
        return fail(MDC_ERR_UNSUPPORTED, "mdc_profile_plan_create: too many bins");
    }
    p->g.total = (int)total;
    int rc = p->group_end.alloc(n_groups);
    if (rc == MDC_OK) rc = p->counts.alloc(total);
    cudaError_t e = cudaSuccess;
    if (rc == MDC_OK) e = cudaMemcpy(p->group_end.p, ends.data(), n_groups * sizeof(int), cudaMemcpyHostToDevice);
    if (rc == MDC_OK && e == cudaSuccess) e = cudaMemset(p->counts.p, 0, total * sizeof(unsigned long long));
    if (rc == MDC_OK && e != cudaSuccess) rc = fail(MDC_ERR_CUDA, std::string("mdc_profile_plan_create: ") + cudaGetErrorString(e));
    if (rc != MDC_OK) {
        p->group_end.release();
        p->counts.release();
        delete p;
        return rc;
    }
    *out = p;
    return MDC_OK;
}

extern "C" int mdc_profile_plan_destroy(mdc_profile_plan *p)
{
    if (!p) return MDC_OK;
    cudaSetDevice(p->ctx->device);
    cudaDeviceSynchronize();
    p->group_end.release();
    p->raw.release();
    p->raw64.release();
    p->frame_counts.release();
    p->counts.release();
    delete p;
    return MDC_OK;
}

extern "C" int64_t mdc_profile_num_counts(mdc_profile_plan *p) { return p ? p->g.total : -1; }

namespace {

// T positions (n_frames, N, 3), groups concatenated in order.  per_frame_out: NULL, or host i64 (n_frames, total)
// that receives the counts of every frame (average=False in the reference).
template <typename T>
int profile_push(mdc_profile_plan *p, const T *pos, int n_frames, int on_device, int64_t *per_frame_out, DevBuf<T> &raw)
{
    if (!p || !pos) return fail(MDC_ERR_INVALID, "mdc_profile_push: NULL argument");
    if (n_frames <= 0) return MDC_OK;
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t st = p->ctx->compute;
    const int total = p->g.total;
    const int blk = (int)std::max<int64_t>(1, std::min<int64_t>(64, ((int64_t)1 << 28) / (p->N * 3 * (int64_t)sizeof(T))));
    std::vector<unsigned> host;
    for (int f0 = 0; f0 < n_frames; f0 += blk) {
        const int F = std::min(blk, n_frames - f0);
        const T *src = pos + (size_t)f0 * p->N * 3;
        if (!on_device) {
            MDC_CHECK(raw.alloc((size_t)blk * p->N * 3));
            MDC_CUDA(cudaMemcpyAsync(raw.p, src, (size_t)F * p->N * 3 * sizeof(T), cudaMemcpyHostToDevice, st));
            src = raw.p;
        }
        MDC_CHECK(p->frame_counts.alloc((size_t)blk * total));
        MDC_CUDA(cudaMemsetAsync(p->frame_counts.p, 0, (size_t)F * total * sizeof(unsigned), st));
        profile_hist<T><<<(unsigned)ceil_div(p->N * F, 256), 256, 0, st>>>(src, p->N, F, p->group_end.p, p->g, p->frame_counts.p);
        profile_accum<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(p->frame_counts.p, F, total, p->counts.p);
        MDC_CUDA(cudaGetLastError());
        if (per_frame_out) {
            host.resize((size_t)F * total);
            MDC_CUDA(cudaMemcpyAsync(host.data(), p->frame_counts.p, host.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            MDC_CUDA(cudaStreamSynchronize(st));
            int64_t *dst = per_frame_out + (size_t)f0 * total;
            for (size_t i = 0; i < host.size(); ++i) dst[i] = (int64_t)host[i];
        } else if (!on_device) {
            MDC_CUDA(cudaStreamSynchronize(st));   // the staging buffer is reused by the next block, the caller's by the caller
        }
    }
    p->n_frames += n_frames;
    return MDC_OK;
}

}  // namespace

extern "C" int mdc_profile_push(mdc_profile_plan *p, const float *pos, int n_frames, int on_device, int64_t *per_frame_out)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_profile_push: NULL plan");
    return profile_push<float>(p, pos, n_frames, on_device, per_frame_out, p->raw);
}

// float64 positions: what DensityProfile holds after recentring (profile.py:1137-1165, unwrap + centre-of-mass
// shift, host bookkeeping with frame-to-frame state); wrapped and binned on the device like the float32 ones
extern "C" int mdc_profile_push_f64(mdc_profile_plan *p, const double *pos, int n_frames, int on_device,
                                    int64_t *per_frame_out)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_profile_push_f64: NULL plan");
    return profile_push<double>(p, pos, n_frames, on_device, per_frame_out, p->raw64);
}

extern "C" int mdc_profile_read(mdc_profile_plan *p, int64_t *counts_out, int64_t *n_frames_out)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_profile_read: NULL plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->ctx->compute));
    if (counts_out)
        MDC_CUDA(cudaMemcpy(counts_out, p->counts.p, (size_t)p->g.total * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (n_frames_out) *n_frames_out = p->n_frames;
    return MDC_OK;
}

extern "C" int mdc_profile_reset(mdc_profile_plan *p)
{
    if (!p) return fail(MDC_ERR_INVALID, "mdc_profile_reset: NULL plan");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->ctx->compute));
    MDC_CUDA(cudaMemset(p->counts.p, 0, (size_t)p->g.total * sizeof(unsigned long long)));
    p->n_frames = 0;
    return MDC_OK;
}

extern "C" int mdc_profile_accum_ptr(mdc_profile_plan *p, void **dev_ptr, int64_t *n_elems)
{
    if (!p || !dev_ptr || !n_elems) return fail(MDC_ERR_INVALID, "mdc_profile_accum_ptr: NULL argument");
    MDC_CUDA(cudaSetDevice(p->ctx->device));
    MDC_CUDA(cudaStreamSynchronize(p->ctx->compute));
    *dev_ptr = p->counts.p;
    *n_elems = p->g.total;
    return MDC_OK;
}
