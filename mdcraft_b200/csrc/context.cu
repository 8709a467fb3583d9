// context.cu — library context, error reporting and on-box issue-rate / accuracy probes.
//
// The probes exist because the S(q) and g(r) rooflines are ISSUE-RATE rooflines
// (SURVEY.md §8d): the denominators (MUFU lanes, FP32 / FP64 / ALU lanes, shared
// atomics per clock) are architectural assumptions until measured on the box.
#include "common.cuh"

#include <math.h>
#include <vector>

namespace mdc {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

}  // namespace mdc

using namespace mdc;

extern "C" int mdc_version(void) { return 100; }

extern "C" const char *mdc_last_error(void) { return g_last_error.c_str(); }

static int create_impl(int device, int high_priority, mdc_ctx **out);

extern "C" int mdc_create(int device, mdc_ctx **out) { return create_impl(device, 0, out); }

extern "C" int mdc_create_with_priority(int device, int high_priority, mdc_ctx **out)
{
    return create_impl(device, high_priority, out);
}

static int create_impl(int device, int high_priority, mdc_ctx **out)
{
    if (!out) return fail(MDC_ERR_INVALID, "mdc_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(MDC_ERR_CUDA, std::string("mdc_create: no CUDA device (") +
                                      cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device < 0 || device >= n) return fail(MDC_ERR_INVALID, "mdc_create: device index out of range");
    MDC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MDC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(MDC_ERR_UNSUPPORTED, "mdc_create: device is not sm_100 (library is built for sm_100a only)");
    mdc_ctx *c = new mdc_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    c->clock_khz = khz;
    c->global_mem = (int64_t)prop.totalGlobalMem;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->smem_per_sm = prop.sharedMemPerMultiprocessor;
    // high priority: when thread blocks of another context's kernels retire, this context's blocks are placed first
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // numerically lower = higher priority
    const int prio = high_priority ? prio_hi : prio_lo;
    cudaError_t e2 = cudaStreamCreateWithPriority(&c->copy, cudaStreamNonBlocking, prio);
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithPriority(&c->compute, cudaStreamNonBlocking, prio);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->mark[0]);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->mark[1]);
    if (e2 != cudaSuccess) {
        delete c;
        return fail(MDC_ERR_CUDA, std::string("mdc_create: ") + cudaGetErrorString(e2));
    }
    *out = c;
    return MDC_OK;
}

extern "C" int mdc_destroy(mdc_ctx *ctx)
{
    if (!ctx) return MDC_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->mark[0]) cudaEventDestroy(ctx->mark[0]);
    if (ctx->mark[1]) cudaEventDestroy(ctx->mark[1]);
    if (ctx->copy) cudaStreamDestroy(ctx->copy);
    if (ctx->compute) cudaStreamDestroy(ctx->compute);
    delete ctx;
    return MDC_OK;
}

extern "C" int mdc_mark(mdc_ctx *ctx, int which)
{
    if (!ctx || which < 0 || which > 1) return fail(MDC_ERR_INVALID, "mdc_mark: bad argument");
    MDC_CUDA(cudaSetDevice(ctx->device));
    // the compute stream must not start (stop) the clock before earlier (later) copies are ordered:
    // copies are joined into `compute` through per-plan events, so one record on `compute` suffices
    MDC_CUDA(cudaEventRecord(ctx->mark[which], ctx->compute));
    return MDC_OK;
}

extern "C" int mdc_elapsed_ms(mdc_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return fail(MDC_ERR_INVALID, "mdc_elapsed_ms: NULL argument");
    MDC_CUDA(cudaSetDevice(ctx->device));
    MDC_CUDA(cudaEventSynchronize(ctx->mark[1]));
    MDC_CUDA(cudaEventElapsedTime(ms, ctx->mark[0], ctx->mark[1]));
    return MDC_OK;
}

extern "C" int mdc_device_info(mdc_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                               int *sm_clock_khz, int64_t *global_mem_bytes)
{
    if (!ctx) return fail(MDC_ERR_INVALID, "mdc_device_info: ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (sm_clock_khz) *sm_clock_khz = ctx->clock_khz;
    if (global_mem_bytes) *global_mem_bytes = ctx->global_mem;
    return MDC_OK;
}

extern "C" int mdc_synchronize(mdc_ctx *ctx)
{
    if (!ctx) return fail(MDC_ERR_INVALID, "mdc_synchronize: ctx is NULL");
    MDC_CUDA(cudaSetDevice(ctx->device));
    MDC_CUDA(cudaDeviceSynchronize());
    return MDC_OK;
}

// ---------------------------------------------------------------------------
// Issue-rate probes.  Each kernel runs ITER iterations of UNROLL independent
// chains per thread so that the pipe under test, not latency, is the limiter.
// ---------------------------------------------------------------------------

namespace {

constexpr int PROBE_THREADS = 256;
constexpr int PROBE_ITER = 4096;

__global__ void probe_ffma(float *out, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < PROBE_ITER; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// packed FP32x2 FMA (sm_100 only)
__global__ void probe_ffma2(float2 *out, float a, float b)
{
    unsigned long long x[8], A, B;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float v = threadIdx.x + k;
        asm("mov.b64 %0, {%1, %1};" : "=l"(x[k]) : "f"(v));
    }
    for (int i = 0; i < PROBE_ITER; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(A), "l"(B));
    }
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k]));
        acc.x += lo;
        acc.y += hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// packed FMA with the operand forms of sq_lattice_factored: scalar (broadcast) x register pair
__global__ void probe_ffma2_scalar(float2 *out, float a, float b)
{
    unsigned long long x[8], W, V;
    asm("mov.b64 %0, {%1, %2};" : "=l"(W) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(V) : "f"(-b), "f"(a));
    float s0 = a * 0.5f, s1 = b * 0.25f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float v = threadIdx.x + k;
        asm("mov.b64 %0, {%1, %1};" : "=l"(x[k]) : "f"(v));
    }
    for (int i = 0; i < PROBE_ITER; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned long long S;
            asm("mov.b64 %0, {%1, %1};" : "=l"(S) : "f"((k & 1) ? s0 : s1));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x[k]) : "l"(S), "l"((k & 2) ? W : V));
        }
    }
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k]));
        acc.x += lo;
        acc.y += hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// packed FMAs in the exact operand pattern of sq_lattice_factored's inner loop: 16 accumulator pairs,
// 4 register-pair operands (W0, W1, V0, V1) and 16 scalar operands per "particle"; ORDER selects the
// instruction order (0: as in the kernel, 1: grouped by pair operand)
template <int ORDER>
__global__ void probe_ffma2_loop(float2 *out, float a, float b)
{
    unsigned long long acc[2][8], w[2], v[2];
    float z[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) z[k] = a * (k + 1) + threadIdx.x * 1e-6f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(w[0]) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(w[1]) : "f"(b), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[0]) : "f"(-b), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[1]) : "f"(-a), "f"(b));
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[0][k] = acc[1][k] = 0ull;
    for (int i = 0; i < PROBE_ITER / 4; ++i) {
        if (ORDER == 0) {
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                unsigned long long zr0, zi0, zr1, zi1;
                asm("mov.b64 %0, {%1, %1};" : "=l"(zr0) : "f"(z[2 * c]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(zi0) : "f"(z[2 * c + 1]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(zr1) : "f"(z[2 * c + 2]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(zi1) : "f"(z[2 * c + 3]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[0][c]) : "l"(zr0), "l"(w[0]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[1][c]) : "l"(zr0), "l"(w[1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[0][c + 1]) : "l"(zr1), "l"(w[0]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[1][c + 1]) : "l"(zr1), "l"(w[1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[0][c]) : "l"(zi0), "l"(v[0]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[1][c]) : "l"(zi0), "l"(v[1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[0][c + 1]) : "l"(zi1), "l"(v[0]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[1][c + 1]) : "l"(zi1), "l"(v[1]));
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    unsigned long long zr;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(zr) : "f"(z[2 * c]));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[h][c]) : "l"(zr), "l"(w[h]));
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    unsigned long long zi;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(zi) : "f"(z[2 * c + 1]));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[h][c]) : "l"(zi), "l"(v[h]));
                }
            }
        }
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[0][k] ^ acc[1][k]));
        r.x += lo;
        r.y += hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void probe_mufu(float *out, float a)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + .1f, x2 = x0 + .2f, x3 = x0 + .3f;
    float acc = 0.f;
    for (int i = 0; i < PROBE_ITER; ++i) {
        float s0, c0, s1, c1, s2, c2, s3, c3;
        asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(x0));
        asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(x0));
        asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(x1));
        asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(x1));
        asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(s2) : "f"(x2));
        asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(c2) : "f"(x2));
        asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(s3) : "f"(x3));
        asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(c3) : "f"(x3));
        x0 = s0 + a; x1 = c1 + a; x2 = s2 + a; x3 = c3 + a;
        acc += c0 + s1 + c2 + s3;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + x0 + x1 + x2 + x3;
}

__global__ void probe_dfma(double *out, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < PROBE_ITER; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void probe_f2f(double *out, float a)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    for (int i = 0; i < PROBE_ITER; ++i) {
        double d0, d1, d2, d3;
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d0) : "f"(x0));
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d1) : "f"(x1));
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d2) : "f"(x2));
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d3) : "f"(x3));
        // keep the conversions live without adding FP64 arithmetic per conversion:
        // xor the high words into an integer accumulator
        acc0 = __hiloint2double(__double2hiint(acc0) ^ __double2hiint(d0), __double2loint(d0));
        acc1 = __hiloint2double(__double2hiint(acc1) ^ __double2hiint(d1), __double2loint(d1));
        acc2 = __hiloint2double(__double2hiint(acc2) ^ __double2hiint(d2), __double2loint(d2));
        acc3 = __hiloint2double(__double2hiint(acc3) ^ __double2hiint(d3), __double2loint(d3));
        x0 += a; x1 += a; x2 += a; x3 += a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

// shared-memory atomics: each lane hits its own bank, address changes every iteration
__global__ void probe_atoms(unsigned *out, int stride)
{
    __shared__ unsigned h[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) h[i] = 0;
    __syncthreads();
    unsigned idx = threadIdx.x;
    for (int i = 0; i < PROBE_ITER; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            atomicAdd(&h[idx & 4095], 1u);
            idx += stride;
        }
    }
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory atomics with a g(r)-like address distribution: bin ~ 200 * u^(1/3)
__global__ void probe_atoms_hist(unsigned *out, unsigned seed)
{
    __shared__ unsigned h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    unsigned s = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u);
    unsigned *hw = h[threadIdx.x >> 5];
    for (int i = 0; i < PROBE_ITER; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = s * 1664525u + 1013904223u;
            float u = (float)(s >> 8) * (1.0f / 16777216.0f);
            int b = (int)(200.0f * cbrtf(u));
            atomicAdd(&hw[b], 1u);
        }
    }
    __syncthreads();
    unsigned t = 0;
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) t += (&h[0][0])[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

// same address generation without the atomic: subtract to isolate the atomic cost
__global__ void probe_atoms_hist_base(unsigned *out, unsigned seed)
{
    unsigned s = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u);
    unsigned t = 0;
    for (int i = 0; i < PROBE_ITER; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = s * 1664525u + 1013904223u;
            float u = (float)(s >> 8) * (1.0f / 16777216.0f);
            int b = (int)(200.0f * cbrtf(u));
            t += b;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

__global__ void probe_imad(unsigned *out, unsigned a, unsigned b)
{
    unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < PROBE_ITER; ++i) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

__global__ void probe_alu(unsigned *out, unsigned a, unsigned b)
{
    unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < PROBE_ITER; ++i) {
        x0 = (x0 ^ a) + b; x1 = (x1 ^ a) + b; x2 = (x2 ^ a) + b; x3 = (x3 ^ a) + b;
        x4 = (x4 ^ a) + b; x5 = (x5 ^ a) + b; x6 = (x6 ^ a) + b; x7 = (x7 ^ a) + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

__global__ void probe_i2f(float *out, unsigned a)
{
    unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    unsigned acc = 0;
    for (int i = 0; i < PROBE_ITER; ++i) {
        float f0, f1, f2, f3;
        asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f0) : "r"(x0));
        asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f1) : "r"(x1));
        asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f2) : "r"(x2));
        asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f3) : "r"(x3));
        acc ^= __float_as_uint(f0) ^ __float_as_uint(f1) ^ __float_as_uint(f2) ^ __float_as_uint(f3);
        x0 += a; x1 += a; x2 += a; x3 += a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

template <typename F>
int time_probe(mdc_ctx *ctx, F launch, double ops_per_thread, double *ops_per_s)
{
    const int blocks = ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    MDC_CUDA(cudaEventCreate(&e0));
    MDC_CUDA(cudaEventCreate(&e1));
    launch(blocks);  // warm-up
    MDC_CUDA(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        MDC_CUDA(cudaEventRecord(e0));
        launch(blocks);
        MDC_CUDA(cudaEventRecord(e1));
        MDC_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MDC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    MDC_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_s = ops_per_thread * (double)blocks * PROBE_THREADS / (best * 1e-3);
    return MDC_OK;
}

}  // namespace

extern "C" int mdc_probe_rate(mdc_ctx *ctx, int which, double *ops_per_s)
{
    if (!ctx || !ops_per_s) return fail(MDC_ERR_INVALID, "mdc_probe_rate: NULL argument");
    MDC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<double> buf;
    MDC_CHECK(buf.alloc((size_t)ctx->sm_count * 8 * PROBE_THREADS));
    void *o = buf.p;
    int rc = MDC_OK;
    const double per8 = 8.0 * PROBE_ITER;
    switch (which) {
    case 0:
        rc = time_probe(ctx, [&](int b) { probe_ffma<<<b, PROBE_THREADS>>>((float *)o, 1.0001f, 0.5f); }, per8, ops_per_s);
        break;
    case 1:  // counts sin+cos PAIRS
        rc = time_probe(ctx, [&](int b) { probe_mufu<<<b, PROBE_THREADS>>>((float *)o, 0.25f); }, 4.0 * PROBE_ITER, ops_per_s);
        break;
    case 2:
        rc = time_probe(ctx, [&](int b) { probe_dfma<<<b, PROBE_THREADS>>>((double *)o, 1.0001, 0.5); }, per8, ops_per_s);
        break;
    case 3:
        rc = time_probe(ctx, [&](int b) { probe_f2f<<<b, PROBE_THREADS>>>((double *)o, 0.37f); }, 4.0 * PROBE_ITER, ops_per_s);
        break;
    case 4:
        rc = time_probe(ctx, [&](int b) { probe_atoms<<<b, PROBE_THREADS>>>((unsigned *)o, 33); }, per8, ops_per_s);
        break;
    case 5:
        rc = time_probe(ctx, [&](int b) { probe_imad<<<b, PROBE_THREADS>>>((unsigned *)o, 3u, 7u); }, per8, ops_per_s);
        break;
    case 6:  // counts packed instructions (2 FMAs each)
        rc = time_probe(ctx, [&](int b) { probe_ffma2<<<b, PROBE_THREADS>>>((float2 *)o, 1.0001f, 0.5f); }, per8, ops_per_s);
        break;
    case 7:  // two dependent ALU ops (LOP3 + IADD3) per counted op
        rc = time_probe(ctx, [&](int b) { probe_alu<<<b, PROBE_THREADS>>>((unsigned *)o, 0x5bd1e995u, 7u); }, 2 * per8, ops_per_s);
        break;
    case 8:
        rc = time_probe(ctx, [&](int b) { probe_i2f<<<b, PROBE_THREADS>>>((float *)o, 977u); }, 4.0 * PROBE_ITER, ops_per_s);
        break;
    case 9:  // g(r)-like histogram atomics; includes address generation
        rc = time_probe(ctx, [&](int b) { probe_atoms_hist<<<b, PROBE_THREADS>>>((unsigned *)o, 12345u); }, per8, ops_per_s);
        break;
    case 10:  // address generation alone (subtract from 9)
        rc = time_probe(ctx, [&](int b) { probe_atoms_hist_base<<<b, PROBE_THREADS>>>((unsigned *)o, 12345u); }, per8, ops_per_s);
        break;
    case 11:  // packed FMA, scalar x pair operands
        rc = time_probe(ctx, [&](int b) { probe_ffma2_scalar<<<b, PROBE_THREADS>>>((float2 *)o, 1.0001f, 0.5f); }, per8, ops_per_s);
        break;
    case 12:  // packed FMAs, kernel operand pattern and order
        rc = time_probe(ctx, [&](int b) { probe_ffma2_loop<0><<<b, PROBE_THREADS>>>((float2 *)o, 1.0001f, 0.5f); }, per8, ops_per_s);
        break;
    case 13:  // packed FMAs, kernel operand pattern, grouped by pair operand
        rc = time_probe(ctx, [&](int b) { probe_ffma2_loop<1><<<b, PROBE_THREADS>>>((float2 *)o, 1.0001f, 0.5f); }, per8, ops_per_s);
        break;
    default:
        rc = fail(MDC_ERR_INVALID, "mdc_probe_rate: unknown probe");
    }
    buf.release();
    return rc;
}
