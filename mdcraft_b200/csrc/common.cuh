// common.cuh — shared helpers for libmdcraft_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mdcraft_b200.h"

namespace mdc {

// thread-local last error (mdc_last_error)
void set_error(const std::string &msg);
int  fail(int code, const std::string &msg);

#define MDC_CUDA(expr)                                                          \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) {                                                \
            return ::mdc::fail(MDC_ERR_CUDA, std::string(#expr) + ": " +        \
                                                 cudaGetErrorString(_e));       \
        }                                                                       \
    } while (0)

#define MDC_CHECK(expr)                                                         \
    do {                                                                        \
        int _r = (expr);                                                        \
        if (_r != MDC_OK) return _r;                                            \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// One context per device.  All plans of a context share its two streams: host->device copies
// and the layout kernels that follow them run on `copy`, the analysis kernels on `compute`.
struct Ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    int clock_khz = 0;
    int64_t global_mem = 0;
    size_t smem_optin = 0;
    size_t smem_per_sm = 0;   // sharedMemPerMultiprocessor
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t mark[2] = {nullptr, nullptr};   // mdc_mark / mdc_elapsed_ms
};

// Raises a kernel's dynamic shared-memory limit to the device maximum (minus the kernel's static shared memory).
// The limit is a per-function, per-device attribute shared by every live plan, so it is never set to one plan's own
// footprint: a later, smaller plan would lower it under an earlier plan's launches.
template <typename K>
static inline cudaError_t raise_smem_limit(K kernel, size_t smem_optin)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    const size_t room = smem_optin > fa.sharedSizeBytes ? smem_optin - fa.sharedSizeBytes : 0;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room);
}

// Device buffer helper (plans free in their destroy).
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        if (p && count <= n) return MDC_OK;
        release();
        if (count == 0) return MDC_OK;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            n = 0;
            cudaGetLastError();
            return fail(e == cudaErrorMemoryAllocation ? MDC_ERR_NOMEM : MDC_ERR_CUDA,
                        std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
        n = count;
        return MDC_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

// Two-slot staging of frame blocks: block k+1 is uploaded (and laid out) on the context's copy
// stream while block k is being processed on its compute stream.
struct Streams {
    cudaStream_t copy = nullptr, compute = nullptr;   // borrowed from the context
    cudaEvent_t staged[2] = {nullptr, nullptr};       // slot's inputs are in their final layout
    cudaEvent_t consumed[2] = {nullptr, nullptr};     // kernels reading the slot have finished
    cudaEvent_t k0[8], k1[8];                         // around the dominant kernel of each block of a push
    int n_timed = 0;
    bool used[2] = {false, false};
    int next = 0;
    int init(const Ctx *ctx) {
        copy = ctx->copy;
        compute = ctx->compute;
        for (int i = 0; i < 2; ++i) {
            MDC_CUDA(cudaEventCreateWithFlags(&staged[i], cudaEventDisableTiming));
            MDC_CUDA(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 8; ++i) {
            k0[i] = k1[i] = nullptr;
            MDC_CUDA(cudaEventCreate(&k0[i]));
            MDC_CUDA(cudaEventCreate(&k1[i]));
        }
        return MDC_OK;
    }
    void destroy() {
        for (int i = 0; i < 2; ++i) {
            if (staged[i]) cudaEventDestroy(staged[i]);
            if (consumed[i]) cudaEventDestroy(consumed[i]);
        }
        for (int i = 0; i < 8; ++i) {
            if (k0[i]) cudaEventDestroy(k0[i]);
            if (k1[i]) cudaEventDestroy(k1[i]);
        }
    }
    // device time of the dominant kernels of the last push (first 8 blocks), in ms
    int kernel_ms(float *out) {
        *out = 0.f;
        for (int i = 0; i < n_timed; ++i) {
            float ms = 0.f;
            MDC_CUDA(cudaEventSynchronize(k1[i]));
            MDC_CUDA(cudaEventElapsedTime(&ms, k0[i], k1[i]));
            *out += ms;
        }
        return MDC_OK;
    }
};

}  // namespace mdc

struct mdc_ctx : public mdc::Ctx {};
