// transforms.cu — g(r)-derived radial transforms on the device.
//
// Replaces the dense (n_q x n_r) NumPy evaluations of
//   /root/reference/src/mdcraft/analysis/structure.py:176-215  radial_fourier_transform
//       rft(q) = (4 pi / q) simpson(f r sin(q r), x = r),      rft(0) = 4 pi simpson(f r^2, x = r)
//   /root/reference/src/mdcraft/analysis/structure.py:134-173  zeroth_order_hankel_transform
//       ht(q)  = 2 pi simpson(f r J0(q r), x = r)
// Simpson's rule is linear in its integrand, so the caller passes the quadrature weights w (simpson applied to
// the identity: exactly SciPy's composite rule, including its end correction for an even number of samples) and
// the kernel evaluates  sum_i w_i f_i r_i K(q r_i)  in FP64, one block per wavenumber, fixed reduction order.
#include "common.cuh"

#include <math.h>

using namespace mdc;

namespace {

constexpr int RT_THREADS = 128;

__global__ void __launch_bounds__(RT_THREADS)
radial_transform_kernel(int kind, const double *__restrict__ r, const double *__restrict__ f,
                        const double *__restrict__ w, int n_r, const double *__restrict__ q, int n_q,
                        double *__restrict__ out)
{
    __shared__ double sh[RT_THREADS];
    const int iq = blockIdx.x;
    if (iq >= n_q) return;
    const double qq = q[iq];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_r; i += RT_THREADS) {
        const double ri = r[i], a = w[i] * f[i] * ri;
        double k;
        if (kind == MDC_TRANSFORM_FOURIER_3D) k = (qq == 0.0) ? ri : sin(qq * ri);
        else k = (qq == 0.0) ? 1.0 : j0(qq * ri);
        acc += a * k;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = RT_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double two_pi = 6.283185307179586;
        if (kind == MDC_TRANSFORM_FOURIER_3D) out[iq] = (qq == 0.0) ? 2.0 * two_pi * sh[0] : 2.0 * two_pi * sh[0] / qq;
        else out[iq] = two_pi * sh[0];
    }
}

}  // namespace

extern "C" int mdc_radial_transform(mdc_ctx *ctx, int kind, const double *r, const double *f, const double *weights,
                                    int64_t n_r, const double *q, int64_t n_q, double *out)
{
    if (!ctx || !r || !f || !weights || !q || !out) return fail(MDC_ERR_INVALID, "mdc_radial_transform: NULL argument");
    if (kind != MDC_TRANSFORM_FOURIER_3D && kind != MDC_TRANSFORM_HANKEL_2D)
        return fail(MDC_ERR_INVALID, "mdc_radial_transform: unknown transform");
    if (n_r <= 0 || n_q <= 0 || n_r > 2147483647ll || n_q > 2147483647ll)
        return fail(MDC_ERR_INVALID, "mdc_radial_transform: bad sizes");
    MDC_CUDA(cudaSetDevice(ctx->device));
    DevBuf<double> buf;
    MDC_CHECK(buf.alloc((size_t)3 * n_r + 2 * n_q));
    double *dr = buf.p, *df = dr + n_r, *dw = df + n_r, *dq = dw + n_r, *dout = dq + n_q;
    int rc = MDC_OK;
    cudaError_t e;
    do {
        if ((e = cudaMemcpyAsync(dr, r, n_r * sizeof(double), cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(df, f, n_r * sizeof(double), cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dw, weights, n_r * sizeof(double), cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dq, q, n_q * sizeof(double), cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess) break;
        radial_transform_kernel<<<(unsigned)n_q, RT_THREADS, 0, ctx->compute>>>(kind, dr, df, dw, (int)n_r, dq, (int)n_q, dout);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(out, dout, n_q * sizeof(double), cudaMemcpyDeviceToHost, ctx->compute)) != cudaSuccess) break;
        e = cudaStreamSynchronize(ctx->compute);
    } while (0);
    if (e != cudaSuccess) rc = fail(MDC_ERR_CUDA, std::string("mdc_radial_transform: ") + cudaGetErrorString(e));
    buf.release();
    return rc;
}
