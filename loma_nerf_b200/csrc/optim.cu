// optim.cu -- the two optimisers the reference hosts apply to the padded weight arrays.
//   Adam: /root/reference/train_nerf.py:133-161, INCLUDING its double bias correction
//         (lr_t already carries sqrt(1-b2^t)/(1-b1^t), and m_hat/v_hat divide again).
//   SGD : /root/reference/fit_img.py:512-513.
// The reference evaluates these in numpy float32 arrays with Python-float (double) scalars; the
// scalar factors are computed in double on the host here and passed down as floats.
#include <math.h>

#include "lnb_internal.h"

namespace {
__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g,
                            float *__restrict__ m, float *__restrict__ v, long long n, float b1,
                            float b2, float omb1, float omb2, float lr_t, float c1,
                            float c2, float eps)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i];
    float mi = b1 * m[i] + omb1 * gi;
    float vi = b2 * v[i] + omb2 * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    float m_hat = mi / c1, v_hat = vi / c2;
    p[i] -= lr_t * m_hat / (sqrtf(v_hat) + eps);
}
// device-resident step counter (CUDA-graph friendly): t = t_dev[0] + 1, scalars derived in double
__global__ void adam_dev_kernel(float *__restrict__ p, const float *__restrict__ g,
                                float *__restrict__ m, float *__restrict__ v, long long n,
                                const int *__restrict__ t_dev, double lr, double b1, double b2, double eps)
{
    __shared__ float sc[6];
    if (threadIdx.x == 0) {
        const int t = t_dev[0] + 1;
        const double c1 = 1.0 - pow(b1, (double)t), c2 = 1.0 - pow(b2, (double)t);
        sc[0] = (float)(lr * (sqrt(c2) / c1));
        sc[1] = (float)c1; sc[2] = (float)c2;
        sc[3] = (float)(1.0 - b1); sc[4] = (float)(1.0 - b2);
        sc[5] = (float)eps;
    }
    __syncthreads();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i];
    float mi = (float)b1 * m[i] + sc[3] * gi;
    float vi = (float)b2 * v[i] + sc[4] * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] -= sc[0] * (mi / sc[1]) / (sqrtf(vi / sc[2]) + sc[5]);
}
__global__ void incr_kernel(int *t) { t[0] += 1; }
// same as adam_dev_kernel but the counter has already been advanced: t = t_dev[0]
__global__ void adam_at_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                               float *__restrict__ v, long long n, const int *__restrict__ t_dev, double lr,
                               double b1, double b2, double eps)
{
    __shared__ float sc[6];
    if (threadIdx.x == 0) {
        const int t = t_dev[0];
        const double c1 = 1.0 - pow(b1, (double)t), c2 = 1.0 - pow(b2, (double)t);
        sc[0] = (float)(lr * (sqrt(c2) / c1));
        sc[1] = (float)c1; sc[2] = (float)c2;
        sc[3] = (float)(1.0 - b1); sc[4] = (float)(1.0 - b2);
        sc[5] = (float)eps;
    }
    __syncthreads();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i];
    float mi = (float)b1 * m[i] + sc[3] * gi;
    float vi = (float)b2 * v[i] + sc[4] * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] -= sc[0] * (mi / sc[1]) / (sqrtf(vi / sc[2]) + sc[5]);
}
__global__ void sgd_kernel(float *__restrict__ p, const float *__restrict__ g, long long n, float lr)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] -= lr * g[i];
}
} // namespace

int lnb_launch_adam(lnb_ctx *ctx, float *p, const float *g, float *m, float *v, long long n, int t,
                    double lr, double b1, double b2, double eps)
{
    if (n <= 0) return LNB_OK;
    // python-float scalars meet float32 arrays: each scalar is rounded to float32 once
    double c1 = 1.0 - pow(b1, t), c2 = 1.0 - pow(b2, t);
    double lr_t = lr * (sqrt(c2) / c1);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        p, g, m, v, n, (float)b1, (float)b2, (float)(1.0 - b1), (float)(1.0 - b2), (float)lr_t,
        (float)c1, (float)c2, (float)eps);
    LNB_CHECK_LAUNCH();
    return LNB_OK;
}

int lnb_launch_adam_dev(lnb_ctx *ctx, float *p, const float *g, float *m, float *v, long long n,
                        int *t_dev, double lr, double b1, double b2, double eps)
{
    if (n > 0) {
        adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, g, m, v, n, t_dev, lr, b1, b2, eps);
        LNB_CHECK_LAUNCH();
    }
    incr_kernel<<<1, 1, 0, ctx->stream>>>(t_dev);
    LNB_CHECK_LAUNCH();
    return LNB_OK;
}

int lnb_launch_incr(lnb_ctx *ctx, int *t_dev)
{
    incr_kernel<<<1, 1, 0, ctx->stream>>>(t_dev);
    LNB_CHECK_LAUNCH();
    return LNB_OK;
}

int lnb_launch_adam_at(lnb_ctx *ctx, float *p, const float *g, float *m, float *v, long long n,
                       const int *t_dev, double lr, double b1, double b2, double eps)
{
    if (n <= 0) return LNB_OK;
    adam_at_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, g, m, v, n, t_dev, lr, b1, b2, eps);
    LNB_CHECK_LAUNCH();
    return LNB_OK;
}

int lnb_launch_sgd(lnb_ctx *ctx, float *p, const float *g, long long n, double lr)
{
    if (n <= 0) return LNB_OK;
    sgd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, g, n, (float)lr);
    LNB_CHECK_LAUNCH();
    return LNB_OK;
}
